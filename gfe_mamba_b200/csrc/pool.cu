// pool.cu -- final residual add fused with the mean over L.
//
// The classifier head pools the Mamba stack's output over the sequence (mamba_transformer.py:122-123:
//     x = self.transformer(x); x = torch.mean(x, dim=1, keepdims=True)),
// and the stack's last operation is the residual add of its last layer (mamba.py:103).  Unfused that is one pass writing the
// (B, L, D) sum and one pass reading it back; here the two (B, L, D) operands are read once and only (B, D) is written:
//   fwd   out[b, d] = (1 / L) * sum_t (a[b, t, d] + r[b, t, d])          reads 2 B L D, writes B D
//   bwd   da[b, t, d] = dr[b, t, d] = dout[b, d] / L                      writes B L D once (both operands share it)
// Deterministic: per-(row tile) partial sums in fp32, added in tile order by a second tiny kernel.
#include "common.cuh"

namespace gfe {

constexpr int kPoolTile = 64;   // rows per partial sum
constexpr int kPoolNT = 128;

template <typename T>
__global__ void __launch_bounds__(kPoolNT) add_mean_pool_partial_kernel(const T *__restrict__ a, const T *__restrict__ r, float *__restrict__ part,
                                                                         int L, int D, int ntile) {
    const int d = blockIdx.x * kPoolNT + threadIdx.x;
    const int tile = blockIdx.y, b = blockIdx.z;
    if (d >= D) return;
    const int t0 = tile * kPoolTile, t1 = min(L, t0 + kPoolTile);
    const T *pa = a + ((size_t)b * L + t0) * D + d;
    const T *pr = r ? r + ((size_t)b * L + t0) * D + d : nullptr;
    float acc0 = 0.f, acc1 = 0.f;
    int t = t0;
    for (; t + 1 < t1; t += 2) {   // two independent chains: the loads of both rows are in flight together
        acc0 += to_f(ld_stream(pa)) + (pr ? to_f(ld_stream(pr)) : 0.f);
        acc1 += to_f(ld_stream(pa + D)) + (pr ? to_f(ld_stream(pr + D)) : 0.f);
        pa += 2 * (size_t)D;
        if (pr) pr += 2 * (size_t)D;
    }
    if (t < t1) acc0 += to_f(ld_stream(pa)) + (pr ? to_f(ld_stream(pr)) : 0.f);
    part[((size_t)b * ntile + tile) * D + d] = acc0 + acc1;
}

template <typename T>
__global__ void __launch_bounds__(kPoolNT) add_mean_pool_final_kernel(const float *__restrict__ part, T *__restrict__ out, int L, int D, int ntile) {
    const int d = blockIdx.x * kPoolNT + threadIdx.x, b = blockIdx.y;
    if (d >= D) return;
    float acc = 0.f;
    for (int i = 0; i < ntile; ++i) acc += part[((size_t)b * ntile + i) * D + d];
    out[(size_t)b * D + d] = from_f<T>(acc / (float)L);
}

template <typename T>
__global__ void __launch_bounds__(kPoolNT) mean_pool_bwd_kernel(const T *__restrict__ dout, T *__restrict__ da, int L, int D) {
    const int d = blockIdx.x * kPoolNT + threadIdx.x;
    const int tile = blockIdx.y, b = blockIdx.z;
    if (d >= D) return;
    const T v = from_f<T>(to_f(dout[(size_t)b * D + d]) / (float)L);
    const int t0 = tile * kPoolTile, t1 = min(L, t0 + kPoolTile);
    T *p = da + ((size_t)b * L + t0) * D + d;
    for (int t = t0; t < t1; ++t, p += D) st_stream(p, v);
}

template <typename T>
static int pool_fwd_t(const void *a, const void *r, void *out, int B, int L, int D, float *part, cudaStream_t st) {
    const int ntile = (L + kPoolTile - 1) / kPoolTile;
    { ScopedKernelTimer tm(K_POOL_FWD, st);
      add_mean_pool_partial_kernel<T><<<dim3((D + kPoolNT - 1) / kPoolNT, ntile, B), kPoolNT, 0, st>>>(
          reinterpret_cast<const T *>(a), reinterpret_cast<const T *>(r), part, L, D, ntile);
      add_mean_pool_final_kernel<T><<<dim3((D + kPoolNT - 1) / kPoolNT, B), kPoolNT, 0, st>>>(part, reinterpret_cast<T *>(out), L, D, ntile); }
    return check_launch("add_mean_pool_fwd");
}

template <typename T>
static int pool_bwd_t(const void *dout, void *da, int B, int L, int D, cudaStream_t st) {
    const int ntile = (L + kPoolTile - 1) / kPoolTile;
    { ScopedKernelTimer tm(K_POOL_BWD, st);
      mean_pool_bwd_kernel<T><<<dim3((D + kPoolNT - 1) / kPoolNT, ntile, B), kPoolNT, 0, st>>>(reinterpret_cast<const T *>(dout), reinterpret_cast<T *>(da), L, D); }
    return check_launch("mean_pool_bwd");
}

}  // namespace gfe

extern "C" {

GFE_API size_t gfe_add_mean_pool_workspace_bytes(int B, int L, int D) {
    if (B <= 0 || L <= 0 || D <= 0) return 0;
    return (size_t)B * ((L + gfe::kPoolTile - 1) / gfe::kPoolTile) * D * sizeof(float);
}

GFE_API int gfe_add_mean_pool_fwd(const void *a, const void *r, void *out, int B, int L, int D, int dtype, void *ws, size_t ws_bytes,
                                  void *stream) {
    using namespace gfe;
    if (!a || !out || B <= 0 || L <= 0 || D <= 0 || B > 65535) { set_error("add_mean_pool_fwd: bad argument"); return GFE_ERR_ARG; }
    if (!ws || ws_bytes < gfe_add_mean_pool_workspace_bytes(B, L, D)) { set_error("add_mean_pool_fwd: workspace too small"); return GFE_ERR_WORKSPACE; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (dtype) {
        case GFE_F32: return pool_fwd_t<float>(a, r, out, B, L, D, reinterpret_cast<float *>(ws), st);
        case GFE_BF16: return pool_fwd_t<__nv_bfloat16>(a, r, out, B, L, D, reinterpret_cast<float *>(ws), st);
        case GFE_F16: return pool_fwd_t<__half>(a, r, out, B, L, D, reinterpret_cast<float *>(ws), st);
        default: set_error("add_mean_pool_fwd: bad dtype %d", dtype); return GFE_ERR_DTYPE;
    }
}

GFE_API int gfe_mean_pool_bwd(const void *dout, void *da, int B, int L, int D, int dtype, void *stream) {
    using namespace gfe;
    if (!dout || !da || B <= 0 || L <= 0 || D <= 0 || B > 65535) { set_error("mean_pool_bwd: bad argument"); return GFE_ERR_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (dtype) {
        case GFE_F32: return pool_bwd_t<float>(dout, da, B, L, D, st);
        case GFE_BF16: return pool_bwd_t<__nv_bfloat16>(dout, da, B, L, D, st);
        case GFE_F16: return pool_bwd_t<__half>(dout, da, B, L, D, st);
        default: set_error("mean_pool_bwd: bad dtype %d", dtype); return GFE_ERR_DTYPE;
    }
}

}  // extern "C"
