// conv1d.cu -- causal depthwise conv1d + bias + SiLU, channel-last, forward and backward.
//
// Replaces nn.Conv1d(groups=ED, kernel_size=K, padding=K-1)(x.transpose(1,2))[:, :, :L].transpose(1,2)
// followed by F.silu (cross_atten/mamba.py:128-131, 208-212): no transposes, the strided x half of
// in_proj's output is read in place, the K-wide window slides through registers so every input element
// is loaded once per time tile (+ K-1 halo).  lane = channel -> coalesced 64/128-byte rows.
//   forward : read xin, write u                      (2 * ED * s bytes / token)
//   backward: read xin, du, write dxin               (3 * ED * s bytes / token), dw/dbias via
//             per-(b, tile) partial sums + a deterministic finalize kernel.
#include "common.cuh"

namespace gfe {

constexpr int kConvTile = 64;   // time steps per thread

struct ConvParams {
    const void *xin, *du;
    void *u, *dxin;
    const float *w, *bias;
    float *dw, *dbias, *part;   // part: [B*ntiles][K+1][ED]
    int64_t x_bs, x_rs, u_bs, u_rs, du_bs, du_rs, dx_bs, dx_rs;
    int B, L, ED, ntiles;
};

template <typename T, int K>
__global__ void __launch_bounds__(128) conv1d_silu_fwd_kernel(ConvParams p) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.ED) return;
    const int tile = blockIdx.y, b = blockIdx.z;
    const int t0 = tile * kConvTile, t1 = min(p.L, t0 + kConvTile);
    const T *x = reinterpret_cast<const T *>(p.xin) + (int64_t)b * p.x_bs + c;
    T *u = reinterpret_cast<T *>(p.u) + (int64_t)b * p.u_bs + c;
    float w[K];
#pragma unroll
    for (int k = 0; k < K; ++k) w[k] = __ldg(p.w + (size_t)c * K + k);
    const float bias = p.bias ? __ldg(p.bias + c) : 0.f;

    float win[K];   // win[k] = xin[t - (K-1) + k]
#pragma unroll
    for (int k = 0; k < K - 1; ++k) {
        const int t = t0 - (K - 1) + k;
        win[k + 1] = t >= 0 ? to_f(ld_stream(x + (int64_t)t * p.x_rs)) : 0.f;
    }
    constexpr int U = 8;
    for (int tb = t0; tb < t1; tb += U) {
        T xr[U];
#pragma unroll
        for (int j = 0; j < U; ++j) xr[j] = ld_stream(x + (int64_t)min(tb + j, t1 - 1) * p.x_rs);
#pragma unroll
        for (int j = 0; j < U; ++j) {
            if (tb + j < t1) {
#pragma unroll
                for (int k = 0; k < K - 1; ++k) win[k] = win[k + 1];
                win[K - 1] = to_f(xr[j]);
                float v = bias;
#pragma unroll
                for (int k = 0; k < K; ++k) v = fmaf(w[k], win[k], v);
                st_stream(u + (int64_t)(tb + j) * p.u_rs, from_f<T>(v * sigmoid_fast(v)));
            }
        }
    }
}

template <typename T, int K>
__global__ void __launch_bounds__(128) conv1d_silu_bwd_kernel(ConvParams p) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.ED) return;
    const int tile = blockIdx.y, b = blockIdx.z;
    const int t0 = tile * kConvTile, t1 = min(p.L, t0 + kConvTile);
    const T *x = reinterpret_cast<const T *>(p.xin) + (int64_t)b * p.x_bs + c;
    const T *du = reinterpret_cast<const T *>(p.du) + (int64_t)b * p.du_bs + c;
    T *dx = reinterpret_cast<T *>(p.dxin) + (int64_t)b * p.dx_bs + c;
    float w[K];
#pragma unroll
    for (int k = 0; k < K; ++k) w[k] = __ldg(p.w + (size_t)c * K + k);
    const float bias = p.bias ? __ldg(p.bias + c) : 0.f;

    float win[K], dvw[K], dw[K], db = 0.f;   // dvw[k] = dv[t - (K-1) + k]
#pragma unroll
    for (int k = 0; k < K; ++k) { dvw[k] = 0.f; dw[k] = 0.f; }
#pragma unroll
    for (int k = 0; k < K - 1; ++k) {
        const int t = t0 - (K - 1) + k;
        win[k + 1] = t >= 0 ? to_f(ld_stream(x + (int64_t)t * p.x_rs)) : 0.f;
    }
    // dv is needed K-1 steps past the tile to finish dxin of the tile's last steps
    const int tend = t1 + (K - 1);
    constexpr int U = 8;
    for (int tb = t0; tb < tend; tb += U) {
        T xr[U], gr[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int t = min(tb + j, p.L - 1);
            xr[j] = ld_stream(x + (int64_t)t * p.x_rs);
            gr[j] = ld_stream(du + (int64_t)t * p.du_rs);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int t = tb + j;
            if (t < tend) {
#pragma unroll
                for (int k = 0; k < K - 1; ++k) { win[k] = win[k + 1]; dvw[k] = dvw[k + 1]; }
                float dv = 0.f;
                if (t < p.L) {
                    win[K - 1] = to_f(xr[j]);
                    float v = bias;
#pragma unroll
                    for (int k = 0; k < K; ++k) v = fmaf(w[k], win[k], v);
                    const float s = sigmoid_fast(v);
                    dv = to_f(gr[j]) * s * fmaf(v, 1.0f - s, 1.0f);
                    if (t < t1) {   // parameter gradients: own tile only
                        db += dv;
#pragma unroll
                        for (int k = 0; k < K; ++k) dw[k] = fmaf(dv, win[k], dw[k]);
                    }
                }
                dvw[K - 1] = dv;
                const int s_out = t - (K - 1);   // dxin[s] = sum_k w[k] dv[s + K-1 - k]
                if (s_out >= t0 && s_out < t1) {
                    float acc = 0.f;
#pragma unroll
                    for (int k = 0; k < K; ++k) acc = fmaf(w[k], dvw[K - 1 - k], acc);
                    st_stream(dx + (int64_t)s_out * p.dx_rs, from_f<T>(acc));
                }
            }
        }
    }
    float *part = p.part + ((size_t)(b * p.ntiles + tile) * (K + 1)) * p.ED + c;
#pragma unroll
    for (int k = 0; k < K; ++k) part[(size_t)k * p.ED] = dw[k];
    part[(size_t)K * p.ED] = db;
}

template <int K>
__global__ void __launch_bounds__(128) conv1d_bwd_finalize_kernel(ConvParams p) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;   // 0..K
    if (c >= p.ED) return;
    float acc = 0.f;
    const int n = p.B * p.ntiles;
    for (int i = 0; i < n; ++i) acc += p.part[((size_t)i * (K + 1) + k) * p.ED + c];
    if (k < K) p.dw[(size_t)c * K + k] = acc;
    else if (p.dbias) p.dbias[c] = acc;
}

template <typename T, int K>
static int conv_fwd_launch(ConvParams &p, cudaStream_t st) {
    const dim3 block(128), grid((p.ED + 127) / 128, p.ntiles, p.B);
    { ScopedKernelTimer tm(K_CONV_FWD, st);
      conv1d_silu_fwd_kernel<T, K><<<grid, block, 0, st>>>(p); }
    return check_launch("conv1d_silu_fwd");
}

template <typename T, int K>
static int conv_bwd_launch(ConvParams &p, cudaStream_t st) {
    const dim3 block(128), grid((p.ED + 127) / 128, p.ntiles, p.B);
    { ScopedKernelTimer tm(K_CONV_BWD, st);
      conv1d_silu_bwd_kernel<T, K><<<grid, block, 0, st>>>(p); }
    int rc = check_launch("conv1d_silu_bwd");
    if (rc != GFE_OK) return rc;
    { ScopedKernelTimer tm(K_CONV_BWD_FIN, st);
      conv1d_bwd_finalize_kernel<K><<<dim3((p.ED + 127) / 128, K + 1), 128, 0, st>>>(p); }
    return check_launch("conv1d_bwd_finalize");
}

template <typename T>
static int conv_dispatch_k(bool bwd, int K, ConvParams &p, cudaStream_t st) {
    switch (K) {
        case 2: return bwd ? conv_bwd_launch<T, 2>(p, st) : conv_fwd_launch<T, 2>(p, st);
        case 3: return bwd ? conv_bwd_launch<T, 3>(p, st) : conv_fwd_launch<T, 3>(p, st);
        case 4: return bwd ? conv_bwd_launch<T, 4>(p, st) : conv_fwd_launch<T, 4>(p, st);
        default: set_error("conv1d: d_conv=%d unsupported (2..4 compiled)", K); return GFE_ERR_UNSUPPORTED;
    }
}

static int conv_dispatch(bool bwd, int K, int dtype, ConvParams &p, cudaStream_t st) {
    switch (dtype) {
        case GFE_F32: return conv_dispatch_k<float>(bwd, K, p, st);
        case GFE_BF16: return conv_dispatch_k<__nv_bfloat16>(bwd, K, p, st);
        case GFE_F16: return conv_dispatch_k<__half>(bwd, K, p, st);
        default: set_error("conv1d: bad dtype %d", dtype); return GFE_ERR_DTYPE;
    }
}

}  // namespace gfe

extern "C" {

GFE_API int gfe_conv1d_silu_fwd(const void *xin, int64_t x_bs, int64_t x_rs, const float *w, const float *bias,
                                void *u, int64_t u_bs, int64_t u_rs, int B, int L, int ED, int K, int dtype,
                                void *stream) {
    using namespace gfe;
    if (!xin || !w || !u) { set_error("conv1d_fwd: NULL pointer"); return GFE_ERR_ARG; }
    if (B <= 0 || L <= 0 || ED <= 0 || B > 65535) { set_error("conv1d_fwd: bad shape"); return GFE_ERR_ARG; }
    ConvParams p{};
    p.xin = xin; p.u = u; p.w = w; p.bias = bias;
    p.x_bs = x_bs; p.x_rs = x_rs; p.u_bs = u_bs; p.u_rs = u_rs;
    p.B = B; p.L = L; p.ED = ED; p.ntiles = (L + kConvTile - 1) / kConvTile;
    return conv_dispatch(false, K, dtype, p, reinterpret_cast<cudaStream_t>(stream));
}

GFE_API size_t gfe_conv1d_bwd_workspace_bytes(int B, int L, int ED, int K) {
    if (B <= 0 || L <= 0 || ED <= 0 || K <= 0) return 0;
    const size_t ntiles = (L + gfe::kConvTile - 1) / gfe::kConvTile;
    return (size_t)B * ntiles * (K + 1) * ED * sizeof(float);
}

GFE_API int gfe_conv1d_silu_bwd(const void *xin, int64_t x_bs, int64_t x_rs, const float *w, const float *bias,
                                const void *du, int64_t du_bs, int64_t du_rs,
                                void *dxin, int64_t dx_bs, int64_t dx_rs, float *dw, float *dbias,
                                int B, int L, int ED, int K, int dtype, void *ws, size_t ws_bytes, void *stream) {
    using namespace gfe;
    if (!xin || !w || !du || !dxin || !dw) { set_error("conv1d_bwd: NULL pointer"); return GFE_ERR_ARG; }
    if (B <= 0 || L <= 0 || ED <= 0 || B > 65535) { set_error("conv1d_bwd: bad shape"); return GFE_ERR_ARG; }
    const size_t need = gfe_conv1d_bwd_workspace_bytes(B, L, ED, K);
    if (!ws || ws_bytes < need) { set_error("conv1d_bwd: workspace too small (%zu < %zu)", ws ? ws_bytes : (size_t)0, need); return GFE_ERR_WORKSPACE; }
    ConvParams p{};
    p.xin = xin; p.du = du; p.dxin = dxin; p.w = w; p.bias = bias; p.dw = dw; p.dbias = dbias;
    p.part = reinterpret_cast<float *>(ws);
    p.x_bs = x_bs; p.x_rs = x_rs; p.du_bs = du_bs; p.du_rs = du_rs; p.dx_bs = dx_bs; p.dx_rs = dx_rs;
    p.B = B; p.L = L; p.ED = ED; p.ntiles = (L + kConvTile - 1) / kConvTile;
    return conv_dispatch(true, K, dtype, p, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
