// addnorm.cu -- fused residual add + RMSNorm, forward and backward.
// Replaces, per layer, the residual add of ResidualBlock.forward (mamba.py:103: mixer(norm(x)) + x) fused with the NEXT
// layer's RMSNorm (mamba.py:408-418: x * rsqrt(mean(x^2) + eps) * weight):
//     resid = x (+ a)                      one pass over the residual stream instead of add kernel + pow + mean + ...
//     y     = (resid * rstd) * w,  rstd = rsqrt(mean(resid^2) + eps)
// Backward (dy from the mixer, dres = gradient already flowing down the residual stream):
//     g = dy * w;  dx = dres + rstd * (g - resid * rstd^2 * mean(g * resid));  dw = sum_rows dy * resid * rstd.
// HBM-bound: forward reads x, a and writes resid, y (4 D s bytes per row), backward reads resid, dy, dres and writes dx.
// One warp per row, the row cached in registers (D <= 32 * 8 vectors of 16 bytes), 16-byte coalesced accesses, warp
// shuffle reductions; dw is accumulated per lane over a grid-stride loop of rows, reduced per CTA in shared memory and
// written as one partial row per CTA; a second tiny kernel adds the partial rows (deterministic, no atomics).
#include "common.cuh"

namespace gfe {

constexpr int kNormWarps = 8;          // rows in flight per CTA
constexpr int kNormMaxVec = 8;         // 16-byte vectors per lane held in registers

template <typename T> struct Vec16 { static constexpr int n = 16 / (int)sizeof(T); };

// N consecutive elements as ONE access of N * sizeof(T) bytes (16, or 8 when a 16-bit branch tensor rides along an fp32 stream)
template <int BYTES> struct RawN;
template <> struct RawN<16> { using type = uint4; };
template <> struct RawN<8> { using type = uint2; };
template <typename T, int N>
__device__ __forceinline__ void loadN(const T *p, float (&v)[N]) {
    using R = typename RawN<N * (int)sizeof(T)>::type;
    const R raw = *reinterpret_cast<const R *>(p);
    const T *e = reinterpret_cast<const T *>(&raw);
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = to_f(e[i]);
}
template <typename T, int N>
__device__ __forceinline__ void storeN(T *p, const float (&v)[N]) {
    using R = typename RawN<N * (int)sizeof(T)>::type;
    R raw;
    T *e = reinterpret_cast<T *>(&raw);
#pragma unroll
    for (int i = 0; i < N; ++i) e[i] = from_f<T>(v[i]);
    *reinterpret_cast<R *>(p) = raw;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// NV = vectors per lane (compile-time), D == any multiple of the vector width with D <= 32 * NV * VEC.
// TR = element type of the residual stream (x, resid; backward: dres, dx), TB = element type of the branch side (a, y;
// backward: dy, da).  TB == TR is the plain case; TR = float with a 16-bit TB is the stack under autocast: the fp32 stream
// takes the mixer's 16-bit output and hands the next mixer a 16-bit normed input with no cast kernels in between.
template <typename TR, typename TB, int NV, bool HAS_A>
__global__ void __launch_bounds__(32 * kNormWarps) add_rmsnorm_fwd_kernel(const TR *__restrict__ x, const TB *__restrict__ a,
                                                                          const float *__restrict__ w, TR *__restrict__ resid,
                                                                          TB *__restrict__ y, float *__restrict__ rstd_out,
                                                                          int64_t rows, int D, float eps) {
    constexpr int VEC = Vec16<TR>::n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = D / VEC;
    float wv[NV][VEC];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + 32 * i;
#pragma unroll
        for (int e = 0; e < VEC; ++e) wv[i][e] = v < nvec ? __ldg(w + v * VEC + e) : 0.f;
    }
    for (int64_t r = (int64_t)blockIdx.x * kNormWarps + warp; r < rows; r += (int64_t)gridDim.x * kNormWarps) {
        float xv[NV][VEC];
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + 32 * i;
            if (v < nvec) {
                loadN<TR, VEC>(x + r * D + v * VEC, xv[i]);
                if (HAS_A) {
                    float av[VEC];
                    loadN<TB, VEC>(a + r * D + v * VEC, av);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) xv[i][e] = to_f(from_f<TR>(xv[i][e] + av[e]));   // the stream is stored in TR
                    storeN<TR, VEC>(resid + r * D + v * VEC, xv[i]);
                }
#pragma unroll
                for (int e = 0; e < VEC; ++e) ss = fmaf(xv[i][e], xv[i][e], ss);
            }
        }
        ss = warp_sum(ss);
        const float rstd = rsqrtf(ss / (float)D + eps);
        if (lane == 0 && rstd_out != nullptr) rstd_out[r] = rstd;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + 32 * i;
            if (v < nvec) {
                float o[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) o[e] = (xv[i][e] * rstd) * wv[i][e];   // the reference's order: (x * rstd) * w
                storeN<TB, VEC>(y + r * D + v * VEC, o);
            }
        }
    }
}

// da (optional): the gradient of the branch operand `a` in ITS element type -- the same values as dx, rounded -- so that a
// 16-bit branch gets its gradient from this pass instead of a cast kernel over dx.
template <typename TR, typename TB, int NV, bool HAS_DRES>
__global__ void __launch_bounds__(32 * kNormWarps) add_rmsnorm_bwd_kernel(const TR *__restrict__ resid, const float *__restrict__ w,
                                                                          const float *__restrict__ rstd_in, const TB *__restrict__ dy,
                                                                          const TR *__restrict__ dres, TR *__restrict__ dx,
                                                                          TB *__restrict__ da, float *__restrict__ dw_part, int64_t rows, int D) {
    constexpr int VEC = Vec16<TR>::n;
    __shared__ float s_dw[kNormWarps][32 * VEC + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = D / VEC;
    float dwv[NV][VEC];   // w is re-read through the read-only path every row (L1 hit): registers go to x, g and dw
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int e = 0; e < VEC; ++e) dwv[i][e] = 0.f;
    for (int64_t r = (int64_t)blockIdx.x * kNormWarps + warp; r < rows; r += (int64_t)gridDim.x * kNormWarps) {
        float xv[NV][VEC], gv[NV][VEC];
        uint4 dv_raw[HAS_DRES ? NV : 1];   // the third stream is requested together with the other two (kept packed)
        const float rstd = rstd_in[r];
        float dot = 0.f;
        if (HAS_DRES) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int v = lane + 32 * i;
                if (v < nvec) dv_raw[i] = *reinterpret_cast<const uint4 *>(dres + r * D + v * VEC);
            }
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + 32 * i;
            if (v < nvec) {
                loadN<TR, VEC>(resid + r * D + v * VEC, xv[i]);
                loadN<TB, VEC>(dy + r * D + v * VEC, gv[i]);
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    dwv[i][e] = fmaf(gv[i][e], xv[i][e] * rstd, dwv[i][e]);
                    gv[i][e] *= __ldg(w + v * VEC + e);
                    dot = fmaf(gv[i][e], xv[i][e], dot);
                }
            }
        }
        dot = warp_sum(dot);
        const float c = dot / (float)D * rstd * rstd;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + 32 * i;
            if (v < nvec) {
                float o[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) o[e] = rstd * (gv[i][e] - xv[i][e] * c);
                if (HAS_DRES) {
                    const TR *dv = reinterpret_cast<const TR *>(&dv_raw[i]);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) o[e] += to_f(dv[e]);
                }
                storeN<TR, VEC>(dx + r * D + v * VEC, o);
                if (da != nullptr) storeN<TB, VEC>(da + r * D + v * VEC, o);
            }
        }
    }
    // dw: add the CTA's warps, one partial row per CTA
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        __syncthreads();
#pragma unroll
        for (int e = 0; e < VEC; ++e) s_dw[warp][lane * VEC + e] = dwv[i][e];
        __syncthreads();
        for (int j = threadIdx.x; j < 32 * VEC; j += 32 * kNormWarps) {
            const int col = (32 * i) * VEC + j;   // vector (lane' + 32 i), element e -> column (lane' + 32 i) * VEC + e
            if (col < D) {
                float s = 0.f;
#pragma unroll
                for (int wq = 0; wq < kNormWarps; ++wq) s += s_dw[wq][j];
                dw_part[(size_t)blockIdx.x * D + col] = s;
            }
        }
    }
}

// ---- rows wider than the register-resident form (D > 256 vectors): two passes over the row, the second one from L1/L2 ----
// Same arithmetic and element types; one warp per row.  dw comes from a separate column pass (dw_wide) because per-lane
// accumulators for a whole row no longer fit in registers.
template <typename TR, typename TB, bool HAS_A>
__global__ void __launch_bounds__(32 * kNormWarps) add_rmsnorm_fwd_wide_kernel(const TR *__restrict__ x, const TB *__restrict__ a,
                                                                               const float *__restrict__ w, TR *__restrict__ resid,
                                                                               TB *__restrict__ y, float *__restrict__ rstd_out,
                                                                               int64_t rows, int D, float eps) {
    constexpr int VEC = Vec16<TR>::n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = D / VEC;
    for (int64_t r = (int64_t)blockIdx.x * kNormWarps + warp; r < rows; r += (int64_t)gridDim.x * kNormWarps) {
        float ss = 0.f;
        for (int v = lane; v < nvec; v += 32) {
            float xv[VEC];
            loadN<TR, VEC>(x + r * D + v * VEC, xv);
            if (HAS_A) {
                float av[VEC];
                loadN<TB, VEC>(a + r * D + v * VEC, av);
#pragma unroll
                for (int e = 0; e < VEC; ++e) xv[e] = to_f(from_f<TR>(xv[e] + av[e]));
                storeN<TR, VEC>(resid + r * D + v * VEC, xv);
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) ss = fmaf(xv[e], xv[e], ss);
        }
        ss = warp_sum(ss);
        const float rstd = rsqrtf(ss / (float)D + eps);
        if (lane == 0 && rstd_out != nullptr) rstd_out[r] = rstd;
        __syncwarp();   // this warp's stores of resid are read back below by other lanes' loops only through the same addresses
        const TR *src = HAS_A ? resid : x;
        for (int v = lane; v < nvec; v += 32) {
            float xv[VEC], o[VEC];
            loadN<TR, VEC>(src + r * D + v * VEC, xv);
#pragma unroll
            for (int e = 0; e < VEC; ++e) o[e] = (xv[e] * rstd) * __ldg(w + v * VEC + e);
            storeN<TB, VEC>(y + r * D + v * VEC, o);
        }
    }
}

template <typename TR, typename TB, bool HAS_DRES>
__global__ void __launch_bounds__(32 * kNormWarps) add_rmsnorm_bwd_wide_kernel(const TR *__restrict__ resid, const float *__restrict__ w,
                                                                               const float *__restrict__ rstd_in, const TB *__restrict__ dy,
                                                                               const TR *__restrict__ dres, TR *__restrict__ dx,
                                                                               TB *__restrict__ da, int64_t rows, int D) {
    constexpr int VEC = Vec16<TR>::n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = D / VEC;
    for (int64_t r = (int64_t)blockIdx.x * kNormWarps + warp; r < rows; r += (int64_t)gridDim.x * kNormWarps) {
        const float rstd = rstd_in[r];
        float dot = 0.f;
        for (int v = lane; v < nvec; v += 32) {
            float xv[VEC], gv[VEC];
            loadN<TR, VEC>(resid + r * D + v * VEC, xv);
            loadN<TB, VEC>(dy + r * D + v * VEC, gv);
#pragma unroll
            for (int e = 0; e < VEC; ++e) dot = fmaf(gv[e] * __ldg(w + v * VEC + e), xv[e], dot);
        }
        dot = warp_sum(dot);
        const float c = dot / (float)D * rstd * rstd;
        for (int v = lane; v < nvec; v += 32) {
            float xv[VEC], gv[VEC], o[VEC];
            loadN<TR, VEC>(resid + r * D + v * VEC, xv);
            loadN<TB, VEC>(dy + r * D + v * VEC, gv);
#pragma unroll
            for (int e = 0; e < VEC; ++e) o[e] = rstd * (gv[e] * __ldg(w + v * VEC + e) - xv[e] * c);
            if (HAS_DRES) {
                float dv[VEC];
                loadN<TR, VEC>(dres + r * D + v * VEC, dv);
#pragma unroll
                for (int e = 0; e < VEC; ++e) o[e] += dv[e];
            }
            storeN<TR, VEC>(dx + r * D + v * VEC, o);
            if (da != nullptr) storeN<TB, VEC>(da + r * D + v * VEC, o);
        }
    }
}

// partial rows of dw[j] = sum_r dy[r, j] resid[r, j] rstd[r]: block (32 columns, 8 row lanes), gridDim.y row slices
template <typename TR, typename TB>
__global__ void __launch_bounds__(256) add_rmsnorm_dw_wide_kernel(const TR *__restrict__ resid, const float *__restrict__ rstd,
                                                                  const TB *__restrict__ dy, float *__restrict__ part, int64_t rows, int D) {
    __shared__ float s_acc[8][33];
    const int col = blockIdx.x * 32 + threadIdx.x;
    const int64_t per = (rows + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = (int64_t)blockIdx.y * per, r1 = r0 + per < rows ? r0 + per : rows;
    float s = 0.f;
    if (col < D)
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s = fmaf(to_f(dy[r * D + col]) * to_f(resid[r * D + col]), rstd[r], s);
    s_acc[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && col < D) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += s_acc[q][threadIdx.x];
        part[(size_t)blockIdx.y * D + col] = t;
    }
}

// block (32 columns, 8 slices of the partial rows): coalesced 128-byte reads, 8-way split of the serial sum
__global__ void add_rmsnorm_dw_finalize_kernel(const float *__restrict__ part, float *__restrict__ dw, int nparts, int D) {
    __shared__ float s_acc[8][33];
    const int col = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (col < D)
        for (int p = threadIdx.y; p < nparts; p += 8) s += part[(size_t)p * D + col];
    s_acc[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && col < D) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += s_acc[q][threadIdx.x];
        dw[col] = t;
    }
}

static int norm_grid(int64_t rows) {
    const int64_t want = (rows + kNormWarps - 1) / kNormWarps;
    const int64_t cap = (int64_t)sm_count() * 4;   // persistent: 4 CTAs of 8 warps per SM (fewer partial dw rows to add)
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

template <typename T>
static int nv_for(int D) {
    const int VEC = Vec16<T>::n;
    if (D % VEC != 0) return 0;
    const int nvec = D / VEC;
    for (int nv : {1, 2, 3, 4, 6, 8})
        if (nvec <= 32 * nv) return nv;
    return 0;
}

template <typename TR, typename TB>
static int launch_fwd_t(const void *x, const void *a, const float *w, void *resid, void *y, float *rstd, int64_t rows, int D,
                        float eps, cudaStream_t st) {
    const int nv = nv_for<TR>(D);
    if (D % Vec16<TR>::n != 0) {
        set_error("add_rmsnorm: d_model=%d must be a multiple of %d for this dtype", D, Vec16<TR>::n);
        return GFE_ERR_ARG;
    }
    const int grid = norm_grid(rows);
    ScopedKernelTimer tm(K_ADDNORM_FWD, st);
    if (nv == 0) {   // wider than the register-resident form
        if (a) add_rmsnorm_fwd_wide_kernel<TR, TB, true><<<grid, 32 * kNormWarps, 0, st>>>((const TR *)x, (const TB *)a, w, (TR *)resid, (TB *)y, rstd, rows, D, eps);
        else add_rmsnorm_fwd_wide_kernel<TR, TB, false><<<grid, 32 * kNormWarps, 0, st>>>((const TR *)x, nullptr, w, nullptr, (TB *)y, rstd, rows, D, eps);
        return check_launch("add_rmsnorm_fwd (wide)");
    }
#define GFE_NF(NVv)                                                                                                          \
    if (a) add_rmsnorm_fwd_kernel<TR, TB, NVv, true><<<grid, 32 * kNormWarps, 0, st>>>((const TR *)x, (const TB *)a, w, (TR *)resid, (TB *)y, rstd, rows, D, eps); \
    else add_rmsnorm_fwd_kernel<TR, TB, NVv, false><<<grid, 32 * kNormWarps, 0, st>>>((const TR *)x, nullptr, w, nullptr, (TB *)y, rstd, rows, D, eps)
    switch (nv) {
        case 1: GFE_NF(1); break;
        case 2: GFE_NF(2); break;
        case 3: GFE_NF(3); break;
        case 4: GFE_NF(4); break;
        case 6: GFE_NF(6); break;
        default: GFE_NF(8); break;
    }
#undef GFE_NF
    return check_launch("add_rmsnorm_fwd");
}

template <typename TR, typename TB>
static int launch_bwd_t(const void *resid, const float *w, const float *rstd, const void *dy, const void *dres, void *dx, void *da,
                        float *dw, int64_t rows, int D, void *ws, size_t ws_bytes, cudaStream_t st) {
    const int nv = nv_for<TR>(D);
    if (D % Vec16<TR>::n != 0) {
        set_error("add_rmsnorm: d_model=%d must be a multiple of %d for this dtype", D, Vec16<TR>::n);
        return GFE_ERR_ARG;
    }
    const int grid = norm_grid(rows);
    const size_t need = (size_t)grid * D * sizeof(float);
    if (ws == nullptr || ws_bytes < need) {
        set_error("add_rmsnorm_bwd: workspace too small (%zu < %zu)", ws ? ws_bytes : (size_t)0, need);
        return GFE_ERR_WORKSPACE;
    }
    float *part = reinterpret_cast<float *>(ws);
    if (nv == 0) {   // wider than the register-resident form: dx (+ da) per row, dw from a column pass over row slices
        const int slices = grid < 64 ? grid : 64;
        {
            ScopedKernelTimer tm(K_ADDNORM_BWD, st);
            if (dres) add_rmsnorm_bwd_wide_kernel<TR, TB, true><<<grid, 32 * kNormWarps, 0, st>>>((const TR *)resid, w, rstd, (const TB *)dy, (const TR *)dres, (TR *)dx, (TB *)da, rows, D);
            else add_rmsnorm_bwd_wide_kernel<TR, TB, false><<<grid, 32 * kNormWarps, 0, st>>>((const TR *)resid, w, rstd, (const TB *)dy, nullptr, (TR *)dx, (TB *)da, rows, D);
            add_rmsnorm_dw_wide_kernel<TR, TB><<<dim3((D + 31) / 32, slices), dim3(32, 8), 0, st>>>((const TR *)resid, rstd, (const TB *)dy, part, rows, D);
        }
        int rcw = check_launch("add_rmsnorm_bwd (wide)");
        if (rcw != GFE_OK) return rcw;
        {
            ScopedKernelTimer tm(K_ADDNORM_BWD_FIN, st);
            add_rmsnorm_dw_finalize_kernel<<<(D + 31) / 32, dim3(32, 8), 0, st>>>(part, dw, slices, D);
        }
        return check_launch("add_rmsnorm_dw_finalize");
    }
    {
        ScopedKernelTimer tm(K_ADDNORM_BWD, st);
#define GFE_NB(NVv)                                                                                                          \
    if (dres) add_rmsnorm_bwd_kernel<TR, TB, NVv, true><<<grid, 32 * kNormWarps, 0, st>>>((const TR *)resid, w, rstd, (const TB *)dy, (const TR *)dres, (TR *)dx, (TB *)da, part, rows, D); \
    else add_rmsnorm_bwd_kernel<TR, TB, NVv, false><<<grid, 32 * kNormWarps, 0, st>>>((const TR *)resid, w, rstd, (const TB *)dy, nullptr, (TR *)dx, (TB *)da, part, rows, D)
        switch (nv) {
            case 1: GFE_NB(1); break;
            case 2: GFE_NB(2); break;
            case 3: GFE_NB(3); break;
            case 4: GFE_NB(4); break;
            case 6: GFE_NB(6); break;
            default: GFE_NB(8); break;
        }
#undef GFE_NB
    }
    int rc = check_launch("add_rmsnorm_bwd");
    if (rc != GFE_OK) return rc;
    {
        ScopedKernelTimer tm(K_ADDNORM_BWD_FIN, st);
        add_rmsnorm_dw_finalize_kernel<<<(D + 31) / 32, dim3(32, 8), 0, st>>>(part, dw, grid, D);
    }
    return check_launch("add_rmsnorm_dw_finalize");
}

}  // namespace gfe

using namespace gfe;

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" {

GFE_API size_t gfe_add_rmsnorm_bwd_workspace_bytes(int64_t rows, int D) {
    return (size_t)norm_grid(rows) * (size_t)(D > 0 ? D : 0) * sizeof(float);
}

GFE_API int gfe_add_rmsnorm_fwd_mixed(const void *x, const void *a, const float *w, void *resid, void *y, float *rstd, int64_t rows,
                                      int D, float eps, int res_dtype, int io_dtype, void *stream) {
    if (x == nullptr || w == nullptr || y == nullptr || rows < 0 || D <= 0 || (a != nullptr && resid == nullptr)) {
        set_error("add_rmsnorm_fwd: bad argument");
        return GFE_ERR_ARG;
    }
    if (!aligned16(x) || !aligned16(a) || !aligned16(resid) || !aligned16(y)) {
        set_error("add_rmsnorm_fwd: tensors must be 16-byte aligned and row-contiguous");
        return GFE_ERR_ARG;
    }
    if (rows == 0) return GFE_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (res_dtype == io_dtype) {
        switch (res_dtype) {
            case GFE_F32: return launch_fwd_t<float, float>(x, a, w, resid, y, rstd, rows, D, eps, st);
            case GFE_BF16: return launch_fwd_t<__nv_bfloat16, __nv_bfloat16>(x, a, w, resid, y, rstd, rows, D, eps, st);
            case GFE_F16: return launch_fwd_t<__half, __half>(x, a, w, resid, y, rstd, rows, D, eps, st);
            default: break;
        }
    } else if (res_dtype == GFE_F32) {
        if (io_dtype == GFE_BF16) return launch_fwd_t<float, __nv_bfloat16>(x, a, w, resid, y, rstd, rows, D, eps, st);
        if (io_dtype == GFE_F16) return launch_fwd_t<float, __half>(x, a, w, resid, y, rstd, rows, D, eps, st);
    }
    set_error("add_rmsnorm_fwd: unsupported dtype pair (stream %d, branch %d)", res_dtype, io_dtype);
    return GFE_ERR_DTYPE;
}

GFE_API int gfe_add_rmsnorm_fwd(const void *x, const void *a, const float *w, void *resid, void *y, float *rstd, int64_t rows,
                                int D, float eps, int dtype, void *stream) {
    return gfe_add_rmsnorm_fwd_mixed(x, a, w, resid, y, rstd, rows, D, eps, dtype, dtype, stream);
}

GFE_API int gfe_add_rmsnorm_bwd_mixed(const void *resid, const float *w, const float *rstd, const void *dy, const void *dres, void *dx,
                                      void *da, float *dw, int64_t rows, int D, int res_dtype, int io_dtype, void *ws,
                                      size_t ws_bytes, void *stream) {
    if (resid == nullptr || w == nullptr || rstd == nullptr || dy == nullptr || dx == nullptr || dw == nullptr || rows <= 0 || D <= 0) {
        set_error("add_rmsnorm_bwd: bad argument");
        return GFE_ERR_ARG;
    }
    if (!aligned16(resid) || !aligned16(dy) || !aligned16(dres) || !aligned16(dx) || !aligned16(da)) {
        set_error("add_rmsnorm_bwd: tensors must be 16-byte aligned and row-contiguous");
        return GFE_ERR_ARG;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (res_dtype == io_dtype) {
        switch (res_dtype) {
            case GFE_F32: return launch_bwd_t<float, float>(resid, w, rstd, dy, dres, dx, da, dw, rows, D, ws, ws_bytes, st);
            case GFE_BF16: return launch_bwd_t<__nv_bfloat16, __nv_bfloat16>(resid, w, rstd, dy, dres, dx, da, dw, rows, D, ws, ws_bytes, st);
            case GFE_F16: return launch_bwd_t<__half, __half>(resid, w, rstd, dy, dres, dx, da, dw, rows, D, ws, ws_bytes, st);
            default: break;
        }
    } else if (res_dtype == GFE_F32) {
        if (io_dtype == GFE_BF16) return launch_bwd_t<float, __nv_bfloat16>(resid, w, rstd, dy, dres, dx, da, dw, rows, D, ws, ws_bytes, st);
        if (io_dtype == GFE_F16) return launch_bwd_t<float, __half>(resid, w, rstd, dy, dres, dx, da, dw, rows, D, ws, ws_bytes, st);
    }
    set_error("add_rmsnorm_bwd: unsupported dtype pair (stream %d, branch %d)", res_dtype, io_dtype);
    return GFE_ERR_DTYPE;
}

GFE_API int gfe_add_rmsnorm_bwd(const void *resid, const float *w, const float *rstd, const void *dy, const void *dres, void *dx,
                                float *dw, int64_t rows, int D, int dtype, void *ws, size_t ws_bytes, void *stream) {
    return gfe_add_rmsnorm_bwd_mixed(resid, w, rstd, dy, dres, dx, nullptr, dw, rows, D, dtype, dtype, ws, ws_bytes, stream);
}

}  // extern "C"
