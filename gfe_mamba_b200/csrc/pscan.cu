// pscan.cu -- the reference's pscan(A, X) API (cross_atten/pscan.py:152-224) as HBM-streaming kernels.
//
// H[t] = A[t] * H[t-1] + X[t] per (b, d, n).  The reference runs a Blelloch tree with 2*log2(L) strided
// passes over two cloned, power-of-two padded tensors; here every thread owns 4 adjacent (d, n)
// elements (one 128-bit lane of the contiguous D*N axis), walks L sequentially with the running state
// in registers and touches every input/output element exactly once:
//   forward : read A, X            write H          (12 B / element)
//   backward: read A, dH, H        write dA, dX     (20 B / element)
// Padding to a power of two is unnecessary (it is appended after L-1 and never changes [0, L)).
// HBM needs ~12 MB of loads in flight; a launch has B*D*N*4*nloads*U bytes in flight (U = steps whose loads are issued before
// the first is used), so small problems run the SAME single pass with narrower vectors and a deeper unroll -- forward (V, U)
// = (4, 16) or (1, 32), backward (4, 8), (2, 16) or (1, 32), chosen from measurements (profiles/r02_pscan.txt) -- instead of
// paying a second pass.  Only when even
// that cannot fill the GPU (B*D*N below ~25 k elements, e.g. one long sequence with few channels) is L split into segments:
// pass 1 reduces each segment to its (product of A, local end state) pair, pass 2 starts each segment from the combined carry.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace gfe {

constexpr int kPsUnrollMax = 32;

struct PscanPlan {
    int nseg, seg_len;
    int V, U;   // elements per thread, steps per load batch
};

// nloads = tensors read per step: 2 forward (A, X), 3 backward (A, dH, H): backward reaches the same bytes in flight with
// fewer steps, so it keeps the wider vectors longer (measured at B=8, L=1024, D=1024: (4, 8) 98 % of the HBM peak, (2, 16) 88 %).
static PscanPlan pscan_plan(int B, int L, int64_t DN, int nloads) {
    PscanPlan p;
    double target = 12e6 * sm_count() / 148.0;                  // bytes in flight that saturate HBM (Little: ~6.5 TB/s x ~1.5 us)
#ifdef GFE_EXPERIMENTS
    if (const char *e = getenv("GFE_PSCAN_TARGET_MB")) target = atof(e) * 1e6;   // A/B measurements only
#endif
    const double per_u = (double)B * (double)DN * 4.0 * nloads;  // bytes requested per unrolled step
    if (nloads == 2) {   // forward: (4, 16) when that alone fills the pipe, else the narrowest vectors and the deepest unroll
        p.V = 4; p.U = 16;
        if (DN % 4 != 0 || per_u * 16 < target) { p.V = 1; p.U = 32; }
    } else {             // backward
        p.V = 4; p.U = 8;
        if (DN % 4 != 0 || per_u * 8 < target) { p.V = 2; p.U = 16; }
        if (DN % 2 != 0 || per_u * 16 < target) { p.V = 1; p.U = 32; }
    }
#ifdef GFE_EXPERIMENTS
    if (const char *e = getenv("GFE_PSCAN_VU")) { int v = 0, u = 0; if (sscanf(e, "%d,%d", &v, &u) == 2 && nloads == 2) { p.V = v; p.U = u; } }
#endif
    int S = 1;
    if (per_u * p.U * 2 < target) {   // still less than half of it: split L (re-reads A and X once more)
        S = (int)(target / (per_u * p.U));
        const int max_by_len = L / 64;
        if (S > max_by_len) S = max_by_len;
        if (S > kMaxSeg) S = kMaxSeg;
        if (S < 1) S = 1;
    }
    p.seg_len = (int)ceil_div64(ceil_div64(L, S), kPsUnrollMax) * kPsUnrollMax;
    p.nseg = (int)ceil_div64(L, p.seg_len);
    return p;
}

template <int V> struct Vec;
template <> struct Vec<4> {
    using type = float4;
    static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ float4 one() { return make_float4(1.f, 1.f, 1.f, 1.f); }
    static __device__ __forceinline__ float4 fma(float4 a, float4 b, float4 c) {
        return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
    }
    static __device__ __forceinline__ float4 mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
    static __device__ __forceinline__ float4 add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
};
template <> struct Vec<2> {
    using type = float2;
    static __device__ __forceinline__ float2 zero() { return make_float2(0.f, 0.f); }
    static __device__ __forceinline__ float2 one() { return make_float2(1.f, 1.f); }
    static __device__ __forceinline__ float2 fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
    static __device__ __forceinline__ float2 mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
    static __device__ __forceinline__ float2 add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
};
template <> struct Vec<1> {
    using type = float;
    static __device__ __forceinline__ float zero() { return 0.f; }
    static __device__ __forceinline__ float one() { return 1.f; }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
};

struct PscanParams {
    const float *A, *X, *H, *dH;
    float *Hout, *dA, *dX;
    float *segP, *segS;   // [B][nseg][DN]
    int B, L, nseg, seg_len;
    int64_t DN;           // D*N
};

// ---- forward -----------------------------------------------------------------------------------------
template <int V, int U, bool SUMMARY>
__global__ void __launch_bounds__(128) pscan_fwd_kernel(PscanParams p) {
    constexpr int kPsUnroll = U;
    using VT = typename Vec<V>::type;
    const int64_t nvec = p.DN / V;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvec) return;
    const int seg = blockIdx.y, b = blockIdx.z;
    const int t0 = seg * p.seg_len, t1 = min(p.L, t0 + p.seg_len);
    const VT *A = reinterpret_cast<const VT *>(p.A + (size_t)b * p.L * p.DN) + i;
    const VT *X = reinterpret_cast<const VT *>(p.X + (size_t)b * p.L * p.DN) + i;
    VT *H = SUMMARY ? nullptr : reinterpret_cast<VT *>(p.Hout + (size_t)b * p.L * p.DN) + i;

    VT h = Vec<V>::zero(), P = Vec<V>::one();
    if (!SUMMARY) {
        for (int s = 0; s < seg; ++s) {
            const size_t o = ((size_t)b * p.nseg + s) * p.DN;
            h = Vec<V>::fma(reinterpret_cast<const VT *>(p.segP + o)[i], h, reinterpret_cast<const VT *>(p.segS + o)[i]);
        }
    }
    for (int tb = t0; tb < t1; tb += kPsUnroll) {   // (double-buffered load batches were measured slower: 150+ registers)
        VT a[kPsUnroll], x[kPsUnroll];
#pragma unroll
        for (int j = 0; j < kPsUnroll; ++j) {
            const int t = min(tb + j, t1 - 1);
            a[j] = __ldcs(A + (size_t)t * nvec);
            x[j] = __ldcs(X + (size_t)t * nvec);
        }
#pragma unroll
        for (int j = 0; j < kPsUnroll; ++j) {
            if (tb + j < t1) {
                h = Vec<V>::fma(a[j], h, x[j]);
                if (SUMMARY) P = Vec<V>::mul(P, a[j]);
                else __stcs(H + (size_t)(tb + j) * nvec, h);
            }
        }
    }
    if (SUMMARY) {
        const size_t o = ((size_t)b * p.nseg + seg) * p.DN;
        reinterpret_cast<VT *>(p.segP + o)[i] = P;
        reinterpret_cast<VT *>(p.segS + o)[i] = h;
    }
}

// ---- backward ----------------------------------------------------------------------------------------
// G(t) = A[t] * g[t] is the carry handed to step t-1;  g[t] = dH[t] + G(t+1).
template <int V, int U, bool SUMMARY>
__global__ void __launch_bounds__(128) pscan_bwd_kernel(PscanParams p) {
    constexpr int kPsUnroll = U;
    using VT = typename Vec<V>::type;
    const int64_t nvec = p.DN / V;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvec) return;
    const int seg = SUMMARY ? blockIdx.y + 1 : blockIdx.y, b = blockIdx.z;
    const int t0 = seg * p.seg_len, t1 = min(p.L, t0 + p.seg_len);
    const VT *A = reinterpret_cast<const VT *>(p.A + (size_t)b * p.L * p.DN) + i;
    const VT *dH = reinterpret_cast<const VT *>(p.dH + (size_t)b * p.L * p.DN) + i;
    const VT *H = reinterpret_cast<const VT *>(p.H + (size_t)b * p.L * p.DN) + i;
    VT *dA = SUMMARY ? nullptr : reinterpret_cast<VT *>(p.dA + (size_t)b * p.L * p.DN) + i;
    VT *dX = SUMMARY ? nullptr : reinterpret_cast<VT *>(p.dX + (size_t)b * p.L * p.DN) + i;

    VT G = Vec<V>::zero(), P = Vec<V>::one();
    if (!SUMMARY) {
        for (int s = p.nseg - 1; s > seg; --s) {
            const size_t o = ((size_t)b * p.nseg + s) * p.DN;
            G = Vec<V>::fma(reinterpret_cast<const VT *>(p.segP + o)[i], G, reinterpret_cast<const VT *>(p.segS + o)[i]);
        }
    }
    // walk [t0, t1) backwards in blocks of kPsUnroll aligned to t0
    const int nblk = (t1 - t0 + kPsUnroll - 1) / kPsUnroll;
    for (int blk = nblk - 1; blk >= 0; --blk) {
        const int tb = t0 + blk * kPsUnroll;
        VT a[kPsUnroll], g[kPsUnroll], hp[kPsUnroll];
#pragma unroll
        for (int j = 0; j < kPsUnroll; ++j) {
            const int t = min(tb + j, t1 - 1);
            a[j] = __ldcs(A + (size_t)t * nvec);
            g[j] = __ldcs(dH + (size_t)t * nvec);
            if (!SUMMARY) hp[j] = t > 0 ? __ldcs(H + (size_t)(t - 1) * nvec) : Vec<V>::zero();
        }
#pragma unroll
        for (int j = kPsUnroll - 1; j >= 0; --j) {
            if (tb + j < t1) {
                const VT gt = Vec<V>::add(g[j], G);
                if (!SUMMARY) {
                    __stcs(dX + (size_t)(tb + j) * nvec, gt);
                    __stcs(dA + (size_t)(tb + j) * nvec, Vec<V>::mul(hp[j], gt));
                } else {
                    P = Vec<V>::mul(P, a[j]);
                }
                G = Vec<V>::mul(a[j], gt);
            }
        }
    }
    if (SUMMARY) {
        const size_t o = ((size_t)b * p.nseg + seg) * p.DN;
        reinterpret_cast<VT *>(p.segP + o)[i] = P;
        reinterpret_cast<VT *>(p.segS + o)[i] = G;
    }
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int pscan_check(const void *a, const void *b, const void *c, int B, int L, int D, int N) {
    if (!a || !b || !c) { set_error("pscan: NULL pointer"); return GFE_ERR_ARG; }
    if (B <= 0 || L <= 0 || D <= 0 || N <= 0) { set_error("pscan: non-positive shape"); return GFE_ERR_ARG; }
    if (B > 65535) { set_error("pscan: batch %d exceeds grid limit", B); return GFE_ERR_UNSUPPORTED; }
    return GFE_OK;
}

}  // namespace gfe

extern "C" {

GFE_API size_t gfe_pscan_workspace_bytes(int B, int L, int D, int N) {
    if (B <= 0 || L <= 0 || D <= 0 || N <= 0) return 0;
    const int64_t DN = (int64_t)D * N;
    const gfe::PscanPlan pl = gfe::pscan_plan(B, L, DN, 2);   // the forward plan never has fewer segments than the backward plan
    if (pl.nseg <= 1) return 0;
    return 2 * gfe::align_up((size_t)B * pl.nseg * DN * sizeof(float), 256);
}

GFE_API int gfe_pscan_fwd(const float *A, const float *X, float *H, int B, int L, int D, int N,
                          void *ws, size_t ws_bytes, void *stream) {
    using namespace gfe;
    int rc = pscan_check(A, X, H, B, L, D, N);
    if (rc != GFE_OK) return rc;
    const int64_t DN = (int64_t)D * N;
    const PscanPlan pl = pscan_plan(B, L, DN, 2);
    const size_t need = gfe_pscan_workspace_bytes(B, L, D, N);
    if (need > 0 && (ws == nullptr || ws_bytes < need)) { set_error("pscan_fwd: workspace too small (%zu < %zu)", ws ? ws_bytes : (size_t)0, need); return GFE_ERR_WORKSPACE; }
    PscanParams p{};
    p.A = A; p.X = X; p.Hout = H; p.B = B; p.L = L; p.DN = DN; p.nseg = pl.nseg; p.seg_len = pl.seg_len;
    p.segP = reinterpret_cast<float *>(ws);
    p.segS = reinterpret_cast<float *>(reinterpret_cast<char *>(ws) + need / 2);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool al = aligned16(A) && aligned16(X) && aligned16(H) && (need == 0 || aligned16(ws));
    const int V = al ? pl.V : 1;                                  // rows of D*N fp32 keep 8 / 16-byte alignment when DN % V == 0
    const int64_t nvec = DN / V;
    const dim3 block(128), grid((unsigned)ceil_div64(nvec, 128), pl.nseg, B);
    const bool deep = pl.U == 16;   // (4, 16): twice the loads in flight at the widest request
    if (pl.nseg > 1) {
        { ScopedKernelTimer tm(K_PSCAN_FWD_SUMMARY, st);
          if (V == 4 && deep) pscan_fwd_kernel<4, 16, true><<<grid, block, 0, st>>>(p);
          else if (V == 4) pscan_fwd_kernel<4, 8, true><<<grid, block, 0, st>>>(p);
          else if (V == 2) pscan_fwd_kernel<2, 16, true><<<grid, block, 0, st>>>(p);
          else pscan_fwd_kernel<1, 32, true><<<grid, block, 0, st>>>(p); }
        rc = check_launch("pscan_fwd_summary");
        if (rc != GFE_OK) return rc;
    }
    { ScopedKernelTimer tm(K_PSCAN_FWD, st);
      if (V == 4 && deep) pscan_fwd_kernel<4, 16, false><<<grid, block, 0, st>>>(p);
      else if (V == 4) pscan_fwd_kernel<4, 8, false><<<grid, block, 0, st>>>(p);
      else if (V == 2) pscan_fwd_kernel<2, 16, false><<<grid, block, 0, st>>>(p);
      else pscan_fwd_kernel<1, 32, false><<<grid, block, 0, st>>>(p); }
    return check_launch("pscan_fwd");
}

GFE_API int gfe_pscan_bwd(const float *A, const float *H, const float *dH, float *dA, float *dX,
                          int B, int L, int D, int N, void *ws, size_t ws_bytes, void *stream) {
    using namespace gfe;
    int rc = pscan_check(A, H, dH, B, L, D, N);
    if (rc != GFE_OK) return rc;
    if (!dA || !dX) { set_error("pscan_bwd: NULL output pointer"); return GFE_ERR_ARG; }
    const int64_t DN = (int64_t)D * N;
    const PscanPlan pl = pscan_plan(B, L, DN, 3);
    const size_t need = gfe_pscan_workspace_bytes(B, L, D, N);
    if (need > 0 && (ws == nullptr || ws_bytes < need)) { set_error("pscan_bwd: workspace too small (%zu < %zu)", ws ? ws_bytes : (size_t)0, need); return GFE_ERR_WORKSPACE; }
    PscanParams p{};
    p.A = A; p.H = H; p.dH = dH; p.dA = dA; p.dX = dX; p.B = B; p.L = L; p.DN = DN; p.nseg = pl.nseg; p.seg_len = pl.seg_len;
    p.segP = reinterpret_cast<float *>(ws);
    p.segS = reinterpret_cast<float *>(reinterpret_cast<char *>(ws) + need / 2);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool al = aligned16(A) && aligned16(H) && aligned16(dH) && aligned16(dA) && aligned16(dX) && (need == 0 || aligned16(ws));
    const int V = al ? pl.V : 1;
    const int64_t nvec = DN / V;
    const dim3 block(128);
    if (pl.nseg > 1) {
        const dim3 grid((unsigned)ceil_div64(nvec, 128), pl.nseg - 1, B);
        { ScopedKernelTimer tm(K_PSCAN_BWD_SUMMARY, st);
          if (V == 4) pscan_bwd_kernel<4, 8, true><<<grid, block, 0, st>>>(p);
          else if (V == 2) pscan_bwd_kernel<2, 16, true><<<grid, block, 0, st>>>(p);
          else pscan_bwd_kernel<1, 32, true><<<grid, block, 0, st>>>(p); }
        rc = check_launch("pscan_bwd_summary");
        if (rc != GFE_OK) return rc;
    }
    const dim3 grid((unsigned)ceil_div64(nvec, 128), pl.nseg, B);
    { ScopedKernelTimer tm(K_PSCAN_BWD, st);
      if (V == 4) pscan_bwd_kernel<4, 8, false><<<grid, block, 0, st>>>(p);
      else if (V == 2) pscan_bwd_kernel<2, 16, false><<<grid, block, 0, st>>>(p);
      else pscan_bwd_kernel<1, 32, false><<<grid, block, 0, st>>>(p); }
    return check_launch("pscan_bwd");
}

}  // extern "C"
