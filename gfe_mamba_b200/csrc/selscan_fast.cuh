// selscan_fast.cuh -- the tuned selective-scan kernels (included by selscan.cu).
//
// Same math and work decomposition as the generic kernels in selscan.cu (lane = channel, warp = 32
// channels, 16-step chunks, state pairs as float2), restructured around what the first ncu captures
// showed (profiles/r01_*): the generic kernels issued ~2.4x the necessary instructions (64-bit address
// math, per-step branches), ran at IPC 0.25-0.37 with one warp per scheduler, and overflowed the
// instruction cache.  Here:
//   * every per-channel tensor (u, delta, z, dout) and the B|C rows travel HBM -> shared memory with
//     cp.async (16-byte pieces when pointers/strides allow, else 4-byte), two chunk stages per warp, so the
//     next chunk is in flight while the current one is computed and no register holds prefetched data;
//   * the step loop is rolled (groups of 4 steps) and reads its scalars from the staged tile, which keeps
//     the hot loop inside the instruction cache and leaves the scheduler 4 independent steps to overlap
//     MUFU latency with the FFMA2 chains;
//   * no divergent branches in the hot loop: softplus picks its log1p formulation per 4-step group with a
//     warp-uniform vote, padding steps are handled by selects (delta = 0 -> a = 1, bx = 0);
//   * forward: a configurable number of state pairs take exp2 on the FMA pipe (degree-5 polynomial,
//     packed FFMA2) instead of MUFU.EX2, because MUFU (0.5 warp-instr/clk/SM, profiles/r01_microbench_pipes.txt)
//     is the binding pipe of the forward recurrence.
// Preconditions (checked on the host, else the generic kernels run): ED % 32 == 0; for 16-bit activations
// even strides and 4-byte aligned bases.
#pragma once

#include "common.cuh"

namespace gfe {

// ---- cp.async helpers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst, const void *src) {
    if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Copy up to kChunk rows of ROW_ELEMS contiguous elements (row stride rs elements) into a dense shared tile.
template <typename T, int CPB, int ROW_ELEMS>
__device__ __forceinline__ void tile_issue(uint32_t dst, const T *src, int64_t rs, int nrows, int lane, int dst_row_bytes,
                                           int dst_col_byte_off) {
    constexpr int RB = ROW_ELEMS * (int)sizeof(T);   // bytes per source row
    constexpr int PPR = RB / CPB;                     // pieces per row
    constexpr int TOTAL = kChunk * PPR;
    const char *s = reinterpret_cast<const char *>(src);
#pragma unroll
    for (int i0 = 0; i0 < TOTAL; i0 += 32) {
        const int i = i0 + lane;
        const int row = i / PPR, piece = i % PPR;
        if ((TOTAL % 32 == 0 || i < TOTAL) && row < nrows)
            cp_async<CPB>(dst + row * dst_row_bytes + dst_col_byte_off + piece * CPB, s + (int64_t)row * rs * (int64_t)sizeof(T) + piece * CPB);
    }
}

// exp2 of two non-positive arguments on the FMA pipe: Cody-Waite split + degree-5 minimax (max rel err 2.3e-7
// in fp32, same class as ex2.approx), exponent inserted with integer adds.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
    x.x = fmaxf(x.x, -125.0f);
    x.y = fmaxf(x.y, -125.0f);
    const float2 magic = make_float2(12582912.0f, 12582912.0f);
    const float2 t = fadd2(x, magic);
    const float2 n = fadd2(t, make_float2(-12582912.0f, -12582912.0f));
    const float2 f = fadd2(x, make_float2(-n.x, -n.y));
    float2 p = make_float2(0.001327647129073739f, 0.001327647129073739f);
    p = ffma2(p, f, make_float2(0.009675540961325169f, 0.009675540961325169f));
    p = ffma2(p, f, make_float2(0.05550713092088699f, 0.05550713092088699f));
    p = ffma2(p, f, make_float2(0.24022120237350464f, 0.24022120237350464f));
    p = ffma2(p, f, make_float2(0.6931469440460205f, 0.6931469440460205f));
    p = ffma2(p, f, make_float2(1.0000001192092896f, 1.0000001192092896f));
    return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
                       __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}

// log1p(e) for 0 <= e < 0.5 via 2 atanh(e / (2 + e))
__device__ __forceinline__ float log1p_small(float e) {
    const float s = e * rcp_approx(2.0f + e);
    const float s2 = s * s;
    float p = fmaf(s2, 1.0f / 9.0f, 1.0f / 7.0f);
    p = fmaf(s2, p, 1.0f / 5.0f);
    p = fmaf(s2, p, 1.0f / 3.0f);
    p = fmaf(s2, p, 1.0f);
    return 2.0f * s * p;
}

// softplus for a group of G steps, branch-free per lane; the log1p formulation is chosen by a warp-uniform vote.
// x[i] -> dl[i]; optionally sig[i] = sigmoid(x[i]).
template <int G, bool WANT_SIG>
__device__ __forceinline__ void softplus_group(const float (&x)[G], float (&dl)[G], float (&sig)[G]) {
    float e[G];
    bool any_big = false;
#pragma unroll
    for (int i = 0; i < G; ++i) {
        e[i] = ex2_approx(fminf(x[i], 30.0f) * kLog2e);
        any_big |= e[i] >= 0.5f;
    }
#pragma unroll
    for (int i = 0; i < G; ++i) dl[i] = log1p_small(fminf(e[i], 0.5f));
    if (__any_sync(0xffffffffu, any_big)) {
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const float big = kLn2 * lg2_approx(1.0f + e[i]);
            dl[i] = e[i] >= 0.5f ? big : dl[i];
        }
    }
#pragma unroll
    for (int i = 0; i < G; ++i) {
        if (WANT_SIG) sig[i] = x[i] > 20.0f ? 1.0f : e[i] * rcp_approx(1.0f + e[i]);
        dl[i] = x[i] > 20.0f ? x[i] : dl[i];
    }
}

constexpr int kFwdPolyPairs = 2;   // state pairs whose exp2 runs on the FMA pipe in the forward kernel
constexpr int kGroup = 4;          // steps per rolled-loop iteration

template <typename T> struct FastCfg {
    static constexpr int kTileBytes = kChunk * 32 * (int)sizeof(T);   // one staged (16 x 32) tile
};

// =====================================================================================================
// Forward
// =====================================================================================================
template <typename T, bool HAS_Z>
__host__ __device__ constexpr int fwd_fast_smem_per_warp() {
    // 2 stages x (u, delta, [z], B|C raw) + fp32 B|C tile when T is 16-bit
    return 2 * ((HAS_Z ? 4 : 3) * FastCfg<T>::kTileBytes) + (sizeof(T) == 4 ? 0 : kChunk * 32 * 4);
}

template <typename T, bool HAS_Z, int CPB>
__global__ void __launch_bounds__(128) selscan_fwd_fast_kernel(ScanParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TILE = FastCfg<T>::kTileBytes;
    constexpr int NT = HAS_Z ? 4 : 3;            // tiles per stage: u, delta, [z], bc
    constexpr int STAGE = NT * TILE;
    constexpr bool CONVERT_BC = sizeof(T) != 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * (blockDim.x >> 5) + warp;
    if (g * 32 >= p.ED) return;
    const int c = g * 32 + lane;                 // ED % 32 == 0: every lane owns a real channel
    const int seg = blockIdx.y, b = blockIdx.z;
    const int t0 = seg * p.seg_len, t1 = min(p.L, t0 + p.seg_len);

    unsigned char *sw = smem_raw + (size_t)warp * fwd_fast_smem_per_warp<T, HAS_Z>();
    const uint32_t sw_u32 = smem_u32(sw);
    float *sBCf = CONVERT_BC ? reinterpret_cast<float *>(sw + 2 * STAGE) : nullptr;

    const T *ub = reinterpret_cast<const T *>(p.u) + (int64_t)b * p.u_bs + g * 32;
    const T *db = reinterpret_cast<const T *>(p.delta) + (int64_t)b * p.d_bs + g * 32;
    const T *zb = HAS_Z ? reinterpret_cast<const T *>(p.z) + (int64_t)b * p.z_bs + g * 32 : nullptr;
    const T *Bb = reinterpret_cast<const T *>(p.Bm) + (int64_t)b * p.B_bs;
    const T *Cb = reinterpret_cast<const T *>(p.Cm) + (int64_t)b * p.C_bs;
    T *ob = reinterpret_cast<T *>(p.out) + (int64_t)b * p.o_bs + c;
    const bool sp = p.flags & GFE_FLAG_DELTA_SOFTPLUS;
    const float bias = p.dt_bias ? __ldg(p.dt_bias + c) : 0.0f;
    const float Dc = __ldg(p.D + c);

    auto issue = [&](int tb, int stage) {
        const int nrows = min(kChunk, t1 - tb);
        const uint32_t s = sw_u32 + stage * STAGE;
        tile_issue<T, CPB, 32>(s, ub + (int64_t)tb * p.u_rs, p.u_rs, nrows, lane, 32 * sizeof(T), 0);
        tile_issue<T, CPB, 32>(s + TILE, db + (int64_t)tb * p.d_rs, p.d_rs, nrows, lane, 32 * sizeof(T), 0);
        if (HAS_Z) tile_issue<T, CPB, 32>(s + 2 * TILE, zb + (int64_t)tb * p.z_rs, p.z_rs, nrows, lane, 32 * sizeof(T), 0);
        tile_issue<T, CPB, 16>(s + (NT - 1) * TILE, Bb + (int64_t)tb * p.B_rs, p.B_rs, nrows, lane, 32 * sizeof(T), 0);
        tile_issue<T, CPB, 16>(s + (NT - 1) * TILE, Cb + (int64_t)tb * p.C_rs, p.C_rs, nrows, lane, 32 * sizeof(T), 16 * sizeof(T));
        cp_async_commit();
    };

    issue(t0, 0);
    if (t0 + kChunk < t1) issue(t0 + kChunk, 1);
    else cp_async_commit();

    float2 A2[kPairs], h[kPairs];
    load_A2(A2, p.A_log, c);
#pragma unroll
    for (int q = 0; q < kPairs; ++q) h[q] = make_float2(0.f, 0.f);
    for (int s = 0; s < seg; ++s) {   // carry-in from earlier segments
        const float sd = p.seg_sd[(size_t)(b * p.nseg + s) * p.ED + c];
        const float2 *src = p.seg_h + ((size_t)(b * p.nseg + s) * kPairs) * p.ED + c;
        const float2 sd2 = splat2(sd);
#pragma unroll
        for (int q = 0; q < kPairs; ++q) h[q] = ffma2(ex2_2(fmul2(sd2, A2[q])), h[q], src[(size_t)q * p.ED]);
    }

    int stage = 0;
    for (int tb = t0; tb < t1; tb += kChunk, stage ^= 1) {
        cp_async_wait<1>();
        __syncwarp();
        const unsigned char *st = sw + stage * STAGE;
        const T *sU = reinterpret_cast<const T *>(st);
        const T *sD = reinterpret_cast<const T *>(st + TILE);
        const T *sZ = reinterpret_cast<const T *>(st + 2 * TILE);
        const float *sBC;
        if (CONVERT_BC) {
            const T *raw = reinterpret_cast<const T *>(st + (NT - 1) * TILE);
#pragma unroll
            for (int j = 0; j < kChunk; ++j) sBCf[j * 32 + lane] = to_f(raw[j * 32 + lane]);
            __syncwarp();
            sBC = sBCf;
        } else {
            sBC = reinterpret_cast<const float *>(st + (NT - 1) * TILE);
        }

        if (p.ckpt != nullptr) {   // state at the start of this chunk, for backward
            float2 *dst = p.ckpt + ((size_t)(b * p.nchunks + tb / kChunk) * kPairs) * p.ED + c;
#pragma unroll
            for (int q = 0; q < kPairs; ++q) __stcs(dst + (size_t)q * p.ED, h[q]);
        }

#pragma unroll 1
        for (int j0 = 0; j0 < kChunk; j0 += kGroup) {
            float x[kGroup], uj[kGroup], dl[kGroup], gate[kGroup], sgdummy[kGroup];
#pragma unroll
            for (int i = 0; i < kGroup; ++i) {
                const int j = j0 + i;
                x[i] = to_f(sD[j * 32 + lane]) + bias;
                uj[i] = to_f(sU[j * 32 + lane]);
                if (HAS_Z) {
                    const float zj = to_f(sZ[j * 32 + lane]);
                    gate[i] = zj * sigmoid_fast(zj);
                }
            }
            if (sp) {
                softplus_group<kGroup, false>(x, dl, sgdummy);
            } else {
#pragma unroll
                for (int i = 0; i < kGroup; ++i) dl[i] = x[i];
            }
#pragma unroll
            for (int i = 0; i < kGroup; ++i) {
                const bool valid = tb + j0 + i < t1;   // warp-uniform
                dl[i] = valid ? dl[i] : 0.f;
                uj[i] = valid ? uj[i] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < kGroup; ++i) {
                const int j = j0 + i;
                const float2 dl2 = splat2(dl[i]), du2 = splat2(dl[i] * uj[i]);
                float2 y2 = make_float2(0.f, 0.f);
                const float4 *sb = reinterpret_cast<const float4 *>(sBC + j * 32);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const float4 Bq = sb[q4], Cq = sb[4 + q4];
                    const float2 x0 = fmul2(dl2, A2[2 * q4]), x1 = fmul2(dl2, A2[2 * q4 + 1]);
                    const float2 a0 = (2 * q4 < kFwdPolyPairs) ? ex2_poly2(x0) : ex2_2(x0);
                    const float2 a1 = (2 * q4 + 1 < kFwdPolyPairs) ? ex2_poly2(x1) : ex2_2(x1);
                    h[2 * q4] = ffma2(a0, h[2 * q4], fmul2(du2, make_float2(Bq.x, Bq.y)));
                    y2 = ffma2(h[2 * q4], make_float2(Cq.x, Cq.y), y2);
                    h[2 * q4 + 1] = ffma2(a1, h[2 * q4 + 1], fmul2(du2, make_float2(Bq.z, Bq.w)));
                    y2 = ffma2(h[2 * q4 + 1], make_float2(Cq.z, Cq.w), y2);
                }
                float y = fmaf(Dc, uj[i], y2.x + y2.y);
                if (HAS_Z) y *= gate[i];
                if (tb + j < t1) st_stream(ob + (int64_t)(tb + j) * p.o_rs, from_f<T>(y));
            }
        }
        __syncwarp();   // every lane is done with this stage before it is refilled
        if (tb + 2 * kChunk < t1) issue(tb + 2 * kChunk, stage);
        else cp_async_commit();
    }

    if (p.last_state != nullptr && seg == p.nseg - 1) {
        float2 *dst = reinterpret_cast<float2 *>(p.last_state + ((size_t)b * p.ED + c) * kNState);
#pragma unroll
        for (int q = 0; q < kPairs; ++q) dst[q] = h[q];
    }
}

// =====================================================================================================
// Backward
// =====================================================================================================
template <typename T, bool HAS_Z>
__host__ __device__ constexpr int bwd_fast_smem_per_warp() {
    // 2 stages x (u, delta, dout, [z], B|C raw) + fp32 B|C tile (16-bit only) + reduced tile
    // + 4 per-(pair, lane) float2 arrays (A, G, dA, H) + 2 per-(step, lane) float arrays (f, sigmoid)
    return 2 * ((HAS_Z ? 5 : 4) * FastCfg<T>::kTileBytes) + (sizeof(T) == 4 ? 0 : kChunk * 32 * 4) +
           kChunk * kRedStride * 4 + 4 * (kPairs * 32 * 8) + 2 * (kChunk * 32 * 4);
}

template <typename T, bool HAS_Z, int CPB>
__global__ void __launch_bounds__(128) selscan_bwd_fast_kernel(ScanParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TILE = FastCfg<T>::kTileBytes;
    constexpr int NT = HAS_Z ? 5 : 4;            // tiles per stage: u, delta, dout, [z], bc
    constexpr int STAGE = NT * TILE;
    constexpr bool CONVERT_BC = sizeof(T) != 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * (blockDim.x >> 5) + warp;
    if (g * 32 >= p.ED) return;
    const int c = g * 32 + lane;
    const int seg = blockIdx.y, b = blockIdx.z;
    const int t0 = seg * p.seg_len, t1 = min(p.L, t0 + p.seg_len);

    unsigned char *sw = smem_raw + (size_t)warp * bwd_fast_smem_per_warp<T, HAS_Z>();
    const uint32_t sw_u32 = smem_u32(sw);
    unsigned char *cur = sw + 2 * STAGE;
    float *sBCf = reinterpret_cast<float *>(cur);
    if (CONVERT_BC) cur += kChunk * 32 * 4;
    float *sRed = reinterpret_cast<float *>(cur);
    cur += kChunk * kRedStride * 4;
    float2 *sA = reinterpret_cast<float2 *>(cur);
    float2 *sG = sA + kPairs * 32;
    float2 *sdA = sG + kPairs * 32;
    float2 *sH = sdA + kPairs * 32;
    float *sF = reinterpret_cast<float *>(sH + kPairs * 32);
    float *sSg = sF + kChunk * 32;

    const T *ub = reinterpret_cast<const T *>(p.u) + (int64_t)b * p.u_bs + g * 32;
    const T *db = reinterpret_cast<const T *>(p.delta) + (int64_t)b * p.d_bs + g * 32;
    const T *gb = reinterpret_cast<const T *>(p.dout) + (int64_t)b * p.do_bs + g * 32;
    const T *zb = HAS_Z ? reinterpret_cast<const T *>(p.z) + (int64_t)b * p.z_bs + g * 32 : nullptr;
    const T *Bb = reinterpret_cast<const T *>(p.Bm) + (int64_t)b * p.B_bs;
    const T *Cb = reinterpret_cast<const T *>(p.Cm) + (int64_t)b * p.C_bs;
    T *dub = reinterpret_cast<T *>(p.du) + (int64_t)b * p.du_bs + c;
    T *ddb = reinterpret_cast<T *>(p.ddelta) + (int64_t)b * p.dd_bs + c;
    T *dzb = HAS_Z ? reinterpret_cast<T *>(p.dz) + (int64_t)b * p.dz_bs + c : nullptr;
    const bool sp = p.flags & GFE_FLAG_DELTA_SOFTPLUS;
    const float bias = p.dt_bias ? __ldg(p.dt_bias + c) : 0.0f;
    const float Dc = __ldg(p.D + c);

    const int first_chunk = t0 / kChunk, last_chunk = (t1 - 1) / kChunk;

    auto issue = [&](int k, int stage) {
        const int tb = k * kChunk;
        const int nrows = min(kChunk, t1 - tb);
        const uint32_t s = sw_u32 + stage * STAGE;
        tile_issue<T, CPB, 32>(s, ub + (int64_t)tb * p.u_rs, p.u_rs, nrows, lane, 32 * sizeof(T), 0);
        tile_issue<T, CPB, 32>(s + TILE, db + (int64_t)tb * p.d_rs, p.d_rs, nrows, lane, 32 * sizeof(T), 0);
        tile_issue<T, CPB, 32>(s + 2 * TILE, gb + (int64_t)tb * p.do_rs, p.do_rs, nrows, lane, 32 * sizeof(T), 0);
        if (HAS_Z) tile_issue<T, CPB, 32>(s + 3 * TILE, zb + (int64_t)tb * p.z_rs, p.z_rs, nrows, lane, 32 * sizeof(T), 0);
        tile_issue<T, CPB, 16>(s + (NT - 1) * TILE, Bb + (int64_t)tb * p.B_rs, p.B_rs, nrows, lane, 32 * sizeof(T), 0);
        tile_issue<T, CPB, 16>(s + (NT - 1) * TILE, Cb + (int64_t)tb * p.C_rs, p.C_rs, nrows, lane, 32 * sizeof(T), 16 * sizeof(T));
        cp_async_commit();
    };
    issue(last_chunk, 0);
    if (last_chunk - 1 >= first_chunk) issue(last_chunk - 1, 1);
    else cp_async_commit();

    {   // per-lane constants and carries
        const float4 *row = reinterpret_cast<const float4 *>(p.A_log + (size_t)c * kNState);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = __ldg(row + q);
            sA[(2 * q) * 32 + lane] = make_float2(-expf(v.x), -expf(v.y));
            sA[(2 * q + 1) * 32 + lane] = make_float2(-expf(v.z), -expf(v.w));
        }
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
            sG[q * 32 + lane] = make_float2(0.f, 0.f);
            sdA[q * 32 + lane] = make_float2(0.f, 0.f);
        }
        for (int s = p.nseg - 1; s > seg; --s) {
            const float sd = p.seg_sd[(size_t)(b * p.nseg + s) * p.ED + c];
            const float2 *src = p.seg_h + ((size_t)(b * p.nseg + s) * kPairs) * p.ED + c;
            const float2 sd2 = splat2(sd * kLog2e);
#pragma unroll
            for (int q = 0; q < kPairs; ++q)
                sG[q * 32 + lane] = ffma2(ex2_2(fmul2(sd2, sA[q * 32 + lane])), sG[q * 32 + lane], src[(size_t)q * p.ED]);
        }
    }
    float dD_acc = 0.f, dbias_acc = 0.f;

    int stage = 0;
    for (int k = last_chunk; k >= first_chunk; --k, stage ^= 1) {
        const int tb = k * kChunk;
        {   // this chunk's checkpoint: issue the loads before waiting on the tile
            const float2 *ck = p.ckpt + ((size_t)(b * p.nchunks + k) * kPairs) * p.ED + c;
            float2 hk[kPairs];
#pragma unroll
            for (int q = 0; q < kPairs; ++q) hk[q] = __ldcs(ck + (size_t)q * p.ED);
            cp_async_wait<1>();
            __syncwarp();
#pragma unroll
            for (int q = 0; q < kPairs; ++q) sH[q * 32 + lane] = hk[q];
        }
        const unsigned char *st = sw + stage * STAGE;
        const T *sU = reinterpret_cast<const T *>(st);
        const T *sD = reinterpret_cast<const T *>(st + TILE);
        const T *sDo = reinterpret_cast<const T *>(st + 2 * TILE);
        const T *sZ = reinterpret_cast<const T *>(st + 3 * TILE);
        const float *sBC;
        if (CONVERT_BC) {
            const T *raw = reinterpret_cast<const T *>(st + (NT - 1) * TILE);
#pragma unroll
            for (int j = 0; j < kChunk; ++j) sBCf[j * 32 + lane] = to_f(raw[j * 32 + lane]);
            sBC = sBCf;
        } else {
            sBC = reinterpret_cast<const float *>(st + (NT - 1) * TILE);
        }

        // ---- per-step scalars of this lane's channel ----
        float dl[kChunk], dlu[kChunk], dy[kChunk];
#pragma unroll
        for (int j0 = 0; j0 < kChunk; j0 += kGroup) {
            float x[kGroup], dlg[kGroup], sg[kGroup];
#pragma unroll
            for (int i = 0; i < kGroup; ++i) x[i] = to_f(sD[(j0 + i) * 32 + lane]) + bias;
            if (sp) {
                softplus_group<kGroup, true>(x, dlg, sg);
            } else {
#pragma unroll
                for (int i = 0; i < kGroup; ++i) { dlg[i] = x[i]; sg[i] = 1.0f; }
            }
#pragma unroll
            for (int i = 0; i < kGroup; ++i) {
                const int j = j0 + i;
                const bool valid = tb + j < t1;
                const float uj = valid ? to_f(sU[j * 32 + lane]) : 0.f;
                const float doj = valid ? to_f(sDo[j * 32 + lane]) : 0.f;
                dl[j] = valid ? dlg[i] : 0.f;
                dlu[j] = dl[j] * uj;
                float f = 0.f;
                if (HAS_Z) {
                    const float zj = to_f(sZ[j * 32 + lane]);
                    const float sz = sigmoid_fast(zj);
                    dy[j] = doj * (zj * sz);
                    f = doj * sz * fmaf(zj, 1.0f - sz, 1.0f);
                } else {
                    dy[j] = doj;
                }
                sF[j * 32 + lane] = f;
                sSg[j * 32 + lane] = sg[i];
            }
        }
        __syncwarp();   // converted B|C tile visible to all lanes

        float S1[kChunk], S2[kChunk], yv[kChunk];
#pragma unroll
        for (int j = 0; j < kChunk; ++j) S1[j] = S2[j] = yv[j] = 0.f;

        // ---- one state pair at a time: forward sweep (recompute), reverse sweep (gradients) ----
#pragma unroll 1
        for (int q = 0; q < kPairs; ++q) {
            const float2 Aq = sA[q * 32 + lane];
            const float2 A2q = fmul2(Aq, splat2(kLog2e));
            const float *bq = sBC + 2 * q;
            float2 h = sH[q * 32 + lane];
            float2 a[kChunk], hp[kChunk];
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                const float2 Bq = *reinterpret_cast<const float2 *>(bq + j * 32);
                const float2 Cq = *reinterpret_cast<const float2 *>(bq + j * 32 + 16);
                a[j] = ex2_2(fmul2(splat2(dl[j]), A2q));
                hp[j] = h;
                h = ffma2(a[j], h, fmul2(splat2(dlu[j]), Bq));
                yv[j] = fmaf(h.x, Cq.x, fmaf(h.y, Cq.y, yv[j]));
            }
            float2 G = sG[q * 32 + lane];
            float2 dA = sdA[q * 32 + lane];
            float v[64];   // [0,32): dB contributions (step-major, pair element minor); [32,64): dC
#pragma unroll
            for (int j = kChunk - 1; j >= 0; --j) {
                const float2 Bq = *reinterpret_cast<const float2 *>(bq + j * 32);
                const float2 Cq = *reinterpret_cast<const float2 *>(bq + j * 32 + 16);
                const float2 gg = ffma2(Cq, splat2(dy[j]), G);
                const float2 dc = fmul2(splat2(dy[j]), h);
                const float2 dbv = fmul2(gg, splat2(dlu[j]));
                v[2 * j] = dbv.x;
                v[2 * j + 1] = dbv.y;
                v[32 + 2 * j] = dc.x;
                v[32 + 2 * j + 1] = dc.y;
                S1[j] = fmaf(gg.x, Bq.x, fmaf(gg.y, Bq.y, S1[j]));
                G = fmul2(a[j], gg);
                const float2 t1v = fmul2(G, hp[j]);
                S2[j] = fmaf(t1v.x, Aq.x, fmaf(t1v.y, Aq.y, S2[j]));
                dA = ffma2(t1v, splat2(dl[j]), dA);
                h = hp[j];
            }
            sG[q * 32 + lane] = G;
            sdA[q * 32 + lane] = dA;

            transpose_reduce_step<32>(v, lane);
            transpose_reduce_step<16>(v, lane);
            transpose_reduce_step<8>(v, lane);
            transpose_reduce_step<4>(v, lane);
            transpose_reduce_step<2>(v, lane);
            *reinterpret_cast<float2 *>(sRed + (lane & 15) * kRedStride + (lane >> 4) * 16 + 2 * q) = make_float2(v[0], v[1]);
        }
        __syncwarp();

        // ---- per-warp partial rows of dB|dC ----
        {
            float *dst = p.part_bc + (((size_t)g * p.B + b) * p.L + tb) * 32 + lane;
#pragma unroll
            for (int j = 0; j < kChunk; ++j)
                if (tb + j < t1) __stcs(dst + (size_t)j * 32, sRed[j * kRedStride + lane]);
        }

        // ---- per-(t, c) outputs ----
#pragma unroll
        for (int j = 0; j < kChunk; ++j) {
            if (tb + j < t1) {
                const float uj = to_f(sU[j * 32 + lane]);
                const float ddl = fmaf(S1[j], uj, S2[j]);
                const float draw = ddl * sSg[j * 32 + lane];
                st_stream(dub + (int64_t)(tb + j) * p.du_rs, from_f<T>(fmaf(dl[j], S1[j], Dc * dy[j])));
                st_stream(ddb + (int64_t)(tb + j) * p.dd_rs, from_f<T>(draw));
                if (HAS_Z) st_stream(dzb + (int64_t)(tb + j) * p.dz_rs, from_f<T>(sF[j * 32 + lane] * fmaf(Dc, uj, yv[j])));
                dD_acc = fmaf(dy[j], uj, dD_acc);
                dbias_acc += draw;
            }
        }
        __syncwarp();   // stage, sBCf and sRed are free again
        if (k - 2 >= first_chunk) issue(k - 2, stage);
        else cp_async_commit();
    }

    {
        float *dst = p.part_par + ((size_t)(b * p.nseg + seg) * 18) * p.ED + c;
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
            const float2 dA = sdA[q * 32 + lane];
            dst[(size_t)(2 * q) * p.ED] = dA.x;
            dst[(size_t)(2 * q + 1) * p.ED] = dA.y;
        }
        dst[(size_t)16 * p.ED] = dD_acc;
        dst[(size_t)17 * p.ED] = dbias_acc;
    }
}

}  // namespace gfe
