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
//   * P = 1 or 2 lanes per channel: with P = 2 a warp covers 16 channels and each lane owns 4 of the 8 state
//     pairs, which doubles the number of warps for shapes whose B*ED/32 cannot fill the schedulers (cfg3 has
//     1.3 warps per scheduler at P = 1).  The per-step scalar work is not duplicated: the two lanes of a
//     channel pre-process alternate time steps and exchange the results (shared slots / shuffles).
//   * forward saves y (pre-gate) next to the chunk checkpoints; backward reads it instead of recomputing the
//     C contraction (HBM has slack, the issue slots do not).
// Preconditions (checked on the host, else the generic kernels run): ED % 32 == 0; for 16-bit activations
// even strides and 4-byte aligned bases.
#pragma once

#include "common.cuh"
#include "selscan_shared.cuh"

namespace gfe {

constexpr int kGroup = 4;          // steps per rolled-loop iteration

template <typename T, int P> struct FastCfg {
    static constexpr int kCh = 32 / P;                                 // channels per warp
    static constexpr int kNP = kPairs / P;                             // state pairs per lane
    static constexpr int kTileBytes = kChunk * kCh * (int)sizeof(T);   // one staged (16 x kCh) tile
    static constexpr int kBCBytes = kChunk * 32 * (int)sizeof(T);      // staged B|C rows (16 x 32)
    static constexpr int kFwdPoly = P == 1 ? 2 : 1;                    // pairs per lane with exp2 on the FMA pipe
};

// =====================================================================================================
// Forward
// =====================================================================================================
template <typename T, bool HAS_Z, int P>
__host__ __device__ constexpr int fwd_fast_smem_per_warp() {
    // 2 stages x (u, delta, [z] tiles + B|C raw rows) + fp32 B|C tile when T is 16-bit
    return 2 * ((HAS_Z ? 3 : 2) * FastCfg<T, P>::kTileBytes + FastCfg<T, P>::kBCBytes) + (sizeof(T) == 4 ? 0 : kChunk * 32 * 4);
}

template <typename T, bool HAS_Z, int CPB, int P>
__global__ void __launch_bounds__(128) selscan_fwd_fast_kernel(ScanParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Cfg = FastCfg<T, P>;
    constexpr int CH = Cfg::kCh, NP = Cfg::kNP, TILE = Cfg::kTileBytes;
    constexpr int NTILE = HAS_Z ? 3 : 2;
    constexpr int STAGE = NTILE * TILE + Cfg::kBCBytes;
    constexpr bool CONVERT_BC = sizeof(T) != 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * (blockDim.x >> 5) + warp;       // channel group of CH channels
    if (g * CH >= p.ED) return;
    const int cl = P == 1 ? lane : (lane >> 1);               // channel within the tile
    const int half = P == 1 ? 0 : (lane & 1);                 // which NP pairs / which alternate steps
    const int c = g * CH + cl;
    const int seg = blockIdx.y, b = blockIdx.z;
    const int t0 = seg * p.seg_len, t1 = min(p.L, t0 + p.seg_len);

    unsigned char *sw = smem_raw + (size_t)warp * fwd_fast_smem_per_warp<T, HAS_Z, P>();
    const uint32_t sw_u32 = smem_u32(sw);
    float *sBCf = CONVERT_BC ? reinterpret_cast<float *>(sw + 2 * STAGE) : nullptr;

    const T *ub = reinterpret_cast<const T *>(p.u) + (int64_t)b * p.u_bs + g * CH;
    const T *db = reinterpret_cast<const T *>(p.delta) + (int64_t)b * p.d_bs + g * CH;
    const T *zb = HAS_Z ? reinterpret_cast<const T *>(p.z) + (int64_t)b * p.z_bs + g * CH : nullptr;
    const T *Bb = reinterpret_cast<const T *>(p.Bm) + (int64_t)b * p.B_bs;
    const T *Cb = reinterpret_cast<const T *>(p.Cm) + (int64_t)b * p.C_bs;
    T *ob = reinterpret_cast<T *>(p.out) + (int64_t)b * p.o_bs + c;
    T *yb = p.ysave ? reinterpret_cast<T *>(p.ysave) + ((int64_t)b * p.L) * p.ED + c : nullptr;
    const bool sp = p.flags & GFE_FLAG_DELTA_SOFTPLUS;
    const float bias = p.dt_bias ? __ldg(p.dt_bias + c) : 0.0f;
    const float Dc = __ldg(p.D + c);

    auto issue = [&](int tb, int stage) {
        const int nrows = min(kChunk, t1 - tb);
        const uint32_t s = sw_u32 + stage * STAGE;
        tile_issue<T, CPB, CH>(s, ub + (int64_t)tb * p.u_rs, p.u_rs, nrows, lane, CH * sizeof(T), 0);
        tile_issue<T, CPB, CH>(s + TILE, db + (int64_t)tb * p.d_rs, p.d_rs, nrows, lane, CH * sizeof(T), 0);
        if (HAS_Z) tile_issue<T, CPB, CH>(s + 2 * TILE, zb + (int64_t)tb * p.z_rs, p.z_rs, nrows, lane, CH * sizeof(T), 0);
        tile_issue<T, CPB, 16>(s + NTILE * TILE, Bb + (int64_t)tb * p.B_rs, p.B_rs, nrows, lane, 32 * sizeof(T), 0);
        tile_issue<T, CPB, 16>(s + NTILE * TILE, Cb + (int64_t)tb * p.C_rs, p.C_rs, nrows, lane, 32 * sizeof(T), 16 * sizeof(T));
        cp_async_commit();
    };

    issue(t0, 0);
    if (t0 + kChunk < t1) issue(t0 + kChunk, 1);
    else cp_async_commit();

    // this lane's NP state pairs: global pair index half * NP + q
    float2 A2[NP], h[NP];
    {
        const float4 *row = reinterpret_cast<const float4 *>(p.A_log + (size_t)c * kNState + half * (2 * NP));
#pragma unroll
        for (int q4 = 0; q4 < NP / 2; ++q4) {
            const float4 v = __ldg(row + q4);
            A2[2 * q4] = make_float2(-expf(v.x) * kLog2e, -expf(v.y) * kLog2e);
            A2[2 * q4 + 1] = make_float2(-expf(v.z) * kLog2e, -expf(v.w) * kLog2e);
        }
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) h[q] = make_float2(0.f, 0.f);
#pragma unroll 4   // the loads of four segments are in flight together: the summaries come from L2 and the loop is latency-bound
    for (int s = 0; s < seg; ++s) {   // carry-in from earlier segments
        const float sd = p.seg_sd[(size_t)(b * p.nseg + s) * p.ED + c];
        const float2 *src = p.seg_h + ((size_t)(b * p.nseg + s) * kPairs + half * NP) * p.ED + c;
        const float2 sd2 = splat2(sd);
#pragma unroll
        for (int q = 0; q < NP; ++q) h[q] = ffma2(ex2_2(fmul2(sd2, A2[q])), h[q], src[(size_t)q * p.ED]);
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
            const T *raw = reinterpret_cast<const T *>(st + NTILE * TILE);
#pragma unroll
            for (int j = 0; j < kChunk; ++j) sBCf[j * 32 + lane] = to_f(raw[j * 32 + lane]);
            __syncwarp();
            sBC = sBCf;
        } else {
            sBC = reinterpret_cast<const float *>(st + NTILE * TILE);
        }
        sBC += half * (2 * NP);   // this lane's states inside the B (and, +16, C) row

        if (p.ckpt != nullptr) {   // state at the start of this chunk, for backward
            float2 *dst = p.ckpt + ((size_t)(b * p.nchunks + tb / kChunk) * kPairs + half * NP) * p.ED + c;
#pragma unroll
            for (int q = 0; q < NP; ++q) __stcs(dst + (size_t)q * p.ED, h[q]);
        }

#pragma unroll 1
        for (int j0 = 0; j0 < kChunk; j0 += kGroup) {
            // ---- per-step scalars: with P = 2 the two lanes of a channel take alternate steps and swap results
            constexpr int GP = kGroup / P;
            float xo[GP], dlo[GP], uo[GP], gate[GP], sgd[GP];
#pragma unroll
            for (int i = 0; i < GP; ++i) {
                const int j = j0 + i * P + half;
                xo[i] = to_f(sD[j * CH + cl]) + bias;
                uo[i] = to_f(sU[j * CH + cl]);
                if (HAS_Z) {
                    const float zj = to_f(sZ[j * CH + cl]);
                    gate[i] = zj * sigmoid_fast(zj);
                }
            }
            if (sp) {
                softplus_group<GP, false>(xo, dlo, sgd);
            } else {
#pragma unroll
                for (int i = 0; i < GP; ++i) dlo[i] = xo[i];
            }
#pragma unroll
            for (int i = 0; i < GP; ++i) {
                const bool valid = tb + j0 + i * P + half < t1;
                dlo[i] = valid ? dlo[i] : 0.f;     // padded step: a = 1, bx = 0 -> state untouched
                uo[i] = valid ? uo[i] : 0.f;
            }
            float dl[kGroup], uj[kGroup];
            if (P == 1) {
#pragma unroll
                for (int i = 0; i < kGroup; ++i) { dl[i] = dlo[i / P]; uj[i] = uo[i / P]; }
            } else {
#pragma unroll
                for (int i = 0; i < GP; ++i) {
                    const float dlx = __shfl_xor_sync(0xffffffffu, dlo[i], 1);
                    const float ux = __shfl_xor_sync(0xffffffffu, uo[i], 1);
                    dl[2 * i] = half ? dlx : dlo[i];
                    dl[2 * i + 1] = half ? dlo[i] : dlx;
                    uj[2 * i] = half ? ux : uo[i];
                    uj[2 * i + 1] = half ? uo[i] : ux;
                }
            }
            // ---- recurrences of this lane's NP state pairs over the 4 steps
            float ys[kGroup];
#pragma unroll
            for (int i = 0; i < kGroup; ++i) {
                const int j = j0 + i;
                const float2 dl2 = splat2(dl[i]), du2 = splat2(dl[i] * uj[i]);
                float2 y2 = make_float2(0.f, 0.f);
                const float4 *sb = reinterpret_cast<const float4 *>(sBC + j * 32);
#pragma unroll
                for (int q4 = 0; q4 < NP / 2; ++q4) {
                    const float4 Bq = sb[q4], Cq = sb[4 + q4];
                    const float2 x0 = fmul2(dl2, A2[2 * q4]), x1 = fmul2(dl2, A2[2 * q4 + 1]);
                    const float2 a0 = (2 * q4 < Cfg::kFwdPoly) ? ex2_poly2(x0) : ex2_2(x0);
                    const float2 a1 = (2 * q4 + 1 < Cfg::kFwdPoly) ? ex2_poly2(x1) : ex2_2(x1);
                    h[2 * q4] = ffma2(a0, h[2 * q4], fmul2(du2, make_float2(Bq.x, Bq.y)));
                    y2 = ffma2(h[2 * q4], make_float2(Cq.x, Cq.y), y2);
                    h[2 * q4 + 1] = ffma2(a1, h[2 * q4 + 1], fmul2(du2, make_float2(Bq.z, Bq.w)));
                    y2 = ffma2(h[2 * q4 + 1], make_float2(Cq.z, Cq.w), y2);
                }
                ys[i] = y2.x + y2.y;
                if (P == 2) ys[i] += __shfl_xor_sync(0xffffffffu, ys[i], 1);
            }
            // ---- D skip, gate, stores: each lane finishes the steps it pre-processed
#pragma unroll
            for (int i = 0; i < GP; ++i) {
                const int j = j0 + i * P + half;
                float y = fmaf(Dc, uo[i], P == 1 ? ys[i] : (half ? ys[2 * i + 1] : ys[2 * i]));
                if (tb + j < t1) {
                    if (yb != nullptr) st_stream(yb + (int64_t)(tb + j) * p.ED, from_f<T>(y));
                    if (HAS_Z) y *= gate[i];
                    st_stream(ob + (int64_t)(tb + j) * p.o_rs, from_f<T>(y));
                }
            }
        }
        __syncwarp();   // every lane is done with this stage before it is refilled
        if (tb + 2 * kChunk < t1) issue(tb + 2 * kChunk, stage);
        else cp_async_commit();
    }

    if (p.last_state != nullptr && seg == p.nseg - 1) {
        float2 *dst = reinterpret_cast<float2 *>(p.last_state + ((size_t)b * p.ED + c) * kNState) + half * NP;
#pragma unroll
        for (int q = 0; q < NP; ++q) dst[q] = h[q];
    }
}

// =====================================================================================================
// Backward
// =====================================================================================================
template <typename T, bool HAS_Z, int P>
__host__ __device__ constexpr int bwd_fast_smem_per_warp() {
    // 2 stages x (u, delta, dout, y, [z] tiles + B|C raw rows) + fp32 B|C tile (16-bit only) + reduced dB|dC tile
    // + 5 per-(pair, lane) float2 arrays (A, G, dA, H x 2 stages) + 5 per-(step, lane) float arrays (dl, dlu, dy, f, sigmoid)
    return 2 * ((HAS_Z ? 5 : 4) * FastCfg<T, P>::kTileBytes + FastCfg<T, P>::kBCBytes) + (sizeof(T) == 4 ? 0 : kChunk * 32 * 4) +
           kChunk * kRedStride * 4 + 5 * (FastCfg<T, P>::kNP * 32 * 8) + 5 * (kChunk * 32 * 4);
}

template <typename T, bool HAS_Z, int CPB, int P>
__global__ void __launch_bounds__(128) selscan_bwd_fast_kernel(ScanParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Cfg = FastCfg<T, P>;
    constexpr int CH = Cfg::kCh, NP = Cfg::kNP, TILE = Cfg::kTileBytes;
    constexpr int NTILE = HAS_Z ? 5 : 4;         // u, delta, dout, y, [z]
    constexpr int STAGE = NTILE * TILE + Cfg::kBCBytes;
    constexpr bool CONVERT_BC = sizeof(T) != 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * (blockDim.x >> 5) + warp;
    if (g * CH >= p.ED) return;
    const int cl = P == 1 ? lane : (lane >> 1);
    const int half = P == 1 ? 0 : (lane & 1);
    const int c = g * CH + cl;
    const int seg = blockIdx.y, b = blockIdx.z;
    const int t0 = seg * p.seg_len, t1 = min(p.L, t0 + p.seg_len);

    unsigned char *sw = smem_raw + (size_t)warp * bwd_fast_smem_per_warp<T, HAS_Z, P>();
    const uint32_t sw_u32 = smem_u32(sw);
    unsigned char *cur = sw + 2 * STAGE;
    float *sBCf = reinterpret_cast<float *>(cur);
    if (CONVERT_BC) cur += kChunk * 32 * 4;
    float *sRed = reinterpret_cast<float *>(cur);
    cur += kChunk * kRedStride * 4;
    float2 *sA = reinterpret_cast<float2 *>(cur);   // natural A = -exp(A_log), this lane's pairs
    float2 *sG = sA + NP * 32;                      // reverse carry a[t+1] g[t+1]
    float2 *sdA = sG + NP * 32;                     // dA accumulators
    float2 *sH = sdA + NP * 32;                     // checkpointed state at the chunk start, one copy per stage
    float *sDl = reinterpret_cast<float *>(sH + 2 * NP * 32);
    float *sDlu = sDl + kChunk * 32;
    float *sDy = sDlu + kChunk * 32;
    float *sF = sDy + kChunk * 32;
    float *sSg = sF + kChunk * 32;

    const T *ub = reinterpret_cast<const T *>(p.u) + (int64_t)b * p.u_bs + g * CH;
    const T *db = reinterpret_cast<const T *>(p.delta) + (int64_t)b * p.d_bs + g * CH;
    const T *gb = reinterpret_cast<const T *>(p.dout) + (int64_t)b * p.do_bs + g * CH;
    const T *yb = reinterpret_cast<const T *>(p.ysave) + ((int64_t)b * p.L) * p.ED + g * CH;
    const T *zb = HAS_Z ? reinterpret_cast<const T *>(p.z) + (int64_t)b * p.z_bs + g * CH : nullptr;
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
        tile_issue<T, CPB, CH>(s, ub + (int64_t)tb * p.u_rs, p.u_rs, nrows, lane, CH * sizeof(T), 0);
        tile_issue<T, CPB, CH>(s + TILE, db + (int64_t)tb * p.d_rs, p.d_rs, nrows, lane, CH * sizeof(T), 0);
        tile_issue<T, CPB, CH>(s + 2 * TILE, gb + (int64_t)tb * p.do_rs, p.do_rs, nrows, lane, CH * sizeof(T), 0);
        if (HAS_Z) {   // y is only needed for dz
            tile_issue<T, CPB, CH>(s + 3 * TILE, yb + (int64_t)tb * p.ED, p.ED, nrows, lane, CH * sizeof(T), 0);
            tile_issue<T, CPB, CH>(s + 4 * TILE, zb + (int64_t)tb * p.z_rs, p.z_rs, nrows, lane, CH * sizeof(T), 0);
        }
        tile_issue<T, CPB, 16>(s + NTILE * TILE, Bb + (int64_t)tb * p.B_rs, p.B_rs, nrows, lane, 32 * sizeof(T), 0);
        tile_issue<T, CPB, 16>(s + NTILE * TILE, Cb + (int64_t)tb * p.C_rs, p.C_rs, nrows, lane, 32 * sizeof(T), 16 * sizeof(T));
        {   // this lane's chunk-start states (written by the forward pass)
            const float2 *ck = p.ckpt + ((size_t)(b * p.nchunks + k) * kPairs + half * NP) * p.ED + c;
#pragma unroll
            for (int q = 0; q < NP; ++q) cp_async<8>(smem_u32(sH + (stage * NP + q) * 32 + lane), ck + (size_t)q * p.ED);
        }
        cp_async_commit();
    };
    issue(last_chunk, 0);
    if (last_chunk - 1 >= first_chunk) issue(last_chunk - 1, 1);
    else cp_async_commit();

    {   // per-lane constants and carries (this lane's NP pairs: global pair index half * NP + q)
        const float4 *row = reinterpret_cast<const float4 *>(p.A_log + (size_t)c * kNState + half * (2 * NP));
#pragma unroll
        for (int q4 = 0; q4 < NP / 2; ++q4) {
            const float4 v = __ldg(row + q4);
            sA[(2 * q4) * 32 + lane] = make_float2(-expf(v.x), -expf(v.y));
            sA[(2 * q4 + 1) * 32 + lane] = make_float2(-expf(v.z), -expf(v.w));
        }
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            sG[q * 32 + lane] = make_float2(0.f, 0.f);
            sdA[q * 32 + lane] = make_float2(0.f, 0.f);
        }
#pragma unroll 4
        for (int s = p.nseg - 1; s > seg; --s) {
            const float sd = p.seg_sd[(size_t)(b * p.nseg + s) * p.ED + c];
            const float2 *src = p.seg_h + ((size_t)(b * p.nseg + s) * kPairs + half * NP) * p.ED + c;
            const float2 sd2 = splat2(sd * kLog2e);
#pragma unroll
            for (int q = 0; q < NP; ++q)
                sG[q * 32 + lane] = ffma2(ex2_2(fmul2(sd2, sA[q * 32 + lane])), sG[q * 32 + lane], src[(size_t)q * p.ED]);
        }
    }
    float dD_acc = 0.f, dbias_acc = 0.f;
    // shared-slot column of the lane that pre-processes even / odd steps of this lane's channel
    const int col_e = P == 1 ? lane : (lane & ~1), col_o = P == 1 ? lane : (lane | 1);

    int stage = 0;
    for (int k = last_chunk; k >= first_chunk; --k, stage ^= 1) {
        const int tb = k * kChunk;
        cp_async_wait<1>();
        __syncwarp();
        const unsigned char *st = sw + stage * STAGE;
        const T *sU = reinterpret_cast<const T *>(st);
        const T *sD = reinterpret_cast<const T *>(st + TILE);
        const T *sDo = reinterpret_cast<const T *>(st + 2 * TILE);
        const T *sY = reinterpret_cast<const T *>(st + 3 * TILE);
        const T *sZ = reinterpret_cast<const T *>(st + 4 * TILE);
        const float *sBC;
        if (CONVERT_BC) {
            const T *raw = reinterpret_cast<const T *>(st + NTILE * TILE);
#pragma unroll
            for (int j = 0; j < kChunk; ++j) sBCf[j * 32 + lane] = to_f(raw[j * 32 + lane]);
            sBC = sBCf;
        } else {
            sBC = reinterpret_cast<const float *>(st + NTILE * TILE);
        }
        sBC += half * (2 * NP);

        // ---- per-step scalars: each lane pre-processes kChunk / P steps (alternate steps when P = 2) ----
        constexpr int GP = kGroup / P;
#pragma unroll
        for (int j0 = 0; j0 < kChunk; j0 += kGroup) {
            float x[GP], dlg[GP], sg[GP];
#pragma unroll
            for (int i = 0; i < GP; ++i) x[i] = to_f(sD[(j0 + i * P + half) * CH + cl]) + bias;
            if (sp) {
                softplus_group<GP, true>(x, dlg, sg);
            } else {
#pragma unroll
                for (int i = 0; i < GP; ++i) { dlg[i] = x[i]; sg[i] = 1.0f; }
            }
#pragma unroll
            for (int i = 0; i < GP; ++i) {
                const int j = j0 + i * P + half;
                const bool valid = tb + j < t1;
                const float uj = valid ? to_f(sU[j * CH + cl]) : 0.f;
                const float doj = valid ? to_f(sDo[j * CH + cl]) : 0.f;
                const float dlj = valid ? dlg[i] : 0.f;
                float f = 0.f, dyj = doj;
                if (HAS_Z) {
                    const float zj = to_f(sZ[j * CH + cl]);
                    const float sz = sigmoid_fast(zj);
                    dyj = doj * (zj * sz);
                    f = doj * sz * fmaf(zj, 1.0f - sz, 1.0f);   // dout * d silu(z)/dz ; dz = f * y
                }
                sDl[j * 32 + lane] = dlj;
                sDlu[j * 32 + lane] = dlj * uj;
                sDy[j * 32 + lane] = dyj;
                sF[j * 32 + lane] = f;
                sSg[j * 32 + lane] = sg[i];
            }
        }
        __syncwarp();   // converted B|C tile and the per-step slots are visible to all lanes
        float dl[kChunk], dlu[kChunk], dy[kChunk];
#pragma unroll
        for (int j = 0; j < kChunk; ++j) {
            const int col = (P == 1 || (j & 1) == 0) ? col_e : col_o;
            dl[j] = sDl[j * 32 + col];
            dlu[j] = sDlu[j * 32 + col];
            dy[j] = sDy[j * 32 + col];
        }

        float2 S1[kChunk], S2[kChunk];   // packed partial sums over the pair's two states
#pragma unroll
        for (int j = 0; j < kChunk; ++j) S1[j] = S2[j] = make_float2(0.f, 0.f);

        // ---- one state pair at a time: forward sweep (recompute), reverse sweep (gradients) ----
#pragma unroll 1
        for (int q = 0; q < NP; ++q) {
            const float2 Aq = sA[q * 32 + lane];
            const float2 A2q = fmul2(Aq, splat2(kLog2e));
            const float *bq = sBC + 2 * q;
            float2 a[kChunk], hs[kChunk + 1];   // hs[j] = state before step j, hs[j+1] = state after it
            hs[0] = sH[(stage * NP + q) * 32 + lane];
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                const float2 Bq = *reinterpret_cast<const float2 *>(bq + j * 32);
                a[j] = ex2_2(fmul2(splat2(dl[j]), A2q));
                hs[j + 1] = ffma2(a[j], hs[j], fmul2(splat2(dlu[j]), Bq));
            }
            float2 G = sG[q * 32 + lane];
            float2 dA = sdA[q * 32 + lane];
            float v[64];   // [0,32): dB contributions (step-major, pair element minor); [32,64): dC
#pragma unroll
            for (int j = kChunk - 1; j >= 0; --j) {
                const float2 Bq = *reinterpret_cast<const float2 *>(bq + j * 32);
                const float2 Cq = *reinterpret_cast<const float2 *>(bq + j * 32 + 16);
                const float2 gg = ffma2(Cq, splat2(dy[j]), G);          // g[t] = C dy + a[t+1] g[t+1]
                const float2 dc = fmul2(splat2(dy[j]), hs[j + 1]);      // dC_t[n] += dy * h[t]
                const float2 dbv = fmul2(gg, splat2(dlu[j]));           // dB_t[n] += g * delta * u
                v[2 * j] = dbv.x;
                v[2 * j + 1] = dbv.y;
                v[32 + 2 * j] = dc.x;
                v[32 + 2 * j + 1] = dc.y;
                S1[j] = ffma2(gg, Bq, S1[j]);                           // sum_n g B
                G = fmul2(a[j], gg);                                    // a[t] g[t]
                const float2 t1v = fmul2(G, hs[j]);                     // g a h[t-1]  (= d a * a)
                S2[j] = ffma2(t1v, Aq, S2[j]);                          // sum_n (da a) A
                dA = ffma2(t1v, splat2(dl[j]), dA);                     // dA[c,n] += (da a) delta
            }
            sG[q * 32 + lane] = G;
            sdA[q * 32 + lane] = dA;

            const int qg = half * NP + q;   // global pair index
            transpose_reduce_step<32>(v, lane);
            transpose_reduce_step<16>(v, lane);
            transpose_reduce_step<8>(v, lane);
            transpose_reduce_step<4>(v, lane);
            if (P == 1) {
                transpose_reduce_step<2>(v, lane);
                // lane l owns values 2l, 2l+1: kind = l >> 4 (0: dB, 1: dC), step = l & 15, states 2q, 2q+1
                *reinterpret_cast<float2 *>(sRed + (lane & 15) * kRedStride + (lane >> 4) * 16 + 2 * qg) = make_float2(v[0], v[1]);
            } else {
                // reduced over the 16 lanes of equal parity; lane owns values 4L'..4L'+3, L' = lane >> 1:
                // kind = L' >> 3, steps 2 (L' & 7) and 2 (L' & 7) + 1
                const int Lp = lane >> 1, jj = 2 * (Lp & 7);
                float *dst = sRed + jj * kRedStride + (Lp >> 3) * 16 + 2 * qg;
                *reinterpret_cast<float2 *>(dst) = make_float2(v[0], v[1]);
                *reinterpret_cast<float2 *>(dst + kRedStride) = make_float2(v[2], v[3]);
            }
        }
        __syncwarp();

        // ---- per-warp partial rows of dB|dC ----
        {
            float *dst = p.part_bc + (((size_t)g * p.B + b) * p.L + tb) * 32 + lane;
#pragma unroll
            for (int j = 0; j < kChunk; ++j)
                if (tb + j < t1) __stcs(dst + (size_t)j * 32, sRed[j * kRedStride + lane]);
        }

        // ---- per-(t, c) outputs: each lane finishes the steps it pre-processed (no divergence) ----
#pragma unroll
        for (int i = 0; i < kChunk / P; ++i) {
            float s1, s2, dlj, dyj;
            if (P == 1) {
                s1 = S1[i].x + S1[i].y;
                s2 = S2[i].x + S2[i].y;
                dlj = dl[i];
                dyj = dy[i];
            } else {   // totals over both lanes of the channel for steps 2i and 2i+1, then keep the own one
                float s1e = S1[2 * i].x + S1[2 * i].y, s1o = S1[2 * i + 1].x + S1[2 * i + 1].y;
                float s2e = S2[2 * i].x + S2[2 * i].y, s2o = S2[2 * i + 1].x + S2[2 * i + 1].y;
                s1e += __shfl_xor_sync(0xffffffffu, s1e, 1);
                s1o += __shfl_xor_sync(0xffffffffu, s1o, 1);
                s2e += __shfl_xor_sync(0xffffffffu, s2e, 1);
                s2o += __shfl_xor_sync(0xffffffffu, s2o, 1);
                s1 = half ? s1o : s1e;
                s2 = half ? s2o : s2e;
                dlj = half ? dl[2 * i + 1] : dl[2 * i];
                dyj = half ? dy[2 * i + 1] : dy[2 * i];
            }
            const int j = i * P + half;
            if (tb + j < t1) {
                const float uj = to_f(sU[j * CH + cl]);
                const float ddl = fmaf(s1, uj, s2);                     // d delta
                const float draw = ddl * sSg[j * 32 + lane];            // through softplus
                st_stream(dub + (int64_t)(tb + j) * p.du_rs, from_f<T>(fmaf(dlj, s1, Dc * dyj)));
                st_stream(ddb + (int64_t)(tb + j) * p.dd_rs, from_f<T>(draw));
                if (HAS_Z) st_stream(dzb + (int64_t)(tb + j) * p.dz_rs, from_f<T>(sF[j * 32 + lane] * to_f(sY[j * CH + cl])));
                dD_acc = fmaf(dyj, uj, dD_acc);
                dbias_acc += draw;
            }
        }
        __syncwarp();   // stage, sBCf, sRed and the slots are free again
        if (k - 2 >= first_chunk) issue(k - 2, stage);
        else cp_async_commit();
    }

    {
        float *dst = p.part_par + ((size_t)(b * p.nseg + seg) * 18) * p.ED + c;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const float2 dA = sdA[q * 32 + lane];
            dst[(size_t)(2 * (half * NP + q)) * p.ED] = dA.x;
            dst[(size_t)(2 * (half * NP + q) + 1) * p.ED] = dA.y;
        }
        if (P == 2) {
            dD_acc += __shfl_xor_sync(0xffffffffu, dD_acc, 1);
            dbias_acc += __shfl_xor_sync(0xffffffffu, dbias_acc, 1);
        }
        if (half == 0) {
            dst[(size_t)16 * p.ED] = dD_acc;
            dst[(size_t)17 * p.ED] = dbias_acc;
        }
    }
}

}  // namespace gfe
