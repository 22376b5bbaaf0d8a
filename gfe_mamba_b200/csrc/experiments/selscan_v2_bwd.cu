// selscan_v2_bwd.cu -- fused selective-scan backward, "v2" (see selscan_v2_fwd.cu for the decomposition).
// Replaces autograd through mamba.py:255-256, 275-284, 220-222 and PScan.backward (pscan.py:189-224).
//
// A CTA of 128 threads serves 32 adjacent channels of one batch row; thread (c, q) owns states 4q..4q+3 of channel c
// (4 lanes per channel) as two float2 pairs that are swept together, so every operand fetched from shared memory
// feeds four states (the first v2 layout, one pair per lane, was shared-memory-bandwidth bound:
// profiles/r01_ncu_v2a_*).  The forward pass checkpoints the state every 8 steps, so the register-resident history
// is a[8], h[9] per pair.  Chunks of 16 steps are staged and walked in reverse, each as two half-chunk sweeps:
//   phase A  (item mapping, one thread per (t, channel pair)): softplus and its derivative, dy = dout * silu(z), dz, the
//            splatted scalars {dl, dl, dl*u, dl*u} and {dy, dy} -> shared slots; B|C rows -> fp32 pair tiles;
//   phase B  (recurrence mapping): a forward sweep re-derives the 8 states of a half chunk from its checkpoint into
//            registers, then the reverse sweep runs g[t] = C dy + a[t+1] g[t+1] and
//            forms every contraction with packed FP32 ops.  dB|dC contributions are reduced over the 8 lanes of the
//            warp that share the state quad in groups of 2 steps (16 values -> 2 per lane, 14 SHFL; the first level
//            needs no selects because odd channels hold their pairs swapped), per-warp rows go to shared memory;
//            the per-(t, c) sums over states leave as 4 partials per channel in bank-conflict-free planes;
//   phase C  (item mapping): du, ddelta from the partials; dD / ddt_bias accumulation; the four warps' dB|dC rows are
//            added and written once per CTA.
// Segments of L are chained exactly as in the forward kernel, in reverse time order.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "selscan_shared.cuh"

namespace gfe {

void v2_bwd_plan(int B, int L, int ED, int &nblk, int &nseg, int &seg_len);
int v2_fill_sched(ChainSched &cs, char *ws, int B, int ED, int nblk, int nseg, int seg_len, cudaStream_t st);
size_t v2_chain_bytes(int B, int ED, int nblk, int nseg);
void v2_fill_params(ScanParams &p, const gfe_selscan_args *a);
int v2_cpb(const gfe_selscan_args *a, bool bwd);
bool v2_pair_stores(const gfe_selscan_args *a, bool bwd);
int v2_check_alignment(const gfe_selscan_args *a);

// finalize kernels live in selscan.cu
void launch_bwd_finalize(const gfe_selscan_args *a, ScanParams &p, cudaStream_t st, int &rc);

#ifndef GFE_SOFTPLUS2
#define GFE_SOFTPLUS2 0        // branch-free packed softplus: -2.8 % in the forward kernel, neutral here (and 3 registers over)
#endif
#ifndef GFE_BWD_PREFETCH
#define GFE_BWD_PREFETCH 1
#endif
#ifndef GFE_BWD_KEEP_A
#define GFE_BWD_KEEP_A 0       // 1: keep exp(delta A) of the half chunk in registers (160 regs, 3 CTAs / SM); 0: recompute (4 CTAs / SM)
#endif
#ifndef GFE_BWD_MINB
#define GFE_BWD_MINB (GFE_BWD_KEEP_A ? 3 : 4)
#endif
constexpr int kBwdCPC = 32;    // channels per CTA
constexpr int kBwdNT = 128;    // 4 lanes per channel: lane (c, q) owns states 4q..4q+3 (two float2 pairs) of channel c
constexpr int kBwdWarps = kBwdNT / 32;
constexpr int kRedRow = 36;    // padded row (floats) of the per-warp dB|dC tile: [t][n]{dB, dC}
constexpr int kSPlane = 36;    // padded plane (float2) of the {S1, S2} partials: [t][q][c]
constexpr int kDDPlane = kChunk * 16 + 4;   // float4 per parity plane of the {dl, dl*u, dy, 0} slots (+64 B skew)
constexpr int kBCPlane = kChunk * 8 + 4;    // float4 per B|C plane; the 64 B skew keeps the natural and the pair-swapped plane (read by
                                            // the even / odd channels of one quarter-warp) on disjoint banks

template <typename T, bool HAS_Z>
struct BwdV2Smem {
    static constexpr int kStages = (sizeof(T) == 4 || !GFE_BWD_KEEP_A) ? 2 : 3;   // 2 stages keep 4 CTAs / SM within 227 KB
    static constexpr int kTile = kChunk * kBwdCPC * (int)sizeof(T);    // one of u, delta, dout, y, z
    static constexpr int kNTile = HAS_Z ? 5 : 3;
    static constexpr int kBCRaw = kChunk * kNState * (int)sizeof(T);    // one of B, C
    static constexpr int kStage = kNTile * kTile + 2 * kBCRaw;
    static constexpr int kOffDD = kStages * kStage;                     // float4 [2 parity][16][16] {dl, dl*u, dy, 0}
    static constexpr int kOffEpi = kOffDD + 2 * kDDPlane * 16;          // float2 [16][32] {u, softplus'}
    static constexpr int kOffBC = kOffEpi + kChunk * 32 * 8;            // float4 [2][16][8] (+64 B skew) B quads | C quads, natural / pair-swapped
    static constexpr int kOffS = kOffBC + 2 * kBCPlane * 16;            // float2 [16][4][36] {S1, S2} partial per lane
    static constexpr int kOffRed = kOffS + kChunk * 4 * kSPlane * 8;    // float  [4 warps][16][36] per-warp dB|dC rows
    static constexpr int kTotal = kOffRed + kBwdWarps * kChunk * kRedRow * 4;
};

template <typename T, bool HAS_Z, int CPB>
__global__ void __launch_bounds__(kBwdNT, GFE_BWD_MINB) selscan_bwd_v2_kernel(ScanParams p, ChainSched cs) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_unit;
    using SM = BwdV2Smem<T, HAS_Z>;
    constexpr int NT = kBwdNT, CPC = kBwdCPC, NST = SM::kStages;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rc = tid >> 2, rq = tid & 3;   // recurrence mapping: channel in block, state quad
    const int sw = rc & 1;                   // odd channels hold their pairs swapped (select-free first reduce level)
    const int ip = tid & 15, ir = tid >> 4;  // item mapping: channels 2 ip, 2 ip + 1; rows ir and ir + 8

    float4 *sDD = reinterpret_cast<float4 *>(smem + SM::kOffDD);
    float4 *sEpi4 = reinterpret_cast<float4 *>(smem + SM::kOffEpi);   // {u0, sg0, u1, sg1}
    float4 *sBC = reinterpret_cast<float4 *>(smem + SM::kOffBC);
    float2 *sS = reinterpret_cast<float2 *>(smem + SM::kOffS);
    float *sRed = reinterpret_cast<float *>(smem + SM::kOffRed);
    const bool sp = p.flags & GFE_FLAG_DELTA_SOFTPLUS;
    const bool vec = p.flags & kFlagPairStores;
    const int per_seg = p.B * cs.nblk;

    // recurrence-side shared pointers (fixed for the whole kernel)
    const float4 *dd_r = sDD + (rc & 1) * kDDPlane + (rc >> 1);
    const float4 *bc_r = sBC + sw * kBCPlane + rq;
    float2 *s_w = sS + rq * kSPlane + rc;
    const bool up8 = (lane & 8) != 0, up16 = (lane & 16) != 0;
    float *red_w = sRed + warp * (kChunk * kRedRow) + (up16 ? kRedRow : 0) + 2 * (4 * rq + (up8 ? 2 : 0) + sw);

    for (;;) {
        __syncthreads();
        if (tid == 0) s_unit = atomicAdd(cs.counter, 1);
        __syncthreads();
        const int unit = s_unit;
        if (unit >= cs.total) break;
        const int rseg = unit / per_seg;           // processing order: last time segment first
        const int seg = cs.nseg - 1 - rseg;
        const int rem = unit - rseg * per_seg;
        const int b = rem / cs.nblk;
        const int blk = rem - b * cs.nblk;
        const int c0 = blk * CPC;
        const int t0 = seg * cs.seg_len, t1 = min(p.L, t0 + cs.seg_len);
        const int kfirst = t0 / kChunk, klast = (t1 - 1) / kChunk;   // global chunk indices, walked klast .. kfirst
        const int nch = klast - kfirst + 1;

        const T *ub = reinterpret_cast<const T *>(p.u) + (int64_t)b * p.u_bs + c0;
        const T *db = reinterpret_cast<const T *>(p.delta) + (int64_t)b * p.d_bs + c0;
        const T *gb = reinterpret_cast<const T *>(p.dout) + (int64_t)b * p.do_bs + c0;
        const T *yb = reinterpret_cast<const T *>(p.ysave) + (int64_t)b * p.L * p.ED + c0;
        const T *zb = HAS_Z ? reinterpret_cast<const T *>(p.z) + (int64_t)b * p.z_bs + c0 : nullptr;
        const T *Bb = reinterpret_cast<const T *>(p.Bm) + (int64_t)b * p.B_bs;
        const T *Cb = reinterpret_cast<const T *>(p.Cm) + (int64_t)b * p.C_bs;
        // this lane's checkpointed quad: [b][t / 8][c][16] fp32
        const float4 *ckq = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(p.ckpt) +
                                                             ((size_t)b * p.nchunks * p.ED + c0 + rc) * kNState) + rq;
        const size_t ck_step = (size_t)p.ED * (kNState / 4);   // float4 between consecutive checkpoints
        T *dub = reinterpret_cast<T *>(p.du) + (int64_t)b * p.du_bs + c0 + 2 * ip;
        T *ddb = reinterpret_cast<T *>(p.ddelta) + (int64_t)b * p.dd_bs + c0 + 2 * ip;
        T *dzb = HAS_Z ? reinterpret_cast<T *>(p.dz) + (int64_t)b * p.dz_bs + c0 + 2 * ip : nullptr;

        auto issue = [&](int i) {   // i-th chunk in processing order (global chunk klast - i) -> stage i % NST
            if (i < nch) {
                const int tb = (klast - i) * kChunk;
                const int nrows = min(kChunk, t1 - tb);
                unsigned char *s = smem + (i % NST) * SM::kStage;
                stage_tile1<T, CPB, CPC, NT>(s, ub + (int64_t)tb * p.u_rs, p.u_rs, nrows, tid);
                stage_tile1<T, CPB, CPC, NT>(s + SM::kTile, db + (int64_t)tb * p.d_rs, p.d_rs, nrows, tid);
                stage_tile1<T, CPB, CPC, NT>(s + 2 * SM::kTile, gb + (int64_t)tb * p.do_rs, p.do_rs, nrows, tid);
                if (HAS_Z) {
                    stage_tile1<T, CPB, CPC, NT>(s + 3 * SM::kTile, yb + (int64_t)tb * p.ED, p.ED, nrows, tid);
                    stage_tile1<T, CPB, CPC, NT>(s + 4 * SM::kTile, zb + (int64_t)tb * p.z_rs, p.z_rs, nrows, tid);
                }
                unsigned char *sb = s + SM::kNTile * SM::kTile;
                stage_tile1<T, CPB, kNState, NT>(sb, Bb + (int64_t)tb * p.B_rs, p.B_rs, nrows, tid);
                stage_tile1<T, CPB, kNState, NT>(sb + SM::kBCRaw, Cb + (int64_t)tb * p.C_rs, p.C_rs, nrows, tid);
            }
            cp_async_commit();
        };
#pragma unroll
        for (int i = 0; i < NST; ++i) issue(i);

        // per-thread constants (pair elements swapped when sw)
        float2 A2[2], G[2], dA[2];
        {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(p.A_log + (size_t)(c0 + rc) * kNState) + rq);
            const float a0 = -expf(v.x) * kLog2e, a1 = -expf(v.y) * kLog2e, a2 = -expf(v.z) * kLog2e, a3 = -expf(v.w) * kLog2e;
            A2[0] = sw ? make_float2(a1, a0) : make_float2(a0, a1);
            A2[1] = sw ? make_float2(a3, a2) : make_float2(a2, a3);
        }
        dA[0] = dA[1] = make_float2(0.f, 0.f);
        const float2 Dc = __ldg(reinterpret_cast<const float2 *>(p.D + c0) + ip);
        const float2 bias = p.dt_bias ? __ldg(reinterpret_cast<const float2 *>(p.dt_bias + c0) + ip) : make_float2(0.f, 0.f);
        float2 dD_acc = make_float2(0.f, 0.f), dbias_acc = make_float2(0.f, 0.f);

        float *carry = cs.carry + ((size_t)b * p.ED + c0 + rc) * kNState + 4 * rq;
        if (rseg > 0) {
            if (tid == 0) {
                const int *f = cs.flags + (unit - per_seg);
                while (ld_acquire(f) == 0) __nanosleep(100);
            }
            __syncthreads();
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(carry));
            G[0] = sw ? make_float2(v.y, v.x) : make_float2(v.x, v.y);
            G[1] = sw ? make_float2(v.w, v.z) : make_float2(v.z, v.w);
        } else {
            G[0] = G[1] = make_float2(0.f, 0.f);
        }

        auto phase_a = [&](int i) {
            const int tb = (klast - i) * kChunk;
            const unsigned char *s = smem + (i % NST) * SM::kStage;
            const T *sU = reinterpret_cast<const T *>(s);
            const T *sD = reinterpret_cast<const T *>(s + SM::kTile);
            const T *sDo = reinterpret_cast<const T *>(s + 2 * SM::kTile);
            const T *sY = reinterpret_cast<const T *>(s + 3 * SM::kTile);
            const T *sZ = reinterpret_cast<const T *>(s + 4 * SM::kTile);
            const T *sBr = reinterpret_cast<const T *>(s + SM::kNTile * SM::kTile);
            float x[4], dl[4], sg[4];
#pragma unroll
            for (int ps = 0; ps < 2; ++ps) {
                const float2 d2 = lds_pair(sD + (ps * 8 + ir) * CPC, ip);
                x[2 * ps] = d2.x + bias.x;
                x[2 * ps + 1] = d2.y + bias.y;
            }
#if GFE_SOFTPLUS2
#pragma unroll
            for (int j = 0; j < 2; ++j) {   // branch-free packed softplus + sigmoid (selscan_shared.cuh)
                float2 sg2;
                const float2 v = softplus2<true>(make_float2(x[2 * j], x[2 * j + 1]), sg2);
                dl[2 * j] = sp ? v.x : x[2 * j];
                dl[2 * j + 1] = sp ? v.y : x[2 * j + 1];
                sg[2 * j] = sp ? sg2.x : 1.0f;
                sg[2 * j + 1] = sp ? sg2.y : 1.0f;
            }
#else
            if (sp) {
                softplus_group<4, true>(x, dl, sg);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) { dl[j] = x[j]; sg[j] = 1.0f; }
            }
#endif
            float2 uq[2], gq[2], zq[2], yq[2];   // every load of the phase before its first store (an LDS is never moved above an STS)
#pragma unroll
            for (int ps = 0; ps < 2; ++ps) {
                const int t = ps * 8 + ir;
                uq[ps] = lds_pair(sU + t * CPC, ip);
                gq[ps] = lds_pair(sDo + t * CPC, ip);
                if (HAS_Z) {
                    zq[ps] = lds_pair(sZ + t * CPC, ip);
                    yq[ps] = lds_pair(sY + t * CPC, ip);
                }
            }
            float4 bcv = make_float4(0.f, 0.f, 0.f, 0.f);
            {
                const int t = tid >> 3, q8 = tid & 7;
                if (tb + t < t1) {
                    const T *src = sBr + (q8 < 4 ? 0 : kChunk * 16) + t * 16 + 4 * (q8 & 3);
                    const float2 lo = lds_pair(src, 0), hi = lds_pair(src, 1);
                    bcv = make_float4(lo.x, lo.y, hi.x, hi.y);
                }
            }
#pragma unroll
            for (int ps = 0; ps < 2; ++ps) {
                const int t = ps * 8 + ir;
                const bool valid = tb + t < t1;
                const float2 u2 = uq[ps];
                const float2 g2 = gq[ps];
                const float u0 = valid ? u2.x : 0.f, u1 = valid ? u2.y : 0.f;
                const float do0 = valid ? g2.x : 0.f, do1 = valid ? g2.y : 0.f;
                const float dl0 = valid ? dl[2 * ps] : 0.f, dl1 = valid ? dl[2 * ps + 1] : 0.f;
                float dy0 = do0, dy1 = do1;
                if (HAS_Z) {
                    const float2 z2 = zq[ps];
                    const float z0 = valid ? z2.x : 0.f, z1 = valid ? z2.y : 0.f;
                    const float sz0 = sigmoid_fast(z0), sz1 = sigmoid_fast(z1);
                    dy0 = do0 * (z0 * sz0);
                    dy1 = do1 * (z1 * sz1);
                    if (valid) {   // dz = dout * d silu(z)/dz * y needs nothing from the sweeps
                        const float2 y2 = yq[ps];
                        const float f0 = do0 * sz0 * fmaf(z0, 1.0f - sz0, 1.0f), f1 = do1 * sz1 * fmaf(z1, 1.0f - sz1, 1.0f);
                        stg_pair<T>(dzb + (int64_t)(tb + t) * p.dz_rs, f0 * y2.x, f1 * y2.y, vec);
                    }
                }
                const float dlu0 = dl0 * u0, dlu1 = dl1 * u1;
                sDD[t * 16 + ip] = make_float4(dl0, dlu0, dy0, 0.f);               // even channel
                sDD[kDDPlane + t * 16 + ip] = make_float4(dl1, dlu1, dy1, 0.f);    // odd channel
                sEpi4[t * 16 + ip] = make_float4(u0, sg[2 * ps], u1, sg[2 * ps + 1]);
            }
            // B|C rows as fp32 quads, natural and pair-swapped order: [sw][t][B quads 0..3 | C quads 0..3]
            sBC[tid] = bcv;
            sBC[kBCPlane + tid] = make_float4(bcv.y, bcv.x, bcv.w, bcv.z);
        };

        cp_async_wait<NST - 1>();
        __syncthreads();
        phase_a(0);
        // checkpoints of the first chunk (states before its steps 0 and 8)
        float4 ck_lo = __ldcs(ckq + (size_t)(2 * klast) * ck_step);
        float4 ck_hi = (klast * kChunk + kCkptV2 < t1) ? __ldcs(ckq + (size_t)(2 * klast + 1) * ck_step) : make_float4(0.f, 0.f, 0.f, 0.f);

        for (int i = 0; i < nch; ++i) {
            const int k = klast - i;
            const int tb = k * kChunk;
            __syncthreads();   // (1) slots of this chunk are complete

            // ---------------------------------------------------------------- phase B
#pragma unroll 1
            for (int half = 1; half >= 0; --half) {   // steps 8..15, then 0..7 (one copy of the sweep code)
                const int jo = half * kCkptV2;
                if (tb + jo >= t1) continue;          // block-uniform: the half lies beyond the sequence
                const float4 *dd_p = dd_r + jo * 16;
                const float4 *bc_p = bc_r + jo * 8;
                float2 *s_p = s_w + jo * (4 * kSPlane);
                float *red_p = red_w + jo * kRedRow;
#if GFE_BWD_KEEP_A
                float2 a0[kCkptV2], a1[kCkptV2];
#endif
                float2 h0[kCkptV2 + 1], h1[kCkptV2 + 1];
                {
                    const float4 ck = half ? ck_hi : ck_lo;
                    h0[0] = sw ? make_float2(ck.y, ck.x) : make_float2(ck.x, ck.y);
                    h1[0] = sw ? make_float2(ck.w, ck.z) : make_float2(ck.z, ck.w);
                }
#pragma unroll
                for (int j = 0; j < kCkptV2; ++j) {   // forward sweep: re-derive the states of this half chunk
                    const float2 dd = *reinterpret_cast<const float2 *>(dd_p + j * 16);   // {dl, dl*u}
                    const float4 B4 = bc_p[j * 8];
                    const float2 dl2 = splat2(dd.x), du2 = splat2(dd.y);   // scalar-broadcast operands (R.F32), no moves
                    const float2 e0 = ex2_2(fmul2(dl2, A2[0])), e1 = ex2_2(fmul2(dl2, A2[1]));
#if GFE_BWD_KEEP_A
                    a0[j] = e0; a1[j] = e1;
#endif
                    h0[j + 1] = ffma2(e0, h0[j], fmul2(du2, make_float2(B4.x, B4.y)));
                    h1[j + 1] = ffma2(e1, h1[j], fmul2(du2, make_float2(B4.z, B4.w)));
                }
#if GFE_BWD_PREFETCH
                // slot loads of step j - 1 are issued before the stores of step j: neither nvcc nor ptxas moves an LDS above
                // a (may-alias) STS, so in plain source order every step waited out a full LDS latency (seen in the SASS)
                float4 dd_n = dd_p[(kCkptV2 - 1) * 16], B_n = bc_p[(kCkptV2 - 1) * 8], C_n = bc_p[(kCkptV2 - 1) * 8 + 4];
#endif
#pragma unroll
                for (int jb = kCkptV2 - 2; jb >= 0; jb -= 2) {
                    float v[16];   // [kind (dB, dC)][step in group (2)][state in quad (4)]
#pragma unroll
                    for (int jj = 1; jj >= 0; --jj) {
                        const int j = jb + jj;
#if GFE_BWD_PREFETCH
                        const float4 dd = dd_n, B4 = B_n, C4 = C_n;   // {dl, dl*u, dy, 0}
                        if (j > 0) { dd_n = dd_p[(j - 1) * 16]; B_n = bc_p[(j - 1) * 8]; C_n = bc_p[(j - 1) * 8 + 4]; }
#else
                        const float4 dd = dd_p[j * 16];   // {dl, dl*u, dy, 0}
                        const float4 B4 = bc_p[j * 8], C4 = bc_p[j * 8 + 4];
#endif
                        const float2 dl2 = splat2(dd.x), du2 = splat2(dd.y), dy2 = splat2(dd.z);
                        const float2 gg0 = ffma2(make_float2(C4.x, C4.y), dy2, G[0]);   // g[t] = C dy + a[t+1] g[t+1]
                        const float2 gg1 = ffma2(make_float2(C4.z, C4.w), dy2, G[1]);
                        const float2 dc0 = fmul2(dy2, h0[j + 1]), dc1 = fmul2(dy2, h1[j + 1]);   // dC_t[n] += dy h[t]
                        const float2 db0 = fmul2(gg0, du2), db1 = fmul2(gg1, du2);               // dB_t[n] += g delta u
                        const float2 sb = ffma2(gg1, make_float2(B4.z, B4.w), fmul2(gg0, make_float2(B4.x, B4.y)));   // sum_n g B
#if GFE_BWD_KEEP_A
                        G[0] = fmul2(a0[j], gg0);                                                // a[t] g[t]
                        G[1] = fmul2(a1[j], gg1);
#else   // the decay factors are re-derived (MUFU has slack in backward) instead of living in 32 registers
                        G[0] = fmul2(ex2_2(fmul2(dl2, A2[0])), gg0);
                        G[1] = fmul2(ex2_2(fmul2(dl2, A2[1])), gg1);
#endif
                        const float2 w0 = fmul2(G[0], h0[j]), w1 = fmul2(G[1], h1[j]);           // (d a) a = g a h[t-1]
                        const float2 sa = ffma2(w1, A2[1], fmul2(w0, A2[0]));                    // sum_n (da a) A log2e
                        dA[0] = ffma2(w0, dl2, dA[0]);                                           // dA[c,n] += (da a) delta
                        dA[1] = ffma2(w1, dl2, dA[1]);
                        s_p[j * (4 * kSPlane)] = make_float2(sb.x + sb.y, sa.x + sa.y);
                        v[4 * jj] = db0.x; v[4 * jj + 1] = db0.y; v[4 * jj + 2] = db1.x; v[4 * jj + 3] = db1.y;
                        v[8 + 4 * jj] = dc0.x; v[8 + 4 * jj + 1] = dc0.y; v[8 + 4 * jj + 2] = dc1.x; v[8 + 4 * jj + 3] = dc1.y;
                    }
                    // reduce the 16 values over the 8 lanes that share this quad (lane bits 2, 3, 4)
                    float r1[8];   // [kind][jj][m]: state 4 rq + 2 m + sw (odd channels hold their pairs swapped)
#pragma unroll
                    for (int m = 0; m < 8; ++m) r1[m] = v[2 * m] + __shfl_xor_sync(0xffffffffu, v[2 * m + 1], 4);
                    float r2[4];   // [kind][jj]: exchange on m
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const float keep = up8 ? r1[2 * m + 1] : r1[2 * m];
                        const float send = up8 ? r1[2 * m] : r1[2 * m + 1];
                        r2[m] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
                    float r3[2];   // [kind]: exchange on jj
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        const float keep = up16 ? r2[2 * m + 1] : r2[2 * m];
                        const float send = up16 ? r2[2 * m] : r2[2 * m + 1];
                        r3[m] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
                    *reinterpret_cast<float2 *>(red_p + jb * kRedRow) = make_float2(r3[0], r3[1]);   // {dB, dC} of (step, state)
                }
            }

            cp_async_wait<NST - 2>();
            __syncthreads();   // (2) partials complete; next chunk visible; this chunk's stage free
            issue(i + NST);
            if (i + 1 < nch) {   // next chunk's checkpoints travel while phases C and A run
                ck_lo = __ldcs(ckq + (size_t)(2 * (k - 1)) * ck_step);
                ck_hi = __ldcs(ckq + (size_t)(2 * (k - 1) + 1) * ck_step);
            }

            // ---------------------------------------------------------------- phase C
#pragma unroll
            for (int ps = 0; ps < 2; ++ps) {
                const int t = ps * 8 + ir;
                if (tb + t < t1) {
                    const float4 *sp4 = reinterpret_cast<const float4 *>(sS + (t * 4) * kSPlane) + ip;   // {S1, S2} of both channels
                    const float4 p0 = sp4[0], p1 = sp4[kSPlane / 2], p2 = sp4[kSPlane], p3 = sp4[3 * (kSPlane / 2)];
                    const float s1a = (p0.x + p1.x) + (p2.x + p3.x), s2a = ((p0.y + p1.y) + (p2.y + p3.y)) * kLn2;
                    const float s1b = (p0.z + p1.z) + (p2.z + p3.z), s2b = ((p0.w + p1.w) + (p2.w + p3.w)) * kLn2;
                    const float4 dd0 = sDD[t * 16 + ip], dd1 = sDD[kDDPlane + t * 16 + ip];
                    const float dl0 = dd0.x, dl1 = dd1.x;
                    const float4 dy4 = make_float4(dd0.z, dd0.z, dd1.z, dd1.z);
                    const float4 e4 = sEpi4[t * 16 + ip];
                    const float draw0 = fmaf(s1a, e4.x, s2a) * e4.y, draw1 = fmaf(s1b, e4.z, s2b) * e4.w;   // d delta through softplus
                    stg_pair<T>(dub + (int64_t)(tb + t) * p.du_rs, fmaf(dl0, s1a, Dc.x * dy4.x), fmaf(dl1, s1b, Dc.y * dy4.z), vec);
                    stg_pair<T>(ddb + (int64_t)(tb + t) * p.dd_rs, draw0, draw1, vec);
                    dD_acc.x = fmaf(dy4.x, e4.x, dD_acc.x);
                    dD_acc.y = fmaf(dy4.z, e4.z, dD_acc.y);
                    dbias_acc.x += draw0;
                    dbias_acc.y += draw1;
                }
            }
            {   // dB|dC rows of this CTA: add the four warps' tiles; row layout {dB[n], dC[n]} interleaved
                const int t = tid >> 3, q8 = tid & 7;
                if (tb + t < t1) {
                    const float4 *r = reinterpret_cast<const float4 *>(sRed + t * kRedRow) + q8;
                    constexpr int W4 = kChunk * kRedRow / 4;
                    const float4 x0 = r[0], x1 = r[W4], x2 = r[2 * W4], x3 = r[3 * W4];
                    const float4 acc = make_float4((x0.x + x1.x) + (x2.x + x3.x), (x0.y + x1.y) + (x2.y + x3.y),
                                                   (x0.z + x1.z) + (x2.z + x3.z), (x0.w + x1.w) + (x2.w + x3.w));
                    __stcs(reinterpret_cast<float4 *>(p.part_bc + (((size_t)blk * p.B + b) * p.L + tb + t) * 32) + q8, acc);
                }
            }
            if (i + 1 < nch) phase_a(i + 1);
        }

        // ---- end of unit: parameter-gradient partials, carry-out ----
        {
            float *dst = p.part_par + ((size_t)(b * cs.nseg + seg) * 18) * p.ED + c0 + rc;
            dst[(size_t)(4 * rq + sw) * p.ED] = dA[0].x;
            dst[(size_t)(4 * rq + 1 - sw) * p.ED] = dA[0].y;
            dst[(size_t)(4 * rq + 2 + sw) * p.ED] = dA[1].x;
            dst[(size_t)(4 * rq + 3 - sw) * p.ED] = dA[1].y;
        }
        __syncthreads();   // sRed is free: reuse it to add the eight item rows of each channel
        *reinterpret_cast<float2 *>(sRed + ir * 32 + 2 * ip) = dD_acc;
        *reinterpret_cast<float2 *>(sRed + 256 + ir * 32 + 2 * ip) = dbias_acc;
        __syncthreads();
        if (tid < 32) {
            float *dst = p.part_par + ((size_t)(b * cs.nseg + seg) * 18) * p.ED + c0 + tid;
            float sD = 0.f, sB = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                sD += sRed[w * 32 + tid];
                sB += sRed[256 + w * 32 + tid];
            }
            dst[(size_t)16 * p.ED] = sD;
            dst[(size_t)17 * p.ED] = sB;
        }
        if (seg > 0) {
            const float4 gv = sw ? make_float4(G[0].y, G[0].x, G[1].y, G[1].x) : make_float4(G[0].x, G[0].y, G[1].x, G[1].y);
            __stcg(reinterpret_cast<float4 *>(carry), gv);
            __threadfence();
            __syncthreads();
            if (tid == 0) st_release(cs.flags + unit, 1);
        }
        cp_async_wait<0>();
    }
}

// ---------------------------------------------------------------------------------------------------- host
struct BwdWs {
    size_t part_bc, part_par, total;
};
static BwdWs bwd_ws(int B, int L, int ED) {
    int nblk, nseg, seg_len;
    v2_bwd_plan(B, L, ED, nblk, nseg, seg_len);
    BwdWs w{};
    size_t off = v2_chain_bytes(B, ED, nblk, nseg);
    w.part_bc = off;
    off += align_up((size_t)nblk * B * L * 32 * sizeof(float), 256);
    w.part_par = off;
    off += align_up((size_t)B * nseg * 18 * ED * sizeof(float), 256);
    w.total = off;
    return w;
}

size_t v2_bwd_workspace_bytes(int B, int L, int ED) { return bwd_ws(B, L, ED).total; }

template <typename T, bool HAS_Z, int CPB>
static void launch_bwd_v2_inst(const ScanParams &p, const ChainSched &cs, cudaStream_t st) {
    auto kernel = selscan_bwd_v2_kernel<T, HAS_Z, CPB>;
    constexpr size_t smem = BwdV2Smem<T, HAS_Z>::kTotal;
    static thread_local int cache_total = -1, cache_grid = 0;
    if (cache_total != cs.total) {
        int per_sm = 0;
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBwdNT, smem) != cudaSuccess || per_sm < 1) {
            (void)cudaGetLastError();
            per_sm = 1;
        }
        const int64_t slots = (int64_t)sm_count() * per_sm;
        cache_grid = (int)(cs.total < slots ? cs.total : slots);
        cache_total = cs.total;
    }
    kernel<<<cache_grid, kBwdNT, smem, st>>>(p, cs);
}

template <typename T>
static int launch_bwd_v2_t(const gfe_selscan_args *a, cudaStream_t st) {
    int nblk, nseg, seg_len;
    v2_bwd_plan(a->batch, a->seqlen, a->d_inner, nblk, nseg, seg_len);
    const BwdWs w = bwd_ws(a->batch, a->seqlen, a->d_inner);
    if (a->ws == nullptr || a->ws_bytes < w.total) {
        set_error("selscan_bwd: workspace too small (%zu < %zu)", a->ws ? a->ws_bytes : (size_t)0, w.total);
        return GFE_ERR_WORKSPACE;
    }
    int rc = v2_check_alignment(a);
    if (rc != GFE_OK) return rc;
    ScanParams p{};
    v2_fill_params(p, a);
    char *ws = reinterpret_cast<char *>(a->ws);
    p.part_bc = reinterpret_cast<float *>(ws + w.part_bc);
    p.part_par = reinterpret_cast<float *>(ws + w.part_par);
    p.dout = a->dout; p.do_bs = a->dout_bs; p.do_rs = a->dout_rs;
    p.du = a->du; p.du_bs = a->du_bs; p.du_rs = a->du_rs;
    p.ddelta = a->ddelta; p.dd_bs = a->ddelta_bs; p.dd_rs = a->ddelta_rs;
    p.dz = a->dz; p.dz_bs = a->dz_bs; p.dz_rs = a->dz_rs;
    p.dBm = a->dBm; p.dB_bs = a->dB_bs; p.dB_rs = a->dB_rs;
    p.dCm = a->dCm; p.dC_bs = a->dC_bs; p.dC_rs = a->dC_rs;
    p.dA_log = a->dA_log; p.dD = a->dD; p.ddt_bias = a->ddt_bias;
    p.nseg = nseg;   // finalize_par sums over B * nseg partial rows
    p.bc_interleaved = 1;
    ChainSched cs{};
    rc = v2_fill_sched(cs, ws, a->batch, a->d_inner, nblk, nseg, seg_len, st);
    if (rc != GFE_OK) return rc;
    const int cpb = v2_cpb(a, true);
    if (v2_pair_stores(a, true)) p.flags |= kFlagPairStores;
    {
        ScopedKernelTimer tm(K_SELSCAN_BWD, st);
        if (a->z != nullptr) {
            if (cpb == 16) launch_bwd_v2_inst<T, true, 16>(p, cs, st);
            else launch_bwd_v2_inst<T, true, 0>(p, cs, st);
        } else {
            if (cpb == 16) launch_bwd_v2_inst<T, false, 16>(p, cs, st);
            else launch_bwd_v2_inst<T, false, 0>(p, cs, st);
        }
    }
    rc = check_launch("selscan_bwd_v2");
    if (rc != GFE_OK) return rc;
    launch_bwd_finalize(a, p, st, rc);
    return rc;
}

int v2_launch_bwd(const gfe_selscan_args *a, cudaStream_t st) {
    switch (a->dtype) {
        case GFE_F32: return launch_bwd_v2_t<float>(a, st);
        case GFE_BF16: return launch_bwd_v2_t<__nv_bfloat16>(a, st);
        default: return launch_bwd_v2_t<__half>(a, st);
    }
}

}  // namespace gfe
