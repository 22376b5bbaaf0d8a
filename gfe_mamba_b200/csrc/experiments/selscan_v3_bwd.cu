// selscan_v3_bwd.cu -- fused selective-scan backward, warp-autonomous ("v3", see selscan_v3.cuh).
// Replaces autograd through mamba.py:255-256, 275-284, 220-222 and PScan.backward (pscan.py:189-224).
//
// lane = channel.  Chunks of 8 steps (= the forward checkpoint interval) are walked in reverse.  Per chunk:
//   S  per-step scalars into registers: delta = softplus(.), its derivative, dy = dout silu(z), dz (stored at once);
//   P  for each of the 8 state pairs: a forward sweep re-derives the pair's 8 states from the checkpoint (a[t], h[t-1]
//      stay in registers), the reverse sweep runs g[t] = C dy + a[t+1] g[t+1] and forms every contraction with packed
//      FP32 ops (8 per state pair and step).  The dB|dC contributions of the 32 channels are summed through shared
//      memory: every lane stores {dB, dB, dC, dC} per step, then lane (step, part) adds 8 lanes' float4 and a
//      3-shuffle transposing exchange leaves one finished value per lane (42 instructions per pass instead of the
//      124-instruction shuffle tree of the first kernels);
//   E  du, ddelta from the per-step sums; outputs leave as 16-byte rows from the consumed input tiles.
// dA, dD, ddt_bias are carried along the chain of segments together with g, so one partial row per batch row remains.
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "selscan_v3.cuh"

namespace gfe {

int v3_fill_sched(ChainSched &cs, char *ws, int B, int ED, int nblk, int nseg, int seg_len, int carry_floats, cudaStream_t st);
void v3_fill_params(ScanParams &p, const gfe_selscan_args *a);
int v3_aligned(const gfe_selscan_args *a, bool bwd);
int v3_warps_per_cta(size_t per_warp_smem);
void launch_bwd_finalize(const gfe_selscan_args *a, ScanParams &p, cudaStream_t st, int &rc);   // selscan.cu

constexpr int kB3Chunk = kCkptV2;   // 8
constexpr int kB3Stages = 4;
constexpr int kB3Carry = 34;        // floats per channel handed between segments: g[16], dA[16], dD, ddt_bias

template <typename T, bool HAS_Z>
struct BwdV3Smem {
    static constexpr int kTile = kB3Chunk * 32 * (int)sizeof(T);      // one of u, delta, dout, y, z
    static constexpr int kNT = HAS_Z ? 5 : 3;
    static constexpr int kBC = kB3Chunk * 16 * (int)sizeof(T);        // one of B, C
    static constexpr int kCk = kPairs * 32 * 8;                        // checkpoint [pair][lane] float2
    static constexpr int kOffB = kNT * kTile;
    static constexpr int kOffCk = kOffB + 2 * kBC;
    static constexpr int kStage = kOffCk + kCk;
    static constexpr int kOffBCf = kB3Stages * kStage;                 // float4 [8 steps][8 pairs] {B, B, C, C}
    static constexpr int kOffRed = kOffBCf + kB3Chunk * kPairs * 16;   // float2 [2 buffers][2 planes dB|dC][8 steps][32 lanes]
    static constexpr int kOffRows = kOffRed + 2 * kB3Chunk * 32 * 16;  // float  [8 steps][32]
    static constexpr int kOffPar = kOffRows + kB3Chunk * 32 * 4;       // float2 [3][8 pairs][32 lanes]: A log2e, g, dA of the unit
    static constexpr int kOffUS = kOffPar + 3 * kPairs * 32 * 8;       // float2 [8 steps][32 lanes] {u, softplus'} of the chunk
    static constexpr int kOffIds = kOffUS + kB3Chunk * 32 * 8;
    static constexpr int kPerWarp = kOffIds + kV3IdRing * 4;
};

template <typename T, bool HAS_Z>
__global__ void __launch_bounds__(32 * kV3MaxWarps, 1) selscan_bwd_v3_kernel(ScanParams p, ChainSched cs, int aligned) {
    extern __shared__ __align__(16) unsigned char smem_all[];
    using SM = BwdV3Smem<T, HAS_Z>;
    constexpr int NST = kB3Stages, CH = kB3Chunk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *smem = smem_all + warp * SM::kPerWarp;
    const uint32_t smem_a = smem_u32(smem);
    int *ids = reinterpret_cast<int *>(smem + SM::kOffIds);
    float4 *sBCf = reinterpret_cast<float4 *>(smem + SM::kOffBCf);
    float2 *sRed = reinterpret_cast<float2 *>(smem + SM::kOffRed);
    float *sRows = reinterpret_cast<float *>(smem + SM::kOffRows);
    // per-lane spill space with static addressing: only the pair being swept lives in registers
    float2 *sA2 = reinterpret_cast<float2 *>(smem + SM::kOffPar) + lane;
    float2 *sGc = sA2 + kPairs * 32, *sdA = sA2 + 2 * kPairs * 32;
    float2 *sUS = reinterpret_cast<float2 *>(smem + SM::kOffUS) + lane;
    const bool sp = p.flags & GFE_FLAG_DELTA_SOFTPLUS;
    const int per_seg = p.B * cs.nblk;
    const int jr = lane >> 2, part = lane & 3;   // reduction mapping: step; part = 2 * plane (dB | dC) + half of the lanes / element kept

    // ---- the chunk stream (reverse time): units in draw order, chunks prefetched NST - 1 ahead ----
    int n_drawn = 0, n_used = 0;
    V3Unit pf;
    int pf_k = 0, pf_slot = 0;
    pf.id = -1; pf.nch = 0;
    bool exhausted = false;
    auto draw = [&]() {
        const int id = exhausted ? cs.total : v3_draw(cs.counter, lane);
        if (id >= cs.total) exhausted = true;
        if (lane == 0) ids[n_drawn % kV3IdRing] = id;
        ++n_drawn;
        return id;
    };
    int next_id = draw();
    auto prefetch_next = [&]() {
        if (pf_k == pf.nch) {
            pf = v3_decode(next_id, cs, p.B, p.L, CH, true);
            pf_k = 0;
        }
        if (pf.id >= 0) {
            const int tb = pf.t0 + (pf.nch - 1 - pf_k) * CH;
            const int nrows = min(CH, pf.t1 - tb);
            const int st = pf_slot % NST;
            const uint32_t s = smem_a + st * SM::kStage;
            const int64_t b = pf.b;
            const T *ub = reinterpret_cast<const T *>(p.u) + b * p.u_bs + pf.c0 + (int64_t)tb * p.u_rs;
            const T *db = reinterpret_cast<const T *>(p.delta) + b * p.d_bs + pf.c0 + (int64_t)tb * p.d_rs;
            const T *gb = reinterpret_cast<const T *>(p.dout) + b * p.do_bs + pf.c0 + (int64_t)tb * p.do_rs;
            const T *yb = reinterpret_cast<const T *>(p.ysave) + (b * p.L + tb) * p.ED + pf.c0;
            const T *zb = HAS_Z ? reinterpret_cast<const T *>(p.z) + b * p.z_bs + pf.c0 + (int64_t)tb * p.z_rs : nullptr;
            const T *Bb = reinterpret_cast<const T *>(p.Bm) + b * p.B_bs + (int64_t)tb * p.B_rs;
            const T *Cb = reinterpret_cast<const T *>(p.Cm) + b * p.C_bs + (int64_t)tb * p.C_rs;
            const float2 *ck = p.ckpt + ((size_t)(b * p.nchunks + tb / CH) * kPairs) * p.ED + pf.c0;   // [pair][ED]
            if (aligned) {
                v3_stage32<T, CH>(s, ub, p.u_rs, nrows, lane);
                v3_stage32<T, CH>(s + SM::kTile, db, p.d_rs, nrows, lane);
                v3_stage32<T, CH>(s + 2 * SM::kTile, gb, p.do_rs, nrows, lane);
                if (HAS_Z) {
                    v3_stage32<T, CH>(s + 3 * SM::kTile, yb, p.ED, nrows, lane);
                    v3_stage32<T, CH>(s + 4 * SM::kTile, zb, p.z_rs, nrows, lane);
                }
                v3_stage16<T, CH>(s + SM::kOffB, Bb, p.B_rs, nrows, lane);
                v3_stage16<T, CH>(s + SM::kOffB + SM::kBC, Cb, p.C_rs, nrows, lane);
                v3_stage_ck(s + SM::kOffCk, ck, p.ED, lane);
            } else {   // misaligned views: plain element loads (correct, slow)
                T *d = reinterpret_cast<T *>(smem + st * SM::kStage);
                constexpr int TE = SM::kTile / (int)sizeof(T);
                for (int r = 0; r < nrows; ++r) {
                    d[r * 32 + lane] = ub[(int64_t)r * p.u_rs + lane];
                    d[TE + r * 32 + lane] = db[(int64_t)r * p.d_rs + lane];
                    d[2 * TE + r * 32 + lane] = gb[(int64_t)r * p.do_rs + lane];
                    if (HAS_Z) {
                        d[3 * TE + r * 32 + lane] = yb[(int64_t)r * p.ED + lane];
                        d[4 * TE + r * 32 + lane] = zb[(int64_t)r * p.z_rs + lane];
                    }
                    T *bc = d + SM::kNT * TE;
                    if (lane < 16) bc[r * 16 + lane] = Bb[(int64_t)r * p.B_rs + lane];
                    else bc[SM::kBC / (int)sizeof(T) + r * 16 + lane - 16] = Cb[(int64_t)r * p.C_rs + lane - 16];
                }
                float2 *dk = reinterpret_cast<float2 *>(smem + st * SM::kStage + SM::kOffCk);
                for (int q = 0; q < kPairs; ++q) dk[q * 32 + lane] = ck[(size_t)q * p.ED + lane];
            }
            if (pf_k == pf.nch - 1) next_id = draw();
            ++pf_k;
        }
        cp_async_commit();
        ++pf_slot;
    };
#pragma unroll 1
    for (int s = 0; s < NST - 1; ++s) prefetch_next();

    int slot = 0;
    for (;;) {
        __syncwarp();
        const V3Unit cur = v3_decode(ids[n_used % kV3IdRing], cs, p.B, p.L, CH, true);
        ++n_used;
        if (cur.id < 0) break;
        const int c = cur.c0 + lane;
        const int rseg = cs.nseg - 1 - cur.seg;

        // ---- unit prologue ----
        {
            const float4 *row = reinterpret_cast<const float4 *>(p.A_log + (size_t)c * kNState);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v = __ldg(row + q);
                sA2[(2 * q) * 32] = make_float2(-expf(v.x) * kLog2e, -expf(v.y) * kLog2e);
                sA2[(2 * q + 1) * 32] = make_float2(-expf(v.z) * kLog2e, -expf(v.w) * kLog2e);
            }
        }
        const float Dc = __ldg(p.D + c);
        const float bias = p.dt_bias ? __ldg(p.dt_bias + c) : 0.f;
        float dD_acc = 0.f, dbias_acc = 0.f;
        float2 *carry = reinterpret_cast<float2 *>(cs.carry) + (size_t)cur.b * (kB3Carry / 2) * p.ED + c;   // [b][17][ED] float2
        if (rseg > 0) {
            if (lane == 0) {
                const int *f = cs.flags + (cur.id - per_seg);
                v3_wait_flag(f);
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < kPairs; ++q) {
                sGc[q * 32] = __ldcg(carry + (size_t)q * p.ED);
                sdA[q * 32] = __ldcg(carry + (size_t)(kPairs + q) * p.ED);
            }
            const float2 t = __ldcg(carry + (size_t)(2 * kPairs) * p.ED);
            dD_acc = t.x;
            dbias_acc = t.y;
        } else {
#pragma unroll
            for (int q = 0; q < kPairs; ++q) sGc[q * 32] = sdA[q * 32] = make_float2(0.f, 0.f);
        }
        T *dub = reinterpret_cast<T *>(p.du) + (int64_t)cur.b * p.du_bs + cur.c0;
        T *ddb = reinterpret_cast<T *>(p.ddelta) + (int64_t)cur.b * p.dd_bs + cur.c0;
        T *dzb = HAS_Z ? reinterpret_cast<T *>(p.dz) + (int64_t)cur.b * p.dz_bs + cur.c0 : nullptr;
        float *pbc = p.part_bc + (((size_t)(cur.c0 / 32) * p.B + cur.b) * p.L) * 32;

        for (int k = 0; k < cur.nch; ++k, ++slot) {
            prefetch_next();
            cp_async_wait<NST - 1>();
            __syncwarp();
            const int tb = cur.t0 + (cur.nch - 1 - k) * CH;
            const int nsteps = min(CH, cur.t1 - tb);
            unsigned char *s = smem + (slot % NST) * SM::kStage;
            T *sU = reinterpret_cast<T *>(s) + lane;
            T *sD = reinterpret_cast<T *>(s + SM::kTile) + lane;
            const T *sG = reinterpret_cast<const T *>(s + 2 * SM::kTile) + lane;
            const T *sY = reinterpret_cast<const T *>(s + 3 * SM::kTile) + lane;
            T *sZ = reinterpret_cast<T *>(s + 4 * SM::kTile) + lane;
            const float2 *sCk = reinterpret_cast<const float2 *>(s + SM::kOffCk) + lane;

            {   // B|C rows -> fp32 {B pair, C pair} quads; rows beyond the sequence are zero
                const T *rb = reinterpret_cast<const T *>(s + SM::kOffB) + jr * 16 + 4 * part;
                const T *rc = reinterpret_cast<const T *>(s + SM::kOffB + SM::kBC) + jr * 16 + 4 * part;
                float4 bq, cq;
                if constexpr (sizeof(T) == 4) {
                    bq = *reinterpret_cast<const float4 *>(rb);
                    cq = *reinterpret_cast<const float4 *>(rc);
                } else {
                    const float2 b0 = lds_pair(rb, 0), b1 = lds_pair(rb, 1), c0 = lds_pair(rc, 0), c1 = lds_pair(rc, 1);
                    bq = make_float4(b0.x, b0.y, b1.x, b1.y);
                    cq = make_float4(c0.x, c0.y, c1.x, c1.y);
                }
                if (jr >= nsteps) bq = cq = make_float4(0.f, 0.f, 0.f, 0.f);
                sBCf[jr * kPairs + 2 * part] = make_float4(bq.x, bq.y, cq.x, cq.y);
                sBCf[jr * kPairs + 2 * part + 1] = make_float4(bq.z, bq.w, cq.z, cq.w);
            }

            // ---------------------------------------------------------------- phase S
            float dl[CH], dy[CH], du[CH];
            const bool full = nsteps == CH;
            {
                float x[CH], sg[CH];
#pragma unroll
                for (int j = 0; j < CH; ++j) x[j] = to_f(sD[j * 32]) + bias;
                if (sp) {
                    v3_softplus<CH, true>(x, dl, sg);
                } else {
#pragma unroll
                    for (int j = 0; j < CH; ++j) { dl[j] = x[j]; sg[j] = 1.0f; }
                }
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const bool live = full || j < nsteps;
                    sUS[j * 32] = make_float2(live ? to_f(sU[j * 32]) : 0.f, sg[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const bool live = full || j < nsteps;
                const float u = to_f(sU[j * 32]), go = to_f(sG[j * 32]);
                float dyj = go;
                if (HAS_Z) {
                    const float z = to_f(sZ[j * 32]), ys = to_f(sY[j * 32]);
                    const float sz = sigmoid_fast(z);
                    dyj = go * (z * sz);
                    sZ[j * 32] = from_f<T>(go * ys * sz * fmaf(z, 1.0f - sz, 1.0f));   // dz, leaves with the tile below
                }
                const float uj = live ? u : 0.f;
                dl[j] = live ? dl[j] : 0.f;
                dy[j] = live ? dyj : 0.f;
                du[j] = dl[j] * uj;
                dD_acc = fmaf(dy[j], uj, dD_acc);
            }
            float2 S1[CH], S2[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) S1[j] = S2[j] = make_float2(0.f, 0.f);
            __syncwarp();   // sBCf complete

            // ---------------------------------------------------------------- phase P
#pragma unroll 1
            for (int q = 0; q < kPairs; ++q) {
                float2 a[CH], hp[CH];
                float2 h = sCk[q * 32];
                const float2 A2q = sA2[q * 32];
                float2 Gq = sGc[q * 32], dAq = sdA[q * 32];
#pragma unroll
                for (int j = 0; j < CH; ++j) {   // forward sweep: states of this pair over the chunk
                    const float2 B2 = *reinterpret_cast<const float2 *>(sBCf + j * kPairs + q);
                    const float2 dl2 = make_float2(dl[j], dl[j]);
                    a[j] = ex2_2(fmul2(dl2, A2q));
                    hp[j] = h;
                    h = ffma2(a[j], h, fmul2(make_float2(du[j], du[j]), B2));
                }
                float2 *red = sRed + (q & 1) * (2 * CH * 32) + lane;
#pragma unroll
                for (int j = CH - 1; j >= 0; --j) {   // reverse sweep
                    const float4 bc = sBCf[j * kPairs + q];
                    const float2 B2 = make_float2(bc.x, bc.y), C2 = make_float2(bc.z, bc.w);
                    const float2 dy2 = make_float2(dy[j], dy[j]), dl2 = make_float2(dl[j], dl[j]);
                    const float2 g = ffma2(C2, dy2, Gq);                   // g[t] = C dy + a[t+1] g[t+1]
                    const float2 dc = fmul2(dy2, h);                       // dC_t[n] += dy h[t]
                    const float2 db = fmul2(g, make_float2(du[j], du[j]));      // dB_t[n] += g delta u
                    S2[j] = ffma2(g, B2, S2[j]);                           // sum_n g B
                    Gq = fmul2(a[j], g);                                   // a[t] g[t]
                    const float2 w = fmul2(Gq, hp[j]);                     // (d a) a = g a h[t-1]
                    S1[j] = ffma2(w, A2q, S1[j]);                          // sum_n (da a) A log2e
                    dAq = ffma2(w, dl2, dAq);                              // dA[c,n] += (da a) delta
                    red[j * 32] = db;
                    red[CH * 32 + j * 32] = dc;
                    h = hp[j];
                }
                sGc[q * 32] = Gq;
                sdA[q * 32] = dAq;
                __syncwarp();
                {   // sum the 32 channels: lane (step, plane, half) adds 16 lanes' pairs, one exchange with the other half
                    const float2 *src = sRed + (q & 1) * (2 * CH * 32) + (part >> 1) * (CH * 32) + jr * 32 + (part & 1);
                    float2 acc = src[2 * (jr & 15)];
#pragma unroll
                    for (int i = 1; i < 16; ++i) acc = fadd2(acc, src[2 * ((i + jr) & 15)]);
                    const bool hi = part & 1;
                    float kk = hi ? acc.y : acc.x;
                    const float ss = hi ? acc.x : acc.y;
                    kk += __shfl_xor_sync(0xffffffffu, ss, 1);
                    sRows[jr * 32 + part * kPairs + q] = kk;   // part 0: dB[2q], 1: dB[2q+1], 2: dC[2q], 3: dC[2q+1]
                }
            }

            // ---------------------------------------------------------------- phase E
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const float s1 = (S1[j].x + S1[j].y) * kLn2, s2 = S2[j].x + S2[j].y;
                const float2 us = sUS[j * 32];
                const float draw = fmaf(us.x, s2, s1) * us.y;   // d delta through softplus
                sU[j * 32] = from_f<T>(fmaf(dl[j], s2, Dc * dy[j]));
                sD[j * 32] = from_f<T>(draw);
                dbias_acc += (full || j < nsteps) ? draw : 0.f;
            }
            __syncwarp();
            v3_store32<T, CH>(reinterpret_cast<const T *>(s), dub + (int64_t)tb * p.du_rs, p.du_rs, nsteps, lane, aligned);
            v3_store32<T, CH>(reinterpret_cast<const T *>(s + SM::kTile), ddb + (int64_t)tb * p.dd_rs, p.dd_rs, nsteps, lane, aligned);
            if (HAS_Z) v3_store32<T, CH>(reinterpret_cast<const T *>(s + 4 * SM::kTile), dzb + (int64_t)tb * p.dz_rs, p.dz_rs, nsteps, lane, aligned);
            v3_store32<float, CH>(sRows, pbc + (size_t)tb * 32, 32, nsteps, lane, 1);
            __syncwarp();   // stage and row tile may be reused
        }

        // ---- unit epilogue: hand g and the parameter-gradient sums to the predecessor segment, or finish ----
        if (cur.seg > 0) {
#pragma unroll
            for (int q = 0; q < kPairs; ++q) {
                __stcg(carry + (size_t)q * p.ED, sGc[q * 32]);
                __stcg(carry + (size_t)(kPairs + q) * p.ED, sdA[q * 32]);
            }
            __stcg(carry + (size_t)(2 * kPairs) * p.ED, make_float2(dD_acc, dbias_acc));
            __threadfence();
            __syncwarp();
            if (lane == 0) st_release(cs.flags + cur.id, 1);
        } else {
            float *dst = p.part_par + ((size_t)cur.b * 18) * p.ED + c;
#pragma unroll
            for (int q = 0; q < kPairs; ++q) {
                const float2 v = sdA[q * 32];
                dst[(size_t)(2 * q) * p.ED] = v.x;
                dst[(size_t)(2 * q + 1) * p.ED] = v.y;
            }
            dst[(size_t)16 * p.ED] = dD_acc;
            dst[(size_t)17 * p.ED] = dbias_acc;
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------------- host
struct V3BwdWs {
    size_t part_bc, part_par, total;
};
static V3BwdWs v3_bwd_ws(int B, int L, int ED) {
    int nblk, nseg, seg_len;
    v3_plan(B, L, ED, nblk, nseg, seg_len);
    V3BwdWs w{};
    size_t off = v3_chain_bytes(B, ED, nblk, nseg, kB3Carry);
    w.part_bc = off;
    off += align_up((size_t)nblk * B * L * 32 * sizeof(float), 256);
    w.part_par = off;
    off += align_up((size_t)B * 18 * ED * sizeof(float), 256);
    w.total = off;
    return w;
}
size_t v3_bwd_workspace_bytes(int B, int L, int ED) { return v3_bwd_ws(B, L, ED).total; }

template <typename T, bool HAS_Z>
static void launch_bwd_v3_inst(const ScanParams &p, const ChainSched &cs, int aligned, cudaStream_t st) {
    auto kernel = selscan_bwd_v3_kernel<T, HAS_Z>;
    constexpr size_t per_warp = BwdV3Smem<T, HAS_Z>::kPerWarp;
    const int W = v3_warps_per_cta(per_warp);
    const size_t smem = per_warp * W;
    static thread_local size_t smem_set = 0;
    if (smem_set != smem) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        smem_set = smem;
    }
    const int64_t want = ceil_div64(cs.total, W);
    const int grid = (int)(want < sm_count() ? want : sm_count());
    kernel<<<grid, 32 * W, smem, st>>>(p, cs, aligned);
}

template <typename T>
static int launch_bwd_v3_t(const gfe_selscan_args *a, cudaStream_t st) {
    int nblk, nseg, seg_len;
    v3_plan(a->batch, a->seqlen, a->d_inner, nblk, nseg, seg_len);
    const V3BwdWs w = v3_bwd_ws(a->batch, a->seqlen, a->d_inner);
    if (a->ws == nullptr || a->ws_bytes < w.total) {
        set_error("selscan_bwd: workspace too small (%zu < %zu)", a->ws ? a->ws_bytes : (size_t)0, w.total);
        return GFE_ERR_WORKSPACE;
    }
    if ((reinterpret_cast<uintptr_t>(a->ws) & 15) != 0) {
        set_error("selscan_bwd: workspace must be 16-byte aligned");
        return GFE_ERR_ARG;
    }
    ScanParams p{};
    v3_fill_params(p, a);
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
    p.nseg = 1;             // one parameter-gradient partial row per batch row (carried along the chain)
    p.bc_interleaved = 2;   // rows are [dB even n | dB odd n | dC even n | dC odd n]
    ChainSched cs{};
    int rc = v3_fill_sched(cs, ws, a->batch, a->d_inner, nblk, nseg, seg_len, kB3Carry, st);
    if (rc != GFE_OK) return rc;
    const int aligned = v3_aligned(a, true);
    {
        ScopedKernelTimer tm(K_SELSCAN_BWD, st);
        if (a->z != nullptr) launch_bwd_v3_inst<T, true>(p, cs, aligned, st);
        else launch_bwd_v3_inst<T, false>(p, cs, aligned, st);
    }
    rc = check_launch("selscan_bwd_v3");
    if (rc != GFE_OK) return rc;
    launch_bwd_finalize(a, p, st, rc);
    return rc;
}

int v3_launch_bwd(const gfe_selscan_args *a, cudaStream_t st) {
    switch (a->dtype) {
        case GFE_F32: return launch_bwd_v3_t<float>(a, st);
        case GFE_BF16: return launch_bwd_v3_t<__nv_bfloat16>(a, st);
        default: return launch_bwd_v3_t<__half>(a, st);
    }
}

}  // namespace gfe
