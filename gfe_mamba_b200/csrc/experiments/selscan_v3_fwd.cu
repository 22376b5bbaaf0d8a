// selscan_v3_fwd.cu -- fused selective-scan forward, warp-autonomous ("v3", see selscan_v3.cuh).
// Replaces mamba.py:255-256 (softplus + bias), :275-284 (discretise, scan, C contraction, D skip) and :220-222 (gate).
//
// Per (lane = channel, step): 3 shared loads (u, delta, z), softplus, then for each of the 8 state pairs
//   x = delta * A2 (FMUL2); a = exp2(x) (2 MUFU, or a degree-5 polynomial on the FMA pipe for kNPoly of the pairs,
//   because MUFU is the binding pipe); h = a h + (delta u) B (FMUL2 + FFMA2); y += h C (FFMA2)
// with B|C read as broadcast LDS.128 from an fp32 tile, then the D skip, the SiLU gate and two 2-byte stores.
// The state before every 8th step is written as a checkpoint for backward ([b][t/8][pair][ED] float2).
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "selscan_v3.cuh"

namespace gfe {

constexpr int kF3Chunk = 16;
constexpr int kF3Stages = 4;
#ifndef GFE_V3_GROUP
#define GFE_V3_GROUP 4
#endif
constexpr int kF3Group = GFE_V3_GROUP;   // steps per unrolled loop body
#ifndef GFE_V3_NPOLY
#define GFE_V3_NPOLY 2
#endif
constexpr int kNPoly = GFE_V3_NPOLY;   // state pairs per step whose exp2 runs on the FMA pipe

template <typename T, bool HAS_Z>
struct FwdV3Smem {
    static constexpr int kTile = kF3Chunk * 32 * (int)sizeof(T);       // one of u, delta, z
    static constexpr int kBC = kF3Chunk * 16 * (int)sizeof(T);         // one of B, C
    static constexpr int kNT = HAS_Z ? 3 : 2;
    static constexpr int kStage = kNT * kTile + 2 * kBC;
    static constexpr int kOffBCf = kF3Stages * kStage;                  // fp32 [16][16] B then [16][16] C (16-bit inputs only)
    static constexpr int kBCf = sizeof(T) == 4 ? 0 : 2 * kF3Chunk * 16 * 4;
    static constexpr int kOffIds = kOffBCf + kBCf;
    static constexpr int kPerWarp = kOffIds + kV3IdRing * 4;
};

template <typename T, bool HAS_Z, bool SAVE>
__global__ void __launch_bounds__(32 * kV3MaxWarps, 1) selscan_fwd_v3_kernel(ScanParams p, ChainSched cs, int aligned) {
    extern __shared__ __align__(16) unsigned char smem_all[];
    using SM = FwdV3Smem<T, HAS_Z>;
    constexpr int NST = kF3Stages;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *smem = smem_all + warp * SM::kPerWarp;
    const uint32_t smem_a = smem_u32(smem);
    int *ids = reinterpret_cast<int *>(smem + SM::kOffIds);
    const bool sp = p.flags & GFE_FLAG_DELTA_SOFTPLUS;

    // ---- the chunk stream: units drawn in order, their chunks prefetched NST - 1 ahead of the compute ----
    int n_drawn = 0, n_used = 0;       // ids pushed / popped
    V3Unit pf;                         // unit whose chunks are being prefetched
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
            pf = v3_decode(next_id, cs, p.B, p.L, kF3Chunk, false);
            pf_k = 0;
        }
        if (pf.id >= 0) {
            const int tb = pf.t0 + pf_k * kF3Chunk;
            const int nrows = min(kF3Chunk, pf.t1 - tb);
            const uint32_t s = smem_a + (pf_slot % NST) * SM::kStage;
            const T *ub = reinterpret_cast<const T *>(p.u) + (int64_t)pf.b * p.u_bs + pf.c0 + (int64_t)tb * p.u_rs;
            const T *db = reinterpret_cast<const T *>(p.delta) + (int64_t)pf.b * p.d_bs + pf.c0 + (int64_t)tb * p.d_rs;
            const T *Bb = reinterpret_cast<const T *>(p.Bm) + (int64_t)pf.b * p.B_bs + (int64_t)tb * p.B_rs;
            const T *Cb = reinterpret_cast<const T *>(p.Cm) + (int64_t)pf.b * p.C_bs + (int64_t)tb * p.C_rs;
            if (aligned) {
                v3_stage32<T, kF3Chunk>(s, ub, p.u_rs, nrows, lane);
                v3_stage32<T, kF3Chunk>(s + SM::kTile, db, p.d_rs, nrows, lane);
                if (HAS_Z) {
                    const T *zb = reinterpret_cast<const T *>(p.z) + (int64_t)pf.b * p.z_bs + pf.c0 + (int64_t)tb * p.z_rs;
                    v3_stage32<T, kF3Chunk>(s + 2 * SM::kTile, zb, p.z_rs, nrows, lane);
                }
                v3_stage16<T, kF3Chunk>(s + SM::kNT * SM::kTile, Bb, p.B_rs, nrows, lane);
                v3_stage16<T, kF3Chunk>(s + SM::kNT * SM::kTile + SM::kBC, Cb, p.C_rs, nrows, lane);
            } else {   // misaligned views: plain element loads (correct, slow)
                T *d = reinterpret_cast<T *>(smem + (pf_slot % NST) * SM::kStage);
                for (int r = 0; r < nrows; ++r) {
                    d[r * 32 + lane] = ub[(int64_t)r * p.u_rs + lane];
                    d[SM::kTile / (int)sizeof(T) + r * 32 + lane] = db[(int64_t)r * p.d_rs + lane];
                    if (HAS_Z) {
                        const T *zb = reinterpret_cast<const T *>(p.z) + (int64_t)pf.b * p.z_bs + pf.c0 + (int64_t)tb * p.z_rs;
                        d[2 * SM::kTile / (int)sizeof(T) + r * 32 + lane] = zb[(int64_t)r * p.z_rs + lane];
                    }
                    T *bc = d + SM::kNT * SM::kTile / (int)sizeof(T);
                    if (lane < 16) bc[r * 16 + lane] = Bb[(int64_t)r * p.B_rs + lane];
                    else bc[SM::kBC / (int)sizeof(T) + r * 16 + lane - 16] = Cb[(int64_t)r * p.C_rs + lane - 16];
                }
            }
            if (pf_k == pf.nch - 1) next_id = draw();   // result needed one chunk-time from now
            ++pf_k;
        }
        cp_async_commit();
        ++pf_slot;
    };
#pragma unroll 1
    for (int s = 0; s < NST - 1; ++s) prefetch_next();

    int slot = 0;   // ring position of the chunk being computed
    for (;;) {
        __syncwarp();
        const V3Unit cur = v3_decode(ids[n_used % kV3IdRing], cs, p.B, p.L, kF3Chunk, false);
        ++n_used;
        if (cur.id < 0) break;
        const int c = cur.c0 + lane;

        // ---- unit prologue: parameters, carry-in ----
        float2 A2[kPairs], h[kPairs];
        {
            const float4 *row = reinterpret_cast<const float4 *>(p.A_log + (size_t)c * kNState);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v = __ldg(row + q);
                A2[2 * q] = make_float2(-expf(v.x) * kLog2e, -expf(v.y) * kLog2e);
                A2[2 * q + 1] = make_float2(-expf(v.z) * kLog2e, -expf(v.w) * kLog2e);
            }
        }
        const float Dc = __ldg(p.D + c);
        const float bias = p.dt_bias ? __ldg(p.dt_bias + c) : 0.f;
        float2 *carry = reinterpret_cast<float2 *>(cs.carry) + (size_t)cur.b * kPairs * p.ED + c;   // [b][pair][ED]
        if (cur.seg > 0) {
            if (lane == 0) {
                const int *f = cs.flags + (cur.id - p.B * cs.nblk);
                v3_wait_flag(f);
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < kPairs; ++q) h[q] = __ldcg(carry + (size_t)q * p.ED);
        } else {
#pragma unroll
            for (int q = 0; q < kPairs; ++q) h[q] = make_float2(0.f, 0.f);
        }
        T *ob = reinterpret_cast<T *>(p.out) + (int64_t)cur.b * p.o_bs + cur.c0;
        T *yb = SAVE ? reinterpret_cast<T *>(p.ysave) + (int64_t)cur.b * p.L * p.ED + cur.c0 : nullptr;
        float2 *ckb = SAVE ? p.ckpt + (size_t)cur.b * p.nchunks * kPairs * p.ED + c : nullptr;   // [b][t/8][pair][ED]

        for (int k = 0; k < cur.nch; ++k, ++slot) {
            prefetch_next();
            cp_async_wait<NST - 1>();
            __syncwarp();
            const int tb = cur.t0 + k * kF3Chunk;
            const int nsteps = min(kF3Chunk, cur.t1 - tb);
            const unsigned char *s = smem + (slot % NST) * SM::kStage;
            const T *sU = reinterpret_cast<const T *>(s) + lane;
            const T *sD = reinterpret_cast<const T *>(s + SM::kTile) + lane;
            const T *sZ = reinterpret_cast<const T *>(s + 2 * SM::kTile) + lane;
            T *sYo = const_cast<T *>(sD);                  // y before the gate overwrites the delta tile (read below, once)
            T *sOo = const_cast<T *>(HAS_Z ? sZ : sU);     // out overwrites z (or u): each lane reads its slot first
            if (nsteps < kF3Chunk) {   // ragged tail: rows beyond the sequence hold stale bytes; dead steps multiply them by 0
                uint32_t *bc = reinterpret_cast<uint32_t *>(smem + (slot % NST) * SM::kStage + SM::kNT * SM::kTile);
                constexpr int WPR = 16 * (int)sizeof(T) / 4;   // words per row
                for (int i = nsteps * WPR + lane; i < kF3Chunk * WPR; i += 32) {
                    bc[i] = 0u;
                    bc[SM::kBC / 4 + i] = 0u;
                }
                __syncwarp();
            }
            const float4 *sB4, *sC4;
            if constexpr (sizeof(T) == 4) {
                sB4 = reinterpret_cast<const float4 *>(s + SM::kNT * SM::kTile);
                sC4 = reinterpret_cast<const float4 *>(s + SM::kNT * SM::kTile + SM::kBC);
            } else {   // B|C rows -> fp32: lane < 16 converts B row `lane`, lane >= 16 C row `lane - 16`
                float v[16];
                v3_cvt16<T>(reinterpret_cast<const T *>(s + SM::kNT * SM::kTile) + lane * 16, v);
                float4 *d = reinterpret_cast<float4 *>(smem + SM::kOffBCf) + lane * 4;
#pragma unroll
                for (int q = 0; q < 4; ++q) d[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                sB4 = reinterpret_cast<const float4 *>(smem + SM::kOffBCf);
                sC4 = sB4 + kF3Chunk * 4;
                __syncwarp();
            }

            // Groups of kF3Group steps: the loop body (~500 instructions) stays inside the instruction cache -- the fully
            // unrolled 16-step body measured 0.67 no-instruction stalls per issue with one warp per scheduler.
            auto group = [&](auto full_tag, int j0) {
                constexpr bool FULL = decltype(full_tag)::value;
                const T *gU = sU + j0 * 32, *gD = sD + j0 * 32, *gZ = sZ + j0 * 32;
                T *gYo = sYo + j0 * 32, *gOo = sOo + j0 * 32;
                const float4 *gB4 = sB4 + j0 * 4, *gC4 = sC4 + j0 * 4;
                float dl[kF3Group];
                {   // delta = softplus(raw + bias)
                    float x[kF3Group], sg[kF3Group];
#pragma unroll
                    for (int jj = 0; jj < kF3Group; ++jj) x[jj] = to_f(gD[jj * 32]) + bias;
                    if (sp) {
                        v3_softplus<kF3Group, false>(x, dl, sg);
                    } else {
#pragma unroll
                        for (int jj = 0; jj < kF3Group; ++jj) dl[jj] = x[jj];
                    }
                }
                if (SAVE && (j0 % kCkptV2) == 0 && (FULL || j0 < nsteps)) {   // state before step tb + j0
                    float2 *ck = ckb + (size_t)((tb + j0) / kCkptV2) * kPairs * p.ED;
#pragma unroll
                    for (int q = 0; q < kPairs; ++q) __stcs(ck + (size_t)q * p.ED, h[q]);
                }
#pragma unroll
                for (int jj = 0; jj < kF3Group; ++jj) {
                    const bool live = FULL || j0 + jj < nsteps;   // ragged tail: dead steps leave the state untouched
                    const float u = live ? to_f(gU[jj * 32]) : 0.f;
                    const float dlj = live ? dl[jj] : 0.f;
                    const float du = dlj * u;
                    const float2 dl2 = make_float2(dlj, dlj), du2 = make_float2(du, du);
                    float2 ya = make_float2(0.f, 0.f), yb2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 B4 = gB4[jj * 4 + q], C4 = gC4[jj * 4 + q];
                        const float2 x0 = fmul2(dl2, A2[2 * q]), x1 = fmul2(dl2, A2[2 * q + 1]);
                        const float2 a0 = (2 * q < kNPoly) ? ex2_poly2(x0) : ex2_2(x0);
                        const float2 a1 = (2 * q + 1 < kNPoly) ? ex2_poly2(x1) : ex2_2(x1);
                        h[2 * q] = ffma2(a0, h[2 * q], fmul2(du2, make_float2(B4.x, B4.y)));
                        h[2 * q + 1] = ffma2(a1, h[2 * q + 1], fmul2(du2, make_float2(B4.z, B4.w)));
                        ya = ffma2(h[2 * q], make_float2(C4.x, C4.y), ya);
                        yb2 = ffma2(h[2 * q + 1], make_float2(C4.z, C4.w), yb2);
                    }
                    const float2 ys = fadd2(ya, yb2);
                    float y = fmaf(Dc, u, ys.x + ys.y);
                    // results go back into the consumed tiles (same lane, same slot), and leave as 16-byte rows below
                    if (SAVE) gYo[jj * 32] = from_f<T>(y);
                    if (HAS_Z) {
                        const float z = to_f(gZ[jj * 32]);
                        y *= z * sigmoid_fast(z);
                    }
                    gOo[jj * 32] = from_f<T>(y);
                }
            };
            if (nsteps == kF3Chunk) {
#pragma unroll 1
                for (int j0 = 0; j0 < kF3Chunk; j0 += kF3Group) group(std::true_type{}, j0);
            } else {   // ragged tail of the sequence
#pragma unroll 1
                for (int j0 = 0; j0 < kF3Chunk; j0 += kF3Group) group(std::false_type{}, j0);
            }
            __syncwarp();
            v3_store32<T, kF3Chunk>(reinterpret_cast<const T *>(s + (HAS_Z ? 2 : 0) * SM::kTile), ob + (int64_t)tb * p.o_rs, p.o_rs, nsteps, lane, aligned);
            if (SAVE) v3_store32<T, kF3Chunk>(reinterpret_cast<const T *>(s + SM::kTile), yb + (int64_t)tb * p.ED, p.ED, nsteps, lane, aligned);
            __syncwarp();   // stage `slot` may be refilled by the next prefetch
        }

        // ---- unit epilogue: carry-out / final state ----
        if (cur.seg == cs.nseg - 1) {
            if (p.last_state != nullptr) {
                float4 *d = reinterpret_cast<float4 *>(p.last_state + ((size_t)cur.b * p.ED + c) * kNState);
#pragma unroll
                for (int q = 0; q < 4; ++q) d[q] = make_float4(h[2 * q].x, h[2 * q].y, h[2 * q + 1].x, h[2 * q + 1].y);
            }
        } else {
#pragma unroll
            for (int q = 0; q < kPairs; ++q) __stcg(carry + (size_t)q * p.ED, h[q]);
            __threadfence();
            __syncwarp();
            if (lane == 0) st_release(cs.flags + cur.id, 1);
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------------- host
bool v3_shape_ok(int B, int L, int ED) {
    if (ED % 32 != 0) return false;
    // Opt-in only (GFE_SELSCAN_V3=1, A/B measurements): one warp per scheduler caps the FP32 pipe at half rate, the
    // v3 pair measured 4.64 ms per cfg3 step against 3.46 ms for v2 (profiles/r01_ab_v2_v3.txt).
    const char *e = getenv("GFE_SELSCAN_V3");
    if (e == nullptr || e[0] != '1') return false;
    return plan_segments(B, L, ED).nseg == 1;   // enough (row, channel) chains; otherwise the L-split kernels
}

// Segments of 256 steps (>= 128): long enough to amortise the hand-off, short enough that the dynamic draw balances
// the chains over the schedulers to ~1 %.
void v3_plan(int B, int L, int ED, int &nblk, int &nseg, int &seg_len) {
    nblk = ED / 32;
    seg_len = 256;
    if (const char *e = getenv("GFE_V3_SEGLEN")) {
        const int v = atoi(e);
        if (v >= 16 && v % 16 == 0) seg_len = v;
    }
    nseg = (L + seg_len - 1) / seg_len;
    (void)B;
}

struct V3ChainLayout {
    size_t counter, flags, carry, total;
};
static V3ChainLayout v3_chain_layout(int B, int ED, int nblk, int nseg, int carry_floats) {
    V3ChainLayout c{};
    size_t off = 0;
    c.counter = off;
    off += 256;
    c.flags = off;
    off += align_up((size_t)nseg * B * nblk * sizeof(int), 256);
    c.carry = off;
    off += align_up((size_t)B * ED * carry_floats * sizeof(float), 256);
    c.total = off;
    return c;
}
size_t v3_chain_bytes(int B, int ED, int nblk, int nseg, int carry_floats) { return v3_chain_layout(B, ED, nblk, nseg, carry_floats).total; }

size_t v3_fwd_workspace_bytes(int B, int L, int ED) {
    int nblk, nseg, seg_len;
    v3_plan(B, L, ED, nblk, nseg, seg_len);
    return v3_chain_bytes(B, ED, nblk, nseg, kNState);
}

int v3_fill_sched(ChainSched &cs, char *ws, int B, int ED, int nblk, int nseg, int seg_len, int carry_floats, cudaStream_t st) {
    const V3ChainLayout cl = v3_chain_layout(B, ED, nblk, nseg, carry_floats);
    cs.counter = reinterpret_cast<int *>(ws + cl.counter);
    cs.flags = reinterpret_cast<int *>(ws + cl.flags);
    cs.carry = reinterpret_cast<float *>(ws + cl.carry);
    cs.nseg = nseg;
    cs.seg_len = seg_len;
    cs.nblk = nblk;
    cs.total = nseg * B * nblk;
    if (cudaMemsetAsync(ws, 0, cl.carry, st) != cudaSuccess) return check_launch("selscan_v3 memset");
    return GFE_OK;
}

void v3_fill_params(ScanParams &p, const gfe_selscan_args *a) {
    p.B = a->batch; p.L = a->seqlen; p.ED = a->d_inner;
    p.nseg = 1; p.seg_len = a->seqlen; p.nchunks = (a->seqlen + kCkptV2 - 1) / kCkptV2;   // checkpoints: [b][t / 8][pair][ED]
    p.flags = a->flags;
    p.u = a->u; p.delta = a->delta; p.z = a->z; p.Bm = a->Bm; p.Cm = a->Cm;
    p.u_bs = a->u_bs; p.u_rs = a->u_rs; p.d_bs = a->delta_bs; p.d_rs = a->delta_rs;
    p.z_bs = a->z_bs; p.z_rs = a->z_rs; p.B_bs = a->B_bs; p.B_rs = a->B_rs; p.C_bs = a->C_bs; p.C_rs = a->C_rs;
    p.A_log = a->A_log; p.D = a->D; p.dt_bias = a->dt_bias;
    p.ckpt = reinterpret_cast<float2 *>(a->ckpt);
    p.ysave = a->ckpt ? reinterpret_cast<char *>(a->ckpt) + (size_t)a->batch * p.nchunks * a->d_inner * kNState * sizeof(float) : nullptr;
    p.G = a->d_inner / 32;
}

// every staged tensor can be moved in 16-byte cp.async pieces
int v3_aligned(const gfe_selscan_args *a, bool bwd) {
    const int64_t s = a->dtype == GFE_F32 ? 4 : 2;
    const void *ptrs[10] = {a->u, a->delta, a->z, a->Bm, a->Cm, bwd ? a->dout : nullptr, bwd ? nullptr : a->out,
                            bwd ? a->du : nullptr, bwd ? a->ddelta : nullptr, bwd ? a->dz : nullptr};
    const int64_t strides[20] = {a->u_bs, a->u_rs, a->delta_bs, a->delta_rs, a->z ? a->z_bs : 0, a->z ? a->z_rs : 0,
                                 a->B_bs, a->B_rs, a->C_bs, a->C_rs, bwd ? a->dout_bs : 0, bwd ? a->dout_rs : 0,
                                 bwd ? 0 : a->out_bs, bwd ? 0 : a->out_rs, bwd ? a->du_bs : 0, bwd ? a->du_rs : 0,
                                 bwd ? a->ddelta_bs : 0, bwd ? a->ddelta_rs : 0, bwd && a->dz ? a->dz_bs : 0, bwd && a->dz ? a->dz_rs : 0};
    bool ok = true;
    for (const void *q : ptrs) ok &= (reinterpret_cast<uintptr_t>(q) & 15) == 0;
    for (int64_t stv : strides) ok &= (stv * s) % 16 == 0;
    if (a->ckpt) ok &= (reinterpret_cast<uintptr_t>(a->ckpt) & 15) == 0;
    if (const char *e = getenv("GFE_SELSCAN_PATH"))
        if (!strcmp(e, "plain")) ok = false;
    return ok ? 1 : 0;
}

int v3_warps_per_cta(size_t per_warp_smem) {
    int w = 4;
    if (const char *e = getenv("GFE_V3_WARPS")) {
        const int v = atoi(e);
        if (v >= 1 && v <= kV3MaxWarps) w = v;
    }
    while (w > 1 && (size_t)w * per_warp_smem > 220 * 1024) --w;
    return w;
}

template <typename T, bool HAS_Z, bool SAVE>
static void launch_fwd_v3_inst(const ScanParams &p, const ChainSched &cs, int aligned, cudaStream_t st) {
    auto kernel = selscan_fwd_v3_kernel<T, HAS_Z, SAVE>;
    constexpr size_t per_warp = FwdV3Smem<T, HAS_Z>::kPerWarp;
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
static int launch_fwd_v3_t(const gfe_selscan_args *a, cudaStream_t st) {
    int nblk, nseg, seg_len;
    v3_plan(a->batch, a->seqlen, a->d_inner, nblk, nseg, seg_len);
    const size_t need = v3_chain_bytes(a->batch, a->d_inner, nblk, nseg, kNState);
    if (a->ws == nullptr || a->ws_bytes < need) {
        set_error("selscan_fwd: workspace too small (%zu < %zu)", a->ws ? a->ws_bytes : (size_t)0, need);
        return GFE_ERR_WORKSPACE;
    }
    if (a->last_state && (reinterpret_cast<uintptr_t>(a->last_state) & 15) != 0) {
        set_error("selscan: last_state must be 16-byte aligned");
        return GFE_ERR_ARG;
    }
    ScanParams p{};
    v3_fill_params(p, a);
    p.out = a->out; p.o_bs = a->out_bs; p.o_rs = a->out_rs; p.last_state = a->last_state;
    ChainSched cs{};
    int rc = v3_fill_sched(cs, reinterpret_cast<char *>(a->ws), a->batch, a->d_inner, nblk, nseg, seg_len, kNState, st);
    if (rc != GFE_OK) return rc;
    const int aligned = v3_aligned(a, false);
    {
        ScopedKernelTimer tm(K_SELSCAN_FWD, st);
        const bool save = a->ckpt != nullptr;
        if (a->z != nullptr) {
            if (save) launch_fwd_v3_inst<T, true, true>(p, cs, aligned, st);
            else launch_fwd_v3_inst<T, true, false>(p, cs, aligned, st);
        } else {
            if (save) launch_fwd_v3_inst<T, false, true>(p, cs, aligned, st);
            else launch_fwd_v3_inst<T, false, false>(p, cs, aligned, st);
        }
    }
    return check_launch("selscan_fwd_v3");
}

int v3_launch_fwd(const gfe_selscan_args *a, cudaStream_t st) {
    switch (a->dtype) {
        case GFE_F32: return launch_fwd_v3_t<float>(a, st);
        case GFE_BF16: return launch_fwd_v3_t<__nv_bfloat16>(a, st);
        default: return launch_fwd_v3_t<__half>(a, st);
    }
}

}  // namespace gfe
