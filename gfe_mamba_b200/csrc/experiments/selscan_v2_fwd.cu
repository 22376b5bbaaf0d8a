// selscan_v2_fwd.cu -- fused selective-scan forward, "v2": CTA-cooperative staging, 4 lanes per channel, persistent
// CTAs with dynamically chained L-segments.  Replaces mamba.py:255-256, 275-284, 220-222 of the reference.
//
// Why this shape (profiles/r01_*): the first kernels were instruction-issue bound -- 191 warp-instructions per
// (32 channels x 1 step) of which only ~50 are the recurrence, at 2.3 warps per scheduler -- and cfg3's 768 warps on
// 592 schedulers quantised badly.  Here
//   * a CTA serves CPC adjacent channels of one batch row with 4 * CPC threads; thread (c, q) owns states 4q..4q+3 of
//     channel c as two float2 pairs, so a channel's per-step scalar work (softplus, SiLU, D skip, conversions, the
//     store) is done ONCE per (t, c) by a separate "item" mapping of the same threads instead of once per lane;
//   * the 16-step tiles of u, delta, z and the B|C rows are staged HBM -> shared memory by the whole CTA (cp.async,
//     3 stages), B|C are shared by all channels of the CTA;
//   * per step a lane issues 3 LDS.128 + 8 packed FP32 ops + 4 exp2 (a quarter of them on the FMA pipe) + 1 STS;
//   * time is cut into segments; persistent CTAs draw (segment, row, block) units from an atomic counter in
//     segment-major order and hand the state to the successor through global memory (ChainSched), which keeps every
//     SM busy to the end whatever B * ED / CPC is.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "selscan_shared.cuh"

namespace gfe {

constexpr int kV2Stages = 3;

template <typename T, bool HAS_Z, int CPC>
struct FwdV2Smem {
    static constexpr int kTile = kChunk * CPC * (int)sizeof(T);      // one of u, delta, z
    static constexpr int kBCRaw = kChunk * kNState * (int)sizeof(T);  // one of B, C
    static constexpr int kStage = (HAS_Z ? 3 : 2) * kTile + 2 * kBCRaw;
    static constexpr int kDDPlane = kChunk * (CPC / 2) + 4;           // float4 per parity plane (+64 B skew)
    static constexpr int kYPlane = CPC + 8;                           // floats per (t, quad) plane of the partial C.h
    static constexpr int kOffDD = kV2Stages * kStage;                 // float4 [2 parity][16][CPC/2] {dl, dl, dl*u, dl*u}
    static constexpr int kOffBC = kOffDD + 2 * kDDPlane * 16;         // float4 [16][8]  B quads | C quads
    static constexpr int kOffY = kOffBC + kChunk * 32 * 4;            // float  [16][4][CPC + 8]
    static constexpr int kTotal = kOffY + kChunk * 4 * kYPlane * 4;
};

// which (step, pair) exponentials run as a polynomial on the FMA pipe instead of MUFU.EX2
__device__ __forceinline__ constexpr bool fwd_v2_poly(int j, int pair) { return pair == 0 && (j & 1) == 1; }

template <typename T, bool HAS_Z, int CPB, int CPC>
__global__ void __launch_bounds__(4 * CPC, CPC == 64 ? 3 : 6) selscan_fwd_v2_kernel(ScanParams p, ChainSched cs) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_unit;
    using SM = FwdV2Smem<T, HAS_Z, CPC>;
    constexpr int NT = 4 * CPC;
    constexpr int NTILE = HAS_Z ? 3 : 2;
    constexpr int NP = CPC / 2;                 // channel pairs per row
    const int tid = threadIdx.x;
    const int rc = tid >> 2, rq = tid & 3;      // recurrence mapping: channel in block, state quad
    const int ip = tid % NP, ir = tid / NP;     // item mapping: channels 2 ip, 2 ip + 1; rows ir and ir + 8

    float4 *sDD = reinterpret_cast<float4 *>(smem + SM::kOffDD);
    float4 *sBC = reinterpret_cast<float4 *>(smem + SM::kOffBC);
    float *sY = reinterpret_cast<float *>(smem + SM::kOffY);
    const bool sp = p.flags & GFE_FLAG_DELTA_SOFTPLUS;
    const bool vec = p.flags & kFlagPairStores;
    const int per_seg = p.B * cs.nblk;

    const float4 *dd_r = sDD + (rc & 1) * SM::kDDPlane + (rc >> 1);
    const float4 *bc_r = sBC + rq;
    float *y_w = sY + rq * SM::kYPlane + rc;

    for (;;) {
        __syncthreads();   // every thread is done with the previous unit's shared memory
        if (tid == 0) s_unit = atomicAdd(cs.counter, 1);
        __syncthreads();
        const int unit = s_unit;
        if (unit >= cs.total) break;
        const int seg = unit / per_seg;
        const int rem = unit - seg * per_seg;
        const int b = rem / cs.nblk;
        const int c0 = (rem - b * cs.nblk) * CPC;
        const int t0 = seg * cs.seg_len, t1 = min(p.L, t0 + cs.seg_len);
        const int nch = (t1 - t0 + kChunk - 1) / kChunk;

        const T *ub = reinterpret_cast<const T *>(p.u) + (int64_t)b * p.u_bs + c0;
        const T *db = reinterpret_cast<const T *>(p.delta) + (int64_t)b * p.d_bs + c0;
        const T *zb = HAS_Z ? reinterpret_cast<const T *>(p.z) + (int64_t)b * p.z_bs + c0 : nullptr;
        const T *Bb = reinterpret_cast<const T *>(p.Bm) + (int64_t)b * p.B_bs;
        const T *Cb = reinterpret_cast<const T *>(p.Cm) + (int64_t)b * p.C_bs;
        T *ob = reinterpret_cast<T *>(p.out) + (int64_t)b * p.o_bs + c0 + 2 * ip;
        T *yb = p.ysave ? reinterpret_cast<T *>(p.ysave) + (int64_t)b * p.L * p.ED + c0 + 2 * ip : nullptr;
        float4 *ckq = p.ckpt ? reinterpret_cast<float4 *>(reinterpret_cast<float *>(p.ckpt) +
                                                          ((size_t)b * p.nchunks * p.ED + c0 + rc) * kNState) + rq : nullptr;
        const size_t ck_step = (size_t)p.ED * (kNState / 4);

        auto issue = [&](int k) {   // chunk k of this segment -> stage k % kV2Stages
            if (k < nch) {
                const int tb = t0 + k * kChunk;
                const int nrows = min(kChunk, t1 - tb);
                unsigned char *s = smem + (k % kV2Stages) * SM::kStage;
                stage_tile<T, CPB, CPC, NT>(s, ub + (int64_t)tb * p.u_rs, p.u_rs, nrows, tid);
                stage_tile<T, CPB, CPC, NT>(s + SM::kTile, db + (int64_t)tb * p.d_rs, p.d_rs, nrows, tid);
                if (HAS_Z) stage_tile<T, CPB, CPC, NT>(s + 2 * SM::kTile, zb + (int64_t)tb * p.z_rs, p.z_rs, nrows, tid);
                stage_tile<T, CPB, kNState, NT>(s + NTILE * SM::kTile, Bb + (int64_t)tb * p.B_rs, p.B_rs, nrows, tid);
                stage_tile<T, CPB, kNState, NT>(s + NTILE * SM::kTile + SM::kBCRaw, Cb + (int64_t)tb * p.C_rs, p.C_rs, nrows, tid);
            }
            cp_async_commit();
        };
#pragma unroll
        for (int k = 0; k < kV2Stages; ++k) issue(k);

        // per-thread constants: A (recurrence mapping), D and bias (item mapping)
        float2 A2[2], h[2];
        {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(p.A_log + (size_t)(c0 + rc) * kNState) + rq);
            A2[0] = make_float2(-expf(v.x) * kLog2e, -expf(v.y) * kLog2e);
            A2[1] = make_float2(-expf(v.z) * kLog2e, -expf(v.w) * kLog2e);
        }
        const float2 Dc = __ldg(reinterpret_cast<const float2 *>(p.D + c0) + ip);
        const float2 bias = p.dt_bias ? __ldg(reinterpret_cast<const float2 *>(p.dt_bias + c0) + ip) : make_float2(0.f, 0.f);

        // carry-in: the state our predecessor segment left behind
        float *carry = cs.carry + ((size_t)b * p.ED + c0 + rc) * kNState + 4 * rq;
        if (seg > 0) {
            if (tid == 0) {
                const int *f = cs.flags + (unit - per_seg);
                while (ld_acquire(f) == 0) __nanosleep(100);
            }
            __syncthreads();
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(carry));
            h[0] = make_float2(v.x, v.y);
            h[1] = make_float2(v.z, v.w);
        } else {
            h[0] = h[1] = make_float2(0.f, 0.f);
        }

        float2 Du[2], gate[2];
        auto phase_a = [&](int k) {   // per-(t, channel pair) scalars of chunk k -> shared slots; B|C rows -> fp32 quads
            const int tb = t0 + k * kChunk;
            const unsigned char *s = smem + (k % kV2Stages) * SM::kStage;
            const T *sU = reinterpret_cast<const T *>(s);
            const T *sD = reinterpret_cast<const T *>(s + SM::kTile);
            const T *sZ = reinterpret_cast<const T *>(s + 2 * SM::kTile);
            const T *sBr = reinterpret_cast<const T *>(s + NTILE * SM::kTile);   // B rows then C rows
            float x[4], dl[4], sg[4];
#pragma unroll
            for (int ps = 0; ps < 2; ++ps) {
                const float2 d2 = lds_pair(sD + (ps * 8 + ir) * CPC, ip);
                x[2 * ps] = d2.x + bias.x;
                x[2 * ps + 1] = d2.y + bias.y;
            }
            if (sp) {
                softplus_group<4, false>(x, dl, sg);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) dl[i] = x[i];
            }
#pragma unroll
            for (int ps = 0; ps < 2; ++ps) {
                const int t = ps * 8 + ir;
                const bool valid = tb + t < t1;
                const float2 u2 = lds_pair(sU + t * CPC, ip);
                const float u0 = valid ? u2.x : 0.f, u1 = valid ? u2.y : 0.f;
                const float dl0 = valid ? dl[2 * ps] : 0.f, dl1 = valid ? dl[2 * ps + 1] : 0.f;   // padded step: a = 1, bx = 0
                const float dlu0 = dl0 * u0, dlu1 = dl1 * u1;
                sDD[t * NP + ip] = make_float4(dl0, dl0, dlu0, dlu0);                  // even channel
                sDD[SM::kDDPlane + t * NP + ip] = make_float4(dl1, dl1, dlu1, dlu1);   // odd channel
                Du[ps] = make_float2(Dc.x * u0, Dc.y * u1);
                if (HAS_Z) {
                    const float2 z2 = lds_pair(sZ + t * CPC, ip);
                    gate[ps] = make_float2(z2.x * sigmoid_fast(z2.x), z2.y * sigmoid_fast(z2.y));
                }
            }
            if (tid < kChunk * 8) {   // B|C rows -> fp32 quads: [t][B quads 0..3 | C quads 0..3]
                const int t = tid >> 3, q8 = tid & 7;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tb + t < t1) {
                    const T *src = sBr + (q8 < 4 ? 0 : kChunk * 16) + t * 16 + 4 * (q8 & 3);
                    const float2 lo = lds_pair(src, 0), hi = lds_pair(src, 1);
                    v = make_float4(lo.x, lo.y, hi.x, hi.y);
                }
                sBC[tid] = v;
            }
        };

        cp_async_wait<kV2Stages - 1>();
        __syncthreads();
        phase_a(0);

        for (int k = 0; k < nch; ++k) {
            const int tb = t0 + k * kChunk;
            __syncthreads();   // (1) slots of chunk k are complete
            {   // ---- the recurrence: 16 steps of this lane's 4 states ----
#pragma unroll
                for (int j = 0; j < kChunk; ++j) {
                    if (j % kCkptV2 == 0 && ckq != nullptr && tb + j < t1)   // state before step tb + j, for backward
                        __stcs(ckq + (size_t)((tb + j) / kCkptV2) * ck_step, make_float4(h[0].x, h[0].y, h[1].x, h[1].y));
                    const float4 dd = dd_r[j * NP];
                    const float4 B4 = bc_r[j * 8], C4 = bc_r[j * 8 + 4];
                    const float2 dl2 = make_float2(dd.x, dd.y), du2 = make_float2(dd.z, dd.w);
                    const float2 x0 = fmul2(dl2, A2[0]), x1 = fmul2(dl2, A2[1]);
                    const float2 a0 = fwd_v2_poly(j, 0) ? ex2_poly2(x0) : ex2_2(x0);
                    const float2 a1 = fwd_v2_poly(j, 1) ? ex2_poly2(x1) : ex2_2(x1);
                    h[0] = ffma2(a0, h[0], fmul2(du2, make_float2(B4.x, B4.y)));
                    h[1] = ffma2(a1, h[1], fmul2(du2, make_float2(B4.z, B4.w)));
                    const float2 y2 = ffma2(h[1], make_float2(C4.z, C4.w), fmul2(h[0], make_float2(C4.x, C4.y)));
                    y_w[j * (4 * SM::kYPlane)] = y2.x + y2.y;
                }
            }
            cp_async_wait<kV2Stages - 2>();   // chunk k + 1 has landed (this thread's pieces)
            __syncthreads();                  // (2) partial sums complete; chunk k + 1 visible; stage k % S free
            issue(k + kV2Stages);

            // ---- per-(t, channel pair) epilogue of chunk k: sum the 4 lanes, D skip, gate, store ----
#pragma unroll
            for (int ps = 0; ps < 2; ++ps) {
                const int t = ps * 8 + ir;
                if (tb + t < t1) {
                    const float2 *yp = reinterpret_cast<const float2 *>(sY + (t * 4) * SM::kYPlane) + ip;
                    const float2 p0 = yp[0], p1 = yp[SM::kYPlane / 2], p2 = yp[SM::kYPlane], p3 = yp[3 * (SM::kYPlane / 2)];
                    float y0 = (p0.x + p1.x) + (p2.x + p3.x) + Du[ps].x;
                    float y1 = (p0.y + p1.y) + (p2.y + p3.y) + Du[ps].y;
                    if (yb != nullptr) stg_pair<T>(yb + (int64_t)(tb + t) * p.ED, y0, y1, true);
                    if (HAS_Z) { y0 *= gate[ps].x; y1 *= gate[ps].y; }
                    stg_pair<T>(ob + (int64_t)(tb + t) * p.o_rs, y0, y1, vec);
                }
            }
            if (k + 1 < nch) phase_a(k + 1);
        }

        // carry-out / final state
        const float4 hv = make_float4(h[0].x, h[0].y, h[1].x, h[1].y);
        if (seg == cs.nseg - 1) {
            if (p.last_state != nullptr)
                *(reinterpret_cast<float4 *>(p.last_state + ((size_t)b * p.ED + c0 + rc) * kNState) + rq) = hv;
        } else {
            __stcg(reinterpret_cast<float4 *>(carry), hv);
            __threadfence();
            __syncthreads();
            if (tid == 0) st_release(cs.flags + unit, 1);
        }
        cp_async_wait<0>();
    }
}

// ---------------------------------------------------------------------------------------------------- host
// Units per launch are sized for ~8 per resident CTA; both numbers are pure functions of the shape so that the
// workspace queries and the launches agree.
static void chain_plan(int B, int L, int nblk, int ctas_per_sm, int &nseg, int &seg_len) {
    const int64_t slots = (int64_t)sm_count() * ctas_per_sm;
    int64_t want = ceil_div64(8 * slots, (int64_t)B * nblk);
    const int64_t max_by_len = L / (8 * kChunk) > 0 ? L / (8 * kChunk) : 1;   // >= 128 steps per segment
    if (want > max_by_len) want = max_by_len;
    if (want > kMaxSeg) want = kMaxSeg;
    if (want < 1) want = 1;
    seg_len = (int)ceil_div64(ceil_div64(L, want), kChunk) * kChunk;
    nseg = (L + seg_len - 1) / seg_len;
}

static int fwd_cpc(int ED) { return ED % 64 == 0 ? 64 : 32; }

template <typename T>
void v4_launch_fwd_kernel(const ScanParams &p, const ChainSched &cs, bool has_z, int cpb, cudaStream_t st);   // selscan_v4_fwd.cu
template <typename T>
bool v5_launch_fwd_kernel(const ScanParams &p, const ChainSched &cs, bool has_z, int cpb, cudaStream_t st);   // selscan_v5_fwd.cu
static bool v5_enabled() {   // opt-in (GFE_SELSCAN_V5=1): measured 0.84 ms against 0.80 ms for v4 on cfg3 (profiles/r01_fwd_variants.txt)
    const char *e = getenv("GFE_SELSCAN_V5");
    return e != nullptr && e[0] == '1';
}
template <typename T>
void v6_launch_fwd_kernel(const ScanParams &p, const ChainSched &cs, bool has_z, int cpb, cudaStream_t st);   // selscan_v6_fwd.cu
static bool v6_enabled() {   // opt-in (GFE_SELSCAN_V6=1): measured 0.88 ms against 0.80 ms for v4 on cfg3 (profiles/r01_fwd_variants.txt)
    const char *e = getenv("GFE_SELSCAN_V6");
    return e != nullptr && e[0] == '1';
}
static bool v4_enabled() {   // GFE_SELSCAN_V4=0: A/B measurements against the v2 forward kernel
    const char *e = getenv("GFE_SELSCAN_V4");
    return e == nullptr || e[0] != '0';
}

void v2_fwd_plan(int B, int L, int ED, int &cpc, int &nblk, int &nseg, int &seg_len) {
    cpc = fwd_cpc(ED);
    nblk = ED / cpc;
    chain_plan(B, L, nblk, cpc == 64 ? 3 : 6, nseg, seg_len);
}
void v2_bwd_plan(int B, int L, int ED, int &nblk, int &nseg, int &seg_len) {
    nblk = ED / 32;
    chain_plan(B, L, nblk, 4, nseg, seg_len);
}

bool v2_applicable(int B, int L, int ED) {
    if (ED % 32 != 0) return false;
    if (const char *e = getenv("GFE_SELSCAN_V2")) {   // A/B measurements only
        if (e[0] == '0') return false;
        if (e[0] == '1') return true;
    }
    return plan_segments(B, L, ED).nseg == 1;   // enough (row, channel) parallelism: no L-split recomputation needed
}

struct ChainLayout {
    size_t counter, flags, carry, total;
};
static ChainLayout chain_layout(int B, int ED, int nblk, int nseg) {
    ChainLayout c{};
    size_t off = 0;
    c.counter = off;
    off += 256;
    c.flags = off;
    off += align_up((size_t)nseg * B * nblk * sizeof(int), 256);
    c.carry = off;
    off += align_up((size_t)B * ED * kNState * sizeof(float), 256);
    c.total = off;
    return c;
}

size_t v2_fwd_workspace_bytes(int B, int L, int ED) {
    int cpc, nblk, nseg, seg_len;
    v2_fwd_plan(B, L, ED, cpc, nblk, nseg, seg_len);
    return chain_layout(B, ED, nblk, nseg).total;
}

int v2_fill_sched(ChainSched &cs, char *ws, int B, int ED, int nblk, int nseg, int seg_len, cudaStream_t st) {
    const ChainLayout cl = chain_layout(B, ED, nblk, nseg);
    cs.counter = reinterpret_cast<int *>(ws + cl.counter);
    cs.flags = reinterpret_cast<int *>(ws + cl.flags);
    cs.carry = reinterpret_cast<float *>(ws + cl.carry);
    cs.nseg = nseg;
    cs.seg_len = seg_len;
    cs.nblk = nblk;
    cs.total = nseg * B * nblk;
    if (cudaMemsetAsync(ws, 0, cl.carry, st) != cudaSuccess) return check_launch("selscan_v2 memset");
    return GFE_OK;
}

size_t v2_chain_bytes(int B, int ED, int nblk, int nseg) { return chain_layout(B, ED, nblk, nseg).total; }

// the v2 kernels read D / dt_bias two channels at a time and the checkpoints as float4
int v2_check_alignment(const gfe_selscan_args *a) {
    if ((reinterpret_cast<uintptr_t>(a->D) & 7) != 0 || (a->dt_bias && (reinterpret_cast<uintptr_t>(a->dt_bias) & 7) != 0)) {
        set_error("selscan: D and dt_bias must be 8-byte aligned");
        return GFE_ERR_ARG;
    }
    if (a->ckpt && (reinterpret_cast<uintptr_t>(a->ckpt) & 15) != 0) {
        set_error("selscan: ckpt must be 16-byte aligned");
        return GFE_ERR_ARG;
    }
    if (a->last_state && (reinterpret_cast<uintptr_t>(a->last_state) & 15) != 0) {
        set_error("selscan: last_state must be 16-byte aligned");
        return GFE_ERR_ARG;
    }
    return GFE_OK;
}

void v2_fill_params(ScanParams &p, const gfe_selscan_args *a) {
    p.B = a->batch; p.L = a->seqlen; p.ED = a->d_inner;
    p.nseg = 1; p.seg_len = a->seqlen; p.nchunks = (a->seqlen + kCkptV2 - 1) / kCkptV2;   // number of checkpoints: [b][t / 8][c][16]
    p.flags = a->flags;
    p.u = a->u; p.delta = a->delta; p.z = a->z; p.Bm = a->Bm; p.Cm = a->Cm;
    p.u_bs = a->u_bs; p.u_rs = a->u_rs; p.d_bs = a->delta_bs; p.d_rs = a->delta_rs;
    p.z_bs = a->z_bs; p.z_rs = a->z_rs; p.B_bs = a->B_bs; p.B_rs = a->B_rs; p.C_bs = a->C_bs; p.C_rs = a->C_rs;
    p.A_log = a->A_log; p.D = a->D; p.dt_bias = a->dt_bias;
    p.ckpt = reinterpret_cast<float2 *>(a->ckpt);
    p.ysave = a->ckpt ? reinterpret_cast<char *>(a->ckpt) + (size_t)a->batch * p.nchunks * a->d_inner * kNState * sizeof(float) : nullptr;
    p.G = a->d_inner / 32;
}

// 16 when every staged tensor can be moved in 16-byte cp.async pieces, else 0 (plain loads)
int v2_cpb(const gfe_selscan_args *a, bool bwd) {
    const int64_t s = a->dtype == GFE_F32 ? 4 : 2;
    const void *ptrs[6] = {a->u, a->delta, a->z, a->Bm, a->Cm, bwd ? a->dout : nullptr};
    const int64_t strides[12] = {a->u_bs, a->u_rs, a->delta_bs, a->delta_rs, a->z ? a->z_bs : 0, a->z ? a->z_rs : 0,
                                 a->B_bs, a->B_rs, a->C_bs, a->C_rs, bwd ? a->dout_bs : 0, bwd ? a->dout_rs : 0};
    bool ok = true;
    for (const void *q : ptrs) ok &= (reinterpret_cast<uintptr_t>(q) & 15) == 0;
    for (int64_t stv : strides) ok &= (stv * s) % 16 == 0;
    if (a->ckpt) ok &= (reinterpret_cast<uintptr_t>(a->ckpt) & 15) == 0;
    if (const char *e = getenv("GFE_SELSCAN_PATH"))
        if (!strcmp(e, "plain")) ok = false;
    return ok ? 16 : 0;
}

// outputs may be written two channels at a time when every base and stride keeps the pair aligned
bool v2_pair_stores(const gfe_selscan_args *a, bool bwd) {
    const int64_t s = a->dtype == GFE_F32 ? 4 : 2;
    const int64_t al = 2 * s;
    bool ok = true;
    auto chk = [&](const void *q, int64_t bs, int64_t rs) {
        if (q == nullptr) return;
        ok &= (reinterpret_cast<uintptr_t>(q) % al) == 0 && (bs * s) % al == 0 && (rs * s) % al == 0;
    };
    if (bwd) {
        chk(a->du, a->du_bs, a->du_rs);
        chk(a->ddelta, a->ddelta_bs, a->ddelta_rs);
        chk(a->dz, a->dz_bs, a->dz_rs);
    } else {
        chk(a->out, a->out_bs, a->out_rs);
    }
    return ok;
}

template <typename K>
static int persistent_grid(K kernel, int nt, size_t smem, int total) {
    int per_sm = 0;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, nt, smem) != cudaSuccess || per_sm < 1) {
        (void)cudaGetLastError();
        per_sm = 1;
    }
    const int64_t slots = (int64_t)sm_count() * per_sm;
    return (int)(total < slots ? total : slots);
}

template <typename T, bool HAS_Z, int CPB, int CPC>
static void launch_fwd_v2_inst(const ScanParams &p, const ChainSched &cs, cudaStream_t st) {
    auto kernel = selscan_fwd_v2_kernel<T, HAS_Z, CPB, CPC>;
    constexpr size_t smem = FwdV2Smem<T, HAS_Z, CPC>::kTotal;
    static thread_local int grid_cache_total = -1, grid_cache = 0;
    if (grid_cache_total != cs.total) {
        grid_cache = persistent_grid(kernel, 4 * CPC, smem, cs.total);
        grid_cache_total = cs.total;
    }
    kernel<<<grid_cache, 4 * CPC, smem, st>>>(p, cs);
}

template <typename T>
static int launch_fwd_v2_t(const gfe_selscan_args *a, cudaStream_t st) {
    int cpc, nblk, nseg, seg_len;
    v2_fwd_plan(a->batch, a->seqlen, a->d_inner, cpc, nblk, nseg, seg_len);
    const size_t need = v2_chain_bytes(a->batch, a->d_inner, nblk, nseg);
    if (a->ws == nullptr || a->ws_bytes < need) {
        set_error("selscan_fwd: workspace too small (%zu < %zu)", a->ws ? a->ws_bytes : (size_t)0, need);
        return GFE_ERR_WORKSPACE;
    }
    int rc = v2_check_alignment(a);
    if (rc != GFE_OK) return rc;
    ScanParams p{};
    v2_fill_params(p, a);
    p.out = a->out; p.o_bs = a->out_bs; p.o_rs = a->out_rs; p.last_state = a->last_state;
    ChainSched cs{};
    rc = v2_fill_sched(cs, reinterpret_cast<char *>(a->ws), a->batch, a->d_inner, nblk, nseg, seg_len, st);
    if (rc != GFE_OK) return rc;
    const bool hz = a->z != nullptr;
    const int cpb = v2_cpb(a, false);
    if (v2_pair_stores(a, false)) p.flags |= kFlagPairStores;
    ScopedKernelTimer tm(K_SELSCAN_FWD, st);
#define GFE_V2F(HZ, CPB, CPC) launch_fwd_v2_inst<T, HZ, CPB, CPC>(p, cs, st)
    if (cpc == 64 && v4_enabled()) {
        // two channels per lane: warp-specialised (selscan_v5_fwd.cu), else the single-role kernel (selscan_v4_fwd.cu)
        if (v6_enabled() && (p.flags & kFlagPairStores)) v6_launch_fwd_kernel<T>(p, cs, hz, cpb, st);   // merged phases (selscan_v6_fwd.cu)
        else if (!(v5_enabled() && v5_launch_fwd_kernel<T>(p, cs, hz, cpb, st))) v4_launch_fwd_kernel<T>(p, cs, hz, cpb, st);
    } else if (cpc == 64) {
        if (hz && cpb == 16) GFE_V2F(true, 16, 64);
        else if (hz) GFE_V2F(true, 0, 64);
        else if (cpb == 16) GFE_V2F(false, 16, 64);
        else GFE_V2F(false, 0, 64);
    } else {
        if (hz && cpb == 16) GFE_V2F(true, 16, 32);
        else if (hz) GFE_V2F(true, 0, 32);
        else if (cpb == 16) GFE_V2F(false, 16, 32);
        else GFE_V2F(false, 0, 32);
    }
#undef GFE_V2F
    return check_launch("selscan_fwd_v2");
}

int v2_launch_fwd(const gfe_selscan_args *a, cudaStream_t st) {
    switch (a->dtype) {
        case GFE_F32: return launch_fwd_v2_t<float>(a, st);
        case GFE_BF16: return launch_fwd_v2_t<__nv_bfloat16>(a, st);
        default: return launch_fwd_v2_t<__half>(a, st);
    }
}

}  // namespace gfe
