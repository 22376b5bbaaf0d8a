// selscan_v6_fwd.cu -- fused selective-scan forward, "v6": the v4 lane layout with the three phases of a chunk merged
// into ONE instruction stream per half chunk.
// Replaces mamba.py:255-256, 275-284, 220-222 of the reference (softplus -> discretise -> scan -> C.h -> D skip -> gate).
//
// Why (profiles/r01_fwd_variants.txt): the chained forward is latency-bound along L.  In v4 a CTA runs item phase ->
// barrier -> 16 recurrence steps -> barrier -> epilogue, and each of the three is a dependent chain the warp's own
// static schedule cannot fill (2.2 issue clocks per instruction in the recurrence).  Moving the item work to other
// warps (v5) did not help: the same schedulers were simply shared by more half-idle warps.  v6 keeps the item work
// in the recurrence warps but removes the phase boundaries: per half chunk (8 steps) a thread executes, as one
// straight-line block that ptxas interleaves freely,
//     R(g)    8 recurrence steps of its 2 channels x 4 states        (slots[g & 1]      -> partial y[g & 1])
//     A(g+1)  the per-(t, channel pair) scalars of the NEXT half      (raw tiles         -> slots[(g + 1) & 1])
//     E(g-1)  quad sum, D skip, gate, stores of the PREVIOUS half     (partial y[(g-1)&1] -> HBM)
// with double-buffered slots and partial sums and ONE __syncthreads per half chunk.  softplus is the branch-free
// packed form (selscan_shared.cuh: softplus2) so that A has no vote / no branch to split the block.
// Lane layout, chaining, checkpoints and the saved y are those of v4, so the v2 backward consumes what this writes.
// MEASURED (profiles/r01_fwd_variants.txt): parity green, but 0.88 ms against v4's 0.80 ms on cfg3 -- ptxas still splits
// the block at the predicated stores and schedules every part at ~2 issue clocks per instruction, so the merge buys
// no overlap and costs the double buffers.  Kept opt-in (GFE_SELSCAN_V6=1) as a recorded experiment.
#include "common.cuh"
#include "selscan_shared.cuh"

namespace gfe {

constexpr int kV6CPC = 64;        // channels per CTA
constexpr int kV6NT = 128;        // lane (pair, quad) owns 2 channels x 4 states
constexpr int kV6H = 8;           // steps per merged block (== checkpoint interval)
constexpr int kV6YPlane = 36;     // float2 per (t, quad) plane of the partial C.h (32 pairs + 32 B skew)
#ifndef GFE_V6_MINB
#define GFE_V6_MINB 3
#endif
static_assert(kV6H == kCkptV2, "the merged block is one checkpoint interval");

template <typename T, bool HAS_Z>
struct FwdV6Smem {
    static constexpr int kStages = sizeof(T) == 2 ? 4 : 3;   // a chunk's tiles are re-read by E one block after its last step
    static constexpr int kNTile = HAS_Z ? 3 : 2;
    static constexpr int kTile = kChunk * kV6CPC * (int)sizeof(T);      // one of u, delta, z (16 rows)
    static constexpr int kBCRaw = kChunk * kNState * (int)sizeof(T);     // one of B, C
    static constexpr int kStage = kNTile * kTile + 2 * kBCRaw;
    static constexpr int kDD = kV6H * (kV6CPC / 2) * 16;                 // float4 [8][32] {dl0, dl0*u0, dl1, dl1*u1}
    static constexpr int kBC = kV6H * 8 * 16;                            // float4 [8][8]  B quads | C quads
    static constexpr int kY = kV6H * 4 * kV6YPlane * 8;                  // float2 [8][4][36]
    static constexpr int kOffDD = kStages * kStage;                      // x2 buffers
    static constexpr int kOffBC = kOffDD + 2 * kDD;
    static constexpr int kOffY = kOffBC + 2 * kBC;
    static constexpr int kTotal = kOffY + 2 * kY;
};

template <typename T, bool HAS_Z, int CPB>
__global__ void __launch_bounds__(kV6NT, GFE_V6_MINB) selscan_fwd_v6_kernel(ScanParams p, ChainSched cs) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_unit;
    using SM = FwdV6Smem<T, HAS_Z>;
    constexpr int NT = kV6NT, CPC = kV6CPC, NST = SM::kStages, NTILE = SM::kNTile, H = kV6H;
    constexpr int RB = CPC * (int)sizeof(T);          // bytes per tile row
    constexpr int PPT = kChunk * RB / 16 / NT;        // 16-byte pieces per thread per activation tile (1 bf16, 2 fp32)
    constexpr int BCP = kChunk * kNState * (int)sizeof(T) / 16;   // pieces per B (or C) tile: 32 bf16, 64 fp32
    const int tid = threadIdx.x;
    const int rp = tid >> 2, rq = tid & 3;     // recurrence mapping: channel pair in block (0..31), state quad
    const int ip = tid & 31, ir = tid >> 5;    // item mapping: channel pair, rows ir and ir + 4 of a half chunk

    float4 *sDD = reinterpret_cast<float4 *>(smem + SM::kOffDD);   // [buf][t][pair] {dl0, dl0*u0, dl1, dl1*u1}
    float4 *sBC = reinterpret_cast<float4 *>(smem + SM::kOffBC);   // [buf][t][B quads 0..3 | C quads 0..3]
    float2 *sY = reinterpret_cast<float2 *>(smem + SM::kOffY);     // [buf][t][quad][pair]
    const bool sp = p.flags & GFE_FLAG_DELTA_SOFTPLUS;
    const int per_seg = p.B * cs.nblk;

    const float4 *dd_r = sDD + rp;
    const float4 *bc_r = sBC + rq;
    float2 *y_w = sY + rq * kV6YPlane + rp;

    const int srow = (tid * PPT) / (RB / 16), spiece = (tid * PPT) % (RB / 16);
    const bool bc_thread = tid < 2 * BCP;
    const int bcsel = tid / BCP;
    const int bcrow = (tid % BCP) / (BCP / kChunk), bcpiece = (tid % BCP) % (BCP / kChunk);

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
        const int G = 2 * nch;                                  // half chunks of this unit

        const T *ub = reinterpret_cast<const T *>(p.u) + (int64_t)b * p.u_bs + c0;
        const T *db = reinterpret_cast<const T *>(p.delta) + (int64_t)b * p.d_bs + c0;
        const T *zb = HAS_Z ? reinterpret_cast<const T *>(p.z) + (int64_t)b * p.z_bs + c0 : nullptr;
        const T *Bb = reinterpret_cast<const T *>(p.Bm) + (int64_t)b * p.B_bs;
        const T *Cb = reinterpret_cast<const T *>(p.Cm) + (int64_t)b * p.C_bs;
        T *ob = reinterpret_cast<T *>(p.out) + (int64_t)b * p.o_bs + c0 + 2 * ip;
        T *yb = p.ysave ? reinterpret_cast<T *>(p.ysave) + (int64_t)b * p.L * p.ED + c0 + 2 * ip : nullptr;
        float4 *ckq = p.ckpt ? reinterpret_cast<float4 *>(reinterpret_cast<float *>(p.ckpt) +
                                                          ((size_t)b * p.nchunks * p.ED + c0 + 2 * rp) * kNState) + rq : nullptr;
        const size_t ck_step = (size_t)p.ED * (kNState / 4);
        const uint32_t dst_act = smem_u32(smem) + srow * RB + spiece * 16;
        const uint32_t dst_bc = smem_u32(smem) + NTILE * SM::kTile + bcsel * SM::kBCRaw + (tid % BCP) * 16;

        auto issue = [&](int k) {   // chunk k of this segment -> stage k % NST
            if (k < nch) {
                const int tb = t0 + k * kChunk;
                const int nrows = min(kChunk, t1 - tb);
                const uint32_t so = (k % NST) * SM::kStage;
                if constexpr (CPB == 16) {
                    if (srow < nrows) {
                        const int64_t r = tb + srow;
#pragma unroll
                        for (int i = 0; i < PPT; ++i) {
                            const int o = (spiece + i) * 16;
                            cp_async<16>(dst_act + so + i * 16, reinterpret_cast<const char *>(ub + r * p.u_rs) + o);
                            cp_async<16>(dst_act + so + SM::kTile + i * 16, reinterpret_cast<const char *>(db + r * p.d_rs) + o);
                            if (HAS_Z) cp_async<16>(dst_act + so + 2 * SM::kTile + i * 16, reinterpret_cast<const char *>(zb + r * p.z_rs) + o);
                        }
                    }
                    if (bc_thread && bcrow < nrows) {
                        const int64_t r = tb + bcrow;
                        cp_async<16>(dst_bc + so, reinterpret_cast<const char *>(bcsel ? Cb + r * p.C_rs : Bb + r * p.B_rs) + bcpiece * 16);
                    }
                } else {
                    unsigned char *s = smem + so;
                    stage_tile<T, 0, CPC, NT>(s, ub + (int64_t)tb * p.u_rs, p.u_rs, nrows, tid);
                    stage_tile<T, 0, CPC, NT>(s + SM::kTile, db + (int64_t)tb * p.d_rs, p.d_rs, nrows, tid);
                    if (HAS_Z) stage_tile<T, 0, CPC, NT>(s + 2 * SM::kTile, zb + (int64_t)tb * p.z_rs, p.z_rs, nrows, tid);
                    stage_tile<T, 0, kNState, NT>(s + NTILE * SM::kTile, Bb + (int64_t)tb * p.B_rs, p.B_rs, nrows, tid);
                    stage_tile<T, 0, kNState, NT>(s + NTILE * SM::kTile + SM::kBCRaw, Cb + (int64_t)tb * p.C_rs, p.C_rs, nrows, tid);
                }
            }
            cp_async_commit();
        };
#pragma unroll
        for (int k = 0; k < NST; ++k) issue(k);

        float2 A2[2][2], h[2][2];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(p.A_log + (size_t)(c0 + 2 * rp + ch) * kNState) + rq);
            A2[ch][0] = make_float2(-expf(v.x) * kLog2e, -expf(v.y) * kLog2e);
            A2[ch][1] = make_float2(-expf(v.z) * kLog2e, -expf(v.w) * kLog2e);
        }
        const float2 Dc = __ldg(reinterpret_cast<const float2 *>(p.D + c0) + ip);
        const float2 bias = p.dt_bias ? __ldg(reinterpret_cast<const float2 *>(p.dt_bias + c0) + ip) : make_float2(0.f, 0.f);

        float *carry = cs.carry + ((size_t)b * p.ED + c0 + 2 * rp) * kNState + 4 * rq;
        if (seg > 0) {
            if (tid == 0) {
                const int *f = cs.flags + (unit - per_seg);
                while (ld_acquire(f) == 0) __nanosleep(100);
            }
            __syncthreads();
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const float4 v = __ldcg(reinterpret_cast<const float4 *>(carry + ch * kNState));
                h[ch][0] = make_float2(v.x, v.y);
                h[ch][1] = make_float2(v.z, v.w);
            }
        } else {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) h[ch][0] = h[ch][1] = make_float2(0.f, 0.f);
        }

        // ---- A: slots of half chunk g ({dl, dl*u} per channel pair, fp32 B|C quads), branch-free ----
        auto phase_a = [&](int g, int buf) {   // buf == g & 1 except for the harmless extra call of the last block
            const int k = g >> 1, r0 = (g & 1) * H;
            const int tb = t0 + k * kChunk + r0;
            const unsigned char *s = smem + (k % NST) * SM::kStage;
            const T *sU = reinterpret_cast<const T *>(s) + r0 * CPC;
            const T *sD = reinterpret_cast<const T *>(s + SM::kTile) + r0 * CPC;
            const T *sBr = reinterpret_cast<const T *>(s + NTILE * SM::kTile) + r0 * kNState;   // B rows; C rows 16 * 16 later
            float4 *dd_w = sDD + buf * (SM::kDD / 16);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int t = ir + 4 * i;
                const bool valid = tb + t < t1;
                const float2 d2 = lds_pair(sD + t * CPC, ip);
                const float2 u2 = lds_pair(sU + t * CPC, ip);
                const float2 x = make_float2(d2.x + bias.x, d2.y + bias.y);
                float2 sg;
                const float2 spv = softplus2<false>(x, sg);
                const float dl0 = valid ? (sp ? spv.x : x.x) : 0.f, dl1 = valid ? (sp ? spv.y : x.y) : 0.f;   // padded step: a = 1, bx = 0
                const float u0 = valid ? u2.x : 0.f, u1 = valid ? u2.y : 0.f;
                dd_w[t * (CPC / 2) + ip] = make_float4(dl0, dl0 * u0, dl1, dl1 * u1);
            }
            if (tid < H * 8) {   // B|C rows -> fp32 quads (warps 0, 1)
                const int t = tid >> 3, q8 = tid & 7;
                const bool valid = tb + t < t1;
                const T *src = sBr + (q8 < 4 ? 0 : kChunk * 16) + t * 16 + 4 * (q8 & 3);
                const float2 lo = lds_pair(src, 0), hi = lds_pair(src, 1);
                sBC[buf * (SM::kBC / 16) + tid] = valid ? make_float4(lo.x, lo.y, hi.x, hi.y) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        // ---- E: quad sum, D skip, gate, stores of half chunk g ----
        auto epilogue = [&](int g, bool store) {
            const int k = g >> 1, r0 = (g & 1) * H, buf = g & 1;
            const int tb = t0 + k * kChunk + r0;
            const unsigned char *s = smem + (k % NST) * SM::kStage;
            const T *sU = reinterpret_cast<const T *>(s) + r0 * CPC;
            const T *sZ = reinterpret_cast<const T *>(s + 2 * SM::kTile) + r0 * CPC;
            const float2 *y_r = sY + buf * (SM::kY / 8) + ip;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int t = ir + 4 * i;
                const float2 *yp = y_r + (t * 4) * kV6YPlane;
                const float2 p0 = yp[0], p1 = yp[kV6YPlane], p2 = yp[2 * kV6YPlane], p3 = yp[3 * kV6YPlane];
                const float2 u2 = lds_pair(sU + t * CPC, ip);
                float y0 = fmaf(Dc.x, u2.x, (p0.x + p1.x) + (p2.x + p3.x));
                float y1 = fmaf(Dc.y, u2.y, (p0.y + p1.y) + (p2.y + p3.y));
                float g0 = 1.f, g1 = 1.f;
                if (HAS_Z) {
                    const float2 z2 = lds_pair(sZ + t * CPC, ip);
                    g0 = z2.x * sigmoid_fast(z2.x);
                    g1 = z2.y * sigmoid_fast(z2.y);
                }
                if (store && tb + t < t1) {   // pair stores are guaranteed by the dispatch (kFlagPairStores)
                    if (yb != nullptr) stg_pair<T>(yb + (int64_t)(tb + t) * p.ED, y0, y1, true);
                    stg_pair<T>(ob + (int64_t)(tb + t) * p.o_rs, y0 * g0, y1 * g1, true);
                }
            }
        };

        cp_async_wait<NST - 1>();
        __syncthreads();
        phase_a(0, 0);

        for (int g = 0; g < G; ++g) {
            const int buf = g & 1;
            const int tb = t0 + g * H;
            if (g & 1) cp_async_wait<NST - 3>();   // chunk (g + 1) / 2 has landed (this thread's pieces)
            __syncthreads();   // slots(g) and partial y(g - 1) are complete; the tiles A(g + 1) reads are visible
            if (g >= 3 && (g & 1)) issue((g - 3) / 2 + NST);   // chunk (g - 3) / 2 was last read by E(g - 2): its stage is free
            if (ckq != nullptr && tb < t1) {   // states before step tb, for backward
                float4 *dst = ckq + (size_t)(tb / kCkptV2) * ck_step;
                __stcs(dst, make_float4(h[0][0].x, h[0][0].y, h[0][1].x, h[0][1].y));
                __stcs(dst + kNState / 4, make_float4(h[1][0].x, h[1][0].y, h[1][1].x, h[1][1].y));
            }
            const float4 *dd_p = dd_r + buf * (SM::kDD / 16);
            const float4 *bc_p = bc_r + buf * (SM::kBC / 16);
            float2 *y_p = y_w + buf * (SM::kY / 8);
            // ---- one straight-line block: R(g) with A(g + 1) and E(g - 1) placed inside it ----
#pragma unroll
            for (int j = 0; j < H; ++j) {
                const float4 dd = dd_p[j * (CPC / 2)];
                const float4 B4 = bc_p[j * 8], C4 = bc_p[j * 8 + 4];
                const float2 B01 = make_float2(B4.x, B4.y), B23 = make_float2(B4.z, B4.w);
                const float2 C01 = make_float2(C4.x, C4.y), C23 = make_float2(C4.z, C4.w);
                float yv[2];
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    const float dl = ch ? dd.z : dd.x, du = ch ? dd.w : dd.y;
                    const float2 x0 = fmul2(splat2(dl), A2[ch][0]), x1 = fmul2(splat2(dl), A2[ch][1]);
                    const float2 a0 = (ch == 0 && (j & 1)) ? ex2_poly2(x0) : ex2_2(x0);   // 2 of the 8 exps of a step on the FMA pipe
                    const float2 a1 = (ch == 1 && !(j & 1)) ? ex2_poly2(x1) : ex2_2(x1);
                    h[ch][0] = ffma2(a0, h[ch][0], fmul2(splat2(du), B01));
                    h[ch][1] = ffma2(a1, h[ch][1], fmul2(splat2(du), B23));
                    const float2 y2 = ffma2(h[ch][1], C23, fmul2(h[ch][0], C01));
                    yv[ch] = y2.x + y2.y;
                }
                y_p[j * (4 * kV6YPlane)] = make_float2(yv[0], yv[1]);
                // unconditional so that the block stays straight-line: the last A rewrites a slot buffer nobody reads any
                // more, the first E reads a buffer still being written and stores nothing
                if (j == 1) phase_a(min(g + 1, G - 1), buf ^ 1);
                if (j == 4) epilogue(max(g - 1, 0), g >= 1);
            }
        }
        __syncthreads();
        epilogue(G - 1, true);

        // carry-out / final state
        if (seg == cs.nseg - 1) {
            if (p.last_state != nullptr) {
#pragma unroll
                for (int ch = 0; ch < 2; ++ch)
                    *(reinterpret_cast<float4 *>(p.last_state + ((size_t)b * p.ED + c0 + 2 * rp + ch) * kNState) + rq) =
                        make_float4(h[ch][0].x, h[ch][0].y, h[ch][1].x, h[ch][1].y);
            }
        } else {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
                __stcg(reinterpret_cast<float4 *>(carry + ch * kNState), make_float4(h[ch][0].x, h[ch][0].y, h[ch][1].x, h[ch][1].y));
            __threadfence();
            __syncthreads();
            if (tid == 0) st_release(cs.flags + unit, 1);
        }
        cp_async_wait<0>();
    }
}

// ---------------------------------------------------------------------------------------------------- host
template <typename T, bool HAS_Z, int CPB>
static void launch_fwd_v6_inst(const ScanParams &p, const ChainSched &cs, cudaStream_t st) {
    auto kernel = selscan_fwd_v6_kernel<T, HAS_Z, CPB>;
    constexpr size_t smem = FwdV6Smem<T, HAS_Z>::kTotal;
    static thread_local int cache_total = -1, cache_grid = 0;
    if (cache_total != cs.total) {
        int per_sm = 0;
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kV6NT, smem) != cudaSuccess || per_sm < 1) {
            (void)cudaGetLastError();
            per_sm = 1;
        }
        const int64_t slots = (int64_t)sm_count() * per_sm;
        cache_grid = (int)(cs.total < slots ? cs.total : slots);
        cache_total = cs.total;
    }
    kernel<<<cache_grid, kV6NT, smem, st>>>(p, cs);
}

template <typename T>
void v6_launch_fwd_kernel(const ScanParams &p, const ChainSched &cs, bool has_z, int cpb, cudaStream_t st) {
    if (has_z) {
        if (cpb == 16) launch_fwd_v6_inst<T, true, 16>(p, cs, st);
        else launch_fwd_v6_inst<T, true, 0>(p, cs, st);
    } else {
        if (cpb == 16) launch_fwd_v6_inst<T, false, 16>(p, cs, st);
        else launch_fwd_v6_inst<T, false, 0>(p, cs, st);
    }
}
template void v6_launch_fwd_kernel<float>(const ScanParams &, const ChainSched &, bool, int, cudaStream_t);
template void v6_launch_fwd_kernel<__nv_bfloat16>(const ScanParams &, const ChainSched &, bool, int, cudaStream_t);
template void v6_launch_fwd_kernel<__half>(const ScanParams &, const ChainSched &, bool, int, cudaStream_t);

}  // namespace gfe
