// selscan_v5_fwd.cu -- fused selective-scan forward, "v5": the v4 lane layout with WARP-SPECIALISED roles.
// Replaces mamba.py:255-256, 275-284, 220-222 of the reference (softplus -> discretise -> scan -> C.h -> D skip -> gate).
//
// Why (profiles/r01_small_batch.txt): a chained CTA is a serial pipeline -- item phase, barrier, 16 recurrence steps,
// barrier, epilogue -- and takes ~4500 clocks per 16-step chunk even alone on an SM (cfg3 with B = 1 runs as long as
// B = 4).  The time along L is latency, not throughput, and cfg3 has only 2.6 chains per SM to hide it.  v5 cuts the
// chain's critical path to the recurrence alone:
//   * warps 0-3 ("R", 128 threads, setmaxnreg 112) do nothing but the recurrence: lane (pair, quad) owns 4 states of
//     two adjacent channels, reads {dl, dl*u} x2 + B quad + C quad per step, writes the partial C.h;
//   * warps 4-7 ("I", 128 threads, setmaxnreg 48) stage the tiles (cp.async), run the per-(t, channel) scalar work
//     (softplus, delta*u, fp32 B|C) one half chunk AHEAD of R, and finish the half chunk BEHIND R (sum of the quads,
//     D skip, SiLU gate, stores);
//   * the two roles meet only through double-buffered shared slots and named barriers (full / empty per buffer),
//     8 steps at a time; there is no CTA-wide barrier inside a unit.
// Segment chaining, checkpoints ([b][t/8][c][16] fp32, written by R) and the saved y are those of v2/v4.
#include "common.cuh"
#include "selscan_shared.cuh"

namespace gfe {

constexpr int kV5CPC = 64;        // channels per CTA
constexpr int kV5NT = 256;        // 4 R warps + 4 I warps
constexpr int kV5H = 8;           // steps per hand-off (== checkpoint interval)
constexpr int kV5YPlane = 36;     // float2 per (t, quad) plane of the partial C.h (32 pairs + 32 B skew)
constexpr int kV5RegsLaunch = 80, kV5RegsR = 112, kV5RegsI = 48;   // 128 * 112 + 128 * 48 == 256 * 80
static_assert(kV5H == kCkptV2, "the hand-off granularity is the checkpoint interval");

enum : int { kBarFullS = 1, kBarEmptyS = 3, kBarFullY = 5, kBarEmptyY = 7, kBarI = 9 };   // + buffer index; 0 = __syncthreads

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <typename T, bool HAS_Z>
struct FwdV5Smem {
    static constexpr int kStages = 3;
    static constexpr int kNTile = HAS_Z ? 3 : 2;
    static constexpr int kTile = kChunk * kV5CPC * (int)sizeof(T);      // one of u, delta, z (16 rows)
    static constexpr int kBCRaw = kChunk * kNState * (int)sizeof(T);     // one of B, C
    static constexpr int kStage = kNTile * kTile + 2 * kBCRaw;
    static constexpr int kDD = kV5H * (kV5CPC / 2) * 16;                 // float4 [8][32] {dl0, dl0*u0, dl1, dl1*u1}
    static constexpr int kBC = kV5H * 8 * 16;                            // float4 [8][8]  B quads | C quads
    static constexpr int kY = kV5H * 4 * kV5YPlane * 8;                  // float2 [8][4][36]
    static constexpr int kOffDD = kStages * kStage;                      // x2 buffers
    static constexpr int kOffBC = kOffDD + 2 * kDD;
    static constexpr int kOffY = kOffBC + 2 * kBC;
    static constexpr int kTotal = kOffY + 2 * kY;
};

template <typename T, bool HAS_Z, int CPB>
__global__ void __launch_bounds__(kV5NT, 3) selscan_fwd_v5_kernel(ScanParams p, ChainSched cs) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_unit;
    using SM = FwdV5Smem<T, HAS_Z>;
    constexpr int CPC = kV5CPC, NST = SM::kStages, NTILE = SM::kNTile, H = kV5H;
    const int tid = threadIdx.x;
    const int per_seg = p.B * cs.nblk;

    if (tid < 128) {
        // =============================================================== R: the recurrence
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kV5RegsR));
        const int rp = tid >> 2, rq = tid & 3;     // channel pair in block (0..31), state quad
        const float4 *dd_r = reinterpret_cast<const float4 *>(smem + SM::kOffDD) + rp;
        const float4 *bc_r = reinterpret_cast<const float4 *>(smem + SM::kOffBC) + rq;
        float2 *y_w = reinterpret_cast<float2 *>(smem + SM::kOffY) + rq * kV5YPlane + rp;
        for (;;) {
            __syncthreads();
            if (tid == 0) s_unit = atomicAdd(cs.counter, 1);
            __syncthreads();
            const int unit = s_unit;
            if (unit >= cs.total) break;
            const int seg = unit / per_seg;
            const int rem = unit - seg * per_seg;
            const int b = rem / cs.nblk;
            const int c0 = (rem - b * cs.nblk) * CPC;
            const int t0 = seg * cs.seg_len, t1 = min(p.L, t0 + cs.seg_len);
            const int G = 2 * ((t1 - t0 + kChunk - 1) / kChunk);   // half chunks of this unit
            float4 *ckq = p.ckpt ? reinterpret_cast<float4 *>(reinterpret_cast<float *>(p.ckpt) +
                                                              ((size_t)b * p.nchunks * p.ED + c0 + 2 * rp) * kNState) + rq : nullptr;
            const size_t ck_step = (size_t)p.ED * (kNState / 4);

            float2 A2[2][2], h[2][2];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(p.A_log + (size_t)(c0 + 2 * rp + ch) * kNState) + rq);
                A2[ch][0] = make_float2(-expf(v.x) * kLog2e, -expf(v.y) * kLog2e);
                A2[ch][1] = make_float2(-expf(v.z) * kLog2e, -expf(v.w) * kLog2e);
            }
            float *carry = cs.carry + ((size_t)b * p.ED + c0 + 2 * rp) * kNState + 4 * rq;
            if (seg > 0) {
                if (tid == 0) {
                    const int *f = cs.flags + (unit - per_seg);
                    while (ld_acquire(f) == 0) __nanosleep(100);
                }
                bar_sync(kBarI + 1, 128);   // R-internal: the predecessor's carry is published
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

            for (int g = 0; g < G; ++g) {
                const int buf = g & 1;
                const int tb = t0 + g * H;
                bar_sync(kBarFullS + buf, 256);                 // the slots of half chunk g are written
                if (g >= 2) bar_sync(kBarEmptyY + buf, 256);    // the epilogue of half chunk g - 2 has read this y buffer
                if (ckq != nullptr && tb < t1) {                // states before step tb, for backward
                    float4 *dst = ckq + (size_t)(tb / kCkptV2) * ck_step;
                    __stcs(dst, make_float4(h[0][0].x, h[0][0].y, h[0][1].x, h[0][1].y));
                    __stcs(dst + kNState / 4, make_float4(h[1][0].x, h[1][0].y, h[1][1].x, h[1][1].y));
                }
                const float4 *dd_p = dd_r + buf * (SM::kDD / 16);
                const float4 *bc_p = bc_r + buf * (SM::kBC / 16);
                float2 *y_p = y_w + buf * (SM::kY / 8);
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
                    y_p[j * (4 * kV5YPlane)] = make_float2(yv[0], yv[1]);
                }
                if (g + 2 < G) bar_arrive(kBarEmptyS + buf, 256);   // the slots may be overwritten
                bar_arrive(kBarFullY + buf, 256);                   // the partial sums are complete
            }

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
                bar_sync(kBarI + 1, 128);   // R-internal: every lane's carry is out
                if (tid == 0) st_release(cs.flags + unit, 1);
            }
        }
    } else {
        // =============================================================== I: staging, item phase, epilogue
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kV5RegsI));
        constexpr int NT = 128;
        constexpr int RB = CPC * (int)sizeof(T);          // bytes per tile row
        constexpr int PPT = kChunk * RB / 16 / NT;        // 16-byte pieces per thread per activation tile (1 bf16, 2 fp32)
        constexpr int BCP = kChunk * kNState * (int)sizeof(T) / 16;   // pieces per B (or C) tile: 32 bf16, 64 fp32
        const int it = tid - 128;
        const int ip = it & 31, ir = it >> 5;      // item mapping: channel pair, rows ir and ir + 4 of a half chunk
        float4 *sDD = reinterpret_cast<float4 *>(smem + SM::kOffDD);
        float4 *sBC = reinterpret_cast<float4 *>(smem + SM::kOffBC);
        const float2 *sY = reinterpret_cast<const float2 *>(smem + SM::kOffY);
        const bool sp = p.flags & GFE_FLAG_DELTA_SOFTPLUS;
        const bool vec = p.flags & kFlagPairStores;
        const int srow = (it * PPT) / (RB / 16), spiece = (it * PPT) % (RB / 16);
        const bool bc_thread = it < 2 * BCP;
        const int bcsel = it / BCP;
        const int bcrow = (it % BCP) / (BCP / kChunk), bcpiece = (it % BCP) % (BCP / kChunk);
        for (;;) {
            __syncthreads();
            __syncthreads();
            const int unit = s_unit;
            if (unit >= cs.total) break;
            const int seg = unit / per_seg;
            const int rem = unit - seg * per_seg;
            const int b = rem / cs.nblk;
            const int c0 = (rem - b * cs.nblk) * CPC;
            const int t0 = seg * cs.seg_len, t1 = min(p.L, t0 + cs.seg_len);
            const int nch = (t1 - t0 + kChunk - 1) / kChunk;
            const int G = 2 * nch;

            const T *ub = reinterpret_cast<const T *>(p.u) + (int64_t)b * p.u_bs + c0;
            const T *db = reinterpret_cast<const T *>(p.delta) + (int64_t)b * p.d_bs + c0;
            const T *zb = HAS_Z ? reinterpret_cast<const T *>(p.z) + (int64_t)b * p.z_bs + c0 : nullptr;
            const T *Bb = reinterpret_cast<const T *>(p.Bm) + (int64_t)b * p.B_bs;
            const T *Cb = reinterpret_cast<const T *>(p.Cm) + (int64_t)b * p.C_bs;
            T *ob = reinterpret_cast<T *>(p.out) + (int64_t)b * p.o_bs + c0 + 2 * ip;
            T *yb = p.ysave ? reinterpret_cast<T *>(p.ysave) + (int64_t)b * p.L * p.ED + c0 + 2 * ip : nullptr;
            const uint32_t dst_act = smem_u32(smem) + srow * RB + spiece * 16;
            const uint32_t dst_bc = smem_u32(smem) + NTILE * SM::kTile + bcsel * SM::kBCRaw + (it % BCP) * 16;

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
                        stage_tile<T, 0, CPC, NT>(s, ub + (int64_t)tb * p.u_rs, p.u_rs, nrows, it);
                        stage_tile<T, 0, CPC, NT>(s + SM::kTile, db + (int64_t)tb * p.d_rs, p.d_rs, nrows, it);
                        if (HAS_Z) stage_tile<T, 0, CPC, NT>(s + 2 * SM::kTile, zb + (int64_t)tb * p.z_rs, p.z_rs, nrows, it);
                        stage_tile<T, 0, kNState, NT>(s + NTILE * SM::kTile, Bb + (int64_t)tb * p.B_rs, p.B_rs, nrows, it);
                        stage_tile<T, 0, kNState, NT>(s + NTILE * SM::kTile + SM::kBCRaw, Cb + (int64_t)tb * p.C_rs, p.C_rs, nrows, it);
                    }
                }
                cp_async_commit();
            };
#pragma unroll
            for (int k = 0; k < NST; ++k) issue(k);

            const float2 Dc = __ldg(reinterpret_cast<const float2 *>(p.D + c0) + ip);
            const float2 bias = p.dt_bias ? __ldg(reinterpret_cast<const float2 *>(p.dt_bias + c0) + ip) : make_float2(0.f, 0.f);

            auto phase_a = [&](int g) {   // slots of half chunk g: {dl, dl*u} per channel pair, fp32 B|C quads
                const int k = g >> 1, r0 = (g & 1) * H, buf = g & 1;
                const int tb = t0 + k * kChunk + r0;
                const unsigned char *s = smem + (k % NST) * SM::kStage;
                const T *sU = reinterpret_cast<const T *>(s) + r0 * CPC;
                const T *sD = reinterpret_cast<const T *>(s + SM::kTile) + r0 * CPC;
                const T *sBr = reinterpret_cast<const T *>(s + NTILE * SM::kTile) + r0 * kNState;   // B rows; C rows 16 * 16 later
                float x[4], dl[4], sg[4];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float2 d2 = lds_pair(sD + (ir + 4 * i) * CPC, ip);
                    x[2 * i] = d2.x + bias.x;
                    x[2 * i + 1] = d2.y + bias.y;
                }
                if (sp) {
                    softplus_group<4, false>(x, dl, sg);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) dl[i] = x[i];
                }
                float4 *dd_w = sDD + buf * (SM::kDD / 16);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int t = ir + 4 * i;
                    const bool valid = tb + t < t1;
                    const float2 u2 = lds_pair(sU + t * CPC, ip);
                    const float dl0 = valid ? dl[2 * i] : 0.f, dl1 = valid ? dl[2 * i + 1] : 0.f;   // padded step: a = 1, bx = 0
                    const float u0 = valid ? u2.x : 0.f, u1 = valid ? u2.y : 0.f;
                    dd_w[t * (CPC / 2) + ip] = make_float4(dl0, dl0 * u0, dl1, dl1 * u1);
                }
                if (it < H * 8) {   // B|C rows -> fp32 quads: [t][B quads 0..3 | C quads 0..3]
                    const int t = it >> 3, q8 = it & 7;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (tb + t < t1) {
                        const T *src = sBr + (q8 < 4 ? 0 : kChunk * 16) + t * 16 + 4 * (q8 & 3);
                        const float2 lo = lds_pair(src, 0), hi = lds_pair(src, 1);
                        v = make_float4(lo.x, lo.y, hi.x, hi.y);
                    }
                    sBC[buf * (SM::kBC / 16) + it] = v;
                }
            };
            auto epilogue = [&](int g) {   // sum the 4 quads, D skip, gate, stores of half chunk g
                const int k = g >> 1, r0 = (g & 1) * H, buf = g & 1;
                const int tb = t0 + k * kChunk + r0;
                const unsigned char *s = smem + (k % NST) * SM::kStage;
                const T *sU = reinterpret_cast<const T *>(s) + r0 * CPC;
                const T *sZ = reinterpret_cast<const T *>(s + 2 * SM::kTile) + r0 * CPC;
                const float2 *y_r = sY + buf * (SM::kY / 8) + ip;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int t = ir + 4 * i;
                    if (tb + t < t1) {
                        const float2 *yp = y_r + (t * 4) * kV5YPlane;
                        const float2 p0 = yp[0], p1 = yp[kV5YPlane], p2 = yp[2 * kV5YPlane], p3 = yp[3 * kV5YPlane];
                        const float2 u2 = lds_pair(sU + t * CPC, ip);
                        float y0 = fmaf(Dc.x, u2.x, (p0.x + p1.x) + (p2.x + p3.x));
                        float y1 = fmaf(Dc.y, u2.y, (p0.y + p1.y) + (p2.y + p3.y));
                        if (yb != nullptr) stg_pair<T>(yb + (int64_t)(tb + t) * p.ED, y0, y1, true);
                        if (HAS_Z) {
                            const float2 z2 = lds_pair(sZ + t * CPC, ip);
                            y0 *= z2.x * sigmoid_fast(z2.x);
                            y1 *= z2.y * sigmoid_fast(z2.y);
                        }
                        stg_pair<T>(ob + (int64_t)(tb + t) * p.o_rs, y0, y1, vec);
                    }
                }
            };

            cp_async_wait<NST - 1>();
            bar_sync(kBarI, 128);           // chunk 0 is visible to every I thread
            phase_a(0);
            bar_arrive(kBarFullS + 0, 256);
            for (int g = 0; g < G; ++g) {
                if (g + 1 < G) {            // run one half chunk ahead of R
                    if (((g + 1) & 1) == 0) {               // first half of a new chunk: its tiles must have landed
                        cp_async_wait<NST - 2>();
                        bar_sync(kBarI, 128);
                    }
                    if (g + 1 >= 2) bar_sync(kBarEmptyS + ((g + 1) & 1), 256);   // R is done with that slot buffer
                    phase_a(g + 1);
                    bar_arrive(kBarFullS + ((g + 1) & 1), 256);
                }
                bar_sync(kBarFullY + (g & 1), 256);          // R has finished half chunk g
                epilogue(g);
                if (g + 2 < G) bar_arrive(kBarEmptyY + (g & 1), 256);
                if (g & 1) {                                 // chunk g >> 1 is finished: its stage may be refilled
                    bar_sync(kBarI, 128);
                    issue((g >> 1) + NST);
                }
            }
            cp_async_wait<0>();
        }
    }
}

// ---------------------------------------------------------------------------------------------------- host
template <typename T, bool HAS_Z, int CPB>
static bool launch_fwd_v5_inst(const ScanParams &p, const ChainSched &cs, cudaStream_t st) {
    auto kernel = selscan_fwd_v5_kernel<T, HAS_Z, CPB>;
    constexpr size_t smem = FwdV5Smem<T, HAS_Z>::kTotal;
    static thread_local int cache_total = -1, cache_grid = 0;
    if (cache_total != cs.total) {
        // The register hand-over (setmaxnreg 112 / 48) only balances if the kernel was built with 80 registers per thread.
        cudaFuncAttributes fa{};
        if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess || fa.numRegs != kV5RegsLaunch) {
            (void)cudaGetLastError();
            return false;
        }
        int per_sm = 0;
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kV5NT, smem) != cudaSuccess || per_sm < 1) {
            (void)cudaGetLastError();
            return false;
        }
        const int64_t slots = (int64_t)sm_count() * per_sm;
        cache_grid = (int)(cs.total < slots ? cs.total : slots);
        cache_total = cs.total;
    }
    kernel<<<cache_grid, kV5NT, smem, st>>>(p, cs);
    return true;
}

// called by launch_fwd_v2_t when the channel block is 64 wide; false = not launched (caller falls back to v4)
template <typename T>
bool v5_launch_fwd_kernel(const ScanParams &p, const ChainSched &cs, bool has_z, int cpb, cudaStream_t st) {
    if (has_z) return cpb == 16 ? launch_fwd_v5_inst<T, true, 16>(p, cs, st) : launch_fwd_v5_inst<T, true, 0>(p, cs, st);
    return cpb == 16 ? launch_fwd_v5_inst<T, false, 16>(p, cs, st) : launch_fwd_v5_inst<T, false, 0>(p, cs, st);
}
template bool v5_launch_fwd_kernel<float>(const ScanParams &, const ChainSched &, bool, int, cudaStream_t);
template bool v5_launch_fwd_kernel<__nv_bfloat16>(const ScanParams &, const ChainSched &, bool, int, cudaStream_t);
template bool v5_launch_fwd_kernel<__half>(const ScanParams &, const ChainSched &, bool, int, cudaStream_t);

}  // namespace gfe
