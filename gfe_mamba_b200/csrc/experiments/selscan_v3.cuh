// selscan_v3.cuh -- plumbing shared by the warp-autonomous ("v3") selective-scan kernels.
//
// Decomposition (why: profiles/r01_v2_ncu_selscan_cfg3_summary.txt -- the CTA-cooperative v2 kernels executed 190 (fwd)
// and 415 (bwd) warp-instructions per (32 channels x 1 step) against ~120 / ~230 of essential work, and stalled on
// block barriers and shared-memory round trips of per-channel scalars):
//   * one warp = 32 adjacent channels of one batch row, lane = channel, all 16 states of the channel in the lane's
//     registers as 8 float2 pairs: the per-(t, channel) scalar work (softplus, SiLU, D skip, conversions, stores) is
//     done exactly once by the lane that needs it, the sum over states never leaves the lane, and no block-wide
//     barrier exists anywhere (warps only ever __syncwarp);
//   * every warp stages its own 16-step (forward) / 8-step (backward) tiles HBM -> shared memory with cp.async in a
//     ring of stages that runs ACROSS work units, so the pipeline never drains;
//   * time is cut into segments; warps draw (segment, batch row, channel block) units from an atomic counter in
//     dependency order and hand the state to the successor through global memory (ChainSched), which balances the
//     B * ED / 32 chains over the 4 * SMs schedulers whatever their ratio is.
#pragma once

#include "common.cuh"
#include "selscan_shared.cuh"

namespace gfe {

constexpr int kV3MaxWarps = 8;      // warps per CTA upper bound (launch uses 4 or 8)
constexpr int kV3IdRing = 8;        // per-warp FIFO of drawn unit ids (prefetch runs at most kStages - 1 chunks ahead)

struct V3Unit {
    int id, seg, b, c0, t0, t1, nch;   // nch = chunks of this unit; id < 0: no more work
};

#ifdef __CUDACC__
__device__ __forceinline__ int v3_draw(int *counter, int lane) {
    int id = 0;
    if (lane == 0) id = atomicAdd(counter, 1);
    return __shfl_sync(0xffffffffu, id, 0);
}

// Wait for the predecessor unit.  Units are drawn in dependency order, so the wait is bounded by one unit's run time
// (tens of microseconds); the trap turns a scheduling bug into a launch error instead of a hung GPU.
__device__ __forceinline__ void v3_wait_flag(const int *f) {
    unsigned spins = 0;
    while (ld_acquire(f) == 0) {
        __nanosleep(64);
        if (++spins > (1u << 25)) __trap();
    }
}

// forward order: segment-major ascending time
__device__ __forceinline__ V3Unit v3_decode(int id, const ChainSched &cs, int B, int L, int chunk, bool reverse) {
    V3Unit u;
    if (id >= cs.total) {
        u.id = -1; u.seg = u.b = u.c0 = u.t0 = u.t1 = u.nch = 0;
        return u;
    }
    const int per_seg = B * cs.nblk;
    const int rseg = id / per_seg;
    const int rem = id - rseg * per_seg;
    u.id = id;
    u.seg = reverse ? cs.nseg - 1 - rseg : rseg;
    u.b = rem / cs.nblk;
    u.c0 = (rem - u.b * cs.nblk) * 32;
    u.t0 = u.seg * cs.seg_len;
    u.t1 = min(L, u.t0 + cs.seg_len);
    u.nch = (u.t1 - u.t0 + chunk - 1) / chunk;
    return u;
}

// one warp copies `nrows` rows of 32 elements (row stride rs elements) into a dense [rows][32] tile, 16-byte pieces
template <typename T, int ROWS>
__device__ __forceinline__ void v3_stage32(uint32_t dst, const T *src, int64_t rs, int nrows, int lane) {
    constexpr int PPR = 32 * (int)sizeof(T) / 16;   // pieces per row: 4 (16-bit) or 8 (fp32)
    constexpr int RPI = 32 / PPR;                   // rows per warp instruction
    const int piece = lane % PPR, r0 = lane / PPR;
    const char *s = reinterpret_cast<const char *>(src) + piece * 16;
#pragma unroll
    for (int r = 0; r < ROWS; r += RPI) {
        const int row = r + r0;
        if (row < nrows) cp_async<16>(dst + (row * PPR + piece) * 16, s + (int64_t)row * rs * (int64_t)sizeof(T));
    }
}
// rows of 16 elements (B or C)
template <typename T, int ROWS>
__device__ __forceinline__ void v3_stage16(uint32_t dst, const T *src, int64_t rs, int nrows, int lane) {
    constexpr int PPR = 16 * (int)sizeof(T) / 16;   // 2 or 4
    constexpr int RPI = 32 / PPR;
    const int piece = lane % PPR, r0 = lane / PPR;
    const char *s = reinterpret_cast<const char *>(src) + piece * 16;
#pragma unroll
    for (int r = 0; r < ROWS; r += RPI) {
        const int row = r + r0;
        if ((ROWS % RPI == 0 || row < ROWS) && row < nrows)
            cp_async<16>(dst + (row * PPR + piece) * 16, s + (int64_t)row * rs * (int64_t)sizeof(T));
    }
}

// checkpoint of one chunk: 8 pair rows of 32 float2 (row stride ED float2) -> dense [pair][lane] float2
__device__ __forceinline__ void v3_stage_ck(uint32_t dst, const float2 *src, int64_t ED, int lane) {
    const int piece = lane & 15, r0 = lane >> 4;
    const char *s = reinterpret_cast<const char *>(src) + piece * 16;
#pragma unroll
    for (int r = 0; r < kPairs; r += 2) {
        const int row = r + r0;
        cp_async<16>(dst + (row * 16 + piece) * 16, s + (int64_t)row * ED * 8);
    }
}

// one warp writes a dense [rows][32] shared tile to global rows (row stride rs elements): 16-byte pieces when aligned
template <typename T, int ROWS>
__device__ __forceinline__ void v3_store32(const T *tile, T *dst, int64_t rs, int nrows, int lane, int aligned) {
    if (aligned) {
        constexpr int PPR = 32 * (int)sizeof(T) / 16;
        constexpr int RPI = 32 / PPR;
        const int piece = lane % PPR, r0 = lane / PPR;
        char *d = reinterpret_cast<char *>(dst) + piece * 16;
#pragma unroll
        for (int r = 0; r < ROWS; r += RPI) {
            const int row = r + r0;
            if (row < nrows)
                __stcs(reinterpret_cast<float4 *>(d + (int64_t)row * rs * (int64_t)sizeof(T)),
                       reinterpret_cast<const float4 *>(tile)[row * PPR + piece]);
        }
    } else {
        for (int r = 0; r < nrows; ++r) st_stream(dst + (int64_t)r * rs + lane, tile[r * 32 + lane]);
    }
}

// 16 consecutive 16-bit elements of shared memory -> fp32
template <typename T>
__device__ __forceinline__ void v3_cvt16(const T *src, float (&v)[16]) {
    const uint4 a = reinterpret_cast<const uint4 *>(src)[0], b = reinterpret_cast<const uint4 *>(src)[1];
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if constexpr (sizeof(T) == 2) {
            typename Pair<T>::type pr;
            memcpy(&pr, &w[i], 4);
            const float2 f = pair_to_f(pr);
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
}

// softplus for a batch of G pre-activations held by one lane: branch-free main path (log1p through 2 atanh), one
// warp vote per batch for the rare large-argument formulation.  sig[i] = sigmoid(x[i]) when WANT_SIG.
template <int G, bool WANT_SIG>
__device__ __forceinline__ void v3_softplus(const float (&x)[G], float (&dl)[G], float (&sig)[G]) {
    softplus_group<G, WANT_SIG>(x, dl, sig);
}
#endif  // __CUDACC__

// host (selscan_v3_fwd.cu / selscan_v3_bwd.cu)
bool v3_applicable(const gfe_selscan_args *a, bool bwd);
bool v3_shape_ok(int B, int L, int ED);
void v3_plan(int B, int L, int ED, int &nblk, int &nseg, int &seg_len);
size_t v3_chain_bytes(int B, int ED, int nblk, int nseg, int carry_floats);
size_t v3_fwd_workspace_bytes(int B, int L, int ED);
size_t v3_bwd_workspace_bytes(int B, int L, int ED);
int v3_launch_fwd(const gfe_selscan_args *a, cudaStream_t st);
int v3_launch_bwd(const gfe_selscan_args *a, cudaStream_t st);

}  // namespace gfe
