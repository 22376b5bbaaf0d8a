// api.cu -- version / error plumbing and the host-side launch planner shared by all ops.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace gfe {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char *what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        (void)cudaGetLastError();   // clear the sticky launch error so the next call starts clean
        return GFE_ERR_CUDA;
    }
    return GFE_OK;
}

int sm_count() {
    static thread_local int cached_dev = -2, cached_sms = 148;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        (void)cudaGetLastError();
        return 148;   // no device (build container): size queries still answer for a B200
    }
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached_sms = n;
        else (void)cudaGetLastError();
        cached_dev = dev;
    }
    return cached_sms;
}

// One warp serves 32 channels of one batch row.  B*ceil(ED/32) warps is all the parallelism the
// recurrence offers without splitting L; splitting costs a second (cheaper) pass over the segment, so
// it is only done when the GPU would otherwise be less than ~half full.
SegPlan plan_segments(int B, int L, int ED) {
    SegPlan p;
    p.nchunks = (L + kChunk - 1) / kChunk;
    const int64_t smsp = (int64_t)sm_count() * 4;
    const int64_t nwarps = (int64_t)B * ((ED + 31) / 32);
    int S = 1;
    if (nwarps * 2 < smsp) {
        S = (int)ceil_div64(2 * smsp, nwarps);
        const int max_by_len = L / (8 * kChunk);      // keep >= 128 steps per segment
        if (S > max_by_len) S = max_by_len;
        if (S > kMaxSeg) S = kMaxSeg;
        if (S < 1) S = 1;
    }
    int seg_len = (int)ceil_div64(ceil_div64(L, S), kChunk) * kChunk;
    if (seg_len < kChunk) seg_len = kChunk;
    p.seg_len = seg_len;
    p.nseg = (L + seg_len - 1) / seg_len;
    if (p.nseg < 1) p.nseg = 1;
    return p;
}

// ---- optional per-kernel timing ------------------------------------------------------------------
static std::atomic<int> g_timing_on{0};
static std::mutex g_timing_mu;
struct TimingRec { int id; cudaEvent_t a, b; };
static std::vector<TimingRec> g_timing;
static const char *const kKernelNames[K_COUNT] = {
    "selscan_fwd_summary", "selscan_fwd", "selscan_bwd_summary", "selscan_bwd", "selscan_bwd_finalize_bc",
    "selscan_bwd_finalize_par", "pscan_fwd_summary", "pscan_fwd", "pscan_bwd_summary", "pscan_bwd",
    "conv1d_silu_fwd", "conv1d_silu_bwd", "conv1d_bwd_finalize", "conv1d_step", "ssm_step",
    "add_rmsnorm_fwd", "add_rmsnorm_bwd", "add_rmsnorm_dw_finalize", "clip_adam_sumsq", "clip_adam_coef", "clip_adam_update", "add_mean_pool_fwd", "mean_pool_bwd"};

void timing_mark(int id, cudaStream_t st, bool begin) {
    if (!g_timing_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    if (begin) {
        TimingRec r{id, nullptr, nullptr};
        if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) { (void)cudaGetLastError(); return; }
        cudaEventRecord(r.a, st);
        g_timing.push_back(r);
    } else {
        for (size_t i = g_timing.size(); i-- > 0;)
            if (g_timing[i].id == id) { cudaEventRecord(g_timing[i].b, st); break; }
    }
}

}  // namespace gfe

extern "C" {

GFE_API int gfe_timing_enable(int on) {
    const int prev = gfe::g_timing_on.exchange(on ? 1 : 0);
    return prev;
}

GFE_API int gfe_timing_kernel_count(void) { return gfe::K_COUNT; }

GFE_API const char *gfe_timing_kernel_name(int id) { return (id >= 0 && id < gfe::K_COUNT) ? gfe::kKernelNames[id] : ""; }

GFE_API int gfe_timing_collect(double *total_ms, int64_t *launches, int n) {
    if (!total_ms || !launches || n < gfe::K_COUNT) { gfe::set_error("timing_collect: need arrays of %d entries", gfe::K_COUNT); return GFE_ERR_ARG; }
    std::lock_guard<std::mutex> lk(gfe::g_timing_mu);
    for (auto &r : gfe::g_timing) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            total_ms[r.id] += ms;
            launches[r.id] += 1;
        } else {
            (void)cudaGetLastError();
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    gfe::g_timing.clear();
    return GFE_OK;
}

GFE_API int gfe_version(void) { return GFE_VERSION; }

GFE_API const char *gfe_last_error_string(void) { return gfe::g_err; }

}  // extern "C"
