// api.cu -- version / error plumbing and the host-side launch planner shared by all ops.
#include <stdarg.h>
#include <stdio.h>

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
    const int64_t nwarps = (int64_t)B * ((ED + 31) / 32);
    const int64_t smsp = (int64_t)sm_count() * 4;
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

}  // namespace gfe

extern "C" {

GFE_API int gfe_version(void) { return GFE_VERSION; }

GFE_API const char *gfe_last_error_string(void) { return gfe::g_err; }

}  // extern "C"
