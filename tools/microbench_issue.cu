// microbench_issue.cu -- does a packed FP32 op (FFMA2 / FMUL2) hold the issue port for its two pipe cycles?
// Mixes of FFMA2 with ALU (IADD3/LOP3) and MUFU work at 1..8 warps per scheduler; prints warp-instructions per clock per
// SMSP for each mix.  If FFMA2 kept the port busy for 2 cycles, FFMA2 + IADD 1:1 could not exceed 0.67 instr/clk/SMSP.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_issue microbench_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define CH 8
__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void probe(float *out, float s0, float s1, int iters, int seed) {
    float2 w[CH], a[CH], b[CH]; float v[CH]; int q[CH];
    for (int i = 0; i < CH; ++i) {
        w[i] = make_float2(threadIdx.x * 1e-3f + i, 0.5f + i); a[i] = make_float2(s0 + i * 1e-6f, s0); b[i] = make_float2(s1, s1 + i * 1e-6f);
        v[i] = 0.25f * i - threadIdx.x * 1e-3f; q[i] = seed + i + threadIdx.x;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0 || MODE == 2 || MODE == 3 || MODE == 5 || MODE == 6) w[i] = __ffma2_rn(w[i], a[i], b[i]);   // 3 distinct register pairs
            if (MODE == 1 || MODE == 2 || MODE == 6) q[i] = (q[i] + seed) ^ i;                                       // ALU
            if (MODE == 3) { q[i] = (q[i] + seed) ^ i; q[i] = (q[i] * 3 + it) | 1; }                                  // 2 more int ops (one may be IMAD: fma pipe)
            if (MODE == 4) { w[i] = __fmul2_rn(w[i], make_float2(s0, s0)); q[i] = (q[i] + seed) ^ i; }                // FMUL2 bcast + ALU
            if ((MODE == 5 || MODE == 6) && (i & 1)) v[i] = ex2a(v[i]);                                              // 2 FFMA2 : 1 MUFU
            if (MODE == 7) { v[i] = fmaf(v[i], a[i].x, b[i].x); w[i].x = fmaf(w[i].x, a[i].y, b[i].y); if (i & 1) w[i].y = ex2a(w[i].y); }   // 4 FFMA : 1 MUFU
            if (MODE == 8) { w[i] = __ffma2_rn(w[i], a[i], b[i]); a[i] = __fmul2_rn(a[i], make_float2(s0, s0)); }      // FFMA2 + FMUL2
            if (MODE == 9) { v[i] = ex2a(v[i]); q[i] = (q[i] + seed) ^ i; }                                           // MUFU + ALU 1:1
        }
    }
    float acc = 0.f;
    for (int i = 0; i < CH; ++i) acc += v[i] + w[i].x + w[i].y + a[i].x + (float)q[i];
    if (acc == 12345.678f) out[0] = acc;
}
template <int MODE>
static void run(const char *name, int wps, int sms, double ipi) {
    float *out; cudaMalloc(&out, 4);
    const int threads = 128 * wps > 1024 ? 1024 : 128 * wps;   // wps warps per scheduler -> 4 * wps warps per SM
    const int bps = (128 * wps + threads - 1) / threads;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<sms * bps, threads>>>(out, 0.999f, 0.001f, 16, 3);
    cudaEventRecord(e0);
    probe<MODE><<<sms * bps, threads>>>(out, 0.999f, 0.001f, ITERS, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double instr = (double)ITERS * CH * ipi * wps;   // per SMSP
    printf("%-34s warps/SMSP=%d  %7.3f ms  %.3f warp-instr/clk/SMSP\n", name, wps, ms, instr / (ms * 1e-3 * khz * 1e3));
    cudaFree(out);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int w : {1, 2, 3, 4, 8}) {
        run<0>("FFMA2 (3 reg pairs)", w, sms, 1);
        run<1>("ALU (IADD+LOP)", w, sms, 2);
        run<2>("FFMA2 + 2 ALU", w, sms, 3);
        run<3>("FFMA2 + 4 int", w, sms, 5);
        run<4>("FMUL2 bcast + 2 ALU", w, sms, 3);
        run<5>("2 FFMA2 : 1 MUFU", w, sms, 1.5);
        run<6>("2 FFMA2 : 1 MUFU : 4 ALU", w, sms, 3.5);
        run<7>("4 FFMA : 1 MUFU", w, sms, 2.5);
        run<8>("FFMA2 + FMUL2", w, sms, 2);
        run<9>("MUFU + 2 ALU", w, sms, 3);
    }
    return 0;
}
