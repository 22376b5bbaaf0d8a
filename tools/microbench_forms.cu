// microbench_forms.cu -- cost of the FP32 instruction forms the scan recurrences can be written in (B200, sm_100a):
// packed FFMA2/FMUL2 with three register pairs vs a scalar-broadcast operand vs plain scalar FFMA, and the forward
// recurrence step (x = dl A; a = ex2 x; h = a h + du B; y += C h) written packed vs scalar.
// Prints clocks per warp-instruction per SMSP (at the attribute clock) for 1..6 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float2 sp(float v) { return make_float2(v, v); }
template <int MODE>
__global__ void probe(float *out, const float *in, int iters) {
    float2 h[4], A[4], Bv[4], Cv[4], y[4];
    float dl = in[0], du = in[1];
    for (int i = 0; i < 4; ++i) {
        h[i] = make_float2(in[2] + i, in[3] - i); A[i] = make_float2(in[4] * (i + 1), in[4] * (i + 1.5f));
        Bv[i] = make_float2(in[5] + i, in[5] - i); Cv[i] = make_float2(in[6] * i, in[6] + i); y[i] = make_float2(0.f, 0.f);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {   // 4 unrolled "steps" of 4 state pairs
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (MODE == 0) { h[i].x = fmaf(h[i].x, A[i].x, Bv[i].x); h[i].y = fmaf(h[i].y, A[i].y, Bv[i].y); }        // 2 FFMA rrr
                if (MODE == 1) h[i] = __ffma2_rn(h[i], A[i], Bv[i]);                                                   // FFMA2 rrr
                if (MODE == 2) h[i] = __ffma2_rn(sp(du), A[i], h[i]);                                                  // FFMA2 bcast,r,r
                if (MODE == 3) h[i] = __fmul2_rn(h[i], A[i]);                                                          // FMUL2 rr
                if (MODE == 4) h[i] = __fmul2_rn(sp(dl), h[i]);                                                        // FMUL2 bcast
                if (MODE == 5) h[i] = __fadd2_rn(h[i], A[i]);                                                          // FADD2 rr
                if (MODE == 6) {   // forward step, packed (as selscan_v4_fwd.cu)
                    const float2 x = __fmul2_rn(sp(dl), A[i]);
                    const float2 a = make_float2(ex2a(x.x), ex2a(x.y));
                    h[i] = __ffma2_rn(a, h[i], __fmul2_rn(sp(du), Bv[i]));
                    y[i] = __ffma2_rn(h[i], Cv[i], y[i]);
                }
                if (MODE == 7) {   // forward step, scalar FFMA for the three-register forms
                    const float2 x = __fmul2_rn(sp(dl), A[i]);
                    const float2 a = make_float2(ex2a(x.x), ex2a(x.y));
                    const float2 bx = __fmul2_rn(sp(du), Bv[i]);
                    h[i].x = fmaf(a.x, h[i].x, bx.x); h[i].y = fmaf(a.y, h[i].y, bx.y);
                    y[i].x = fmaf(h[i].x, Cv[i].x, y[i].x); y[i].y = fmaf(h[i].y, Cv[i].y, y[i].y);
                }
                if (MODE == 8) {   // forward step, all scalar
                    const float a0 = ex2a(dl * A[i].x), a1 = ex2a(dl * A[i].y);
                    h[i].x = fmaf(a0, h[i].x, du * Bv[i].x); h[i].y = fmaf(a1, h[i].y, du * Bv[i].y);
                    y[i].x = fmaf(h[i].x, Cv[i].x, y[i].x); y[i].y = fmaf(h[i].y, Cv[i].y, y[i].y);
                }
                if (MODE == 9) {   // forward step packed, without the exponentials (FMA pipe alone)
                    const float2 x = __fmul2_rn(sp(dl), A[i]);
                    h[i] = __ffma2_rn(x, h[i], __fmul2_rn(sp(du), Bv[i]));
                    y[i] = __ffma2_rn(h[i], Cv[i], y[i]);
                }
                if (MODE == 10) {  // the same, scalar three-register forms
                    const float2 x = __fmul2_rn(sp(dl), A[i]);
                    const float2 bx = __fmul2_rn(sp(du), Bv[i]);
                    h[i].x = fmaf(x.x, h[i].x, bx.x); h[i].y = fmaf(x.y, h[i].y, bx.y);
                    y[i].x = fmaf(h[i].x, Cv[i].x, y[i].x); y[i].y = fmaf(h[i].y, Cv[i].y, y[i].y);
                }
            }
            dl += 1e-7f; du -= 1e-7f;   // keep the per-step scalars changing
        }
    }
    float acc = 0.f;
    for (int i = 0; i < 4; ++i) acc += h[i].x + h[i].y + y[i].x + y[i].y;
    if (acc == 12345.678f) out[0] = acc;
}
template <int MODE>
static void run(const char *name, int wps, int sms, const float *in, double per_pair_step) {
    float *out; cudaMalloc(&out, 4);
    const int threads = 128 * wps > 1024 ? 1024 : 128 * wps;
    const int bps = (128 * wps + threads - 1) / threads;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<sms * bps, threads>>>(out, in, 16);
    cudaEventRecord(e0);
    probe<MODE><<<sms * bps, threads>>>(out, in, ITERS);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double pairsteps = (double)ITERS * 16 * wps;   // (state pair, step) items per SMSP
    const double clks = ms * 1e-3 * khz * 1e3;
    printf("%-44s warps/SMSP=%d  %7.3f ms  %6.2f clk per (pair, step) per SMSP  (%4.2f clk per instr)\n", name, wps, ms, clks / pairsteps,
           clks / pairsteps / per_pair_step);
    cudaFree(out);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float hin[8] = {0.01f, 0.02f, 0.5f, 0.25f, -0.3f, 0.7f, 0.9f, 0.f}, *in;
    cudaMalloc(&in, sizeof(hin)); cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    for (int w : {1, 2, 3, 4, 6}) {
        run<0>("2 FFMA r,r,r", w, sms, in, 2);
        run<1>("FFMA2 r,r,r", w, sms, in, 1);
        run<2>("FFMA2 bcast,r,r", w, sms, in, 1);
        run<3>("FMUL2 r,r", w, sms, in, 1);
        run<4>("FMUL2 bcast,r", w, sms, in, 1);
        run<5>("FADD2 r,r", w, sms, in, 1);
        run<6>("fwd step packed (2 FMUL2b 2 MUFU 2 FFMA2)", w, sms, in, 6);
        run<7>("fwd step mixed  (2 FMUL2b 2 MUFU 4 FFMA)", w, sms, in, 8);
        run<8>("fwd step scalar (4 FMUL 2 MUFU 4 FFMA)", w, sms, in, 10);
        run<9>("fwd step packed, no MUFU", w, sms, in, 4);
        run<10>("fwd step mixed, no MUFU", w, sms, in, 6);
    }
    return 0;
}
