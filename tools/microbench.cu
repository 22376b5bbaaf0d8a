// microbench.cu -- issue-rate probes for the pipes the selective scan leans on (B200, sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu ; run on the GPU box.
// Prints warp-instructions per clock per SM for: FFMA (3 distinct regs), FFMA2, FMUL2, MUFU.EX2, SHFL.BFLY,
// SEL, broadcast LDS.128, and mixed FFMA2+MUFU.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define CHAINS 8

__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void probe(float *out, float s0, float s1, int iters) {
    __shared__ float4 sm[64];
    float v[CHAINS]; float2 w[CHAINS];
    for (int i = 0; i < CHAINS; ++i) { v[i] = threadIdx.x * 1e-3f + i; w[i] = make_float2(v[i], v[i] + 0.5f); }
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(s0, s1, s0, s1);
    __syncthreads();
    const float2 a2 = make_float2(s0, s0), b2 = make_float2(s1, s1);
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0) v[i] = fmaf(v[i], s0, s1);                       // FFMA (2 uniform-ish operands)
            if (MODE == 1) w[i] = __ffma2_rn(w[i], a2, b2);                 // FFMA2
            if (MODE == 2) v[i] = ex2a(v[i]);                               // MUFU.EX2
            if (MODE == 3) v[i] = __shfl_xor_sync(0xffffffffu, v[i], 16);   // SHFL
            if (MODE == 4) v[i] = (lane & 16) ? v[i] : v[(i + 1) % CHAINS]; // SEL
            if (MODE == 5) { float4 t = sm[(it + i) & 63]; v[i] += t.x; }   // broadcast LDS.128 + FADD
            if (MODE == 6) { w[i] = __ffma2_rn(w[i], a2, b2); if ((i & 3) == 0) v[i] = ex2a(v[i]); }   // 4 FFMA2 : 1 MUFU
            if (MODE == 7) v[i] = fmaf(v[i], v[(i + 1) % CHAINS], v[(i + 2) % CHAINS]);               // FFMA 3 distinct regs
            if (MODE == 8) w[i] = __ffma2_rn(w[i], w[(i + 1) % CHAINS], w[(i + 2) % CHAINS]);         // FFMA2 3 distinct
            if (MODE == 9) w[i] = __fmul2_rn(w[i], a2);                     // FMUL2
        }
    }
    float acc = 0.f;
    for (int i = 0; i < CHAINS; ++i) acc += v[i] + w[i].x + w[i].y;
    if (acc == 12345.678f) out[0] = acc;
}

template <int MODE>
static void run(const char *name, int warps_per_sm, int sms, double instr_per_iter_chain, float s0, float s1) {
    float *out; cudaMalloc(&out, 4);
    const int threads = 32 * warps_per_sm > 1024 ? 1024 : 32 * warps_per_sm;
    const int blocks_per_sm = (32 * warps_per_sm + threads - 1) / threads;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<sms * blocks_per_sm, threads>>>(out, s0, s1, 16);
    cudaEventRecord(e0);
    probe<MODE><<<sms * blocks_per_sm, threads>>>(out, s0, s1, ITERS);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double warp_instr = (double)ITERS * CHAINS * instr_per_iter_chain * warps_per_sm;   // per SM
    const double clks = ms * 1e-3 * clk_khz * 1e3;
    printf("%-28s warps/SM=%2d  %.3f ms  %.3f warp-instr/clk/SM (at max clock %d MHz)\n", name, warps_per_sm, ms,
           warp_instr / clks, clk_khz / 1000);
    cudaFree(out);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs: %d\n", sms);
    for (int w : {4, 8, 16, 32}) {
        run<0>("FFMA r,u,u", w, sms, 1, 0.999f, 0.001f);
        run<7>("FFMA 3 regs", w, sms, 1, 0.999f, 0.001f);
        run<1>("FFMA2 r,u,u", w, sms, 1, 0.999f, 0.001f);
        run<8>("FFMA2 3 regs", w, sms, 1, 0.999f, 0.001f);
        run<9>("FMUL2", w, sms, 1, 0.999f, 0.001f);
        run<2>("MUFU.EX2", w, sms, 1, 0.999f, 0.001f);
        run<3>("SHFL.BFLY", w, sms, 1, 0.999f, 0.001f);
        run<4>("SEL", w, sms, 1, 0.999f, 0.001f);
        run<5>("LDS.128 bcast + FADD", w, sms, 2, 0.999f, 0.001f);
        run<6>("4 FFMA2 : 1 MUFU", w, sms, 1.25, 0.999f, 0.001f);
    }
    return 0;
}
