// microbench_lds.cu -- shared-memory wavefront cost of LDS.32/64/128 address patterns (B200).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_lds microbench_lds.cu ; run on the GPU box.
// Reports SM clocks per LDS warp-instruction with 32 warps/SM (the smem data pipe moves 1 wavefront / clk / SM).
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int W>   // bytes per lane: 4, 8, 16
__global__ void probe(float *out, const int *tbl, int iters) {
    __shared__ __align__(16) float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 1e-3f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int off = tbl[lane];                 // in units of W bytes
    float acc = 0.f;
    int rot = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = (off + rot + u * 64) & (8192 * 4 / W - 1);
            if (W == 16) { float4 v = reinterpret_cast<const float4 *>(sm)[idx]; acc += v.x + v.w; }
            if (W == 8) { float2 v = reinterpret_cast<const float2 *>(sm)[idx]; acc += v.x + v.y; }
            if (W == 4) { acc += sm[idx]; }
        }
        rot = (rot + 8) & 255;   // multiples of 8 units keep the bank pattern of the table
    }
    if (acc == 12345.678f) out[0] = acc;
}

template <int W>
static void run(const char *name, const int *h_tbl, int sms) {
    int *tbl; float *out;
    cudaMalloc(&tbl, 128); cudaMalloc(&out, 4);
    cudaMemcpy(tbl, h_tbl, 128, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<W><<<sms, 1024>>>(out, tbl, 16);
    cudaEventRecord(e0);
    probe<W><<<sms, 1024>>>(out, tbl, ITERS);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double instr = (double)ITERS * 8 * 32;   // per SM
    printf("LDS.%-3d %-44s %.3f ms  %.2f clk/instr/SM\n", W * 8, name, ms, ms * 1e-3 * clk_khz * 1e3 / instr);
    cudaFree(tbl); cudaFree(out);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int t[32];
    auto fill = [&](auto f) { for (int l = 0; l < 32; ++l) t[l] = f(l); };
    fill([](int l) { return 0; });            run<16>("uniform", t, sms);
    fill([](int l) { return l / 16; });       run<16>("2 distinct: lane/16", t, sms);
    fill([](int l) { return l & 1; });        run<16>("2 distinct: lane%2", t, sms);
    fill([](int l) { return l / 8; });        run<16>("4 distinct: lane/8 (quarter-uniform)", t, sms);
    fill([](int l) { return l & 3; });        run<16>("4 distinct: lane%4 (Q-layout B4)", t, sms);
    fill([](int l) { return l / 4; });        run<16>("8 distinct: lane/4 (Q-layout dd)", t, sms);
    fill([](int l) { return l & 7; });        run<16>("8 distinct: lane%8", t, sms);
    fill([](int l) { return l; });            run<16>("32 distinct contiguous", t, sms);
    fill([](int l) { return (l & 3) + 8 * (l >> 2 & 1); }); run<16>("8 distinct: lane%4 + 128B*(lane/4%2) (bank conflict)", t, sms);
    fill([](int l) { return (l & 3) + 12 * (l >> 2 & 1); }); run<16>("8 distinct: lane%4 + 192B*(lane/4%2) (skewed)", t, sms);
    fill([](int l) { return 0; });            run<8>("uniform", t, sms);
    fill([](int l) { return l & 1; });        run<8>("2 distinct: lane%2", t, sms);
    fill([](int l) { return l / 4; });        run<8>("8 distinct: lane/4", t, sms);
    fill([](int l) { return l & 3; });        run<8>("4 distinct: lane%4", t, sms);
    fill([](int l) { return l; });            run<8>("32 distinct contiguous", t, sms);
    fill([](int l) { return 0; });            run<4>("uniform", t, sms);
    fill([](int l) { return l / 4; });        run<4>("8 distinct: lane/4", t, sms);
    fill([](int l) { return l; });            run<4>("32 distinct contiguous", t, sms);
    return 0;
}
