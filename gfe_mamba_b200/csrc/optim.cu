// optim.cu -- multi-tensor gradient clipping + Adam in three launches for a whole parameter list.
//
// Replaces the training loop's per-parameter tail (classify_mamba.py:104-109):
//     for param in all_params: torch.nn.utils.clip_grad_norm_(param, max_norm=1.0)     # one norm + one scale PER TENSOR
//     optimizer.step()                                                                 # torch.optim.Adam(lr=1e-4)
//     optimizer.zero_grad()
// which costs ~6 tiny kernels per parameter tensor (norm, add eps, div, clamp, mul, ...) plus the Adam foreach
// groups: at the production shape (B = 2) the step is launch-bound there.  Here the parameter list is described once by a
// chunk table in device memory, and one optimiser step is
//   k1  per chunk:  partial[chunk] = sum g^2                       (fixed-order tree: deterministic)
//   k2  per tensor: coef[t] = min(1, max_norm / (||g_t|| + 1e-6))  (torch.nn.utils.clip_grad_norm_ semantics, per tensor
//                   or over the whole list) and the step counter
//   k3  per chunk:  g *= coef; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps);
//                   optionally g = 0 (the zero_grad of the loop)
// The step counter lives in device memory so that a captured CUDA graph replays the correct bias corrections.
#include "common.cuh"

namespace gfe {

constexpr int kOptChunk = 4096;   // elements per chunk (one CTA of 256 threads, 4 float4 per thread)
constexpr int kOptNT = 256;

struct OptArgs {
    const int64_t *p_ptr, *g_ptr, *m_ptr, *v_ptr;   // [ntensors] device addresses (fp32 tensors)
    const int64_t *numel;                            // [ntensors]
    const int32_t *chunk_tensor;                     // [nchunks]
    const int64_t *chunk_start;                      // [nchunks] element offset inside the tensor
    const int32_t *tensor_chunk0;                    // [ntensors + 1] first chunk of every tensor
    float *partial;                                  // [nchunks]
    float *coef;                                     // [ntensors]
    int32_t *step;                                   // [1]
    int ntensors, nchunks;
    float lr, beta1, beta2, eps, max_norm;
    int global_norm, zero_grad;
};

__device__ __forceinline__ float block_sum(float v, float *sred) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sred[w] = v;
    __syncthreads();
    float r = 0.f;
    if (w == 0) {
        r = l < kOptNT / 32 ? sred[l] : 0.f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;   // valid in warp 0
}

__global__ void __launch_bounds__(kOptNT) opt_sumsq_kernel(OptArgs a) {
    __shared__ float sred[kOptNT / 32];
    const int c = blockIdx.x;
    const int t = a.chunk_tensor[c];
    const int64_t start = a.chunk_start[c];
    const int64_t n = min((int64_t)kOptChunk, a.numel[t] - start);
    const float *g = reinterpret_cast<const float *>(a.g_ptr[t]) + start;
    float acc = 0.f;
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        const int n4 = (int)(n >> 2);
        for (int i = threadIdx.x; i < n4; i += kOptNT) {
            const float4 v = reinterpret_cast<const float4 *>(g)[i];
            acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        for (int i = (n4 << 2) + threadIdx.x; i < n; i += kOptNT) acc += g[i] * g[i];
    } else {
        for (int i = threadIdx.x; i < n; i += kOptNT) acc += g[i] * g[i];
    }
    const float s = block_sum(acc, sred);
    if (threadIdx.x == 0) a.partial[c] = s;
}

__global__ void __launch_bounds__(kOptNT) opt_coef_kernel(OptArgs a) {
    __shared__ float sred[kOptNT / 32];
    if (a.global_norm) {   // one norm over the whole list: a single CTA adds every partial
        float acc = 0.f;
        for (int i = threadIdx.x; i < a.nchunks; i += kOptNT) acc += a.partial[i];
        const float s = block_sum(acc, sred);
        __shared__ float cf;
        if (threadIdx.x == 0) cf = fminf(1.0f, a.max_norm / (sqrtf(s) + 1e-6f));
        __syncthreads();
        for (int t = threadIdx.x; t < a.ntensors; t += kOptNT) a.coef[t] = a.max_norm > 0.f ? cf : 1.0f;
        if (threadIdx.x == 0) a.step[0] += 1;
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * (kOptNT / 32) + warp;   // one warp per tensor
    if (t < a.ntensors) {
        float acc = 0.f;
        for (int i = a.tensor_chunk0[t] + lane; i < a.tensor_chunk0[t + 1]; i += 32) acc += a.partial[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) a.coef[t] = a.max_norm > 0.f ? fminf(1.0f, a.max_norm / (sqrtf(acc) + 1e-6f)) : 1.0f;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) a.step[0] += 1;
}

__global__ void __launch_bounds__(kOptNT) opt_adam_kernel(OptArgs a) {
    const int c = blockIdx.x;
    const int t = a.chunk_tensor[c];
    const int64_t start = a.chunk_start[c];
    const int n = (int)min((int64_t)kOptChunk, a.numel[t] - start);
    float *p = reinterpret_cast<float *>(a.p_ptr[t]) + start;
    float *g = reinterpret_cast<float *>(a.g_ptr[t]) + start;
    float *m = reinterpret_cast<float *>(a.m_ptr[t]) + start;
    float *v = reinterpret_cast<float *>(a.v_ptr[t]) + start;
    const float cf = a.coef[t];
    const float stepf = (float)a.step[0];
    const float bc1 = 1.0f - powf(a.beta1, stepf), bc2 = 1.0f - powf(a.beta2, stepf);
    const float step_size = a.lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    const float b1 = a.beta1, b2 = a.beta2, eps = a.eps;
    auto upd = [&](float &pp, float &gg, float &mm, float &vv) {
        const float gc = gg * cf;
        mm = fmaf(b1, mm, (1.0f - b1) * gc);
        vv = fmaf(b2, vv, (1.0f - b2) * gc * gc);
        pp -= step_size * mm / (sqrtf(vv) * inv_sqrt_bc2 + eps);
        gg = a.zero_grad ? 0.f : gc;
    };
    const bool al = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                      reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (al) {
        const int n4 = n >> 2;
        for (int i = threadIdx.x; i < n4; i += kOptNT) {
            float4 P = reinterpret_cast<float4 *>(p)[i], G = reinterpret_cast<float4 *>(g)[i];
            float4 M = reinterpret_cast<float4 *>(m)[i], V = reinterpret_cast<float4 *>(v)[i];
            upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
            reinterpret_cast<float4 *>(p)[i] = P; reinterpret_cast<float4 *>(g)[i] = G;
            reinterpret_cast<float4 *>(m)[i] = M; reinterpret_cast<float4 *>(v)[i] = V;
        }
        for (int i = (n4 << 2) + threadIdx.x; i < n; i += kOptNT) upd(p[i], g[i], m[i], v[i]);
    } else {
        for (int i = threadIdx.x; i < n; i += kOptNT) upd(p[i], g[i], m[i], v[i]);
    }
}

}  // namespace gfe

extern "C" {

GFE_API int gfe_clip_adam_chunk_elems(void) { return gfe::kOptChunk; }

GFE_API int gfe_clip_adam_step(const int64_t *p_ptr, const int64_t *g_ptr, const int64_t *m_ptr, const int64_t *v_ptr,
                               const int64_t *numel, const int32_t *chunk_tensor, const int64_t *chunk_start,
                               const int32_t *tensor_chunk0, float *partial, float *coef, int32_t *step, int ntensors,
                               int nchunks, float lr, float beta1, float beta2, float eps, float max_norm, int global_norm,
                               int zero_grad, void *stream) {
    using namespace gfe;
    if (ntensors <= 0 || nchunks <= 0) return GFE_OK;
    if (!p_ptr || !g_ptr || !m_ptr || !v_ptr || !numel || !chunk_tensor || !chunk_start || !tensor_chunk0 || !partial || !coef || !step) {
        set_error("clip_adam_step: NULL table pointer");
        return GFE_ERR_ARG;
    }
    if (!(lr >= 0.f) || !(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f) || !(eps >= 0.f)) {
        set_error("clip_adam_step: invalid hyper-parameters");
        return GFE_ERR_ARG;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    OptArgs a{p_ptr, g_ptr, m_ptr, v_ptr, numel, chunk_tensor, chunk_start, tensor_chunk0, partial, coef, step,
              ntensors, nchunks, lr, beta1, beta2, eps, max_norm, global_norm, zero_grad};
    if (max_norm > 0.f) {
        ScopedKernelTimer tm(K_OPT_SUMSQ, st);
        opt_sumsq_kernel<<<nchunks, kOptNT, 0, st>>>(a);
    }
    {
        ScopedKernelTimer tm(K_OPT_COEF, st);
        const int grid = global_norm ? 1 : (ntensors + kOptNT / 32 - 1) / (kOptNT / 32);
        opt_coef_kernel<<<grid, kOptNT, 0, st>>>(a);
    }
    {
        ScopedKernelTimer tm(K_OPT_ADAM, st);
        opt_adam_kernel<<<nchunks, kOptNT, 0, st>>>(a);
    }
    return check_launch("clip_adam_step");
}

}  // extern "C"
