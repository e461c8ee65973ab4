// Fused flat-buffer Adam: ONE launch updates every parameter of the model (reference: torch.optim.Adam as used at
// kernel/train_eval_sgcn_img_snps.py:108,547 -- lr 1e-3, betas (0.9,0.999), eps 1e-8, weight_decay 0, no amsgrad).
// Parameters, gradients and both moments are contiguous fp32 buffers (the gradient buffer is the one the data-parallel
// all-reduce runs on).  `step` and `lr` live in device memory so the launch can sit inside a captured CUDA graph.
#include "common.cuh"

namespace igcn {

__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, const float* __restrict__ step,
                                                        const float* __restrict__ lr, float beta1, float beta2, float eps,
                                                        float grad_scale, int64_t n) {
    const float t = step[0];                       // already incremented for this update (1, 2, ...)
    const float bias1 = 1.f - powf(beta1, t);
    const float bias2 = 1.f - powf(beta2, t);
    const float step_size = lr[0] / bias1;
    const float inv_sqrt_bias2 = rsqrtf(bias2);
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pv = reinterpret_cast<float4*>(p)[i];
        const float4 gv = reinterpret_cast<const float4*>(g)[i];
        float4 mv = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
#define IGCN_ADAM1(C)                                                   \
    {                                                                   \
        const float gg = gv.C * grad_scale;                             \
        mv.C = beta1 * mv.C + (1.f - beta1) * gg;                       \
        vv.C = beta2 * vv.C + (1.f - beta2) * gg * gg;                  \
        pv.C -= step_size * mv.C / (sqrtf(vv.C) * inv_sqrt_bias2 + eps); \
    }
        IGCN_ADAM1(x) IGCN_ADAM1(y) IGCN_ADAM1(z) IGCN_ADAM1(w)
        reinterpret_cast<float4*>(p)[i] = pv;
        reinterpret_cast<float4*>(m)[i] = mv;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gg = g[i] * grad_scale;
        const float mm = beta1 * m[i] + (1.f - beta1) * gg;
        const float vv = beta2 * v[i] + (1.f - beta2) * gg * gg;
        m[i] = mm;
        v[i] = vv;
        p[i] -= step_size * mm / (sqrtf(vv) * inv_sqrt_bias2 + eps);
    }
#undef IGCN_ADAM1
}

}  // namespace igcn

extern "C" int igcn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const float* step,
                              const float* lr, double beta1, double beta2, double eps, double grad_scale, int64_t n, void* stream) {
    using namespace igcn;
    IGCN_REQUIRE(n >= 0, IGCN_ERR_BAD_ARG, "adam_step: negative size");
    if (n == 0) return IGCN_OK;
    IGCN_REQUIRE(params && grads && exp_avg && exp_avg_sq && step && lr, IGCN_ERR_BAD_ARG, "adam_step: null pointer");
    IGCN_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0, IGCN_ERR_BAD_ARG,
                 "adam_step: buffers must be 16-byte aligned");
    int64_t blocks = (n / 4 + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adam_flat_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, step, lr, (float)beta1,
                                                                    (float)beta2, (float)eps, (float)grad_scale, n);
    IGCN_CHECK_LAUNCH("adam_step");
    return IGCN_OK;
}
