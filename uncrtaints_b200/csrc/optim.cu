// Fused Adam step over the flat parameter / gradient buffers.
//
// The reference steps `torch.optim.Adam(netG.parameters(), lr)` (model/src/backbones/base_model.py:48-51,122) -- ~91 small
// tensors, i.e. a multi-tensor launch chain plus a zero_grad chain per step.  Here the 570 010 parameters, their gradients
// and both moments live in four flat fp32 buffers, and ONE kernel does: optional gradient scaling (1/world of the data-parallel
// mean, or loss scaling), the moment updates, the bias-corrected parameter update (torch's default formulation: no amsgrad,
// L2 weight decay added to the gradient) and, optionally, the zeroing of the gradient buffer for the next backward (the C
// ABI accumulates into gradient slots).  ExponentialLR (base_model.py:51) only changes the `lr` argument.
#include "common.cuh"
#include "kernels.h"

namespace ub {

__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, size_t n, float lr, float beta1, float beta2,
                                                         float eps, float weight_decay, float inv_bc1, float inv_sqrt_bc2,
                                                         float grad_scale, int zero_grad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gi = g[i] * grad_scale;
    const float pi = p[i];
    if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
    const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);             // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;                 // (exp_avg_sq.sqrt() / sqrt(bias_correction2)).add_(eps)
    p[i] = pi - (lr * inv_bc1) * (mi / denom);                          // param.addcdiv_(exp_avg, denom, value=-lr / bias_correction1)
    if (zero_grad) g[i] = 0.f;
}

int launch_adam_step(float* p, float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                     float weight_decay, float inv_bc1, float inv_sqrt_bc2, float grad_scale, int zero_grad, cudaStream_t st) {
    if (n == 0) return UB_OK;
    adam_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, inv_bc1,
                                                                 inv_sqrt_bc2, grad_scale, zero_grad);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

}  // namespace ub

// ------------------------------------------------------------------------------------------------------------------
// Device-side batch assembly: prepare_data_multi (model/train_reconstruct.py:161-179) builds the network input
// with torch.stack over the T per-time-point S1 tensors, torch.stack over the S2 tensors and torch.cat of the two stacks --
// three read+write passes over the batch.  One kernel writes x[b][t][c] straight from the 2T source tensors (pointer table in
// device memory): one read, one write.  src[t] -> S1[t] ([B][c1][P]) or null, src[T + t] -> S2[t] ([B][c2][P]).
// ------------------------------------------------------------------------------------------------------------------
namespace ub {

__global__ void __launch_bounds__(256) assemble_input_kernel(const float* const* __restrict__ src, float* __restrict__ x, int T, int c1,
                                                              int c2, int P4 /* H*W/4 */) {
    const int C = c1 + c2;
    const int plane = blockIdx.y;                       // (b*T + t)*C + c
    const int c = plane % C, t = (plane / C) % T, b = plane / (C * T);
    const float4* s = c < c1 ? reinterpret_cast<const float4*>(src[t]) + ((size_t)b * c1 + c) * P4
                             : reinterpret_cast<const float4*>(src[T + t]) + ((size_t)b * c2 + (c - c1)) * P4;
    float4* d = reinterpret_cast<float4*>(x) + (size_t)plane * P4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P4; i += gridDim.x * blockDim.x) d[i] = s[i];
}

int launch_assemble_input(const float* const* src, float* x, int B, int T, int c1, int c2, int P, cudaStream_t st) {
    if (P % 4 != 0 || B < 1 || T < 1 || c2 < 1 || c1 < 0) return UB_ERR_ARG;
    const int P4 = P / 4;
    const int bx = (P4 + 255) / 256 < 8 ? (P4 + 255) / 256 : 8;
    assemble_input_kernel<<<dim3(bx, B * T * (c1 + c2)), 256, 0, st>>>(src, x, T, c1, c2, P4);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

}  // namespace ub
