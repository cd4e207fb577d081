// Squeeze-and-excite gate of the MBConv block (model/src/backbones/uncrtaints.py:82-97):
//   s = sigmoid(F2 * gelu(F1 * mean_hw(g2))),  F1 [32][256], F2 [256][32], no biases.
// One CTA (256 threads) per frame; the 256-vector never leaves shared memory.
//
// Backward additionally finishes the Norm2 backward statistics.  With du = dL/d(g2*s) and
// dz2 = (du*s + dpool/P) * gelu'(z2):
//   sum_p dz2        = s * sum_p du*gelu'(z2)        + dpool/P * sum_p gelu'(z2)
//   sum_p dz2*h2_hat = s * sum_p du*gelu'(z2)*h2_hat + dpool/P * sum_p gelu'(z2)*h2_hat
// where the first factors come from the dX-GEMM epilogue (sums3) and the second from the forward SE
// pooling pass (gp_stats) -- so no extra pass over the 256-channel tensor is needed.
#include "common.cuh"
#include "kernels.h"

namespace ub {

// save layout per frame: m[256] | a1[32] | s[256]
constexpr int SE_SAVE = UB_HID + UB_SE + UB_HID;

__global__ void __launch_bounds__(256) se_fwd_kernel(const double* __restrict__ pool_stats, const float* __restrict__ f1,
                                                      const float* __restrict__ f2, float* __restrict__ save,
                                                      float* __restrict__ gate, double inv_p) {
    __shared__ float m[UB_HID], z1[UB_SE];
    const int n = blockIdx.x, c = threadIdx.x, lane = c % 32, warp = c / 32;
    float* sv = save + (size_t)n * SE_SAVE;
    m[c] = (float)(pool_stats[((size_t)n * UB_HID + c) * 2] * inv_p);
    sv[c] = m[c];
    __syncthreads();
    for (int j = warp; j < UB_SE; j += 8) {
        float a = 0.f;
        for (int k = lane; k < UB_HID; k += 32) a = fmaf(f1[j * UB_HID + k], m[k], a);
        a = warp_sum(a);
        if (lane == 0) { sv[UB_HID + j] = a; z1[j] = gelu_f(a); }
    }
    __syncthreads();
    float a2 = 0.f;
#pragma unroll
    for (int j = 0; j < UB_SE; ++j) a2 = fmaf(f2[c * UB_SE + j], z1[j], a2);
    const float s = sigmoid_f(a2);
    sv[UB_HID + UB_SE + c] = s;
    gate[(size_t)n * UB_HID + c] = s;
}

__global__ void __launch_bounds__(256) se_bwd_kernel(const double* __restrict__ sums3 /* [N][256][3] */,
                                                      const double* __restrict__ gp_stats /* [N][256][2] */,
                                                      const float* __restrict__ f1, const float* __restrict__ f2,
                                                      const float* __restrict__ save, float* df1, float* df2,
                                                      float* __restrict__ dmp, double* __restrict__ bstats2, double inv_p) {
    __shared__ float da2[UB_HID], da1[UB_SE], z1[UB_SE];
    const int n = blockIdx.x, c = threadIdx.x, lane = c % 32, warp = c / 32;
    const float* sv = save + (size_t)n * SE_SAVE;
    const float s = sv[UB_HID + UB_SE + c];
    const double* s3 = sums3 + ((size_t)n * UB_HID + c) * 3;
    const float ds = (float)s3[0];
    da2[c] = ds * s * (1.f - s);
    if (c < UB_SE) z1[c] = gelu_f(sv[UB_HID + c]);
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < UB_SE; ++j) atomicAdd(&df2[c * UB_SE + j], da2[c] * z1[j]);
    for (int j = warp; j < UB_SE; j += 8) {
        float a = 0.f;
        for (int k = lane; k < UB_HID; k += 32) a = fmaf(f2[k * UB_SE + j], da2[k], a);
        a = warp_sum(a);
        if (lane == 0) da1[j] = a * gelu_grad_f(sv[UB_HID + j]);
    }
    __syncthreads();
    float dm = 0.f;
    const float mc = sv[c];
#pragma unroll 4
    for (int j = 0; j < UB_SE; ++j) {
        dm = fmaf(f1[j * UB_HID + c], da1[j], dm);
        atomicAdd(&df1[j * UB_HID + c], da1[j] * mc);
    }
    const double dmp_c = (double)dm * inv_p;
    dmp[(size_t)n * UB_HID + c] = (float)dmp_c;
    const double* gp = gp_stats + ((size_t)n * UB_HID + c) * 2;
    bstats2[((size_t)n * UB_HID + c) * 2 + 0] = (double)s * s3[1] + dmp_c * gp[0];
    bstats2[((size_t)n * UB_HID + c) * 2 + 1] = (double)s * s3[2] + dmp_c * gp[1];
}

int launch_se_fwd(const double* pool_stats, const float* f1, const float* f2, float* save, float* gate, int N, int P,
                  cudaStream_t st) {
    se_fwd_kernel<<<N, 256, 0, st>>>(pool_stats, f1, f2, save, gate, 1.0 / (double)P);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_se_bwd(const double* sums3, const double* gp_stats, const float* f1, const float* f2, const float* save,
                  float* df1, float* df2, float* dmp, double* bstats2, int N, int P, cudaStream_t st) {
    se_bwd_kernel<<<N, 256, 0, st>>>(sums3, gp_stats, f1, f2, save, df1, df2, dmp, bstats2, 1.0 / (double)P);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

}  // namespace ub
