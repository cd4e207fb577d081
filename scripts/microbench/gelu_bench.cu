// Microbenchmark: throughput of the exact-erf GELU (+ derivative) in scalar fp32 form vs packed f32x2 form (FFMA2/FMUL2/FADD2),
// at 8 and 32 resident warps per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gelu_bench gelu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../uncrtaints_b200/csrc/common.cuh"
using namespace ub;
typedef f32x2_t u64;
__device__ __forceinline__ u64 pk(float a, float b) { return pack2(a, b); }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { unpack2(v, a, b); }
__device__ __forceinline__ void gelu_both2(float x0, float x1, float& g0, float& g1, float& p0, float& p1) {
    gelu_pair<true, true>(x0, x1, g0, g1, p0, p1);
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = seed + 0.01f * (threadIdx.x + i);
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { float g, gp; gelu_both(v[i], g, gp); acc += g * gp; v[i] = v[i] * 0.999f + 0.001f; }
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) { float g0, g1, p0, p1; gelu_both2(v[i], v[i + 1], g0, g1, p0, p1); acc += g0 * p0 + g1 * p1; v[i] = v[i] * 0.999f + 0.001f; v[i + 1] = v[i + 1] * 0.999f + 0.001f; }
        } else if (MODE == 4) {      // value-only, scalar two-MUFU form
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc += gelu_f(v[i]); v[i] = v[i] * 0.999f + 0.001f; }
        } else if (MODE == 5) {      // value-only, packed single-MUFU exp-polynomial form
#pragma unroll
            for (int i = 0; i < 8; i += 2) { float g0, g1; gelu_val_pair(v[i], v[i + 1], g0, g1); acc += g0 + g1; v[i] = v[i] * 0.999f + 0.001f; v[i + 1] = v[i + 1] * 0.999f + 0.001f; }
        } else if (MODE == 2) {      // plain FFMA chains: 8 independent
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int r = 0; r < 16; ++r) v[i] = fmaf(v[i], 0.999f, acc);
            }
        } else {                     // FFMA2 chains: 4 independent pairs
            u64 c = pk(acc, acc), m = pk(0.999f, 0.999f);
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                u64 p = pk(v[i], v[i + 1]);
#pragma unroll
                for (int r = 0; r < 16; ++r) p = fma2(p, m, c);
                upk(p, v[i], v[i + 1]);
            }
        }
    }
    float s = acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int ctas_per_sm, double elems_per_iter) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 4000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148 * ctas_per_sm, 256>>>(out, 10, 0.3f);
    cudaEventRecord(a);
    k<MODE><<<148 * ctas_per_sm, 256>>>(out, iters, 0.3f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double total = (double)148 * ctas_per_sm * 256 * iters * elems_per_iter;
    printf("%-28s warps/SM=%2d  %.3f ms  %.1f Gelem/s  (%.2f elem/clk/SM @1.9GHz)\n", name, ctas_per_sm * 8, ms, total / ms / 1e6,
           total / ms / 1e6 / 148 / 1.9);
    cudaFree(out);
}
int main() {
    for (int c : {1, 2, 4}) {
        run<4>("gelu value, 2 MUFU scalar", c, 8); run<5>("gelu value, 1 MUFU packed", c, 8);
        if (c == 1) { run<0>("gelu_both scalar", 1, 8); run<1>("gelu_both packed f32x2", 1, 8); run<2>("FFMA (fma count)", 1, 128); run<3>("FFMA2 (fma count)", 1, 128); }
        if (c == 2) { run<0>("gelu_both scalar", 2, 8); run<1>("gelu_both packed f32x2", 2, 8); run<2>("FFMA (fma count)", 2, 128); run<3>("FFMA2 (fma count)", 2, 128); }
        if (c == 4) { run<0>("gelu_both scalar", 4, 8); run<1>("gelu_both packed f32x2", 4, 8); run<2>("FFMA (fma count)", 4, 128); run<3>("FFMA2 (fma count)", 4, 128); }
    }
    return 0;
}
