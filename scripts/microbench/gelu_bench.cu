// Microbenchmark: throughput of the exact-erf GELU (+ derivative) in scalar fp32 form vs packed f32x2 form (FFMA2/FMUL2/FADD2),
// at 8 and 32 resident warps per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gelu_bench gelu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../uncrtaints_b200/csrc/common.cuh"
using namespace ub;
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

__device__ __forceinline__ void gelu_both2(float x0, float x1, float& g0, float& g1, float& p0, float& p1) {
    const float t0 = fast_rcp(fmaf(0.23164189f, fabsf(x0), 1.0f)), t1 = fast_rcp(fmaf(0.23164189f, fabsf(x1), 1.0f));
    const u64 t = pk(t0, t1), x = pk(x0, x1);
    u64 poly = fma2(t, pk(0.5f * 1.061405429f, 0.5f * 1.061405429f), pk(0.5f * -1.453152027f, 0.5f * -1.453152027f));
    poly = fma2(t, poly, pk(0.5f * 1.421413741f, 0.5f * 1.421413741f));
    poly = fma2(t, poly, pk(0.5f * -0.284496736f, 0.5f * -0.284496736f));
    poly = fma2(t, poly, pk(0.5f * 0.254829592f, 0.5f * 0.254829592f));
    const u64 xs = mul2(x, pk(-0.72134752f, -0.72134752f));
    float a0, a1;
    upk(mul2(xs, x), a0, a1);
    const float e0 = fast_ex2(a0), e1 = fast_ex2(a1);
    const u64 e = pk(e0, e1);
    float q0, q1;
    upk(mul2(mul2(t, poly), e), q0, q1);
    const float c0 = x0 >= 0.f ? 1.0f - q0 : q0, c1 = x1 >= 0.f ? 1.0f - q1 : q1;
    const u64 cdf = pk(c0, c1);
    upk(mul2(x, cdf), g0, g1);
    upk(fma2(mul2(x, pk(0.39894228f, 0.39894228f)), e, cdf), p0, p1);
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
        if (c == 1) { run<0>("gelu_both scalar", 1, 8); run<1>("gelu_both packed f32x2", 1, 8); run<2>("FFMA (fma count)", 1, 128); run<3>("FFMA2 (fma count)", 1, 128); }
        if (c == 2) { run<0>("gelu_both scalar", 2, 8); run<1>("gelu_both packed f32x2", 2, 8); run<2>("FFMA (fma count)", 2, 128); run<3>("FFMA2 (fma count)", 2, 128); }
        if (c == 4) { run<0>("gelu_both scalar", 4, 8); run<1>("gelu_both packed f32x2", 4, 8); run<2>("FFMA (fma count)", 4, 128); run<3>("FFMA2 (fma count)", 4, 128); }
    }
    return 0;
}
