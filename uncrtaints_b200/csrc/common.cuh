// Common device helpers for the UnCRtainTS sm_100a hot path.
//
// Data layout in HBM: every internal activation is pixel-major ("NHWC"):
//   act[n][p][c], n = frame (b*T+t for the encoder, b for the decoder), p = y*W+x, c = channel,
// so the 1x1 convolutions are row-major GEMMs with the reduction dimension contiguous and every
// normalisation statistic is a column sum.  API tensors stay NCHW (converted inside in_conv /
// out_conv, which touch only 15 / 26 channels).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define UB_WIDTH 128      // encoder/decoder width   (uncrtaints.py:235-236 defaults)
#define UB_HID 256        // MBConv hidden = width * expansion(2)  (uncrtaints.py:105,317)
#define UB_SE 32          // SE bottleneck int(128*0.25) (uncrtaints.py:83-87,134)
#define UB_HEADS 16       // n_head (uncrtaints.py:242)
#define UB_LOW 32         // att_down (uncrtaints.py:403)
#define UB_TMAX 8         // sequence lengths with register-resident (templated) L-TAE / aggregation kernels
#define UB_TLONG 64       // longest sequence (run-time loop kernels; the dataset yields up to 30, data/dataLoader.py:200)
#define UB_S2 13          // S2_BANDS (uncrtaints.py:13)
#define UB_MET_ACC 16     // doubles per sample in the image-metrics accumulator (metrics.cu)

#define UB_OK 0
#define UB_ERR_ARG -1
#define UB_ERR_CUDA -2
#define UB_ERR_WORKSPACE -3

namespace ub { extern unsigned long long g_launch_count; }   // kernels launched by this library (model.cu)

#define UB_CHECK_LAUNCH()                                  \
    do {                                                   \
        cudaError_t e__ = cudaGetLastError();              \
        if (e__ != cudaSuccess) return UB_ERR_CUDA;        \
        ++::ub::g_launch_count;                            \
    } while (0)

// Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute is per DEVICE, so the "already done" cache is a
// per-call-site bit mask indexed by the current device (a process may drive several GPUs).
#define UB_SET_SMEM(kern, bytes)                                                                                       \
    do {                                                                                                               \
        static unsigned long long done__ = 0ull;                                                                       \
        int dev__ = 0;                                                                                                 \
        if (cudaGetDevice(&dev__) != cudaSuccess) return UB_ERR_CUDA;                                                  \
        if (dev__ >= 64 || !((done__ >> dev__) & 1ull)) {                                                              \
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)) != cudaSuccess)  \
                return UB_ERR_CUDA;                                                                                    \
            if (dev__ < 64) done__ |= 1ull << dev__;                                                                   \
        }                                                                                                              \
    } while (0)

namespace ub {

// SM count of the CURRENT device (cached per device)
inline int device_sm_count() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (dev >= 0 && dev < 64) cache[dev] = n;
    return n;
}

// ---- per-(frame, channel) coefficient records --------------------------------------------
// forward:  v_norm = v * scale + shift          (GroupNorm and BatchNorm alike)
// saved:    mean, rstd                          (for x_hat = (v-mean)*rstd in backward)
// backward: dv = a*dy + b*v + c                 (normalisation backward, see norm.cu)
struct __align__(8) Coef { float scale, shift; };
struct __align__(8) MeanRstd { float mean, rstd; };
struct __align__(16) BCoef { float a, b, c, pad; };

__device__ __forceinline__ Coef ldg(const Coef* p) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
    return Coef{v.x, v.y};
}
__device__ __forceinline__ MeanRstd ldg(const MeanRstd* p) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
    return MeanRstd{v.x, v.y};
}

// Exact-erf GELU (nn.GELU default; uncrtaints.py:88,128,133) and its derivative.
// erfc(u) = t*(a1 + t*(a2 + t*(a3 + t*(a4 + t*a5))))*exp(-u^2), t = 1/(1 + p*u), u >= 0 (Abramowitz-Stegun 7.1.26,
// |err| <= 1.5e-7): one MUFU.RCP + one MUFU.EX2 + 5 FMA, and the exponential exp(-x^2/2) is shared with the
// derivative's Gaussian term.  Measured max abs error vs fp64: 4.2e-7 (gelu), 3.0e-7 (gelu'), i.e. below torch's own
// fp32 erf path (1.2e-6) -- and ~2x (gelu) to ~4x (gelu + gelu') fewer instructions than erff()/expf().
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ void gelu_both(float x, float& g, float& gp) {
    const float t = fast_rcp(fmaf(0.23164189f, fabsf(x), 1.0f));            // p / sqrt(2) = 0.3275911 * 0.70710678
    float poly = fmaf(t, 1.061405429f, -1.453152027f);
    poly = fmaf(t, poly, 1.421413741f);
    poly = fmaf(t, poly, -0.284496736f);
    poly = fmaf(t, poly, 0.254829592f);
    const float e = fast_ex2(-0.72134752f * x * x);                           // exp(-x^2 / 2)
    const float q = 0.5f * t * poly * e;                                      // 0.5 * erfc(|x| / sqrt(2))
    const float cdf = x >= 0.f ? 1.0f - q : q;
    g = x * cdf;
    gp = fmaf(x * 0.39894228040143267794f, e, cdf);
}
// Packed (f32x2) form of the same formula for two values: the polynomial, the products and the final FMAs are FFMA2 /
// FMUL2 (sm_100 `fma.rn.f32x2` / `mul.rn.f32x2`), the two MUFU ops per value stay scalar: 21 issue slots per PAIR instead
// of 19 per value.  FFMA2 has the same FMA/clk as FFMA on B200 (measured, scripts/microbench/gelu_bench.cu), so this only
// pays in issue-bound kernels.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) { f32x2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <bool WANT_G, bool WANT_GP>
__device__ __forceinline__ void gelu_pair(float x0, float x1, float& g0, float& g1, float& p0, float& p1) {
    const float t0 = fast_rcp(fmaf(0.23164189f, fabsf(x0), 1.0f)), t1 = fast_rcp(fmaf(0.23164189f, fabsf(x1), 1.0f));
    const f32x2_t t = pack2(t0, t1), x = pack2(x0, x1);
    // 0.5 * (a1..a5) of Abramowitz-Stegun 7.1.26: q = 0.5 * erfc(|x| / sqrt(2))
    f32x2_t poly = fma2(t, pack2(0.5307027145f, 0.5307027145f), pack2(-0.7265760135f, -0.7265760135f));
    poly = fma2(t, poly, pack2(0.7107068705f, 0.7107068705f));
    poly = fma2(t, poly, pack2(-0.142248368f, -0.142248368f));
    poly = fma2(t, poly, pack2(0.127414796f, 0.127414796f));
    float a0, a1;
    unpack2(mul2(mul2(x, pack2(-0.72134752f, -0.72134752f)), x), a0, a1);
    const f32x2_t e = pack2(fast_ex2(a0), fast_ex2(a1));                       // exp(-x^2 / 2)
    float q0, q1;
    unpack2(mul2(mul2(t, poly), e), q0, q1);
    const f32x2_t cdf = pack2(x0 >= 0.f ? 1.0f - q0 : q0, x1 >= 0.f ? 1.0f - q1 : q1);
    if (WANT_G) unpack2(mul2(x, cdf), g0, g1);
    if (WANT_GP) unpack2(fma2(mul2(x, pack2(0.39894228040143267794f, 0.39894228040143267794f)), e, cdf), p0, p1);
}
// Value-only GELU with ONE MUFU per value: 0.5 * erfc(|x| / sqrt(2)) = 2^(P(|x|) - 1) with a degree-8 polynomial P (least-squares
// fit of log2(erfc) weighted by erfc on [0, 6.22], coefficients folded for |x|; max |error| of Phi 4e-8, of gelu 6e-7 over
// [-8, 8] in fp32 -- the fp32 rounding of x * Phi itself).  The exact-erf GELU of gelu_pair costs two MUFU (rcp + ex2) and is
// MUFU-bound at ~6 values/clk/SM; the loaders that only need the value use this form: 8 FFMA2 + 2 ex2 per pair.
__device__ __forceinline__ void gelu_val_pair(float x0, float x1, float& g0, float& g1) {
    const float a0 = fminf(fabsf(x0), 6.2225397f), a1 = fminf(fabsf(x1), 6.2225397f);
    const f32x2_t a = pack2(a0, a1);
    f32x2_t p = fma2(a, pack2(-2.763850034e-06f, -2.763850034e-06f), pack2(3.850608118e-05f, 3.850608118e-05f));
    p = fma2(a, p, pack2(-1.819549361e-04f, -1.819549361e-04f));
    p = fma2(a, p, pack2(-1.473483862e-04f, -1.473483862e-04f));
    p = fma2(a, p, pack2(7.077381015e-03f, 7.077381015e-03f));
    p = fma2(a, p, pack2(-5.250628293e-02f, -5.250628293e-02f));
    p = fma2(a, p, pack2(-4.592045248e-01f, -4.592045248e-01f));
    p = fma2(a, p, pack2(-1.151105762e+00f, -1.151105762e+00f));
    p = fma2(a, p, pack2(-1.0f + 2.239119468e-08f, -1.0f + 2.239119468e-08f));      // c0 - 1: the factor 0.5
    float p0, p1;
    unpack2(p, p0, p1);
    const float h0 = fast_ex2(p0), h1 = fast_ex2(p1);                                  // Phi(-|x|)
    unpack2(mul2(pack2(x0, x1), pack2(x0 >= 0.f ? 1.0f - h0 : h0, x1 >= 0.f ? 1.0f - h1 : h1)), g0, g1);
}
// float4 front ends: v = x * scale + shift first (packed as well)
__device__ __forceinline__ void gelu_both4_packed(const float4 z, float4& g, float4& gp) {
    gelu_pair<true, true>(z.x, z.y, g.x, g.y, gp.x, gp.y);
    gelu_pair<true, true>(z.z, z.w, g.z, g.w, gp.z, gp.w);
}
__device__ __forceinline__ float4 gelu4_packed(const float4 z) {
    float4 g;
    gelu_val_pair(z.x, z.y, g.x, g.y);
    gelu_val_pair(z.z, z.w, g.z, g.w);
    return g;
}
__device__ __forceinline__ float4 gelu_grad4_packed(const float4 z) {
    float4 gp; float d0, d1;
    gelu_pair<false, true>(z.x, z.y, d0, d1, gp.x, gp.y);
    gelu_pair<false, true>(z.z, z.w, d0, d1, gp.z, gp.w);
    return gp;
}
__device__ __forceinline__ float gelu_f(float x) { float g, gp; gelu_both(x, g, gp); return g; }
__device__ __forceinline__ float gelu_grad_f(float x) { float g, gp; gelu_both(x, g, gp); return gp; }
// torch.clamp_(min=eps) (losses.py:203-205) keeps NaN; CUDA's fmaxf(NaN, eps) would return eps and hide a diverged
// variance head behind a finite loss
__device__ __forceinline__ float clamp_min_keep_nan(float v, float lo) { return v != v ? v : fmaxf(v, lo); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// streaming (read-once / write-once) variants: keep L1 for the reused weights and coefficients
__device__ __forceinline__ float4 ld4_stream(const float* p) {
    float4 r;
    // not volatile: .nc data is read-only for the whole kernel, so the compiler may hoist and batch these loads
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st4_stream(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- pre-split operand images of the 3x3 convolutions (gemm_tc.cu, norm.cu) ----
// hi / lo halves of a [N][H][W][128] fp32 tensor: one pixel = 512 bytes (128 hi halves, then 128 lo halves, 2 bytes each); image rows
// are padded by ONE halo pixel on each side (x' = x + 1 in [0, W + 2)), which holds the reflected neighbour (activations: x' = 0 <- x = 1,
// x' = W + 1 <- x = W - 2) or zeros (output gradients), so that a tile shifted by a tap's dx is a plain box of the image for the TMA.
#define UB_SPLIT_ROW 512
__host__ __device__ __forceinline__ size_t split_pixel_offset(size_t n, int y, int x, int H, int W) {
    return ((n * (size_t)H + (size_t)y) * (size_t)(W + 2) + (size_t)(x + 1)) * UB_SPLIT_ROW;
}
inline size_t split_image_bytes(size_t N, int H, int W) { return N * (size_t)H * (size_t)(W + 2) * UB_SPLIT_ROW; }

// ---- Philox4x32-10 counter RNG (dropout on the upsampled attention, uncrtaints.py:154,202) ----
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
// Bernoulli keep decision for element `idx` of the [16*B, T, H, W] upsampled attention tensor.
// One Philox call covers 4 consecutive elements; returns 1.0f (keep) or 0.0f (drop).
__device__ __forceinline__ float dropout_keep(uint64_t seed, uint64_t offset, uint64_t idx, float p) {
    const uint64_t blk = (idx >> 2) + offset;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), 0u, 0u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t lane = (uint32_t)(idx & 3);
    const uint32_t bits = lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
    const float u = (float)(bits >> 8) * (1.0f / 16777216.0f);
    return u >= p ? 1.0f : 0.0f;
}

// Bilinear taps of nn.Upsample(mode='bilinear', align_corners=False) (uncrtaints.py:198-200):
// src = max((o+0.5)/s - 0.5, 0); i0 = floor(src); i1 = min(i0+1, in-1); lam = src - i0.
__device__ __forceinline__ void bilinear_tap(int o, float inv_scale, int in_size, int& i0, int& i1, float& lam) {
    float src = ((float)o + 0.5f) * inv_scale - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = (int)src;
    i1 = i0 + 1 < in_size ? i0 + 1 : in_size - 1;
    lam = src - (float)i0;
}

}  // namespace ub
