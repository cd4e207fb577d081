// Row-streaming depthwise 3x3 (reflect) kernels of the MBConv block: forward K2 and the FUSED backward B3
// (model/src/backbones/uncrtaints.py:128-133: GELU -> Conv2d(groups=256, padding_mode='reflect') -> Norm -> GELU).
//
// Layout recap: act[n][y][x][c], c = 256 contiguous channels = 1 KB per pixel.  A run of pixels along x is
// therefore ONE contiguous byte range covering every channel, which is what the TMA bulk-copy engine (cp.async.bulk,
// 1-D, mbarrier complete_tx) moves best: one instruction per image row of a strip, full 1 KB DRAM bursts, no per-thread
// address arithmetic, no registers held across the load.
//
// A CTA owns a strip of RW = 16 output columns x R rows of one frame for ALL 256 channels and walks down the rows:
//   * the (16 + 2)-pixel halo row y+1 arrives by bulk copy into a 5-slot shared-memory ring (2 rows in flight),
//   * all 256 threads transform it in place (forward: Norm1 + GELU; backward: SE-gate backward, gelu'(z2), Norm2
//     backward -> dh2, forced to 0 outside the image),
//   * one __syncthreads later the stencil for row y reads rows y-1, y, y+1 of the ring.  A thread owns one channel quad
//     and 4 adjacent columns: a 6-column window per ring row (6 LDS.128) feeds 4 pixels x 3 taps, the 9 x 4 depthwise
//     taps, the Norm coefficients and the weight-gradient accumulators live in registers.
// The backward kernel is input-stationary (D[tap] gathered per INPUT pixel, so gelu(z1), gelu'(z1)
// are evaluated once per pixel and dW needs no second tile); the adjoint of reflect padding is a rare warp-uniform
// correction.  HBM traffic: forward 2 Hh, backward 4 Hh (du, h2, h1 in, dz1 out) -- the split dh2 + stencil pair moved 6.
#include <cuda_bf16.h>
#include "common.cuh"
#include "kernels.h"

namespace ub {
namespace {

constexpr int RW = 16, RWH = RW + 2, RC = UB_HID, RQ = RC / 4;
constexpr int ROWF = RWH * RC;      // floats per halo row   (18 KB)
constexpr int ROWC = RW * RC;       // floats per centre row (16 KB)
constexpr int ND = 5, NH2 = 3, NH1 = 3;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    const long long t0 = clock64();
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 4000000000ll) __trap();      // ~2 s: a lost bulk copy must fault, not hang the GPU
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// acc += a * b on 4 lanes; F2 = two packed FFMA2 (fma.rn.f32x2, sm_100+) instead of four FFMA
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
template <bool F2>
__device__ __forceinline__ void fma4(float4& acc, const float4 a, const float4 b) {
    if constexpr (F2) {
        unsigned long long d0, d1;
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d0) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)), "l"(pk2(acc.x, acc.y)));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d1) : "l"(pk2(a.z, a.w)), "l"(pk2(b.z, b.w)), "l"(pk2(acc.z, acc.w)));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(d0));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.z), "=f"(acc.w) : "l"(d1));
    } else {
        acc.x = fmaf(a.x, b.x, acc.x); acc.y = fmaf(a.y, b.y, acc.y);
        acc.z = fmaf(a.z, b.z, acc.z); acc.w = fmaf(a.w, b.w, acc.w);
    }
}
// hidden-tensor storage type HT: float, or __nv_bfloat16 (gemm_backend bit 5).  Vector loads / stores of 4 and 2 elements.
typedef __nv_bfloat16 bf16_t;
__device__ __forceinline__ float4 ldh4(const float* p) { return ld4(p); }
__device__ __forceinline__ float4 ldh4(const bf16_t* p) {
    const uint2 w = *reinterpret_cast<const uint2*>(p);
    return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xFFFF0000u), __uint_as_float(w.y << 16), __uint_as_float(w.y & 0xFFFF0000u));
}
__device__ __forceinline__ void sth4(float* p, float4 v) { st4(p, v); }
__device__ __forceinline__ void sth4(bf16_t* p, float4 v) {
    uint2 w;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w.x) : "f"(v.y), "f"(v.x));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w.y) : "f"(v.w), "f"(v.z));
    *reinterpret_cast<uint2*>(p) = w;
}
__device__ __forceinline__ float2 ldh2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 ldh2(const bf16_t* p) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(p);
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
}
__device__ __forceinline__ void sth2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
__device__ __forceinline__ void sth2(bf16_t* p, float2 v) {
    uint32_t w;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(v.y), "f"(v.x));
    *reinterpret_cast<uint32_t*>(p) = w;
}
template <class T> struct is_f32 { static constexpr bool value = false; };
template <> struct is_f32<float> { static constexpr bool value = true; };

__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// reduce the per-thread (s, q) column sums over the 4 column groups of each channel quad -> stats[n][ch][0..1] (fp64)
__device__ __forceinline__ void reduce_sq(float4 ssum, float4 qsum, double* stats_n, float* scratch, int q, int s) {
    float4* red = reinterpret_cast<float4*>(scratch);
    red[s * RQ + q] = ssum;
    red[4 * RQ + s * RQ + q] = qsum;
    __syncthreads();
    {
        const int ch = threadIdx.x;
        double t0 = 0.0, t1 = 0.0;
#pragma unroll
        for (int ss = 0; ss < 4; ++ss) {
            t0 += (double)scratch[ss * RC + ch];
            t1 += (double)scratch[4 * RC + ss * RC + ch];
        }
        atomicAdd(&stats_n[ch * 2 + 0], t0);
        atomicAdd(&stats_n[ch * 2 + 1], t1);
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------------
// K2:  h2 = DW3x3_reflect(gelu(h1 * scale1 + shift1)),  stats2 += column (sum, sumsq) of h2
// grid (W/16, H/R, N), 256 threads, 2 CTAs / SM (92 KB ring each)
// ------------------------------------------------------------------------------------------------------------------
// POOL (eval-mode BatchNorm blocks: the Norm2 coefficients come from running statistics and are known before this kernel runs):
// instead of the (sum, sumsq) of h2 the epilogue accumulates sum_p gelu(h2 * scale2 + shift2), i.e. the squeeze-excite pooling --
// the separate se_pool pass over h2 disappears.  `stats2` then points at the pooling accumulator [N][256][2].
// HT = bf16: the h1 halo rows arrive in a 2-slot bf16 staging area (the bulk copy cannot convert) and the Norm1 + GELU transform
// writes them as fp32 into the ring instead of working in place; h2 is written as bf16 (statistics from the fp32 values).
template <bool F2, bool PG, bool POOL, class HT>
__global__ void __launch_bounds__(256, 2)
dwrows_fwd_kernel(const HT* __restrict__ h1, const Coef* __restrict__ coef1, const float* __restrict__ wdw,
                  HT* __restrict__ h2, double* stats2, const Coef* __restrict__ coef2, int H, int W, int R) {
    constexpr bool F32 = is_f32<HT>::value;
    constexpr int NSLOT = F32 ? ND : 2;                     // slots the bulk copies land in (ring itself / staging)
    extern __shared__ __align__(128) float smem[];
    float* sD = smem;
    HT* sIn = F32 ? reinterpret_cast<HT*>(sD) : reinterpret_cast<HT*>(sD + ND * ROWF);
    const uint32_t barD = F32 ? s32(sD + ND * ROWF) : s32(reinterpret_cast<char*>(sD + ND * ROWF) + 2 * ROWF * sizeof(HT));
    const int tid = threadIdx.x, q = tid & (RQ - 1), s = tid >> 6;
    const int n = blockIdx.z, x0 = blockIdx.x * RW, ybase = blockIdx.y * R;
    const size_t fbase = (size_t)n * H * W * RC;
    float4 wr[9], sc, sh;
    {
        float w_[9][4], a_[4], b_[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const Coef k = coef1[(size_t)n * RC + q * 4 + c];
            a_[c] = k.scale; b_[c] = k.shift;
#pragma unroll
            for (int j = 0; j < 9; ++j) w_[j][c] = wdw[(size_t)(q * 4 + c) * 9 + j];
        }
        sc = make_float4(a_[0], a_[1], a_[2], a_[3]);
        sh = make_float4(b_[0], b_[1], b_[2], b_[3]);
#pragma unroll
        for (int j = 0; j < 9; ++j) wr[j] = make_float4(w_[j][0], w_[j][1], w_[j][2], w_[j][3]);
    }
    float4 sc2 = make_float4(0, 0, 0, 0), sh2 = make_float4(0, 0, 0, 0);
    if constexpr (POOL) {
        const Coef k0 = coef2[(size_t)n * RC + q * 4 + 0], k1 = coef2[(size_t)n * RC + q * 4 + 1],
                   k2 = coef2[(size_t)n * RC + q * 4 + 2], k3 = coef2[(size_t)n * RC + q * 4 + 3];
        sc2 = make_float4(k0.scale, k1.scale, k2.scale, k3.scale);
        sh2 = make_float4(k0.shift, k1.shift, k2.shift, k3.shift);
    }
    if (tid == 0) {
        for (int i = 0; i < NSLOT; ++i) mbar_init(barD + i * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int xs = max(x0 - 1, 0), xe = min(x0 + RW + 1, W);
    const uint32_t rowbytes = (uint32_t)(xe - xs) * RC * sizeof(HT);
    const int dstoff = (xs - (x0 - 1)) * RC;
    const int NIT = R + 2;
    auto issue = [&](int i) {       // thread 0: halo row i (image row reflect(ybase-1+i)) -> slot i % NSLOT
        const int y = reflect1(ybase - 1 + i, H);
        const uint32_t b = barD + (i % NSLOT) * 8;
        HT* dst = sIn + (i % NSLOT) * ROWF;
        const HT* src = h1 + fbase + (size_t)y * W * RC;
        mbar_expect_tx(b, ROWF * sizeof(HT));
        bulk_g2s(s32(dst + dstoff), src + (size_t)xs * RC, rowbytes, b);
        if (x0 == 0) bulk_g2s(s32(dst), src + (size_t)1 * RC, RC * sizeof(HT), b);                                  // x = -1 -> 1
        if (x0 + RW == W) bulk_g2s(s32(dst + (RWH - 1) * RC), src + (size_t)(W - 2) * RC, RC * sizeof(HT), b);      // x = W -> W-2
    };
    if (tid == 0) { issue(0); issue(1); }
    float4 ssum = make_float4(0, 0, 0, 0), qsum = make_float4(0, 0, 0, 0);
    for (int i = 0; i < NIT; ++i) {
        float* d = sD + (i % ND) * ROWF;
        const HT* din = sIn + (i % NSLOT) * ROWF;
        mbar_wait(barD + (i % NSLOT) * 8, (uint32_t)(i / NSLOT) & 1u);
        // Norm1 + GELU of the halo row (in place for fp32 storage): pixel px = s + 4k, channel quad q
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int px = s + 4 * k;
            if (px < RWH) {
                float* ptr = d + px * RC + q * 4;
                const float4 v = ldh4(din + px * RC + q * 4);
                const float4 z = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
                float4 g;
                if constexpr (PG) g = gelu4_packed(z);
                else g = make_float4(gelu_f(z.x), gelu_f(z.y), gelu_f(z.z), gelu_f(z.w));
                st4(ptr, g);
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0 && i + 2 < NIT) issue(i + 2);
        if (i >= 2) {
            const int yc = ybase + i - 2;
            const int woff = (4 * s) * RC + q * 4;
            float4 o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = make_float4(0, 0, 0, 0);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float* rp = sD + ((i - 2 + r) % ND) * ROWF + woff;
                float4 win[6];
#pragma unroll
                for (int wc = 0; wc < 6; ++wc) win[wc] = ld4(rp + wc * RC);
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int jj = 0; jj < 3; ++jj) fma4<F2>(o[j], wr[r * 3 + jj], win[j + jj]);
            }
            HT* op = h2 + fbase + ((size_t)yc * W + x0 + 4 * s) * RC + q * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sth4(op + j * RC, o[j]);
                if constexpr (POOL) {
                    const float4 g = gelu4_packed(make_float4(fmaf(o[j].x, sc2.x, sh2.x), fmaf(o[j].y, sc2.y, sh2.y),
                                                              fmaf(o[j].z, sc2.z, sh2.z), fmaf(o[j].w, sc2.w, sh2.w)));
                    ssum.x += g.x; ssum.y += g.y; ssum.z += g.z; ssum.w += g.w;
                } else {
                    ssum.x += o[j].x; ssum.y += o[j].y; ssum.z += o[j].z; ssum.w += o[j].w;
                    fma4<F2>(qsum, o[j], o[j]);
                }
            }
        }
    }
    __syncthreads();
    reduce_sq(ssum, qsum, stats2 + (size_t)n * RC * 2, sD, q, s);
}

// (A warp-specialised form -- 8 stencil warps at 200 registers + 8 transform warps at 56 via setmaxnreg, mbarrier hand-off
// per row instead of the CTA barrier -- was built and measured at 2.0 ms against 0.96 ms for this kernel at N=16: with the
// 5-slot ring only one row can be in flight behind the three the stencil holds, and the lean transform warps lose their
// ILP.  It was removed)

// ------------------------------------------------------------------------------------------------------------------
// B3 fused:  dz2 = (du*gate + dpool/P) * gelu'(z2);  dh2 = a2*dz2 + b2*h2 + c2;  dg1 = DW3x3^T(dh2) (+ reflect adjoint);
//            dz1 = dg1 * gelu'(z1);  dWdw[c][tap] += sum_p g1[p] * D[tap][p];  bstats1 += (sum dz1, sum dz1 * h1_hat)
// grid (W/16, H/R, N), 512 threads, 1 CTA / SM (200 KB of rings).  A thread owns a channel PAIR (one f32x2 lane pair) and 4
// adjacent columns: 9 taps + 9 dW accumulators + coefficients are 48 registers, the kernel fits 128 registers and 16 warps
// (4 per scheduler) are resident: the row loop is latency-bound (ncu r01: 45 % issue utilisation with 2 warps per scheduler
// in the earlier 256-thread channel-quad form), not throughput-bound.
// ------------------------------------------------------------------------------------------------------------------
constexpr int RP = RC / 2;     // channel pairs
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void st2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
__device__ __forceinline__ void fma2v(float2& acc, const float2 a, const float2 b) {
    unpack2(fma2(pack2(a.x, a.y), pack2(b.x, b.y), pack2(acc.x, acc.y)), acc.x, acc.y);
}

// HT = bf16: du arrives in a 2-slot bf16 staging area and is written into the ring as fp32 dh2; the h2 / h1 staging rings hold bf16
// (half the bytes); dz1 is written as bf16 (statistics from the fp32 values).
template <bool PG, class HT>
__global__ void __launch_bounds__(512, 1)
dwrows_bwd2_kernel(const HT* __restrict__ du, const HT* __restrict__ h2, const HT* __restrict__ h1,
                   const float* __restrict__ gate, const float* __restrict__ dmp, const Coef* __restrict__ coef2,
                   const BCoef* __restrict__ bc2, const Coef* __restrict__ coef1, const MeanRstd* __restrict__ mr1,
                   const float* __restrict__ wdw, HT* __restrict__ dz1, double* bstats1, float* dwdw, int H, int W, int R) {
    constexpr bool F32 = is_f32<HT>::value;
    constexpr int NDU = F32 ? ND : 2;                       // slots the du bulk copies land in (ring itself / staging)
    extern __shared__ __align__(128) float smem[];
    float* sD = smem;                                                    // dh2 ring (fp32), ND halo rows
    HT* sH2 = reinterpret_cast<HT*>(sD + ND * ROWF);                     // h2 staging, NH2 halo rows
    HT* sH1 = sH2 + NH2 * ROWF;                                          // h1 staging, NH1 centre rows
    HT* sDu = F32 ? reinterpret_cast<HT*>(sD) : sH1 + NH1 * ROWC;        // du landing area
    float4* sK2 = reinterpret_cast<float4*>(F32 ? reinterpret_cast<char*>(sH1 + NH1 * ROWC) : reinterpret_cast<char*>(sDu + 2 * ROWF));
                                                                         // [2][128] scale2, shift2, gate, dpool/P   (channel 2p+c at [c][p])
    float4* sB2 = sK2 + RC;                                        // [2][128] a2, b2, c2, -
    const uint32_t barD = s32(sB2 + RC), barH2 = barD + ND * 8, barH1 = barH2 + NH2 * 8;
    const int tid = threadIdx.x, p = tid & (RP - 1), s = tid >> 7;
    const int n = blockIdx.z, x0 = blockIdx.x * RW, ybase = blockIdx.y * R;
    const size_t fbase = (size_t)n * H * W * RC;
    if (tid < RC) {
        const size_t ci = (size_t)n * RC + tid;
        const Coef a = coef2[ci];
        const BCoef bb = bc2[ci];
        const int slot = (tid & 1) * RP + (tid >> 1);
        sK2[slot] = make_float4(a.scale, a.shift, gate[ci], dmp[ci]);
        sB2[slot] = make_float4(bb.a, bb.b, bb.c, 0.f);
    }
    float2 wr[9], sc1, sh1, mu1, rs1;
    {
        const size_t c0 = (size_t)n * RC + p * 2;
        const Coef ka = coef1[c0], kb = coef1[c0 + 1];
        const MeanRstd ma = mr1[c0], mb = mr1[c0 + 1];
        sc1 = make_float2(ka.scale, kb.scale); sh1 = make_float2(ka.shift, kb.shift);
        mu1 = make_float2(ma.mean, mb.mean); rs1 = make_float2(ma.rstd, mb.rstd);
#pragma unroll
        for (int j = 0; j < 9; ++j) wr[j] = make_float2(wdw[(size_t)(p * 2) * 9 + j], wdw[(size_t)(p * 2 + 1) * 9 + j]);
    }
    if (tid == 0) {
        for (int i = 0; i < ND + NH2 + NH1; ++i) mbar_init(barD + i * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int xs = max(x0 - 1, 0), xe = min(x0 + RW + 1, W);
    const uint32_t rowbytes = (uint32_t)(xe - xs) * RC * sizeof(HT);
    const int dstoff = (xs - (x0 - 1)) * RC;
    const int NIT = R + 2;
    auto issue_dh = [&](int i) {
        const int y = ybase - 1 + i;
        const uint32_t bd = barD + (i % NDU) * 8, bh = barH2 + (i % NH2) * 8;
        if (y >= 0 && y < H) {
            const size_t g = fbase + ((size_t)y * W + xs) * RC;
            mbar_expect_tx(bd, rowbytes);
            bulk_g2s(s32(sDu + (i % NDU) * ROWF + dstoff), du + g, rowbytes, bd);
            mbar_expect_tx(bh, rowbytes);
            bulk_g2s(s32(sH2 + (i % NH2) * ROWF + dstoff), h2 + g, rowbytes, bh);
        } else {
            mbar_arrive(bd);
            mbar_arrive(bh);
        }
    };
    auto issue_h1 = [&](int c) {
        const uint32_t b = barH1 + (c % NH1) * 8;
        mbar_expect_tx(b, ROWC * sizeof(HT));
        bulk_g2s(s32(sH1 + (c % NH1) * ROWC), h1 + fbase + ((size_t)(ybase + c) * W + x0) * RC, ROWC * sizeof(HT), b);
    };
    if (tid == 0) { issue_dh(0); issue_dh(1); }
    float2 gw[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) gw[j] = make_float2(0, 0);
    float2 ssum = make_float2(0, 0), qsum = make_float2(0, 0);
    const int xbase = x0 + 4 * s;
    const bool colb = (xbase <= 1 && 1 < xbase + 4) || (xbase <= W - 2 && W - 2 < xbase + 4);
    const int woff = (4 * s) * RC + p * 2;

    for (int i = 0; i < NIT; ++i) {
        {   // ---- du row -> dh2 row in place (pixel px = s + 4k, channel pair p); 0 outside the image ----
            const int y = ybase - 1 + i;
            const bool row_in = (y >= 0 && y < H);
            float* d = sD + (i % ND) * ROWF;
            const HT* dsrc = sDu + (i % NDU) * ROWF;
            const HT* hh = sH2 + (i % NH2) * ROWF;
            mbar_wait(barD + (i % NDU) * 8, (uint32_t)(i / NDU) & 1u);
            mbar_wait(barH2 + (i % NH2) * 8, (uint32_t)(i / NH2) & 1u);
            const float4 k0 = sK2[p], k1 = sK2[RP + p];
            const float4 b0 = sB2[p], b1 = sB2[RP + p];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int px = s + 4 * k;
                if (px < RWH) {
                    const int x = x0 - 1 + px;
                    float2 o = make_float2(0, 0);
                    if (row_in && x >= 0 && x < W) {
                        const float2 dv = ldh2(dsrc + px * RC + p * 2);
                        const float2 hv = ldh2(hh + px * RC + p * 2);
                        const float z0 = fmaf(hv.x, k0.x, k0.y), z1 = fmaf(hv.y, k1.x, k1.y);
                        float gp0, gp1;
                        if constexpr (PG) { float u0, u1; gelu_pair<false, true>(z0, z1, u0, u1, gp0, gp1); }
                        else { gp0 = gelu_grad_f(z0); gp1 = gelu_grad_f(z1); }
                        o.x = fmaf(b0.x, fmaf(dv.x, k0.z, k0.w) * gp0, fmaf(b0.y, hv.x, b0.z));
                        o.y = fmaf(b1.x, fmaf(dv.y, k1.z, k1.w) * gp1, fmaf(b1.y, hv.y, b1.z));
                    }
                    st2(d + px * RC + p * 2, o);
                }
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            if (i + 2 < NIT) issue_dh(i + 2);
            if (i < R) issue_h1(i);
        }
        if (i < 2) continue;
        // ---- input-stationary stencil for image row yc: 4 pixels x 1 channel pair per thread ----
        const int c = i - 2, yc = ybase + c;
        const float* rp0 = sD + ((i - 2) % ND) * ROWF + woff;
        const float* rp1 = sD + ((i - 1) % ND) * ROWF + woff;
        const float* rp2 = sD + (i % ND) * ROWF + woff;
        mbar_wait(barH1 + (c % NH1) * 8, (uint32_t)(c / NH1) & 1u);
        const HT* hp = sH1 + (c % NH1) * ROWC + woff;
        float2 g[4], gp[4], o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 hv = ldh2(hp + j * RC);
            const float z0 = fmaf(hv.x, sc1.x, sh1.x), z1 = fmaf(hv.y, sc1.y, sh1.y);
            if constexpr (PG) gelu_pair<true, true>(z0, z1, g[j].x, g[j].y, gp[j].x, gp[j].y);
            else { gelu_both(z0, g[j].x, gp[j].x); gelu_both(z1, g[j].y, gp[j].y); }
            o[j] = make_float2(0, 0);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float* rp = r == 0 ? rp0 : (r == 1 ? rp1 : rp2);
            const int it = 2 - r;
            float2 win[6];
#pragma unroll
            for (int wc = 0; wc < 6; ++wc) win[wc] = ld2(rp + wc * RC);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    fma2v(o[j], wr[it * 3 + jj], win[j + 2 - jj]);
                    fma2v(gw[it * 3 + jj], g[j], win[j + 2 - jj]);
                }
        }
        const bool y_lo = (yc == 1), y_hi = (yc == H - 2);
        if (y_lo | y_hi | colb) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool x_lo = (xbase + j == 1), x_hi = (xbase + j == W - 2);
                if (y_lo | y_hi | x_lo | x_hi) {
#pragma unroll
                    for (int it = 0; it < 3; ++it)
#pragma unroll
                        for (int jj = 0; jj < 3; ++jj) {
                            const bool ya = (it == 0 && y_lo) || (it == 2 && y_hi);
                            const bool xa = (jj == 0 && x_lo) || (jj == 2 && x_hi);
                            if (ya | xa) {
                                const float* ra = it == 0 ? rp0 : (it == 1 ? rp1 : rp2);
                                const float* rb = it == 0 ? rp2 : (it == 1 ? rp1 : rp0);
                                float2 e = make_float2(0, 0);
                                if (ya) { const float2 tt = ld2(ra + (j + 2 - jj) * RC); e.x += tt.x; e.y += tt.y; }
                                if (xa) { const float2 tt = ld2(rb + (j + jj) * RC); e.x += tt.x; e.y += tt.y; }
                                if (ya && xa) { const float2 tt = ld2(ra + (j + jj) * RC); e.x += tt.x; e.y += tt.y; }
                                fma2v(o[j], wr[it * 3 + jj], e);
                                fma2v(gw[it * 3 + jj], g[j], e);
                            }
                        }
                }
            }
        }
        HT* op = dz1 + fbase + ((size_t)yc * W + xbase) * RC + p * 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 dz = make_float2(o[j].x * gp[j].x, o[j].y * gp[j].y);
            sth2(op + j * RC, dz);
            const float2 hv = ldh2(hp + j * RC);
            ssum.x += dz.x; ssum.y += dz.y;
            qsum.x = fmaf(dz.x, (hv.x - mu1.x) * rs1.x, qsum.x);
            qsum.y = fmaf(dz.y, (hv.y - mu1.y) * rs1.y, qsum.y);
        }
    }
    __syncthreads();
    {   // statistics: sum over the 4 column groups of every channel
        float2* red = reinterpret_cast<float2*>(sD);
        red[s * RP + p] = ssum;
        red[4 * RP + s * RP + p] = qsum;
        __syncthreads();
        if (tid < RC) {
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int ss = 0; ss < 4; ++ss) {
                t0 += (double)sD[ss * RC + tid];
                t1 += (double)sD[4 * RC + ss * RC + tid];
            }
            atomicAdd(&bstats1[((size_t)n * RC + tid) * 2 + 0], t0);
            atomicAdd(&bstats1[((size_t)n * RC + tid) * 2 + 1], t1);
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 9; ++j) red[(j * 4 + s) * RP + p] = gw[j];
        __syncthreads();
        for (int e = tid; e < 9 * RC; e += 512) {
            const int j = e / RC, ch = e % RC;
            float t = 0.f;
#pragma unroll
            for (int ss = 0; ss < 4; ++ss) t += sD[(j * 4 + ss) * RC + ch];
            atomicAdd(&dwdw[(size_t)ch * 9 + j], t);
        }
    }
}

template <class HT> constexpr size_t fwd_smem() {
    return is_f32<HT>::value ? (size_t)ND * ROWF * 4 + ND * 8 : (size_t)ND * ROWF * 4 + 2 * ROWF * sizeof(HT) + 2 * 8;
}
template <class HT> constexpr size_t bwd_smem() {
    return (size_t)ND * ROWF * 4 + (size_t)NH2 * ROWF * sizeof(HT) + (size_t)NH1 * ROWC * sizeof(HT) + (is_f32<HT>::value ? 0 : 2 * ROWF * sizeof(HT)) +
           2 * RC * 16 + (ND + NH2 + NH1) * 8;
}

// rows per CTA: 64 when the grid still fills the GPU several times over (halo rows and the per-CTA prologue / reduction
// epilogue are amortised over more rows), else 32 / 16 / 8
int rows_per_cta(int H, int W, int N) {
    if (H % 64 == 0 && (long long)(W / RW) * (H / 64) * N >= 4 * 148) return 64;
    return H % 32 == 0 ? 32 : (H % 16 == 0 ? 16 : (H % 8 == 0 ? 8 : 0));
}

}  // namespace

template <bool POOL, class HT>
static int launch_fwd_t(const void* h1, const Coef* coef1, const float* wdw, void* h2, double* stats, const Coef* coef2, int N, int H, int W,
                        cudaStream_t st) {
    const int R = rows_per_cta(H, W, N);
    if (W % RW != 0 || R == 0 || H < 4 || W < 4) return UB_ERR_ARG;
    constexpr size_t smem = fwd_smem<HT>();
    UB_SET_SMEM((dwrows_fwd_kernel<true, true, POOL, HT>), smem);
    const dim3 grid(W / RW, H / R, N);
    dwrows_fwd_kernel<true, true, POOL, HT><<<grid, 256, smem, st>>>(static_cast<const HT*>(h1), coef1, wdw, static_cast<HT*>(h2), stats, coef2, H, W, R);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
// hbf != 0: h1 / h2 (and du / dz1 in the backward) are bf16 tensors
int launch_dwconv_fwd(const void* h1, const Coef* coef1, const float* wdw, void* h2, double* stats2, int N, int H, int W, int hbf,
                      cudaStream_t st) {
    return hbf ? launch_fwd_t<false, bf16_t>(h1, coef1, wdw, h2, stats2, nullptr, N, H, W, st)
               : launch_fwd_t<false, float>(h1, coef1, wdw, h2, stats2, nullptr, N, H, W, st);
}
// eval-mode BatchNorm block: depthwise conv + squeeze-excite pooling in one pass (pool[n][c][0] += sum_p gelu(norm2(h2)))
int launch_dwconv_fwd_pool(const void* h1, const Coef* coef1, const float* wdw, void* h2, const Coef* coef2, double* pool, int N,
                           int H, int W, int hbf, cudaStream_t st) {
    return hbf ? launch_fwd_t<true, bf16_t>(h1, coef1, wdw, h2, pool, coef2, N, H, W, st)
               : launch_fwd_t<true, float>(h1, coef1, wdw, h2, pool, coef2, N, H, W, st);
}

template <class HT>
static int launch_bwd_t(const void* du, const void* h2, const void* h1, const float* gate, const float* dmp, const Coef* coef2,
                        const BCoef* bc2, const Coef* coef1, const MeanRstd* mr1, const float* wdw, void* dz1, double* bstats1,
                        float* dwdw, int N, int H, int W, cudaStream_t st) {
    const int R = rows_per_cta(H, W, N);
    if (W % RW != 0 || R == 0 || H < 4 || W < 4) return UB_ERR_ARG;
    constexpr size_t smem = bwd_smem<HT>();
    UB_SET_SMEM((dwrows_bwd2_kernel<true, HT>), smem);
    const dim3 grid(W / RW, H / R, N);
    dwrows_bwd2_kernel<true, HT><<<grid, 512, smem, st>>>(static_cast<const HT*>(du), static_cast<const HT*>(h2), static_cast<const HT*>(h1), gate, dmp,
                                                         coef2, bc2, coef1, mr1, wdw, static_cast<HT*>(dz1), bstats1, dwdw, H, W, R);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_dwconv_bwd(const void* du, const void* h2, const void* h1, const float* gate, const float* dmp, const Coef* coef2,
                      const BCoef* bc2, const Coef* coef1, const MeanRstd* mr1, const float* wdw, void* dz1, double* bstats1,
                      float* dwdw, int N, int H, int W, int hbf, cudaStream_t st) {
    return hbf ? launch_bwd_t<bf16_t>(du, h2, h1, gate, dmp, coef2, bc2, coef1, mr1, wdw, dz1, bstats1, dwdw, N, H, W, st)
               : launch_bwd_t<float>(du, h2, h1, gate, dmp, coef2, bc2, coef1, mr1, wdw, dz1, bstats1, dwdw, N, H, W, st);
}

}  // namespace ub
