// Normalisation finalisers (GroupNorm / BatchNorm2d) and the HBM-bound pixel-major
// elementwise passes of an MBConv block.
//
// Reference semantics: get_norm_layer / PreNorm (model/src/backbones/uncrtaints.py:16-22,72-79),
// nn.GroupNorm(4 groups) in the encoder, nn.BatchNorm2d in the decoder; MBConv residual add
// (uncrtaints.py:142-146).  Statistics are produced by the kernel that writes a tensor (column sums
// in its epilogue, fp64 accumulation) and consumed as per-(frame, channel) scale/shift here.
#include <cuda_bf16.h>
#include "common.cuh"
#include "kernels.h"

namespace ub {

// Reduce two float4 column accumulators over the pixel-rows of a 256-thread block and add them to
// stats_n[c][0..1] (fp64).  Thread layout: c4 = tid % (C/4), row = tid / (C/4).
template <int C>
__device__ __forceinline__ void block_reduce_cols2(float4 a, float4 b, double* stats_n, float* smem) {
    constexpr int Q = C / 4, ROWS = 256 / Q;
    const int c4 = threadIdx.x % Q, r = threadIdx.x / Q;
    float4* sa = reinterpret_cast<float4*>(smem);
    float4* sb = sa + ROWS * Q;
    sa[r * Q + c4] = a;
    sb[r * Q + c4] = b;
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += 256) {
        const int which = i / C, ch = i % C;
        const float* src = reinterpret_cast<const float*>(which ? sb : sa);
        double t = 0.0;
#pragma unroll
        for (int rr = 0; rr < ROWS; ++rr) t += (double)src[rr * C + ch];
        atomicAdd(&stats_n[ch * 2 + which], t);
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// forward finaliser: raw (sum, sumsq) -> scale/shift (+ saved mean/rstd, BN running stats)
// grid = N blocks, C threads.  groups == 0 -> BatchNorm over all frames.
// ------------------------------------------------------------------------------------------
__global__ void norm_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* running_mean, float* running_var,
                                     Coef* __restrict__ coef, MeanRstd* __restrict__ mr, int N, int C, int groups,
                                     double count, float eps, float momentum, int training) {
    extern __shared__ double sm[];
    const int n = blockIdx.x, c = threadIdx.x;
    double mean, var;
    if (groups > 0) {
        sm[c] = stats[((size_t)n * C + c) * 2 + 0];
        sm[C + c] = stats[((size_t)n * C + c) * 2 + 1];
        __syncthreads();
        const int gs = C / groups, g0 = (c / gs) * gs;
        double s = 0.0, q = 0.0;
        for (int j = 0; j < gs; ++j) { s += sm[g0 + j]; q += sm[C + g0 + j]; }
        const double cnt = count * gs;
        mean = s / cnt;
        var = q / cnt - mean * mean;
    } else if (training) {
        double s = 0.0, q = 0.0;
#pragma unroll 8
        for (int j = 0; j < N; ++j) {             // independent loads: batch them, the sum is latency-bound otherwise
            const double2 v = *reinterpret_cast<const double2*>(&stats[((size_t)j * C + c) * 2]);
            s += v.x;
            q += v.y;
        }
        const double cnt = count * N;
        mean = s / cnt;
        var = q / cnt - mean * mean;
        if (n == 0 && running_mean != nullptr) {
            const double unbiased = var * (cnt / (cnt > 1.0 ? cnt - 1.0 : 1.0));
            running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
            running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
        }
    } else {
        mean = running_mean[c];
        var = running_var[c];
    }
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    const float sc = (float)(gamma[c] * rstd);
    coef[(size_t)n * C + c] = Coef{sc, (float)(beta[c] - mean * gamma[c] * rstd)};
    mr[(size_t)n * C + c] = MeanRstd{(float)mean, (float)rstd};
}

// ------------------------------------------------------------------------------------------
// backward finaliser.  bstats[n][c] = (sum_p dy, sum_p dy*x_hat), dy = grad w.r.t. the norm output.
//   dx = r*(g*dy - m1 - x_hat*m2),  m1 = mean_S(g*dy), m2 = mean_S(g*dy*x_hat)
//      = (r g) dy + (-r^2 m2) x + (r^2 m2 mu - r m1)
// S = (group channels x pixels) of one frame for GroupNorm, (all frames x pixels) of one channel for
// BatchNorm in training; eval-mode BatchNorm has no statistics term.
// Also accumulates dgamma[c] += sum_n sum_p dy*x_hat, dbeta[c] += sum_n sum_p dy.
// ------------------------------------------------------------------------------------------
__global__ void norm_finalize_bwd_kernel(const double* __restrict__ bstats, const float* __restrict__ gamma,
                                         const MeanRstd* __restrict__ mr, BCoef* __restrict__ bcoef,
                                         float* dgamma, float* dbeta, int N, int C, int groups, double count,
                                         int training) {
    extern __shared__ double sm[];
    const int n = blockIdx.x, c = threadIdx.x;
    const double g = gamma[c];
    const MeanRstd s = mr[(size_t)n * C + c];
    const double r = s.rstd, mu = s.mean;
    double m1 = 0.0, m2 = 0.0;
    if (groups > 0) {
        sm[c] = g * bstats[((size_t)n * C + c) * 2 + 0];
        sm[C + c] = g * bstats[((size_t)n * C + c) * 2 + 1];
        __syncthreads();
        const int gs = C / groups, g0 = (c / gs) * gs;
        for (int j = 0; j < gs; ++j) { m1 += sm[g0 + j]; m2 += sm[C + g0 + j]; }
        m1 /= count * gs;
        m2 /= count * gs;
    }
    double sdy = 0.0, sdyx = 0.0;
    if (groups == 0 || n == 0) {
#pragma unroll 8
        for (int j = 0; j < N; ++j) {
            const double2 v = *reinterpret_cast<const double2*>(&bstats[((size_t)j * C + c) * 2]);
            sdy += v.x;
            sdyx += v.y;
        }
    }
    if (groups == 0 && training) {
        m1 = g * sdy / (count * N);
        m2 = g * sdyx / (count * N);
    }
    bcoef[(size_t)n * C + c] = BCoef{(float)(r * g), (float)(-r * r * m2), (float)(r * r * m2 * mu - r * m1), 0.f};
    if (n == 0) {
        if (dgamma) dgamma[c] += (float)sdyx;
        if (dbeta) dbeta[c] += (float)sdy;
    }
}

// ------------------------------------------------------------------------------------------
// K5: out = x + (y*scale3 + shift3); epilogue: column sums of `out` for the next block's PreNorm.
// grid (chunks, N); 256 threads; C = 128.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) residual_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            const Coef* __restrict__ coef3, float* __restrict__ out,
                                                            double* out_stats, int P, int chunk) {
    constexpr int C = UB_WIDTH, Q = C / 4, ROWS = 256 / Q;
    __shared__ __align__(16) float smem[2 * ROWS * C];
    const int n = blockIdx.y, c4 = threadIdx.x % Q, r = threadIdx.x / Q;
    const Coef k0 = coef3[(size_t)n * C + c4 * 4 + 0], k1 = coef3[(size_t)n * C + c4 * 4 + 1],
               k2 = coef3[(size_t)n * C + c4 * 4 + 2], k3 = coef3[(size_t)n * C + c4 * 4 + 3];
    const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, P);
    float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
    const size_t base = (size_t)n * P * C + c4 * 4;
#pragma unroll 4
    for (int p = p0 + r; p < p1; p += ROWS) {
        const float4 xv = ld4_stream(x + base + (size_t)p * C);
        const float4 yv = ld4_stream(y + base + (size_t)p * C);
        float4 o;
        o.x = xv.x + fmaf(yv.x, k0.scale, k0.shift);
        o.y = xv.y + fmaf(yv.y, k1.scale, k1.shift);
        o.z = xv.z + fmaf(yv.z, k2.scale, k2.shift);
        o.w = xv.w + fmaf(yv.w, k3.scale, k3.shift);
        st4(out + base + (size_t)p * C, o);
        s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
        q.x += o.x * o.x; q.y += o.y * o.y; q.z += o.z * o.z; q.w += o.w * o.w;
    }
    if (out_stats) block_reduce_cols2<C>(s, q, out_stats + (size_t)n * C * 2, smem);
}

// ------------------------------------------------------------------------------------------
// K3: SE squeeze.  pool[n][c][0] += sum_p gelu(z2), z2 = h2*scale2 + shift2.  C = 256.
// In training it also gathers gp_stats[n][c] += (sum_p gelu'(z2), sum_p gelu'(z2)*h2_hat): the two terms of
// the Norm2 backward statistics that do not depend on the incoming gradient (see se_bwd_kernel).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld4h(const float* p) { return ld4(p); }
__device__ __forceinline__ float4 ld4h(const __nv_bfloat16* p) {
    const uint2 w = *reinterpret_cast<const uint2*>(p);
    return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xFFFF0000u), __uint_as_float(w.y << 16), __uint_as_float(w.y & 0xFFFF0000u));
}
// Row streaming like the depthwise kernels (dwconv_rows.cu): a pixel row is ONE contiguous 1 KB (fp32) run over all 256 channels, so
// thread 0 moves SE_ROWS rows per stage with one cp.async.bulk (TMA, mbarrier complete_tx) into a SE_STAGES-deep shared-memory ring
// and the 256 threads only do LDS.128 + GELU: no registers held across the HBM latency, loads always SE_STAGES - 1 stages ahead.
// The register-loop form this replaces (4 LDG.128 per thread and iteration, consumed right away) sat at 0.58 of the HBM roof with
// long_scoreboard 54 % of the stalls (ncu r02) although its MUFU / issue floor is 0.9 ms per step against 1.3 ms of HBM time.
constexpr int SE_ROWS = 16, SE_STAGES = 3;      // measured at B=16 (ms per step): 3 stages (4 CTAs / SM) 1.78, 4 stages 1.90, 6 stages 1.96; register loop 2.23
__device__ __forceinline__ uint32_t se_s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void se_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    const long long t0 = clock64();
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 4000000000ll) __trap();      // ~2 s: a lost bulk copy must fault, not hang the GPU
    } while (!ok);
}
template <class HT>
__global__ void __launch_bounds__(256) se_pool_kernel(const HT* __restrict__ h2, const Coef* __restrict__ coef2,
                                                       const MeanRstd* __restrict__ mr2, double* pool_stats,
                                                       double* gp_stats, int P, int chunk) {
    constexpr int C = UB_HID, Q = C / 4, ROWS = 256 / Q;
    constexpr uint32_t ROW_BYTES = C * sizeof(HT), STAGE_BYTES = SE_ROWS * ROW_BYTES;
    extern __shared__ __align__(128) unsigned char se_smem[];
    HT* ring = reinterpret_cast<HT*>(se_smem);                                             // SE_STAGES x [SE_ROWS][C]
    float* red = reinterpret_cast<float*>(se_smem);                                        // reused by the final reduction (8 KB)
    const uint32_t bar0 = se_s32(se_smem + SE_STAGES * STAGE_BYTES);
    const int n = blockIdx.y, c4 = threadIdx.x % Q, r = threadIdx.x / Q;
    const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, P);
    const int nst = (p1 - p0 + SE_ROWS - 1) / SE_ROWS;
    const HT* src = h2 + ((size_t)n * P + p0) * C;
    auto issue = [&](int i) {                       // thread 0: stage i -> slot i % SE_STAGES
        const uint32_t rows = (uint32_t)min(SE_ROWS, p1 - p0 - i * SE_ROWS), b = bar0 + (i % SE_STAGES) * 8;
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(b), "r"(rows * ROW_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(se_s32(ring + (size_t)(i % SE_STAGES) * SE_ROWS * C)), "l"(src + (size_t)i * SE_ROWS * C), "r"(rows * ROW_BYTES), "r"(b) : "memory");
    };
    if (threadIdx.x == 0) {
        for (int i = 0; i < SE_STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + i * 8), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < SE_STAGES - 1 && i < nst; ++i) issue(i);
    }
    // per channel PAIR, packed (f32x2): z = v * scale + shift, h_hat = v * rstd - mean * rstd, and the three accumulators -- the pass is
    // issue-bound once the loads are off the threads (ncu: issue 70 %), so the statistics around the GELU are packed as well
    f32x2_t sc2[2], sh2[2], rs2[2], nm2[2], s2[2], g2[2], gh2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const Coef ka = coef2[(size_t)n * C + c4 * 4 + 2 * i], kb = coef2[(size_t)n * C + c4 * 4 + 2 * i + 1];
        const MeanRstd ma = mr2[(size_t)n * C + c4 * 4 + 2 * i], mb = mr2[(size_t)n * C + c4 * 4 + 2 * i + 1];
        sc2[i] = pack2(ka.scale, kb.scale); sh2[i] = pack2(ka.shift, kb.shift);
        rs2[i] = pack2(ma.rstd, mb.rstd);   nm2[i] = pack2(-ma.mean * ma.rstd, -mb.mean * mb.rstd);
        s2[i] = g2[i] = gh2[i] = pack2(0.f, 0.f);
    }
    const f32x2_t one2 = pack2(1.f, 1.f);
    const bool train = gp_stats != nullptr;
    __syncthreads();                                // barrier initialisation visible to every waiter
    for (int i = 0; i < nst; ++i) {
        // slot (i - 1) % SE_STAGES was drained by every thread before the barrier that ended iteration i - 1
        if (threadIdx.x == 0 && i + SE_STAGES - 1 < nst) issue(i + SE_STAGES - 1);
        se_mbar_wait(bar0 + (i % SE_STAGES) * 8, (uint32_t)(i / SE_STAGES) & 1u);
        const HT* st = ring + (size_t)(i % SE_STAGES) * SE_ROWS * C + c4 * 4;
        const int rows = min(SE_ROWS, p1 - p0 - i * SE_ROWS);
#pragma unroll
        for (int j = 0; j < SE_ROWS / ROWS; ++j) {
            const int row = r + j * ROWS;
            if (row < rows) {
                const float4 v = ld4h(st + row * C);
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 2; ++e) {       // packed (f32x2) GELU + derivative: 21 issue slots per channel pair
                    const f32x2_t v2 = pack2(vv[2 * e], vv[2 * e + 1]);
                    float z0, z1, gz0, gz1, gp0, gp1;
                    unpack2(fma2(v2, sc2[e], sh2[e]), z0, z1);
                    if (train) {
                        gelu_pair<true, true>(z0, z1, gz0, gz1, gp0, gp1);
                        const f32x2_t gp2 = pack2(gp0, gp1);
                        g2[e] = fma2(gp2, one2, g2[e]);
                        gh2[e] = fma2(gp2, fma2(v2, rs2[e], nm2[e]), gh2[e]);
                    } else {
                        gelu_val_pair(z0, z1, gz0, gz1);
                    }
                    s2[e] = fma2(pack2(gz0, gz1), one2, s2[e]);
                }
            }
        }
        __syncthreads();                            // every thread is done with this slot: thread 0 may refill it next iteration
    }
    float s[4], g[4], gh[4];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        unpack2(s2[i], s[2 * i], s[2 * i + 1]);
        unpack2(g2[i], g[2 * i], g[2 * i + 1]);
        unpack2(gh2[i], gh[2 * i], gh[2 * i + 1]);
    }
    block_reduce_cols2<C>(make_float4(s[0], s[1], s[2], s[3]), make_float4(0, 0, 0, 0), pool_stats + (size_t)n * C * 2, red);
    if (train)
        block_reduce_cols2<C>(make_float4(g[0], g[1], g[2], g[3]), make_float4(gh[0], gh[1], gh[2], gh[3]),
                              gp_stats + (size_t)n * C * 2, red);
}

// ------------------------------------------------------------------------------------------
// B5a: statistics of the normalisation backward: bstats[n][c] += (sum dy, sum dy * v_hat),
// v_hat = (v - mean) * rstd.  C = 128 (used for the last norm of a block: dy = dOut, v = y).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) norm_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ v,
                                                              const MeanRstd* __restrict__ mr, double* bstats, int P,
                                                              int chunk) {
    constexpr int C = UB_WIDTH, Q = C / 4, ROWS = 256 / Q;
    __shared__ __align__(16) float smem[2 * ROWS * C];
    const int n = blockIdx.y, c4 = threadIdx.x % Q, r = threadIdx.x / Q;
    const MeanRstd m0 = mr[(size_t)n * C + c4 * 4 + 0], m1 = mr[(size_t)n * C + c4 * 4 + 1],
                   m2 = mr[(size_t)n * C + c4 * 4 + 2], m3 = mr[(size_t)n * C + c4 * 4 + 3];
    const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, P);
    float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
    const size_t base = (size_t)n * P * C + c4 * 4;
#pragma unroll 4
    for (int p = p0 + r; p < p1; p += ROWS) {
        const float4 d = ld4_stream(dy + base + (size_t)p * C);
        const float4 w = ld4_stream(v + base + (size_t)p * C);
        s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
        q.x += d.x * (w.x - m0.mean) * m0.rstd;
        q.y += d.y * (w.y - m1.mean) * m1.rstd;
        q.z += d.z * (w.z - m2.mean) * m2.rstd;
        q.w += d.w * (w.w - m3.mean) * m3.rstd;
    }
    block_reduce_cols2<C>(s, q, bstats + (size_t)n * C * 2, smem);
}

// ------------------------------------------------------------------------------------------
// B1: dX = dOut + (a0*dn0 + b0*x + c0)   (residual + PreNorm backward).  C = 128.
// ------------------------------------------------------------------------------------------
// relu_mask != 0 (encoder block only: its input x is the ReLU output of in_conv): dx is also multiplied by [x > 0], i.e. the
// ReLU backward of in_conv is applied here, where x is loaded anyway, and the in_conv gram pass need not read x0 again.
// BELOW: dx is the output gradient of the block below (the producer of x), whose backward starts with the statistics pass B5a over
// (dx, y_below).  Gathering them here costs one more read stream (y_below, A) instead of a separate pass over 2 A.
template <bool BELOW>
__global__ void __launch_bounds__(256, BELOW ? 4 : 0) residual_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ dn0,
                                                            const float* __restrict__ x, const BCoef* __restrict__ bc0,
                                                            float* __restrict__ dx, int P, int chunk, int relu_mask,
                                                            const float* __restrict__ y_below, const MeanRstd* __restrict__ mr_below,
                                                            double* bstats_below) {
    constexpr int C = UB_WIDTH, Q = C / 4, ROWS = 256 / Q;
    __shared__ __align__(16) float smem[BELOW ? 2 * ROWS * C : 4];
    const int n = blockIdx.y, c4 = threadIdx.x % Q, r = threadIdx.x / Q;
    const BCoef k0 = bc0[(size_t)n * C + c4 * 4 + 0], k1 = bc0[(size_t)n * C + c4 * 4 + 1],
                k2 = bc0[(size_t)n * C + c4 * 4 + 2], k3 = bc0[(size_t)n * C + c4 * 4 + 3];
    MeanRstd m0 = {}, m1 = {}, m2 = {}, m3 = {};
    if constexpr (BELOW) {
        m0 = mr_below[(size_t)n * C + c4 * 4 + 0]; m1 = mr_below[(size_t)n * C + c4 * 4 + 1];
        m2 = mr_below[(size_t)n * C + c4 * 4 + 2]; m3 = mr_below[(size_t)n * C + c4 * 4 + 3];
    }
    float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
    const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, P);
    const size_t base = (size_t)n * P * C + c4 * 4;
    auto row = [&](int p) {
        const float4 g = ld4_stream(dout + base + (size_t)p * C);
        const float4 d = ld4_stream(dn0 + base + (size_t)p * C);
        const float4 xv = ld4_stream(x + base + (size_t)p * C);
        float4 o;
        o.x = g.x + fmaf(k0.a, d.x, fmaf(k0.b, xv.x, k0.c));
        o.y = g.y + fmaf(k1.a, d.y, fmaf(k1.b, xv.y, k1.c));
        o.z = g.z + fmaf(k2.a, d.z, fmaf(k2.b, xv.z, k2.c));
        o.w = g.w + fmaf(k3.a, d.w, fmaf(k3.b, xv.w, k3.c));
        if (relu_mask) {
            o.x = xv.x > 0.f ? o.x : 0.f; o.y = xv.y > 0.f ? o.y : 0.f;
            o.z = xv.z > 0.f ? o.z : 0.f; o.w = xv.w > 0.f ? o.w : 0.f;
        }
        st4(dx + base + (size_t)p * C, o);
        if constexpr (BELOW) {
            const float4 w = ld4_stream(y_below + base + (size_t)p * C);
            s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
            q.x += o.x * (w.x - m0.mean) * m0.rstd;
            q.y += o.y * (w.y - m1.mean) * m1.rstd;
            q.z += o.z * (w.z - m2.mean) * m2.rstd;
            q.w += o.w * (w.w - m3.mean) * m3.rstd;
        }
    };
    if constexpr (BELOW) {
#pragma unroll 1                                    // 64 registers (4 CTAs per SM) without spills; the compiler's 4x unrolling needs 70+
        for (int p = p0 + r; p < p1; p += ROWS) row(p);
    } else {
        for (int p = p0 + r; p < p1; p += ROWS) row(p);
    }
    if constexpr (BELOW) block_reduce_cols2<C>(s, q, bstats_below + (size_t)n * C * 2, smem);
}

// ------------------------------------------------------------------------------------------
// Residual blocks (ResidualConvBlock, uncrtaints.py:24-69): ConvLayer = conv3x3 -> norm -> ReLU.  C = 128.
//   R-fwd : out = x + relu(c3*scale3 + shift3)                                     (uncrtaints.py:64-68)
//   R-bwd1: bstats[n][ch] += (sum dz, sum dz * c_hat), dz = dy * [c*scale + shift > 0]   (ReLU + norm backward statistics)
//   R-bwd2: dc = a*dz + b*c + cc (norm backward applied), dbias[ch] += sum dc       (the convolution's output gradient, materialised once
//           -- as a pre-split bf16 hi/lo image -- because its input-gradient GEMM reads every element nine times)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) residual_relu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ c3,
                                                                 const Coef* __restrict__ coef3, float* __restrict__ out, double* out_stats,
                                                                 int P, int chunk) {
    constexpr int C = UB_WIDTH, Q = C / 4, ROWS = 256 / Q;
    __shared__ __align__(16) float smem[2 * ROWS * C];
    const int n = blockIdx.y, c4 = threadIdx.x % Q, r = threadIdx.x / Q;
    const Coef k0 = coef3[(size_t)n * C + c4 * 4 + 0], k1 = coef3[(size_t)n * C + c4 * 4 + 1],
               k2 = coef3[(size_t)n * C + c4 * 4 + 2], k3 = coef3[(size_t)n * C + c4 * 4 + 3];
    const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, P);
    float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
    const size_t base = (size_t)n * P * C + c4 * 4;
#pragma unroll 4
    for (int p = p0 + r; p < p1; p += ROWS) {
        const float4 xv = ld4_stream(x + base + (size_t)p * C);
        const float4 yv = ld4_stream(c3 + base + (size_t)p * C);
        float4 o;
        o.x = xv.x + fmaxf(fmaf(yv.x, k0.scale, k0.shift), 0.f);
        o.y = xv.y + fmaxf(fmaf(yv.y, k1.scale, k1.shift), 0.f);
        o.z = xv.z + fmaxf(fmaf(yv.z, k2.scale, k2.shift), 0.f);
        o.w = xv.w + fmaxf(fmaf(yv.w, k3.scale, k3.shift), 0.f);
        st4(out + base + (size_t)p * C, o);
        s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
        q.x += o.x * o.x; q.y += o.y * o.y; q.z += o.z * o.z; q.w += o.w * o.w;
    }
    if (out_stats) block_reduce_cols2<C>(s, q, out_stats + (size_t)n * C * 2, smem);
}
__global__ void __launch_bounds__(256) relu_norm_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ c,
                                                                   const Coef* __restrict__ coef, const MeanRstd* __restrict__ mr,
                                                                   double* bstats, int P, int chunk) {
    constexpr int C = UB_WIDTH, Q = C / 4, ROWS = 256 / Q;
    __shared__ __align__(16) float smem[2 * ROWS * C];
    const int n = blockIdx.y, c4 = threadIdx.x % Q, r = threadIdx.x / Q;
    Coef k[4]; MeanRstd m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { k[i] = coef[(size_t)n * C + c4 * 4 + i]; m[i] = mr[(size_t)n * C + c4 * 4 + i]; }
    const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, P);
    float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    const size_t base = (size_t)n * P * C + c4 * 4;
#pragma unroll 4
    for (int p = p0 + r; p < p1; p += ROWS) {
        const float4 d4 = ld4_stream(dy + base + (size_t)p * C);
        const float4 c4v = ld4_stream(c + base + (size_t)p * C);
        const float d[4] = {d4.x, d4.y, d4.z, d4.w}, cv[4] = {c4v.x, c4v.y, c4v.z, c4v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float dz = fmaf(cv[i], k[i].scale, k[i].shift) > 0.f ? d[i] : 0.f;
            s[i] += dz;
            q[i] = fmaf(dz, (cv[i] - m[i].mean) * m[i].rstd, q[i]);
        }
    }
    block_reduce_cols2<C>(make_float4(s[0], s[1], s[2], s[3]), make_float4(q[0], q[1], q[2], q[3]), bstats + (size_t)n * C * 2, smem);
}
// dc is written as a PRE-SPLIT image (512-byte rows: 128 bf16 hi halves, then 128 bf16 lo halves; gemm_tc.cu: SPLIT_ROW): both GEMMs
// that consume it (input gradient: nine taps per element; weight gradient: five launches) copy the halves straight into their tiles.
__device__ __forceinline__ void st_split4_bf16(char* row_base, int c0, const float (&v)[4]) {
    uint2 h, l;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h.x) : "f"(v[1]), "f"(v[0]));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h.y) : "f"(v[3]), "f"(v[2]));
    const float r0 = v[0] - __uint_as_float(h.x << 16), r1 = v[1] - __uint_as_float(h.x & 0xFFFF0000u);
    const float r2 = v[2] - __uint_as_float(h.y << 16), r3 = v[3] - __uint_as_float(h.y & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l.x) : "f"(r1), "f"(r0));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l.y) : "f"(r3), "f"(r2));
    *reinterpret_cast<uint2*>(row_base + c0 * 2) = h;
    *reinterpret_cast<uint2*>(row_base + 256 + c0 * 2) = l;
}
__global__ void __launch_bounds__(256) relu_norm_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ c,
                                                                   const Coef* __restrict__ coef, const BCoef* __restrict__ bc,
                                                                   char* __restrict__ dc, float* dbias, int H, int W, int chunk) {
    const int P = H * W;
    constexpr int C = UB_WIDTH, Q = C / 4, ROWS = 256 / Q;
    __shared__ __align__(16) float smem[ROWS * C];
    const int n = blockIdx.y, c4 = threadIdx.x % Q, r = threadIdx.x / Q;
    Coef k[4]; BCoef b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { k[i] = coef[(size_t)n * C + c4 * 4 + i]; b[i] = bc[(size_t)n * C + c4 * 4 + i]; }
    const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, P);
    float s[4] = {0, 0, 0, 0};
    const size_t base = (size_t)n * P * C + c4 * 4;
#pragma unroll 4
    for (int p = p0 + r; p < p1; p += ROWS) {
        const float4 d4 = ld4_stream(dy + base + (size_t)p * C);
        const float4 c4v = ld4_stream(c + base + (size_t)p * C);
        const float d[4] = {d4.x, d4.y, d4.z, d4.w}, cv[4] = {c4v.x, c4v.y, c4v.z, c4v.w};
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float dz = fmaf(cv[i], k[i].scale, k[i].shift) > 0.f ? d[i] : 0.f;
            o[i] = fmaf(b[i].a, dz, fmaf(b[i].b, cv[i], b[i].c));
            s[i] += o[i];
        }
        const int y = p / W, xx = p - y * W;
        st_split4_bf16(dc + split_pixel_offset(n, y, xx, H, W), c4 * 4, o);
        if (xx == 0 || xx == W - 1) {             // zero halo pixel beside the image (the input gradient's operand is zero-extended)
            const float z[4] = {0.f, 0.f, 0.f, 0.f};
            st_split4_bf16(dc + split_pixel_offset(n, y, xx == 0 ? -1 : W, H, W), c4 * 4, z);
        }
    }
    if (dbias) {
        reinterpret_cast<float4*>(smem)[r * Q + c4] = make_float4(s[0], s[1], s[2], s[3]);
        __syncthreads();
        if (threadIdx.x < C) {
            float t = 0.f;
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr) t += smem[rr * C + threadIdx.x];
            atomicAdd(&dbias[threadIdx.x], t);
        }
    }
}

// Adjoint of the reflection of padding_mode='reflect' for the input gradient of a 3x3 convolution.  The main kernel computes the
// transposed convolution g of the zero-extended output gradient on the H x W grid; the same g on the one-pixel ring AROUND the image
// belongs to the pixels the ring reflects onto: din[(1, x')] += g[(-1, x)], din[(H-2, x')] += g[(H, x)], din[(y, 1)] += g[(y, -1)],
// din[(y, W-2)] += g[(y, W)], x' = reflect(x) for x in [-1, W] (corners included in the two rows).
//   g[r][ci] = sum_tap sum_co W[co][ci][tap] * dc[r - tap][co]   over the taps whose source pixel r - tap lies inside the image:
// for a ring pixel of the top / bottom row the three taps of one kernel ROW, for the left / right column those of one kernel COLUMN.
// A CTA owns 16 consecutive ring pixels of one side of one frame: the 18 source pixels they touch are staged in shared memory, each
// of the three 128 x 128 tap matrices is staged once ([tap][co][ci] fp32 image, conv_fold_prep_kernel) and every thread (= ci)
// accumulates its 16 outputs with 64 FMAs per four shared-memory loads.  (The first version, one CTA per ring pixel gathering the
// weights with stride 9 from the [co][ci][3][3] tensor, took 25 ms per step at B=16 -- as long as the weight-gradient GEMMs.)
constexpr int FOLD_PX = 16;
__global__ void conv_fold_prep_kernel(const float* __restrict__ w /* [co][ci][9] */, float* __restrict__ wt /* [9][co][ci] */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9 * UB_WIDTH * UB_WIDTH) return;
    const int tap = i / (UB_WIDTH * UB_WIDTH), r = i % (UB_WIDTH * UB_WIDTH);
    wt[i] = w[(size_t)r * 9 + tap];
}
__global__ void __launch_bounds__(128) conv_fold_kernel(const char* __restrict__ dc /* pre-split bf16 hi/lo image */,
                                                         const float* __restrict__ wt /* [9][co][ci] */, float* din, int H, int W) {
    extern __shared__ __align__(16) float fsm[];
    float* sW = fsm;                                    // [co][ci] of the current tap: 64 KB
    float* sdc = fsm + UB_WIDTH * UB_WIDTH;             // [FOLD_PX + 2][co]: source pixels start-1 .. start+FOLD_PX
    const int n = blockIdx.y, ci = threadIdx.x;
    const int nb_row = (W + 2 + FOLD_PX - 1) / FOLD_PX, nb_col = (H + FOLD_PX - 1) / FOLD_PX;
    int bx = blockIdx.x, side;                          // 0 top, 1 bottom, 2 left, 3 right
    if (bx < 2 * nb_row) { side = bx / nb_row; bx -= side * nb_row; } else { bx -= 2 * nb_row; side = 2 + bx / nb_col; bx -= (side - 2) * nb_col; }
    const bool row_side = side < 2;
    const int len = row_side ? W : H;                   // extent of the image along the side
    const int first = (row_side ? -1 : 0) + bx * FOLD_PX;      // ring coordinate (x for rows: -1 .. W, y for columns: 0 .. H-1) of position 0
    const int fixed = (side == 0 || side == 2) ? 0 : (row_side ? H - 1 : W - 1);   // the image row / column the sources lie in
    // stage the FOLD_PX + 2 source pixels along the side: coordinate first - 1 + j
    for (int j = 0; j < FOLD_PX + 2; ++j) {
        const int s = first - 1 + j;
        float v = 0.f;
        if (s >= 0 && s < len) {
            const size_t off = row_side ? split_pixel_offset(n, fixed, s, H, W) : split_pixel_offset(n, s, fixed, H, W);
            const unsigned short hb = *reinterpret_cast<const unsigned short*>(dc + off + ci * 2);
            const unsigned short lb = *reinterpret_cast<const unsigned short*>(dc + off + 256 + ci * 2);
            v = __uint_as_float((uint32_t)hb << 16) + __uint_as_float((uint32_t)lb << 16);
        }
        sdc[j * UB_WIDTH + ci] = v;
    }
    float acc[FOLD_PX];
#pragma unroll
    for (int i = 0; i < FOLD_PX; ++i) acc[i] = 0.f;
    for (int k = -1; k <= 1; ++k) {                     // the tap offset ALONG the side; across it the offset is fixed by the side
        // ring pixel r = source + tap  =>  across the side: top / left: -1 = 0 + tap -> tap = -1; bottom / right: tap = +1
        const int across = (side == 0 || side == 2) ? -1 : 1;
        const int tap = row_side ? (across + 1) * 3 + (k + 1) : (k + 1) * 3 + (across + 1);
        __syncthreads();
        for (int i = ci; i < UB_WIDTH * UB_WIDTH / 4; i += 128)
            reinterpret_cast<float4*>(sW)[i] = *reinterpret_cast<const float4*>(wt + (size_t)tap * UB_WIDTH * UB_WIDTH + (size_t)i * 4);
        __syncthreads();
#pragma unroll 2
        for (int c4 = 0; c4 < UB_WIDTH / 4; ++c4) {
            const float w0 = sW[(c4 * 4 + 0) * UB_WIDTH + ci], w1 = sW[(c4 * 4 + 1) * UB_WIDTH + ci];
            const float w2 = sW[(c4 * 4 + 2) * UB_WIDTH + ci], w3 = sW[(c4 * 4 + 3) * UB_WIDTH + ci];
#pragma unroll
            for (int i = 0; i < FOLD_PX; ++i) {         // ring position first + i, source coordinate first + i - k -> staged index i + 1 - k
                const float4 d = *reinterpret_cast<const float4*>(sdc + (i + 1 - k) * UB_WIDTH + c4 * 4);
                acc[i] = fmaf(w0, d.x, fmaf(w1, d.y, fmaf(w2, d.z, fmaf(w3, d.w, acc[i]))));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < FOLD_PX; ++i) {
        const int r = first + i;                        // ring coordinate along the side
        if (row_side ? (r > W) : (r >= H)) continue;
        int ty, tx;
        if (row_side) { ty = side == 0 ? 1 : H - 2; tx = r < 0 ? 1 : (r >= W ? W - 2 : r); }
        else { ty = r; tx = side == 2 ? 1 : W - 2; }
        atomicAdd(&din[((size_t)n * H * W + (size_t)ty * W + tx) * UB_WIDTH + ci], acc[i]);
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
#ifndef UB_RB_MAXROWS
#define UB_RB_MAXROWS 1024      // pixels per CTA of the statistics-gathering residual backward pass (whole waves, see launch_residual_bwd)
#endif
static inline int chunk_for(int P) { return P >= 4096 ? 1024 : (P >= 1024 ? 256 : 64); }

int launch_norm_finalize(const double* stats, const float* gamma, const float* beta, float* rm, float* rv, Coef* coef,
                         MeanRstd* mr, int N, int C, int groups, double count, float eps, float momentum, int training,
                         cudaStream_t st) {
    norm_finalize_kernel<<<N, C, 2 * C * sizeof(double), st>>>(stats, gamma, beta, rm, rv, coef, mr, N, C, groups, count,
                                                                 eps, momentum, training);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_norm_finalize_bwd(const double* bstats, const float* gamma, const MeanRstd* mr, BCoef* bcoef, float* dgamma,
                             float* dbeta, int N, int C, int groups, double count, int training, cudaStream_t st) {
    norm_finalize_bwd_kernel<<<N, C, 2 * C * sizeof(double), st>>>(bstats, gamma, mr, bcoef, dgamma, dbeta, N, C, groups,
                                                                     count, training);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_residual_fwd(const float* x, const float* y, const Coef* coef3, float* out, double* out_stats, int N, int P,
                        cudaStream_t st) {
    const int chunk = chunk_for(P);
    residual_fwd_kernel<<<dim3((P + chunk - 1) / chunk, N), 256, 0, st>>>(x, y, coef3, out, out_stats, P, chunk);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
// CTAs per frame such that the whole grid is (just under) a whole number of waves of the resident CTAs: the pass is bound by its
// instruction stream, not by HBM alone, so a ragged last wave costs its full share (1024 CTAs on 592 slots ran 2 waves for 1.73)
static int se_pool_ctas_per_frame(int N, int P, int slots) {
    if (P < 4096) return (P + 255) / 256;
    const long long rows = (long long)N * P;
    int w = 1;
    while (rows / ((long long)slots * w) > 2048) ++w;       // at most ~2048 rows per CTA
    int g = (int)(((long long)slots * w) / N);
    if (g < 1) g = 1;
    if (g > P / SE_ROWS) g = P / SE_ROWS;
    return g;
}
template <class HT>
static int launch_se_pool_t(const void* h2, const Coef* coef2, const MeanRstd* mr2, double* pool_stats, double* gp_stats, int N, int P,
                            cudaStream_t st) {
    constexpr size_t smem = (size_t)SE_STAGES * SE_ROWS * UB_HID * sizeof(HT) + SE_STAGES * 8;
    static_assert(smem >= 2 * (256 / (UB_HID / 4)) * UB_HID * sizeof(float), "the ring doubles as the reduction scratch");
    auto kern = se_pool_kernel<HT>;
    UB_SET_SMEM(kern, smem);
    static int occ_cache[64] = {0};
    int dev = 0, occ = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return UB_ERR_CUDA;
    if (dev >= 0 && dev < 64 && occ_cache[dev] > 0) occ = occ_cache[dev];
    else {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem) != cudaSuccess || occ < 1) occ = 1;
        if (dev >= 0 && dev < 64) occ_cache[dev] = occ;
    }
    const int g = se_pool_ctas_per_frame(N, P, device_sm_count() * occ);
    const int chunk = ((P + g - 1) / g + SE_ROWS - 1) / SE_ROWS * SE_ROWS;
    kern<<<dim3((P + chunk - 1) / chunk, N), 256, smem, st>>>(static_cast<const HT*>(h2), coef2, mr2, pool_stats, gp_stats, P, chunk);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_se_pool(const void* h2, const Coef* coef2, const MeanRstd* mr2, double* pool_stats, double* gp_stats, int N,
                   int P, int hbf, cudaStream_t st) {
    return hbf ? launch_se_pool_t<__nv_bfloat16>(h2, coef2, mr2, pool_stats, gp_stats, N, P, st)
               : launch_se_pool_t<float>(h2, coef2, mr2, pool_stats, gp_stats, N, P, st);
}
int launch_norm_bwd_stats(const float* dy, const float* v, const MeanRstd* mr, double* bstats, int N, int P,
                          cudaStream_t st) {
    const int chunk = chunk_for(P);
    norm_bwd_stats_kernel<<<dim3((P + chunk - 1) / chunk, N), 256, 0, st>>>(dy, v, mr, bstats, P, chunk);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_residual_bwd(const float* dout, const float* dn0, const float* x, const BCoef* bc0, float* dx, int N, int P,
                        int relu_mask, const float* y_below, const MeanRstd* mr_below, double* bstats_below, cudaStream_t st) {
    int chunk = chunk_for(P);
    if (y_below && P >= 4096) {
        // the statistics variant holds fewer CTAs per SM (registers, reduction scratch): 1024 CTAs on 3 x 148 slots ran three waves
        // for 2.3 waves of work (0.8 of the HBM roof) -- size the grid to whole waves of what is resident, <= UB_RB_MAXROWS pixels per CTA (1024: 3.165 ms per step, 2048: 3.18)
        static int occ_cache[64] = {0};
        int dev = 0, occ = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return UB_ERR_CUDA;
        if (dev >= 0 && dev < 64 && occ_cache[dev] > 0) occ = occ_cache[dev];
        else {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, residual_bwd_kernel<true>, 256, 0) != cudaSuccess || occ < 1) occ = 1;
            if (dev >= 0 && dev < 64) occ_cache[dev] = occ;
        }
        const long long slots = (long long)device_sm_count() * occ, rows = (long long)N * P;
        int w = 1;
        while (rows / (slots * w) > UB_RB_MAXROWS) ++w;
        long long g = slots * w / N;
        if (g < 1) g = 1;
        if (g > P / 8) g = P / 8;
        chunk = (int)(((P + g - 1) / g + 7) / 8 * 8);
    }
    const dim3 grid((P + chunk - 1) / chunk, N);
    if (y_below) residual_bwd_kernel<true><<<grid, 256, 0, st>>>(dout, dn0, x, bc0, dx, P, chunk, relu_mask, y_below, mr_below, bstats_below);
    else residual_bwd_kernel<false><<<grid, 256, 0, st>>>(dout, dn0, x, bc0, dx, P, chunk, relu_mask, nullptr, nullptr, nullptr);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

int launch_residual_relu_fwd(const float* x, const float* c3, const Coef* coef3, float* out, double* out_stats, int N, int P, cudaStream_t st) {
    const int chunk = chunk_for(P);
    residual_relu_fwd_kernel<<<dim3((P + chunk - 1) / chunk, N), 256, 0, st>>>(x, c3, coef3, out, out_stats, P, chunk);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_relu_norm_bwd_stats(const float* dy, const float* c, const Coef* coef, const MeanRstd* mr, double* bstats, int N, int P, cudaStream_t st) {
    const int chunk = chunk_for(P);
    relu_norm_bwd_stats_kernel<<<dim3((P + chunk - 1) / chunk, N), 256, 0, st>>>(dy, c, coef, mr, bstats, P, chunk);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_relu_norm_bwd_apply(const float* dy, const float* c, const Coef* coef, const BCoef* bc, void* dc_split, float* dbias, int N, int H,
                               int W, cudaStream_t st) {
    const int P = H * W, chunk = chunk_for(P);
    relu_norm_bwd_apply_kernel<<<dim3((P + chunk - 1) / chunk, N), 256, 0, st>>>(dy, c, coef, bc, static_cast<char*>(dc_split), dbias, H, W, chunk);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_conv_fold(const void* dc_split, const float* w, float* wt_scratch, float* din, int N, int H, int W, cudaStream_t st) {
    conv_fold_prep_kernel<<<(9 * UB_WIDTH * UB_WIDTH + 255) / 256, 256, 0, st>>>(w, wt_scratch);
    UB_CHECK_LAUNCH();
    constexpr size_t smem = (size_t)(UB_WIDTH * UB_WIDTH + (FOLD_PX + 2) * UB_WIDTH) * sizeof(float);
    UB_SET_SMEM(conv_fold_kernel, smem);
    const int nb_row = (W + 2 + FOLD_PX - 1) / FOLD_PX, nb_col = (H + FOLD_PX - 1) / FOLD_PX;
    conv_fold_kernel<<<dim3(2 * nb_row + 2 * nb_col, N), 128, smem, st>>>(static_cast<const char*>(dc_split), wt_scratch, din, H, W);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

}  // namespace ub
