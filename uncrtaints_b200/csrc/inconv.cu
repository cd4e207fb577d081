// in_conv of UnCRtainTS: Conv2d(C_in -> 128, k=1, bias) -> GroupNorm(4) -> ReLU
// (model/src/backbones/utae.py:453-497,500-520; built at uncrtaints.py:310-314), fused with the temporal
// pad-mask test (uncrtaints.py:392-394) and with the NCHW -> pixel-major layout change.
//
// K = C_in = 15 is far too thin for tensor cores, and the 128-channel conv output is 8.5x larger than the
// input, so the conv output is never written.  c0 = W x + b is affine in the 15-channel input, so the GroupNorm
// statistics follow from per-frame input moments (inconv_moments_kernel); the apply pass (inconv_apply_kernel)
// recomputes the conv, normalises, applies ReLU and writes x0 (plus the column sums that the encoder block's
// PreNorm needs).  Backward is ONE gram pass over dX0 plus fp64 finalize kernels (in_conv has no input gradient).
#include "common.cuh"
#include "kernels.h"

namespace ub {

constexpr int IC_MAXC = 16;   // max input channels (13 optical + 2 SAR = 15)
constexpr int IC_PX = 256;    // pixels staged per iteration (one per thread and input channel; next tile prefetched in registers)
constexpr int IC_GROUPS = IC_PX / 32;   // 4-pixel groups per warp and tile

// (Measured and rejected: the lane's 4 x 16 weights in registers as f32x2 pairs with the input tile staged duplicated (x, x), 2 LDS + 8
// FFMA2 per input channel and pixel group instead of 2 LDS + 16 FFMA -- 128 registers, 2 CTAs per SM: 0.69 -> 0.77 ms at N=48; the pass
// lives on resident warps hiding the store stream, not on its instruction count.)
// x0 = relu(gn(W x + b)) written pixel-major, stats += column (sum, sumsq) of x0
__global__ void __launch_bounds__(256) inconv_apply_kernel(const float* __restrict__ x /* [N][Cin][P] */, const float* __restrict__ w /* [128][Cin] */,
                                                            const float* __restrict__ bias, const Coef* __restrict__ coef,
                                                            float* __restrict__ x0, double* stats, int Cin, int P, int chunk) {
    constexpr int C = UB_WIDTH;
    __shared__ __align__(16) float ws[IC_MAXC * C];    // [ci][o]
    __shared__ __align__(16) float xs[IC_MAXC * IC_PX];
    __shared__ __align__(16) float red[2 * 8 * C];
    const int n = blockIdx.y, tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    for (int i = tid; i < Cin * C; i += 256) { const int ci = i / C, o = i % C; ws[i] = w[o * Cin + ci]; }
    const float4 b4 = ld4(bias + lane * 4);
    Coef k[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) k[i] = coef[(size_t)n * C + lane * 4 + i];
    float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
    const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, P);
    float xr[IC_MAXC];            // this thread's pixel of the NEXT tile, all input channels (in flight during the compute)
    auto prefetch = [&](int pb) {
#pragma unroll
        for (int ci = 0; ci < IC_MAXC; ++ci) xr[ci] = (ci < Cin && pb + tid < p1) ? x[((size_t)n * Cin + ci) * P + pb + tid] : 0.f;
    };
    prefetch(p0);
    for (int pb = p0; pb < p1; pb += IC_PX) {
        __syncthreads();
#pragma unroll
        for (int ci = 0; ci < IC_MAXC; ++ci) xs[ci * IC_PX + tid] = xr[ci];
        __syncthreads();
        if (pb + IC_PX < p1) prefetch(pb + IC_PX);
        // each warp owns 32 consecutive pixels of the tile, processed in groups of 4 (4-way ILP; the x values of a
        // group are one broadcast LDS.128 per input channel, the weights one LDS.128 per lane)
#pragma unroll 1
        for (int g4 = 0; g4 < IC_GROUPS; ++g4) {
            const int px0 = warp * (IC_PX / 8) + g4 * 4;
            if (pb + px0 >= p1) break;
            float4 acc[4] = {b4, b4, b4, b4};
            for (int ci = 0; ci < Cin; ++ci) {
                const float4 xv = ld4(xs + ci * IC_PX + px0);
                const float4 wv = ld4(ws + ci * C + lane * 4);
                const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[j].x = fmaf(xx[j], wv.x, acc[j].x); acc[j].y = fmaf(xx[j], wv.y, acc[j].y);
                    acc[j].z = fmaf(xx[j], wv.z, acc[j].z); acc[j].w = fmaf(xx[j], wv.w, acc[j].w);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int px = px0 + j;
                if (pb + px >= p1) continue;
                const size_t row = (size_t)n * P + pb + px;
                float4 o;
                o.x = fmaxf(fmaf(acc[j].x, k[0].scale, k[0].shift), 0.f);
                o.y = fmaxf(fmaf(acc[j].y, k[1].scale, k[1].shift), 0.f);
                o.z = fmaxf(fmaf(acc[j].z, k[2].scale, k[2].shift), 0.f);
                o.w = fmaxf(fmaf(acc[j].w, k[3].scale, k[3].shift), 0.f);
                st4(x0 + row * C + lane * 4, o);
                s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
                q.x += o.x * o.x; q.y += o.y * o.y; q.z += o.z * o.z; q.w += o.w * o.w;
            }
        }
    }
    // reduce over the 8 warps (thread layout c4 = lane, row = warp)
    float4* r4 = reinterpret_cast<float4*>(red);
    __syncthreads();
    r4[warp * 32 + lane] = s;
    r4[256 + warp * 32 + lane] = q;
    __syncthreads();
    {
        const int which = tid / C, ch = tid % C;
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += (double)red[which * 8 * C + r * C + ch];
        atomicAdd(&stats[((size_t)n * C + ch) * 2 + which], t);
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Moment form of the in_conv statistics.  c0 = W x + b is affine in the 15-channel input, so every full-tensor
// statistic of the 128-channel conv output follows from per-frame moments of the INPUT:
//   sum_p c0[o]       = W[o].S1 + P b[o]                       S1[ci]     = sum_p x[ci]
//   sum_p c0[o]^2     = W[o]^T S2 W[o] + 2 b[o] W[o].S1 + P b[o]^2   S2[ci][cj] = sum_p x[ci] x[cj]
// (135 FMAs per pixel instead of the 1920 of recomputing the convolution), and in backward, with dgn = dX0 * [x0 > 0],
//   sum_p dgn[o] c0[o]        = W[o].G[o] + b[o] g0[o]         G[o][ci]   = sum_p dgn[o] x[ci],  g0[o] = sum_p dgn[o]
//   sum_p dc0[o] x[ci]        = a G[o][ci] + b (W[o].S2[:,ci] + b[o] S1[ci]) + c S1[ci]     (dc0 = a dgn + b c0 + c)
// so ONE pass over dX0 (gram kernel) replaces the statistics pass and the weight-gradient pass; the small per-frame
// algebra runs in fp64 finalize kernels.
// ------------------------------------------------------------------------------------------------------------------
constexpr int IM_TRI = IC_MAXC * (IC_MAXC + 1) / 2;      // 136 upper-triangle entries
constexpr int IM_VALS = IC_MAXC + IM_TRI;                // per-thread accumulators
constexpr int IM_STRIDE = IC_MAXC + IC_MAXC * IC_MAXC;   // doubles per frame: S1[16], S2[16][16]
constexpr int IG_STRIDE = IC_MAXC + 1;                   // doubles per (frame, o): G[16], g0

__global__ void __launch_bounds__(256) inconv_moments_kernel(const float* __restrict__ x /* [N][Cin][P] */, double* mom, int* notpad,
                                                              float pad_value, int Cin, int P, int chunk) {
    __shared__ float red[8 * IM_VALS];
    const int n = blockIdx.y, tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    float acc[IM_VALS];
#pragma unroll
    for (int i = 0; i < IM_VALS; ++i) acc[i] = 0.f;
    bool any_nonpad = false;
    const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, P);
    for (int p = p0 + tid; p < p1; p += 256) {
        float v[IC_MAXC];
#pragma unroll
        for (int ci = 0; ci < IC_MAXC; ++ci) {
            v[ci] = ci < Cin ? x[((size_t)n * Cin + ci) * P + p] : 0.f;
            if (ci < Cin && !(v[ci] == pad_value)) any_nonpad = true;
        }
        int k = IC_MAXC;
#pragma unroll
        for (int ci = 0; ci < IC_MAXC; ++ci) {
            acc[ci] += v[ci];
#pragma unroll
            for (int cj = ci; cj < IC_MAXC; ++cj) { acc[k] = fmaf(v[ci], v[cj], acc[k]); ++k; }
        }
    }
    if (any_nonpad) notpad[n] = 1;
#pragma unroll
    for (int i = 0; i < IM_VALS; ++i) {
        const float t = warp_sum(acc[i]);
        if (lane == 0) red[warp * IM_VALS + i] = t;
    }
    __syncthreads();
    if (tid < IM_VALS) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += (double)red[w * IM_VALS + tid];
        double* dst = mom + (size_t)n * IM_STRIDE;
        if (tid < IC_MAXC) atomicAdd(&dst[tid], t);
        else {
            int k = tid - IC_MAXC, ci = 0;                 // upper-triangle index -> (ci, cj)
            while (k >= IC_MAXC - ci) { k -= IC_MAXC - ci; ++ci; }
            const int cj = ci + k;
            atomicAdd(&dst[IC_MAXC + ci * IC_MAXC + cj], t);
            if (cj != ci) atomicAdd(&dst[IC_MAXC + cj * IC_MAXC + ci], t);
        }
    }
}

// stats[n][o] = (sum_p c0, sum_p c0^2) from the input moments; grid N, 128 threads
__global__ void inconv_stats_from_moments_kernel(const double* __restrict__ mom, const float* __restrict__ w, const float* __restrict__ bias,
                                                 double* stats, int Cin, double P) {
    const int n = blockIdx.x, o = threadIdx.x;
    const double* S1 = mom + (size_t)n * IM_STRIDE;
    const double* S2 = S1 + IC_MAXC;
    double wv[IC_MAXC];
    for (int ci = 0; ci < IC_MAXC; ++ci) wv[ci] = ci < Cin ? (double)w[o * Cin + ci] : 0.0;
    const double b = bias[o];
    double lin = 0.0, quad = 0.0;
    for (int ci = 0; ci < Cin; ++ci) {
        lin += wv[ci] * S1[ci];
        double r = 0.0;
        for (int cj = 0; cj < Cin; ++cj) r += wv[cj] * S2[ci * IC_MAXC + cj];
        quad += wv[ci] * r;
    }
    stats[((size_t)n * UB_WIDTH + o) * 2 + 0] = lin + P * b;
    stats[((size_t)n * UB_WIDTH + o) * 2 + 1] = quad + 2.0 * b * lin + P * b * b;
}

// gram pass of the backward: gacc[n][o][ci] += sum_p dgn[p][o] x[p][ci], gacc[n][o][16] += sum_p dgn[p][o],
// dgn = dX0 * [x0 > 0]: either read off the saved forward output x0, or (x0 == nullptr) already applied by the encoder block's
// residual_bwd, which loads x0 anyway.  lane = 4 output channels, warp = pixels.
__global__ void __launch_bounds__(256, 2) inconv_bwd_gram_kernel(const float* __restrict__ x /* [N][Cin][P] */, const float* __restrict__ x0,
                                                               const float* __restrict__ dx0, double* gacc, int Cin, int P, int chunk) {
    constexpr int C = UB_WIDTH;
    __shared__ __align__(16) float xs[IC_MAXC * IC_PX];
    __shared__ __align__(16) float red[8 * C];
    const int n = blockIdx.y, tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    float4 s = make_float4(0, 0, 0, 0);
    float4 gw[IC_MAXC];
#pragma unroll
    for (int ci = 0; ci < IC_MAXC; ++ci) gw[ci] = make_float4(0, 0, 0, 0);
    const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, P);
    float xr[IC_MAXC];
    auto prefetch = [&](int pb) {
#pragma unroll
        for (int ci = 0; ci < IC_MAXC; ++ci) xr[ci] = (ci < Cin && pb + tid < p1) ? x[((size_t)n * Cin + ci) * P + pb + tid] : 0.f;
    };
    prefetch(p0);
    for (int pb = p0; pb < p1; pb += IC_PX) {
        __syncthreads();
#pragma unroll
        for (int ci = 0; ci < IC_MAXC; ++ci) xs[ci * IC_PX + tid] = xr[ci];
        __syncthreads();
        if (pb + IC_PX < p1) prefetch(pb + IC_PX);
#pragma unroll 1
        for (int g4 = 0; g4 < IC_GROUPS; ++g4) {
            const int px0 = warp * (IC_PX / 8) + g4 * 4;
            float4 d[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                d[j] = make_float4(0, 0, 0, 0);
                if (pb + px0 + j < p1) {
                    const size_t row = ((size_t)n * P + pb + px0 + j) * C + lane * 4;
                    const float4 g = ld4_stream(dx0 + row);
                    if (x0 != nullptr) {          // dx0 not yet masked by the ReLU of in_conv
                        const float4 o = ld4_stream(x0 + row);
                        d[j] = make_float4(o.x > 0.f ? g.x : 0.f, o.y > 0.f ? g.y : 0.f, o.z > 0.f ? g.z : 0.f, o.w > 0.f ? g.w : 0.f);
                    } else {
                        d[j] = g;
                    }
                }
                s.x += d[j].x; s.y += d[j].y; s.z += d[j].z; s.w += d[j].w;
            }
#pragma unroll
            for (int ci = 0; ci < IC_MAXC; ++ci) {
                const float4 xv = ld4(xs + ci * IC_PX + px0);
                const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    gw[ci].x = fmaf(d[j].x, xx[j], gw[ci].x); gw[ci].y = fmaf(d[j].y, xx[j], gw[ci].y);
                    gw[ci].z = fmaf(d[j].z, xx[j], gw[ci].z); gw[ci].w = fmaf(d[j].w, xx[j], gw[ci].w);
                }
            }
        }
    }
    float4* r4 = reinterpret_cast<float4*>(red);
#pragma unroll
    for (int ci = 0; ci <= IC_MAXC; ++ci) {              // ci == IC_MAXC: the plain sum g0
        if (ci < IC_MAXC && ci >= Cin) continue;
        __syncthreads();
        r4[warp * 32 + lane] = ci < IC_MAXC ? gw[ci < IC_MAXC ? ci : 0] : s;
        __syncthreads();
        if (tid < C) {
            double t = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r) t += (double)red[r * C + tid];
            atomicAdd(&gacc[((size_t)n * C + tid) * IG_STRIDE + ci], t);
        }
    }
}

// bstats[n][o] = (sum dgn, sum dgn * c0_hat) from the gram sums; grid N, 128 threads
__global__ void inconv_bwd_stats_from_gram_kernel(const double* __restrict__ gacc, const float* __restrict__ w, const float* __restrict__ bias,
                                                  const MeanRstd* __restrict__ mr, double* bstats, int Cin) {
    const int n = blockIdx.x, o = threadIdx.x;
    const double* G = gacc + ((size_t)n * UB_WIDTH + o) * IG_STRIDE;
    const double g0 = G[IC_MAXC];
    double dot = 0.0;
    for (int ci = 0; ci < Cin; ++ci) dot += (double)w[o * Cin + ci] * G[ci];
    const MeanRstd m = mr[(size_t)n * UB_WIDTH + o];
    bstats[((size_t)n * UB_WIDTH + o) * 2 + 0] = g0;
    bstats[((size_t)n * UB_WIDTH + o) * 2 + 1] = (double)m.rstd * (dot + (double)bias[o] * g0 - (double)m.mean * g0);
}

// dW[o][ci] += sum_n a G + b (W[o].S2[:,ci] + bias S1[ci]) + c S1[ci];  db[o] += sum_n a g0 + b (W[o].S1 + P bias) + c P
// grid 128 (o), 8 warps: lane ci < Cin -> dW column, lane 31 -> db; warp w takes frames w, w+8, ... (a frame is a chain of ~17
// dependent fp64 operations behind three dependent loads: one warp walking all N frames took 0.18 ms at N=48), fp64 partials
// are combined through shared memory in a fixed order.
constexpr int IF_WARPS = 8;
__global__ void __launch_bounds__(32 * IF_WARPS) inconv_bwd_finish_kernel(const double* __restrict__ gacc, const double* __restrict__ mom,
                                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                                         const BCoef* __restrict__ bc, float* dw, float* db, int N,
                                                                         int Cin, double P) {
    __shared__ double part[IF_WARPS][32];
    const int o = blockIdx.x, lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    double wv[IC_MAXC];
    for (int ci = 0; ci < IC_MAXC; ++ci) wv[ci] = ci < Cin ? (double)w[o * Cin + ci] : 0.0;
    const double b0 = bias[o];
    double acc = 0.0;
    if (lane < Cin) {
        for (int n = warp; n < N; n += IF_WARPS) {
            const BCoef k = bc[(size_t)n * UB_WIDTH + o];
            const double* S1 = mom + (size_t)n * IM_STRIDE;
            const double* S2 = S1 + IC_MAXC;
            double c0x = b0 * S1[lane];
            for (int cj = 0; cj < Cin; ++cj) c0x += wv[cj] * S2[cj * IC_MAXC + lane];
            acc += (double)k.a * gacc[((size_t)n * UB_WIDTH + o) * IG_STRIDE + lane] + (double)k.b * c0x + (double)k.c * S1[lane];
        }
    } else if (lane == 31) {
        for (int n = warp; n < N; n += IF_WARPS) {
            const BCoef k = bc[(size_t)n * UB_WIDTH + o];
            const double* S1 = mom + (size_t)n * IM_STRIDE;
            double lin = 0.0;
            for (int cj = 0; cj < Cin; ++cj) lin += wv[cj] * S1[cj];
            acc += (double)k.a * gacc[((size_t)n * UB_WIDTH + o) * IG_STRIDE + IC_MAXC] + (double)k.b * (lin + P * b0) + (double)k.c * P;
        }
    }
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < IF_WARPS; ++r) t += part[r][lane];
        if (lane < Cin) atomicAdd(&dw[o * Cin + lane], (float)t);
        else if (lane == 31) atomicAdd(&db[o], (float)t);
    }
}

static inline int ic_chunk(int P) { return P >= 8192 ? 4096 : (P >= 1024 ? 512 : 64); }

int launch_inconv_apply(const float* x, const float* w, const float* b, const Coef* coef, float* x0, double* stats_x0,
                        int N, int Cin, int P, cudaStream_t st) {
    if (Cin > IC_MAXC) return UB_ERR_ARG;
    const int chunk = ic_chunk(P);
    inconv_apply_kernel<<<dim3((P + chunk - 1) / chunk, N), 256, 0, st>>>(x, w, b, coef, x0, stats_x0, Cin, P, chunk);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

size_t inconv_moments_bytes(int N) { return (size_t)N * IM_STRIDE * sizeof(double); }
size_t inconv_gram_bytes(int N) { return (size_t)N * UB_WIDTH * IG_STRIDE * sizeof(double); }

// forward statistics through the input moments (mom must be zero on entry)
int launch_inconv_stats_moments(const float* x, const float* w, const float* b, double* mom, double* stats, int* notpad,
                                float pad_value, int N, int Cin, int P, cudaStream_t st) {
    if (Cin > IC_MAXC) return UB_ERR_ARG;
    const int chunk = P >= 8192 ? 8192 : P;
    inconv_moments_kernel<<<dim3((P + chunk - 1) / chunk, N), 256, 0, st>>>(x, mom, notpad, pad_value, Cin, P, chunk);
    UB_CHECK_LAUNCH();
    inconv_stats_from_moments_kernel<<<N, UB_WIDTH, 0, st>>>(mom, w, b, stats, Cin, (double)P);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
// backward, part 1: gram pass + GroupNorm-backward statistics (gacc must be zero on entry)
int launch_inconv_bwd_gram(const float* x, const float* x0, const float* dx0, const float* w, const float* b, const MeanRstd* mr,
                           double* gacc, double* bstats, int N, int Cin, int P, cudaStream_t st) {
    if (Cin > IC_MAXC) return UB_ERR_ARG;
    const int chunk = ic_chunk(P);
    inconv_bwd_gram_kernel<<<dim3((P + chunk - 1) / chunk, N), 256, 0, st>>>(x, x0, dx0, gacc, Cin, P, chunk);
    UB_CHECK_LAUNCH();
    inconv_bwd_stats_from_gram_kernel<<<N, UB_WIDTH, 0, st>>>(gacc, w, b, mr, bstats, Cin);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
// backward, part 2 (after norm_finalize_bwd produced bc): weight and bias gradients
int launch_inconv_bwd_finish(const double* gacc, const double* mom, const float* w, const float* b, const BCoef* bc, float* dw,
                             float* db, int N, int Cin, int P, cudaStream_t st) {
    if (Cin > IC_MAXC) return UB_ERR_ARG;
    inconv_bwd_finish_kernel<<<UB_WIDTH, 32 * IF_WARPS, 0, st>>>(gacc, mom, w, b, bc, dw, db, N, Cin, (double)P);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

}  // namespace ub
