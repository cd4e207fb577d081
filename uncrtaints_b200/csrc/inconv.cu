// in_conv of UnCRtainTS: Conv2d(C_in -> 128, k=1, bias) -> GroupNorm(4) -> ReLU
// (model/src/backbones/utae.py:453-497,500-520; built at uncrtaints.py:310-314), fused with the temporal
// pad-mask test (uncrtaints.py:392-394) and with the NCHW -> pixel-major layout change.
//
// K = C_in = 15 is far too thin for tensor cores, and the 128-channel conv output is 8.5x larger than the
// input, so the conv output is never written: pass A recomputes it for the GroupNorm statistics, pass B
// recomputes it again, normalises, applies ReLU and writes x0 (plus the column sums that the encoder
// block's PreNorm needs).  Backward mirrors this: pass 1 = GroupNorm-backward statistics, pass 2 = weight
// and bias gradients (in_conv has no input gradient).
#include "common.cuh"
#include "kernels.h"

namespace ub {

constexpr int IC_MAXC = 16;   // max input channels (13 optical + 2 SAR = 15)
constexpr int IC_PX = 64;     // pixels staged per iteration

// MODE 0: stats of c0 (+ pad test) | MODE 1: write x0 = relu(gn(c0)) + stats of x0
// MODE 2: bwd stats: (sum dgn, sum dgn*c0_hat), dgn = dX0 * [x0 > 0]
// MODE 3: bwd weight grads: dc0 = a*dgn + b*c0 + c; dW[o][ci] += dc0*x; db[o] += dc0
template <int MODE>
__global__ void __launch_bounds__(256) inconv_kernel(const float* __restrict__ x /* [N][Cin][P] */, const float* __restrict__ w /* [128][Cin] */,
                                                      const float* __restrict__ bias, const Coef* __restrict__ coef,
                                                      const MeanRstd* __restrict__ mr, const BCoef* __restrict__ bc,
                                                      const float* __restrict__ dx0, float* __restrict__ x0, double* stats,
                                                      int* notpad, float pad_value, float* dw, float* db, int Cin, int P,
                                                      int chunk) {
    constexpr int C = UB_WIDTH;
    __shared__ __align__(16) float ws[IC_MAXC * C];    // [ci][o]
    __shared__ __align__(16) float xs[IC_MAXC * IC_PX];
    __shared__ __align__(16) float red[2 * 8 * C];
    const int n = blockIdx.y, tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    for (int i = tid; i < Cin * C; i += 256) { const int ci = i / C, o = i % C; ws[i] = w[o * Cin + ci]; }
    const float4 b4 = ld4(bias + lane * 4);
    Coef k[4]; MeanRstd m[4]; BCoef bk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (MODE >= 1) k[i] = coef[(size_t)n * C + lane * 4 + i];
        if (MODE >= 2) m[i] = mr[(size_t)n * C + lane * 4 + i];
        if (MODE == 3) bk[i] = bc[(size_t)n * C + lane * 4 + i];
    }
    float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
    float4 gw[MODE == 3 ? IC_MAXC : 1];
    if (MODE == 3) {
#pragma unroll
        for (int ci = 0; ci < IC_MAXC; ++ci) gw[MODE == 3 ? ci : 0] = make_float4(0, 0, 0, 0);
    }
    bool any_nonpad = false;
    const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, P);
    for (int pb = p0; pb < p1; pb += IC_PX) {
        __syncthreads();
        for (int i = tid; i < Cin * IC_PX; i += 256) {
            const int ci = i / IC_PX, px = i % IC_PX;
            const float v = (pb + px < p1) ? x[((size_t)n * Cin + ci) * P + pb + px] : 0.f;
            xs[i] = v;
            if (MODE == 0 && pb + px < p1 && !(v == pad_value)) any_nonpad = true;
        }
        __syncthreads();
        // each warp owns 8 consecutive pixels of the tile, processed as two groups of 4 (4-way ILP; the x values of a
        // group are one broadcast LDS.128 per input channel, the weights one LDS.128 per lane)
        for (int g4 = 0; g4 < 2; ++g4) {
            const int px0 = warp * 8 + g4 * 4;
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
            float dcs[4][4];       // MODE 3: dc0 of the 4 pixels
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int px = px0 + j;
                const bool live = pb + px < p1;
                const size_t row = (size_t)n * P + pb + px;
                const float cv[4] = {acc[j].x, acc[j].y, acc[j].z, acc[j].w};
                if (MODE == 0) {
                    if (live) {
                        s.x += cv[0]; s.y += cv[1]; s.z += cv[2]; s.w += cv[3];
                        q.x += cv[0] * cv[0]; q.y += cv[1] * cv[1]; q.z += cv[2] * cv[2]; q.w += cv[3] * cv[3];
                    }
                } else if (MODE == 1) {
                    float4 o;
                    o.x = fmaxf(fmaf(cv[0], k[0].scale, k[0].shift), 0.f);
                    o.y = fmaxf(fmaf(cv[1], k[1].scale, k[1].shift), 0.f);
                    o.z = fmaxf(fmaf(cv[2], k[2].scale, k[2].shift), 0.f);
                    o.w = fmaxf(fmaf(cv[3], k[3].scale, k[3].shift), 0.f);
                    if (live) {
                        st4(x0 + row * C + lane * 4, o);
                        s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
                        q.x += o.x * o.x; q.y += o.y * o.y; q.z += o.z * o.z; q.w += o.w * o.w;
                    }
                } else {
                    float4 g = make_float4(0, 0, 0, 0);
                    if (live) g = ld4_stream(dx0 + row * C + lane * 4);
                    const float gv[4] = {g.x, g.y, g.z, g.w};
                    float dgn[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) dgn[i] = fmaf(cv[i], k[i].scale, k[i].shift) > 0.f ? gv[i] : 0.f;
                    if (MODE == 2) {
                        if (live) {
                            s.x += dgn[0]; s.y += dgn[1]; s.z += dgn[2]; s.w += dgn[3];
                            q.x += dgn[0] * (cv[0] - m[0].mean) * m[0].rstd;
                            q.y += dgn[1] * (cv[1] - m[1].mean) * m[1].rstd;
                            q.z += dgn[2] * (cv[2] - m[2].mean) * m[2].rstd;
                            q.w += dgn[3] * (cv[3] - m[3].mean) * m[3].rstd;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) dcs[j][i] = live ? fmaf(bk[i].a, dgn[i], fmaf(bk[i].b, cv[i], bk[i].c)) : 0.f;
                        s.x += dcs[j][0]; s.y += dcs[j][1]; s.z += dcs[j][2]; s.w += dcs[j][3];
                    }
                }
            }
            if (MODE == 3) {
#pragma unroll
                for (int ci = 0; ci < IC_MAXC; ++ci) {
                    if (ci < Cin) {
                        const float4 xv = ld4(xs + ci * IC_PX + px0);
                        const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
                        float4& a = gw[MODE == 3 ? ci : 0];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            a.x = fmaf(dcs[j][0], xx[j], a.x); a.y = fmaf(dcs[j][1], xx[j], a.y);
                            a.z = fmaf(dcs[j][2], xx[j], a.z); a.w = fmaf(dcs[j][3], xx[j], a.w);
                        }
                    }
                }
            }
        }
    }
    if (MODE == 0 && any_nonpad) notpad[n] = 1;
    // reduce over the 8 warps (thread layout c4 = lane, row = warp)
    float4* r4 = reinterpret_cast<float4*>(red);
    if (MODE != 3) {
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
    } else {
        // bias gradient, then one input channel at a time
        for (int ci = -1; ci < Cin; ++ci) {
            __syncthreads();
            float4 v = s;
#pragma unroll
            for (int j = 0; j < IC_MAXC; ++j) if (j == ci) v = gw[MODE == 3 ? j : 0];
            r4[warp * 32 + lane] = v;
            __syncthreads();
            if (tid < C) {
                float t = 0.f;
#pragma unroll
                for (int r = 0; r < 8; ++r) t += red[r * C + tid];
                if (ci < 0) atomicAdd(&db[tid], t);
                else atomicAdd(&dw[tid * Cin + ci], t);
            }
        }
    }
}

static inline int ic_chunk(int P) { return P >= 8192 ? 4096 : (P >= 1024 ? 512 : 64); }

int launch_inconv_stats(const float* x, const float* w, const float* b, double* stats, int* notpad, float pad_value,
                        int N, int Cin, int P, cudaStream_t st) {
    if (Cin > IC_MAXC) return UB_ERR_ARG;
    const int chunk = ic_chunk(P);
    inconv_kernel<0><<<dim3((P + chunk - 1) / chunk, N), 256, 0, st>>>(x, w, b, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                     stats, notpad, pad_value, nullptr, nullptr, Cin, P, chunk);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_inconv_apply(const float* x, const float* w, const float* b, const Coef* coef, float* x0, double* stats_x0,
                        int N, int Cin, int P, cudaStream_t st) {
    if (Cin > IC_MAXC) return UB_ERR_ARG;
    const int chunk = ic_chunk(P);
    inconv_kernel<1><<<dim3((P + chunk - 1) / chunk, N), 256, 0, st>>>(x, w, b, coef, nullptr, nullptr, nullptr, x0, stats_x0,
                                                                     nullptr, 0.f, nullptr, nullptr, Cin, P, chunk);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_inconv_bwd_stats(const float* x, const float* w, const float* b, const Coef* coef, const MeanRstd* mr,
                            const float* dx0, double* bstats, int N, int Cin, int P, cudaStream_t st) {
    if (Cin > IC_MAXC) return UB_ERR_ARG;
    const int chunk = ic_chunk(P);
    inconv_kernel<2><<<dim3((P + chunk - 1) / chunk, N), 256, 0, st>>>(x, w, b, coef, mr, nullptr, dx0, nullptr, bstats,
                                                                     nullptr, 0.f, nullptr, nullptr, Cin, P, chunk);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_inconv_bwd_wgrad(const float* x, const float* w, const float* b, const Coef* coef, const MeanRstd* mr,
                            const BCoef* bc, const float* dx0, float* dw, float* db, int N, int Cin, int P,
                            cudaStream_t st) {
    if (Cin > IC_MAXC) return UB_ERR_ARG;
    const int chunk = ic_chunk(P);
    inconv_kernel<3><<<dim3((P + chunk - 1) / chunk, N), 256, 0, st>>>(x, w, b, coef, mr, bc, dx0, nullptr, nullptr, nullptr,
                                                                     0.f, dw, db, Cin, P, chunk);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

}  // namespace ub
