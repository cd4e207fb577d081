// Output head and loss of UnCRtainTS.
//
//   out_conv + head : Conv2d(128 -> 13+covdim, k=1, bias) (model/src/backbones/uncrtaints.py:381,432), then
//                     mean = scale_by * sigmoid(.) (uncrtaints.py:384-385,441), var = softplus(beta=1, thr=20)(.) + eps
//                     (uncrtaints.py:223-228,444), concatenated as [B,1,13+covdim,H,W] NCHW (uncrtaints.py:436-446).
//   MGNLL           : multi_gaussian_nll_loss / multi_diag_gaussian_nll (model/src/losses.py:131-145,149-218), closed form:
//                     loss = 6.5*log(2*pi_f32) + 1/2 mean_hw[ sum_{b,c} log v ] + 1/2 mean_{b,hw}[ max(nan_to_num(sum_c e^2/v), 1e-9) ]
//                     (the log-determinant is summed over the batch, losses.py:138), v = max(var, eps) with identity gradient.
//   covariance      : diag_embed(v) -> [B,1,13,13,H,W] (losses.py:145,211).
//
// Both are HBM-bound byte shufflers over 26-channel planes: thread-per-pixel with plane-coalesced accesses.
#include "common.cuh"
#include "kernels.h"

namespace ub {

constexpr int HD_MAXO = 26;
constexpr int HD_PX = 128;
constexpr int HD_WJ = (HD_MAXO + 3) / 4;   // outputs per warp in the weight-gradient loop of head_bwd_kernel

// ------------------------------------------------------------------------------------------
// out_conv + head forward.  Persistent CTAs over 128-pixel tiles.  The decoder tile is staged by cp.async into a
// pitch-132 shared-memory tile (a lane's LDS.128 along k is conflict-free across the 8 pixels of a quarter warp); a thread
// owns ONE pixel and 13 outputs (half = tid / 128 selects outputs 0-12 or 13-25), reads 4 k-values of its pixel with one
// LDS.128 and the 13 weights of each k as 3 broadcast LDS.128 + 1 LDS.32 from a [k][32] table: 17 LDS per 52 FMA instead of
// the 8 LDS per 7 FMA of a pixel x group-of-7 mapping (ncu r01: mio_throttle + short_scoreboard 35 %, 0.58 ms at B=16).
// ------------------------------------------------------------------------------------------
constexpr int HF_PX = 128, HF_PITCH = UB_WIDTH + 4, HF_HALF = 13, HF_WROW = 32;
constexpr size_t HF_SMEM = (size_t)(HF_PX * HF_PITCH + UB_WIDTH * HF_WROW + 32) * sizeof(float);
__global__ void __launch_bounds__(256, 2) head_fwd_kernel(const float* __restrict__ dec /* [B][P][128] */, const float* __restrict__ w /* [O][128] */,
                                                           const float* __restrict__ bias, float* __restrict__ out /* [B][O][P] */,
                                                           int O, int P, long long total_tiles, float scale_by, int mean_sigmoid,
                                                           float var_eps) {
    constexpr int C = UB_WIDTH;
    extern __shared__ __align__(16) float fsm[];
    float* as = fsm;                                  // [128 px][132]
    float* ws = fsm + HF_PX * HF_PITCH;               // [128 k][32]: outputs 0-12 at columns 0-12, outputs 13-25 at columns 16-28
    float* bs = ws + C * HF_WROW;                     // [32] bias in the same column order
    const int tid = threadIdx.x, px = tid % HF_PX, half = tid / HF_PX;
    for (int i = tid; i < C * HF_WROW; i += 256) {
        const int k = i / HF_WROW, col = i % HF_WROW, o = (col / 16) * HF_HALF + (col % 16);
        ws[i] = ((col % 16) < HF_HALF && o < O) ? w[o * C + k] : 0.f;
    }
    if (tid < HF_WROW) {
        const int o = (tid / 16) * HF_HALF + (tid % 16);
        bs[tid] = ((tid % 16) < HF_HALF && o < O) ? bias[o] : 0.f;
    }
    const int tiles_per_img = P / HF_PX;
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int b = (int)(t / tiles_per_img), p0 = (int)(t % tiles_per_img) * HF_PX;
        __syncthreads();                              // everybody is done with the previous tile (and the tables are written)
        for (int i = tid; i < HF_PX * (C / 4); i += 256) {
            const int r = i / (C / 4), k4 = i % (C / 4);
            const unsigned d = (unsigned)__cvta_generic_to_shared(as + r * HF_PITCH + k4 * 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(dec + ((size_t)b * P + p0 + r) * C + k4 * 4) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        float acc[HF_HALF];
#pragma unroll
        for (int j = 0; j < HF_HALF; ++j) acc[j] = bs[half * 16 + j];
        const float* ap = as + px * HF_PITCH;
        const float* wp = ws + half * 16;
#pragma unroll 2
        for (int k4 = 0; k4 < C / 4; ++k4) {
            const float4 a4 = ld4(ap + k4 * 4);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float* wr = wp + (k4 * 4 + kk) * HF_WROW;
                const float4 w0 = ld4(wr), w1 = ld4(wr + 4), w2 = ld4(wr + 8);
                const float w3 = wr[12];
                const float a = av[kk];
                acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]); acc[2] = fmaf(a, w0.z, acc[2]); acc[3] = fmaf(a, w0.w, acc[3]);
                acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]); acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
                acc[8] = fmaf(a, w2.x, acc[8]); acc[9] = fmaf(a, w2.y, acc[9]); acc[10] = fmaf(a, w2.z, acc[10]); acc[11] = fmaf(a, w2.w, acc[11]);
                acc[12] = fmaf(a, w3, acc[12]);
            }
        }
#pragma unroll
        for (int j = 0; j < HF_HALF; ++j) {
            const int o = half * HF_HALF + j;
            if (o < O) {
                float v = acc[j];
                if (o < UB_S2) { if (mean_sigmoid) v = scale_by * sigmoid_f(v); }
                else v = (v > 20.f ? v : log1pf(expf(v))) + var_eps;
                out[((size_t)b * O + o) * P + p0 + px] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// head backward.  do = dOut * f'(out) from the saved output; dDec = do * W; dW += do^T dec; db += sum do.
// Persistent CTAs (2 per SM).  Per 128-pixel tile: the decoder-output tile is staged in shared memory by cp.async
// while the 26 x 128 `do` tile is built; dDec is produced one pixel per warp iteration; the weight gradient is split by
// output over the warps (warp w owns outputs w, w+8, w+16, w+24: 4 float4 register accumulators, flushed once per CTA).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) head_bwd_kernel(const float* __restrict__ dout /* [B][O][P] */, const float* __restrict__ out,
                                                          const float* __restrict__ dec, const float* __restrict__ w,
                                                          float* __restrict__ ddec, float* dw, float* db, int O, int P,
                                                          long long total_tiles, float scale_by, int mean_sigmoid, float var_eps) {
    constexpr int C = UB_WIDTH;
    extern __shared__ __align__(16) float hsm[];
    float* dect = hsm;                         // [128 px][128 ch]
    float* ws = hsm + HD_PX * C;               // [O][128]
    float* dos = ws + HD_MAXO * C;             // [O][128 px]
    const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    for (int i = tid; i < O * C; i += 256) ws[i] = w[i];
    float4 gw[HD_WJ];
    float gbk[HD_MAXO / 2];   // bias-gradient partials: in the fill loop thread `tid` always sees outputs o = 2k + tid/128
#pragma unroll
    for (int j = 0; j < HD_WJ; ++j) gw[j] = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < HD_MAXO / 2; ++k) gbk[k] = 0.f;
    const int tiles_per_img = P / HD_PX;
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int b = (int)(t / tiles_per_img), p0 = (int)(t % tiles_per_img) * HD_PX;
        const size_t row0 = (size_t)b * P + p0;
        __syncthreads();
        for (int i = tid; i < HD_PX * (C / 4); i += 256) {       // contiguous 64 KB
            const unsigned d = (unsigned)__cvta_generic_to_shared(dect + i * 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(dec + row0 * C + (size_t)i * 4) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll
        for (int k = 0; k < HD_MAXO / 2; ++k) {
            const int i = tid + 256 * k, o = i / HD_PX, px = i % HD_PX;     // HD_PX == 128: o = 2k + tid/128
            if (o < O) {
                const size_t g = ((size_t)b * O + o) * P + p0 + px;
                const float d = dout[g], v = out[g];
                float der;
                if (o < UB_S2) {
                    if (mean_sigmoid) { const float sg = v / scale_by; der = scale_by * sg * (1.f - sg); }
                    else der = 1.f;
                } else {
                    der = 1.f - expf(-(v - var_eps));   // softplus'(x) = sigmoid(x) = 1 - exp(-softplus(x)); == 1 beyond the threshold in fp32
                }
                dos[i] = d * der;
                gbk[k] += d * der;
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // input gradient: 8 consecutive pixels per warp iteration, lane = 4 channels: per output two broadcast LDS.128 of the 8
        // `do` values and one LDS.128 of the weights feed 32 FMAs (4 pixels: 2 LDS per 16 FMAs, the shared-memory pipe as busy
        // as the FMA pipe; 1 pixel: 2 LDS per 4 FMAs, mio_throttle + short_scoreboard 41 %)
        for (int px0 = warp * 8; px0 < HD_PX; px0 += 64) {
            float4 acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = make_float4(0, 0, 0, 0);
#pragma unroll
            for (int o = 0; o < HD_MAXO; ++o) {
                if (o < O) {
                    const float4 d4 = ld4(dos + o * HD_PX + px0), e4 = ld4(dos + o * HD_PX + px0 + 4);
                    const float4 wv = ld4(ws + o * C + lane * 4);
                    const float dd[8] = {d4.x, d4.y, d4.z, d4.w, e4.x, e4.y, e4.z, e4.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        acc[i].x = fmaf(dd[i], wv.x, acc[i].x); acc[i].y = fmaf(dd[i], wv.y, acc[i].y);
                        acc[i].z = fmaf(dd[i], wv.z, acc[i].z); acc[i].w = fmaf(dd[i], wv.w, acc[i].w);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) st4(ddec + (row0 + px0 + i) * C + lane * 4, acc[i]);
        }
        // weight gradient: warp owns outputs (warp % 4) + 4j, j < 7, over one HALF of the tile's pixels (warp / 4): a warp reads half of
        // the staged decoder tile for 7 outputs (all of it for 4 outputs before: the tile reads were 80 % of this loop's LDS traffic)
#pragma unroll 2
        for (int px0 = (warp >> 2) * (HD_PX / 2); px0 < (warp >> 2) * (HD_PX / 2) + HD_PX / 2; px0 += 4) {
            float4 a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = ld4(dect + (px0 + i) * C + lane * 4);
#pragma unroll
            for (int j = 0; j < HD_WJ; ++j) {
                const int o = (warp & 3) + 4 * j;
                if (o < O) {
                    const float4 d4 = ld4(dos + o * HD_PX + px0);
                    const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        gw[j].x = fmaf(dd[i], a[i].x, gw[j].x); gw[j].y = fmaf(dd[i], a[i].y, gw[j].y);
                        gw[j].z = fmaf(dd[i], a[i].z, gw[j].z); gw[j].w = fmaf(dd[i], a[i].w, gw[j].w);
                    }
                }
            }
        }
    }
    // bias gradient: warp-reduce the per-thread partials (a warp lies inside one half of the 256 threads => one output)
#pragma unroll
    for (int k = 0; k < HD_MAXO / 2; ++k) {
        const float t = warp_sum(gbk[k]);
        const int o = 2 * k + tid / HD_PX;
        if (lane == 0 && o < O) atomicAdd(&db[o], t);
    }
#pragma unroll
    for (int j = 0; j < HD_WJ; ++j) {
        const int o = (warp & 3) + 4 * j;
        if (o < O) {
            atomicAdd(&dw[o * C + lane * 4 + 0], gw[j].x); atomicAdd(&dw[o * C + lane * 4 + 1], gw[j].y);
            atomicAdd(&dw[o * C + lane * 4 + 2], gw[j].z); atomicAdd(&dw[o * C + lane * 4 + 3], gw[j].w);
        }
    }
}

// ------------------------------------------------------------------------------------------
// MGNLL forward (+ gradients).  Thread per pixel, loops over batch and channel; planes are coalesced.
// pred/var/target are addressed as base + b*stride_b + c*P + p (slices of the [B,1,26,H,W] network output).
// acc[0] += sum_p sum_{b,c} log v ; acc[1] += sum_p sum_b maha_b ; flag |= any(var < 0)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mgnll_fwd_kernel(const float* __restrict__ pred, long long pred_sb,
                                                         const float* __restrict__ target, long long targ_sb,
                                                         const float* __restrict__ var, long long var_sb, int var_ch,
                                                         float* __restrict__ dpred /* [B][13][P] or null */,
                                                         float* __restrict__ dvar /* [B][var_ch][P] or null */, double* acc,
                                                         int* neg_flag, int B, int P, float eps) {
    // one thread per (sample, pixel): every reduction of the closed form is a plain global sum, so the batch needs no loop
    const int p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    float logdet = 0.f, maha_sum = 0.f;
    bool neg = false;
    if (p < P) {
        const float inv_p = 1.0f / (float)P, inv_bp = 1.0f / ((float)B * (float)P);
        float e[UB_S2], iv[UB_S2], vr[UB_S2];
#pragma unroll
        for (int c = 0; c < UB_S2; ++c) {         // all loads first: 27 (diag) / 15 (iso) independent requests in flight
            if (var_ch != 1 || c == 0) vr[c] = var[b * var_sb + (size_t)c * P + p];
            e[c] = pred[b * pred_sb + (size_t)c * P + p] - target[b * targ_sb + (size_t)c * P + p];
        }
        float maha = 0.f;
#pragma unroll
        for (int c = 0; c < UB_S2; ++c) {
            const float vraw = var_ch == 1 ? vr[0] : vr[c];
            neg |= (vraw < 0.f);
            const float v = clamp_min_keep_nan(vraw, eps);
            iv[c] = 1.0f / v;
            logdet += logf(v);
            maha = fmaf(e[c] * e[c], iv[c], maha);
        }
        const bool isn = maha != maha;
        float m = isn ? 0.f : maha;                       // nan_to_num
        m = fminf(m, 3.4028234663852886e38f);             // +inf -> float max
        const bool live = !isn && m >= 1e-9f && maha <= 3.4028234663852886e38f;   // clamp(min=1e-9) / inf: no gradient
        maha_sum = fmaxf(m, 1e-9f);
        if (dpred) {
            float dv_iso = 0.f;
#pragma unroll
            for (int c = 0; c < UB_S2; ++c) {
                const float gm = live ? e[c] * iv[c] * inv_bp : 0.f;
                dpred[((size_t)b * UB_S2 + c) * P + p] = gm;
                const float gv = 0.5f * iv[c] * inv_p - (live ? 0.5f * e[c] * e[c] * iv[c] * iv[c] * inv_bp : 0.f);
                if (var_ch == 1) dv_iso += gv;
                else dvar[((size_t)b * UB_S2 + c) * P + p] = gv;
            }
            if (var_ch == 1) dvar[(size_t)b * P + p] = dv_iso;
        }
    }
    double l = warp_sum_d((double)logdet), m = warp_sum_d((double)maha_sum);
    __shared__ double sl[8], sm[8];
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    if (lane == 0) { sl[warp] = l; sm[warp] = m; }
    if (neg) *neg_flag = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int i = 0; i < 8; ++i) { a += sl[i]; c += sm[i]; }
        atomicAdd(&acc[0], a);
        atomicAdd(&acc[1], c);
    }
}

// ------------------------------------------------------------------------------------------
// GNLL (gaussian_nll_loss, model/src/losses.py:46-128; `--loss GNLL`, used with covmode 'uni'): element-wise
//   loss = mean over all elements of 1/2 (log v + e^2 / v) [+ 1/2 log(2 pi) if full],  v = max(var, eps) with identity gradient.
// One thread per (sample, pixel) over the 13 channel planes, like the MGNLL kernel; also writes the clamped variance (the
// loss's second return value, losses.py:122-128).  acc[0] += sum of 1/2 (log v + e^2 / v).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gnll_fwd_kernel(const float* __restrict__ pred, long long pred_sb, const float* __restrict__ target,
                                                        long long targ_sb, const float* __restrict__ var, long long var_sb,
                                                        float* __restrict__ dpred /* [B][13][P] or null */, float* __restrict__ dvar,
                                                        float* __restrict__ var_out /* [B][13][P] or null */, double* acc, int* neg_flag,
                                                        int B, int P, float eps) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    float part = 0.f;
    bool neg = false;
    if (p < P) {
        const float inv_n = 1.0f / ((float)B * (float)UB_S2 * (float)P);
        float e[UB_S2], vr[UB_S2];
#pragma unroll
        for (int c = 0; c < UB_S2; ++c) {
            vr[c] = var[b * var_sb + (size_t)c * P + p];
            e[c] = pred[b * pred_sb + (size_t)c * P + p] - target[b * targ_sb + (size_t)c * P + p];
        }
#pragma unroll
        for (int c = 0; c < UB_S2; ++c) {
            neg |= (vr[c] < 0.f);
            const float v = clamp_min_keep_nan(vr[c], eps), iv = 1.0f / v;
            part += 0.5f * (logf(v) + e[c] * e[c] * iv);
            if (var_out) var_out[((size_t)b * UB_S2 + c) * P + p] = v;
            if (dpred) {
                dpred[((size_t)b * UB_S2 + c) * P + p] = e[c] * iv * inv_n;
                dvar[((size_t)b * UB_S2 + c) * P + p] = 0.5f * (iv - e[c] * e[c] * iv * iv) * inv_n;
            }
        }
    }
    double l = warp_sum_d((double)part);
    __shared__ double sl[8];
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    if (lane == 0) sl[warp] = l;
    if (neg) *neg_flag = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int i = 0; i < 8; ++i) a += sl[i];
        atomicAdd(&acc[0], a);
    }
}
__global__ void gnll_finalize_kernel(const double* acc, float* loss, double count, int full) {
    const double cst = full ? 0.5 * 1.8378770664093453 : 0.0;       // 1/2 log(2 pi) (math.log, losses.py:118)
    *loss = (float)(acc[0] / count + cst);
}

__global__ void mgnll_finalize_kernel(const double* acc, float* loss, int B, int P) {
    // 13/2 * log(2*float32(pi)) evaluated in float32 (losses.py:143)
    const float cst = 6.5f * logf(2.0f * 3.14159274101257324f);
    const double v = (double)cst + 0.5 * acc[0] / (double)P + 0.5 * acc[1] / ((double)B * (double)P);
    *loss = (float)v;
}

// ------------------------------------------------------------------------------------------
// reduction='none' forms (losses.py:213-218 / :122-128): per-element losses and their backward for an arbitrary upstream
// gradient.  MGNLL: loss[p][b] (the nested vmap's output order [H][W][B]) = const + 1/2 sum_{b',c} log v[b'][c][p] + 1/2 maha[b][p];
// thread per pixel, loops over the batch (the log-determinant couples the samples of a pixel, losses.py:138).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mgnll_none_fwd_kernel(const float* __restrict__ pred, long long pred_sb, const float* __restrict__ target,
                                                              long long targ_sb, const float* __restrict__ var, long long var_sb, int var_ch,
                                                              float* __restrict__ loss /* [P][B] */, int* neg_flag, int B, int P, float eps) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float cst = 6.5f * logf(2.0f * 3.14159274101257324f);
    float logdet = 0.f;
    bool neg = false;
    for (int b = 0; b < B; ++b) {
        float maha = 0.f;
        for (int c = 0; c < UB_S2; ++c) {
            const float vraw = var[b * var_sb + (size_t)(var_ch == 1 ? 0 : c) * P + p];
            neg |= (vraw < 0.f);
            const float v = clamp_min_keep_nan(vraw, eps);
            const float e = pred[b * pred_sb + (size_t)c * P + p] - target[b * targ_sb + (size_t)c * P + p];
            logdet += logf(v);
            maha = fmaf(e * e, 1.0f / v, maha);
        }
        float m = maha != maha ? 0.f : maha;
        m = fminf(m, 3.4028234663852886e38f);
        loss[(size_t)p * B + b] = fmaxf(m, 1e-9f);
    }
    for (int b = 0; b < B; ++b) loss[(size_t)p * B + b] = cst + 0.5f * logdet + 0.5f * loss[(size_t)p * B + b];
    if (neg) *neg_flag = 1;
}
__global__ void __launch_bounds__(256) mgnll_none_bwd_kernel(const float* __restrict__ pred, long long pred_sb, const float* __restrict__ target,
                                                              long long targ_sb, const float* __restrict__ var, long long var_sb, int var_ch,
                                                              const float* __restrict__ g /* [P][B] */, float* __restrict__ dpred,
                                                              float* __restrict__ dvar, int B, int P, float eps) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float gsum = 0.f;
    for (int b = 0; b < B; ++b) gsum += g[(size_t)p * B + b];
    for (int b = 0; b < B; ++b) {
        const float gb = g[(size_t)p * B + b];
        float e[UB_S2], iv[UB_S2], maha = 0.f;
#pragma unroll
        for (int c = 0; c < UB_S2; ++c) {
            const float v = clamp_min_keep_nan(var[b * var_sb + (size_t)(var_ch == 1 ? 0 : c) * P + p], eps);
            e[c] = pred[b * pred_sb + (size_t)c * P + p] - target[b * targ_sb + (size_t)c * P + p];
            iv[c] = 1.0f / v;
            maha = fmaf(e[c] * e[c], iv[c], maha);
        }
        const bool live = maha == maha && maha >= 1e-9f && maha <= 3.4028234663852886e38f;
        float dv_iso = 0.f;
#pragma unroll
        for (int c = 0; c < UB_S2; ++c) {
            dpred[((size_t)b * UB_S2 + c) * P + p] = live ? gb * e[c] * iv[c] : 0.f;
            const float gv = 0.5f * iv[c] * gsum - (live ? 0.5f * gb * e[c] * e[c] * iv[c] * iv[c] : 0.f);
            if (var_ch == 1) dv_iso += gv;
            else dvar[((size_t)b * UB_S2 + c) * P + p] = gv;
        }
        if (var_ch == 1) dvar[(size_t)b * P + p] = dv_iso;
    }
}
// GNLL, reduction='none': loss[b][c][p] = 1/2 (log v + e^2 / v) [+ 1/2 log(2 pi)]; backward for an element-wise upstream gradient
__global__ void __launch_bounds__(256) gnll_none_kernel(const float* __restrict__ pred, long long pred_sb, const float* __restrict__ target,
                                                         long long targ_sb, const float* __restrict__ var, long long var_sb,
                                                         const float* __restrict__ g /* null: forward */, float* __restrict__ loss,
                                                         float* __restrict__ var_out, float* __restrict__ dpred, float* __restrict__ dvar,
                                                         int* neg_flag, int P, float eps, int full) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (p >= P) return;
    const float cst = full ? 0.5f * 1.8378770664093453f : 0.f;
    bool neg = false;
    for (int c = 0; c < UB_S2; ++c) {
        const size_t o = ((size_t)b * UB_S2 + c) * P + p;
        const float vraw = var[b * var_sb + (size_t)c * P + p];
        neg |= (vraw < 0.f);
        const float v = clamp_min_keep_nan(vraw, eps), iv = 1.0f / v;
        const float e = pred[b * pred_sb + (size_t)c * P + p] - target[b * targ_sb + (size_t)c * P + p];
        if (g) {
            dpred[o] = g[o] * e * iv;
            dvar[o] = g[o] * 0.5f * (iv - e * e * iv * iv);
        } else {
            loss[o] = 0.5f * (logf(v) + e * e * iv) + cst;
            if (var_out) var_out[o] = v;
        }
    }
    if (neg && neg_flag) *neg_flag = 1;
}

// out[i] = in[i] * g[0]   (chain rule with the upstream scalar gradient)
__global__ void scale_by_scalar_kernel(const float* __restrict__ in, const float* __restrict__ g, float* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] * g[0];
}

// covariance[b][0][i][j][p] = (i == j) ? max(var[b][i or 0][p], eps) : 0
__global__ void __launch_bounds__(256) covariance_kernel(const float* __restrict__ var, long long var_sb, int var_ch,
                                                          float* __restrict__ cov, int P, float eps) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (p >= P) return;
    float v[UB_S2];
#pragma unroll
    for (int c = 0; c < UB_S2; ++c) v[c] = clamp_min_keep_nan(var[b * var_sb + (size_t)(var_ch == 1 ? 0 : c) * P + p], eps);
    float* dst = cov + (size_t)b * UB_S2 * UB_S2 * P + p;
#pragma unroll
    for (int i = 0; i < UB_S2; ++i)
#pragma unroll
        for (int j = 0; j < UB_S2; ++j) dst[(size_t)(i * UB_S2 + j) * P] = (i == j) ? v[i] : 0.f;
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
int launch_head_fwd(const float* dec, const float* w, const float* bias, float* out, int B, int O, int P, float scale_by,
                    int mean_sigmoid, float var_eps, cudaStream_t st) {
    if (O > HD_MAXO || O < UB_S2 || P % HF_PX) return UB_ERR_ARG;
    UB_SET_SMEM(head_fwd_kernel, HF_SMEM);
    const long long tiles = (long long)B * (P / HF_PX);
    const int blocks = (int)(tiles < 2LL * device_sm_count() ? tiles : 2LL * device_sm_count());
    head_fwd_kernel<<<blocks, 256, HF_SMEM, st>>>(dec, w, bias, out, O, P, tiles, scale_by, mean_sigmoid, var_eps);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_head_bwd(const float* dout, const float* out, const float* dec, const float* w, float* ddec, float* dw, float* db,
                    int B, int O, int P, float scale_by, int mean_sigmoid, float var_eps, int num_sms, cudaStream_t st) {
    if (O > HD_MAXO || O < UB_S2 || P % HD_PX) return UB_ERR_ARG;
    constexpr size_t smem = (size_t)(HD_PX * UB_WIDTH + HD_MAXO * UB_WIDTH + HD_MAXO * HD_PX) * sizeof(float);
    UB_SET_SMEM(head_bwd_kernel, smem);
    const long long tiles = (long long)B * (P / HD_PX);
    const int blocks = (int)(tiles < 2LL * num_sms ? tiles : 2LL * num_sms);
    head_bwd_kernel<<<blocks, 256, smem, st>>>(dout, out, dec, w, ddec, dw, db, O, P, tiles, scale_by, mean_sigmoid, var_eps);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_mgnll(const float* pred, long long pred_sb, const float* target, long long targ_sb, const float* var,
                 long long var_sb, int var_ch, float* dpred, float* dvar, double* acc, int* neg_flag, float* loss, int B,
                 int P, float eps, cudaStream_t st) {
    if (cudaMemsetAsync(acc, 0, 2 * sizeof(double), st) != cudaSuccess) return UB_ERR_CUDA;
    if (cudaMemsetAsync(neg_flag, 0, sizeof(int), st) != cudaSuccess) return UB_ERR_CUDA;
    mgnll_fwd_kernel<<<dim3((P + 255) / 256, B), 256, 0, st>>>(pred, pred_sb, target, targ_sb, var, var_sb, var_ch, dpred, dvar, acc,
                                                      neg_flag, B, P, eps);
    UB_CHECK_LAUNCH();
    mgnll_finalize_kernel<<<1, 1, 0, st>>>(acc, loss, B, P);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_gnll(const float* pred, long long pred_sb, const float* target, long long targ_sb, const float* var, long long var_sb,
                float* dpred, float* dvar, float* var_out, double* acc, int* neg_flag, float* loss, int B, int P, float eps, int full,
                cudaStream_t st) {
    if (cudaMemsetAsync(acc, 0, 2 * sizeof(double), st) != cudaSuccess) return UB_ERR_CUDA;
    if (cudaMemsetAsync(neg_flag, 0, sizeof(int), st) != cudaSuccess) return UB_ERR_CUDA;
    gnll_fwd_kernel<<<dim3((P + 255) / 256, B), 256, 0, st>>>(pred, pred_sb, target, targ_sb, var, var_sb, dpred, dvar, var_out, acc,
                                                              neg_flag, B, P, eps);
    UB_CHECK_LAUNCH();
    gnll_finalize_kernel<<<1, 1, 0, st>>>(acc, loss, (double)B * UB_S2 * (double)P, full);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_mgnll_none(const float* pred, long long pred_sb, const float* target, long long targ_sb, const float* var, long long var_sb,
                      int var_ch, const float* g, float* loss, float* dpred, float* dvar, int* neg_flag, int B, int P, float eps,
                      cudaStream_t st) {
    if (!g) {
        if (cudaMemsetAsync(neg_flag, 0, sizeof(int), st) != cudaSuccess) return UB_ERR_CUDA;
        mgnll_none_fwd_kernel<<<(P + 255) / 256, 256, 0, st>>>(pred, pred_sb, target, targ_sb, var, var_sb, var_ch, loss, neg_flag, B, P, eps);
    } else {
        mgnll_none_bwd_kernel<<<(P + 255) / 256, 256, 0, st>>>(pred, pred_sb, target, targ_sb, var, var_sb, var_ch, g, dpred, dvar, B, P, eps);
    }
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_gnll_none(const float* pred, long long pred_sb, const float* target, long long targ_sb, const float* var, long long var_sb,
                     const float* g, float* loss, float* var_out, float* dpred, float* dvar, int* neg_flag, int B, int P, float eps,
                     int full, cudaStream_t st) {
    if (!g && cudaMemsetAsync(neg_flag, 0, sizeof(int), st) != cudaSuccess) return UB_ERR_CUDA;
    gnll_none_kernel<<<dim3((P + 255) / 256, B), 256, 0, st>>>(pred, pred_sb, target, targ_sb, var, var_sb, g, loss, var_out, dpred, dvar,
                                                               neg_flag, P, eps, full);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_scale_by_scalar(const float* in, const float* g, float* out, size_t n, cudaStream_t st) {
    scale_by_scalar_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, g, out, n);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_covariance(const float* var, long long var_sb, int var_ch, float* cov, int B, int P, float eps, cudaStream_t st) {
    covariance_kernel<<<dim3((P + 255) / 256, B), 256, 0, st>>>(var, var_sb, var_ch, cov, P, eps);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

}  // namespace ub
