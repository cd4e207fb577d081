// Image metrics of the validation / test loop on the GPU: img_metrics (model/src/learning/metrics.py:20-57)
// and the SSIM it calls (util/pytorch_ssim/__init__.py:17-37, 11x11 Gaussian window, sigma 1.5, zero padding).
//
// The reference evaluates them per sample with ~30 small torch kernels and nine `.cpu().numpy().item()` host syncs
// (train_reconstruct.py:318-353 loops over the batch).  Here a whole batch needs two kernels and one read-back:
//   img_stats_kernel : one thread per pixel over the 13 bands -- squared / absolute / signed error sums (plain and NaN-excluding,
//                      the reference mixes torch.mean and nanmean), spectral angle, variance sums, optional pixel-wise maps
//   ssim_kernel      : separable 11-tap Gaussian moments of (x, y, x^2, y^2, xy) on a 32x32 tile with a 5-pixel halo in shared memory
// Accumulators are fp64 atomics in acc[b][UB_MET_ACC].
#include "common.cuh"
#include "kernels.h"

namespace ub {

// acc layout per sample
enum { M_SE = 0, M_AE, M_SAM, M_NSE, M_NAE, M_NERR, M_NCNT, M_NVAR, M_NVCNT, M_SSIM, M_COUNT_ };
static_assert(M_COUNT_ <= UB_MET_ACC, "accumulator record too small");

__global__ void __launch_bounds__(256) img_stats_kernel(const float* __restrict__ target, const float* __restrict__ pred,
                                                         const float* __restrict__ var /* [B][13][P] or null */, double* acc,
                                                         float* __restrict__ pix /* [B][4][P] (error, ae, se, var) or null */, int P) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (p < P) {
        float tp = 0.f, tt = 0.f, pp = 0.f;
        float se = 0.f, ae = 0.f, nse = 0.f, nae = 0.f, nerr = 0.f, nv = 0.f;
        int cnt = 0, vcnt = 0;
#pragma unroll
        for (int c = 0; c < UB_S2; ++c) {
            const size_t o = ((size_t)b * UB_S2 + c) * P + p;
            const float t = target[o], q = pred[o], e = t - q;
            tp = fmaf(t, q, tp); tt = fmaf(t, t, tt); pp = fmaf(q, q, pp);
            se = fmaf(e, e, se); ae += fabsf(e);
            if (e == e) { nse = fmaf(e, e, nse); nae += fabsf(e); nerr += e; ++cnt; }
            if (var) { const float w = var[o]; if (w == w) { nv += w; ++vcnt; } }
        }
        // spectral angle mapper (metrics.py:25-30): acos(clamp(<t,p> / |t| / |p|, -1, 1)) in degrees
        float m = tp / sqrtf(tt);
        m = m / sqrtf(pp);
        const float cl = m != m ? m : fminf(fmaxf(m, -1.f), 1.f);       // torch.clamp keeps NaN
        const float sam = acosf(cl) * 180.0f / 3.14159274101257324f;
        v[M_SE] = se; v[M_AE] = ae; v[M_SAM] = sam; v[M_NSE] = nse; v[M_NAE] = nae; v[M_NERR] = nerr; v[M_NCNT] = cnt;
        v[M_NVAR] = nv; v[M_NVCNT] = vcnt;
        if (pix) {      // nanmean over the band dimension per pixel (metrics.py:51-54)
            const float nanv = __int_as_float(0x7fc00000);
            float* d = pix + (size_t)b * 4 * P + p;
            d[0] = cnt ? nerr / cnt : nanv;
            d[(size_t)P] = cnt ? nae / cnt : nanv;
            d[(size_t)2 * P] = cnt ? nse / cnt : nanv;
            d[(size_t)3 * P] = vcnt ? nv / vcnt : nanv;
        }
    }
    __shared__ double red[8][9];
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const double t = warp_sum_d(v[k]);
        if (lane == 0) red[warp][k] = t;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        atomicAdd(&acc[(size_t)b * UB_MET_ACC + threadIdx.x], t);
    }
}

constexpr int SS_T = 32, SS_R = 5, SS_IN = SS_T + 2 * SS_R;     // 32x32 outputs, 42x42 inputs
__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ img1, const float* __restrict__ img2, double* acc, int H, int W) {
    __shared__ float s1[SS_IN][SS_IN + 1], s2[SS_IN][SS_IN + 1];
    __shared__ float hz[5][SS_IN][SS_T + 1];
    __shared__ float g[2 * SS_R + 1];
    __shared__ double red[8];
    const int tid = threadIdx.x, c = blockIdx.y, b = blockIdx.z;
    const int tiles_x = (W + SS_T - 1) / SS_T, ty = blockIdx.x / tiles_x, tx = blockIdx.x % tiles_x;
    const int y0 = ty * SS_T - SS_R, x0 = tx * SS_T - SS_R;
    if (tid == 0) {      // gaussian(11, 1.5) normalised in float32 (pytorch_ssim/__init__.py:7-9)
        float w[2 * SS_R + 1], sum = 0.f;
        for (int i = 0; i <= 2 * SS_R; ++i) { w[i] = (float)exp(-(double)((i - SS_R) * (i - SS_R)) / (2.0 * 1.5 * 1.5)); sum += w[i]; }
        for (int i = 0; i <= 2 * SS_R; ++i) g[i] = w[i] / sum;
    }
    const float* a = img1 + ((size_t)b * UB_S2 + c) * H * W;
    const float* bb = img2 + ((size_t)b * UB_S2 + c) * H * W;
    for (int i = tid; i < SS_IN * SS_IN; i += 256) {
        const int r = i / SS_IN, q = i % SS_IN, y = y0 + r, x = x0 + q;
        const bool in = y >= 0 && y < H && x >= 0 && x < W;          // zero padding (F.conv2d padding=5)
        s1[r][q] = in ? a[(size_t)y * W + x] : 0.f;
        s2[r][q] = in ? bb[(size_t)y * W + x] : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < SS_IN * SS_T; i += 256) {
        const int r = i / SS_T, q = i % SS_T;
        float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * SS_R; ++k) {
            const float u = s1[r][q + k], v = s2[r][q + k], w = g[k];
            m1 = fmaf(w, u, m1); m2 = fmaf(w, v, m2);
            e11 = fmaf(w, u * u, e11); e22 = fmaf(w, v * v, e22); e12 = fmaf(w, u * v, e12);
        }
        hz[0][r][q] = m1; hz[1][r][q] = m2; hz[2][r][q] = e11; hz[3][r][q] = e22; hz[4][r][q] = e12;
    }
    __syncthreads();
    double part = 0.0;
    for (int i = tid; i < SS_T * SS_T; i += 256) {
        const int r = i / SS_T, q = i % SS_T, y = ty * SS_T + r, x = tx * SS_T + q;
        if (y >= H || x >= W) continue;
        float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * SS_R; ++k) {
            const float w = g[k];
            m1 = fmaf(w, hz[0][r + k][q], m1); m2 = fmaf(w, hz[1][r + k][q], m2);
            e11 = fmaf(w, hz[2][r + k][q], e11); e22 = fmaf(w, hz[3][r + k][q], e22); e12 = fmaf(w, hz[4][r + k][q], e12);
        }
        const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
        const float v1 = e11 - m11, v2 = e22 - m22, v12 = e12 - m12;
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        part += (double)(((2.f * m12 + C1) * (2.f * v12 + C2)) / ((m11 + m22 + C1) * (v1 + v2 + C2)));
    }
    part = warp_sum_d(part);
    if (tid % 32 == 0) red[tid / 32] = part;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(&acc[(size_t)b * UB_MET_ACC + M_SSIM], t);
    }
}

int launch_img_metrics(const float* target, const float* pred, const float* var, double* acc, float* pix, int B, int H, int W,
                       cudaStream_t st) {
    const int P = H * W;
    if (cudaMemsetAsync(acc, 0, (size_t)B * UB_MET_ACC * sizeof(double), st) != cudaSuccess) return UB_ERR_CUDA;
    img_stats_kernel<<<dim3((P + 255) / 256, B), 256, 0, st>>>(target, pred, var, acc, pix, P);
    UB_CHECK_LAUNCH();
    const int tiles = ((H + SS_T - 1) / SS_T) * ((W + SS_T - 1) / SS_T);
    ssim_kernel<<<dim3(tiles, UB_S2, B), 256, 0, st>>>(target, pred, acc, H, W);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

}  // namespace ub
