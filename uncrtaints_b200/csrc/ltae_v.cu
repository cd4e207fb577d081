// `use_v` variant of the temporal encoder: the full LTAE2d (model/src/backbones/ltae.py:10-141 with MultiHeadAttention :244-307 and
// ScaledDotProductAttention :388-416) whose value output v is upsampled and mixed into the aggregated features by the 1x1
// convolution include_v (model/src/backbones/uncrtaints.py:324-338,414-417).
//
// Everything here lives on the LOW-RESOLUTION grid (B*1024 pixels, <= 50 MB per tensor), so the path is built from small, exact
// fp32 pieces -- the attention weights come from the same folded score kernel as the LTAE2dtiny path (temporal.cu: ltae_fwd):
//   xn = GroupNorm16(pooled) * gamma + beta          [Ne*1024][128]   ltaev_norm_kernel          (ltae.py:103)
//   z  = xn W_in^T + b_in + pe                       [Ne*1024][256]   simt_linear                (ltae.py:105-116)
//   o  = sum_t attn[h(j), t] z[t][j]                 [B*1024][256]    ltaev_attnv_kernel         (ltae.py:286,410; heads = 16-channel slices)
//   m  = o W_m^T + b_m  (+ BatchNorm1d statistics)   [B*1024][128]    simt_linear                (ltae.py:74-83,129)
//   v  = GroupNorm16(dropout(relu(bn(m))))           [B*1024][128]    ltaev_post_fwd_kernel      (ltae.py:129-131)
//   vv = v W_v^T + b                                 [B*1024][128]    simt_linear   (include_v commutes with the bilinear upsampling:
//   mix = agg W_a^T + up(vv)                         [B][P][128]      simt_linear_upadd           conv1x1(cat(agg, up(v))) = W_a agg + up(W_v v + b))
// and the matching backward kernels.  The full-resolution work is ONE 128 -> 128 GEMM per direction plus its weight gradient.
#include "common.cuh"
#include "kernels.h"

namespace ub {

constexpr int LQ = UB_LOW * UB_LOW;     // low-resolution pixels per sample

// GroupNorm(16 groups) statistics of one low-res pixel over (8 channels x T frames); lane l owns channels 4l..4l+3, a group = 2 lanes
__device__ __forceinline__ void v_stats(const float* __restrict__ pooled, int b, int q, int lane, int T, float eps, float& mean, float& rstd) {
    float sum = 0.f;
    for (int t = 0; t < T; ++t) {
        const float4 v = ld4(pooled + (((size_t)(b * T + t)) * LQ + q) * UB_WIDTH + lane * 4);
        sum += v.x + v.y + v.z + v.w;
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    mean = sum / (8.f * T);
    float var = 0.f;
    for (int t = 0; t < T; ++t) {
        const float4 v = ld4(pooled + (((size_t)(b * T + t)) * LQ + q) * UB_WIDTH + lane * 4);
        const float a = v.x - mean, bb = v.y - mean, c = v.z - mean, d = v.w - mean;
        var = fmaf(a, a, var); var = fmaf(bb, bb, var); var = fmaf(c, c, var); var = fmaf(d, d, var);
    }
    var += __shfl_xor_sync(0xffffffffu, var, 1);
    rstd = 1.0f / sqrtf(var / (8.f * T) + eps);
}
__device__ __forceinline__ float4 v_xhat(const float* __restrict__ pooled, int b, int t, int q, int lane, int T, float mean, float rstd) {
    const float4 v = ld4(pooled + (((size_t)(b * T + t)) * LQ + q) * UB_WIDTH + lane * 4);
    return make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd);
}

// xn[(b*T+t)*1024 + q][c] = gamma_c * x_hat + beta_c      (one warp per low-res pixel)
__global__ void __launch_bounds__(256) ltaev_norm_kernel(const float* __restrict__ pooled, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ xn, int B, int T, float eps) {
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int pix = blockIdx.x * 8 + warp;
    if (pix >= B * LQ) return;
    const int b = pix / LQ, q = pix % LQ;
    float mean, rstd;
    v_stats(pooled, b, q, lane, T, eps, mean, rstd);
    const float4 g = ld4(gamma + lane * 4), be = ld4(beta + lane * 4);
    for (int t = 0; t < T; ++t) {
        const float4 xh = v_xhat(pooled, b, t, q, lane, T, mean, rstd);
        st4(xn + (((size_t)(b * T + t)) * LQ + q) * UB_WIDTH + lane * 4,
            make_float4(fmaf(g.x, xh.x, be.x), fmaf(g.y, xh.y, be.y), fmaf(g.z, xh.z, be.z), fmaf(g.w, xh.w, be.w)));
    }
}

// o[pix][j] = sum_t attn[j/16][b][t][q] * z[(b*T+t)*1024 + q][j]       thread = (pix, 4 channels)
__global__ void __launch_bounds__(256) ltaev_attnv_kernel(const float* __restrict__ attn, const float* __restrict__ z, float* __restrict__ o,
                                                           int B, int T) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)B * LQ * (UB_HID / 4)) return;
    const int j4 = (int)(i % (UB_HID / 4)), pix = (int)(i / (UB_HID / 4)), b = pix / LQ, q = pix % LQ, h = j4 / 4;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int t = 0; t < T; ++t) {
        const float a = attn[(((size_t)h * B + b) * T + t) * LQ + q];
        const float4 v = ld4(z + (((size_t)(b * T + t)) * LQ + q) * UB_HID + j4 * 4);
        acc.x = fmaf(a, v.x, acc.x); acc.y = fmaf(a, v.y, acc.y); acc.z = fmaf(a, v.z, acc.z); acc.w = fmaf(a, v.w, acc.w);
    }
    st4(o + (size_t)pix * UB_HID + j4 * 4, acc);
}
// dz[(b*T+t)*1024+q][j] = attn[j/16,b,t,q] * do[pix][j];   dattn[h,b,t,q] += sum_{j in head h} do[pix][j] * z[..][j]
// one warp per low-res pixel, lane owns 8 channels (two lanes per head)
__global__ void __launch_bounds__(256) ltaev_attnv_bwd_kernel(const float* __restrict__ attn, const float* __restrict__ z,
                                                               const float* __restrict__ d_o, float* __restrict__ dz, float* dattn, int B, int T) {
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int pix = blockIdx.x * 8 + warp;
    if (pix >= B * LQ) return;
    const int b = pix / LQ, q = pix % LQ, h = lane / 2;
    const float4 d0 = ld4(d_o + (size_t)pix * UB_HID + lane * 8), d1 = ld4(d_o + (size_t)pix * UB_HID + lane * 8 + 4);
    for (int t = 0; t < T; ++t) {
        const size_t ao = (((size_t)h * B + b) * T + t) * LQ + q;
        const size_t zo = (((size_t)(b * T + t)) * LQ + q) * UB_HID + lane * 8;
        const float a = attn[ao];
        const float4 z0 = ld4(z + zo), z1 = ld4(z + zo + 4);
        st4(dz + zo, make_float4(a * d0.x, a * d0.y, a * d0.z, a * d0.w));
        st4(dz + zo + 4, make_float4(a * d1.x, a * d1.y, a * d1.z, a * d1.w));
        float dot = d0.x * z0.x + d0.y * z0.y + d0.z * z0.z + d0.w * z0.w + d1.x * z1.x + d1.y * z1.y + d1.z * z1.z + d1.w * z1.w;
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        if ((lane & 1) == 0) dattn[ao] += dot;          // one writer per (h, b, t, q)
    }
}

// Bernoulli keep factor (0 or 1/(1-p)) of element idx of the [B*1024][128] MLP output (nn.Dropout(0.2), ltae.py:85,129)
__device__ __forceinline__ float4 v_keep4(const unsigned char* keep_mask, unsigned long long seed, unsigned long long offset, float p,
                                          size_t idx4 /* index of the first of 4 consecutive elements, % 4 == 0 */) {
    if (!(p > 0.f)) return make_float4(1.f, 1.f, 1.f, 1.f);
    const float inv = 1.f / (1.f - p);
    if (keep_mask) {
        const uchar4 k = *reinterpret_cast<const uchar4*>(keep_mask + idx4);
        return make_float4(k.x ? inv : 0.f, k.y ? inv : 0.f, k.z ? inv : 0.f, k.w ? inv : 0.f);
    }
    const unsigned long long blk = (idx4 >> 2) + offset;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), 1u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t bits[4] = {r.x, r.y, r.z, r.w};
    float k[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) k[i] = (((float)(bits[i] >> 8) * (1.0f / 16777216.0f)) >= p) ? inv : 0.f;
    return make_float4(k[0], k[1], k[2], k[3]);
}

// v = out_norm(dropout(relu(m*scale + shift)))     out_norm = GroupNorm(16, 128) on [rows][128]: groups of 8 channels per row
// one warp per row; lane owns 4 channels (a group = 2 lanes)
// The BatchNorm is applied as ((m - mean) * rstd) * gamma_bn + beta_bn, not as m * scale + shift: the ReLU behind it makes the
// SIGN of the result matter (an element that is +1e-7 in the reference and -1e-7 here loses its whole gradient, and in a group
// whose other members are zero the GroupNorm amplifies that gradient by up to 1/sqrt(eps)); the centred form has no cancellation.
__device__ __forceinline__ float bn_relu_in(float m, const MeanRstd& s, float g, float b) { return fmaf((m - s.mean) * s.rstd, g, b); }
__global__ void __launch_bounds__(256) ltaev_post_fwd_kernel(const float* __restrict__ m, const MeanRstd* __restrict__ mr /* [B][128] */,
                                                              const float* __restrict__ bn_g, const float* __restrict__ bn_b,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const unsigned char* __restrict__ keep_mask, unsigned long long seed,
                                                              unsigned long long offset, float drop_p, float* __restrict__ v, int B, float eps) {
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int pix = blockIdx.x * 8 + warp;
    if (pix >= B * LQ) return;
    const int b = pix / LQ;
    const float4 mv = ld4(m + (size_t)pix * UB_WIDTH + lane * 4);
    const MeanRstd* k = mr + (size_t)b * UB_WIDTH + lane * 4;
    const float4 bg = ld4(bn_g + lane * 4), bb = ld4(bn_b + lane * 4);
    const float4 kp = v_keep4(keep_mask, seed, offset, drop_p, (size_t)pix * UB_WIDTH + lane * 4);
    float r[4] = {fmaxf(bn_relu_in(mv.x, k[0], bg.x, bb.x), 0.f) * kp.x, fmaxf(bn_relu_in(mv.y, k[1], bg.y, bb.y), 0.f) * kp.y,
                  fmaxf(bn_relu_in(mv.z, k[2], bg.z, bb.z), 0.f) * kp.z, fmaxf(bn_relu_in(mv.w, k[3], bg.w, bb.w), 0.f) * kp.w};
    float sum = r[0] + r[1] + r[2] + r[3];
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    const float mean = sum * 0.125f;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { r[i] -= mean; var = fmaf(r[i], r[i], var); }
    var += __shfl_xor_sync(0xffffffffu, var, 1);
    const float rstd = 1.0f / sqrtf(var * 0.125f + eps);
    const float4 g = ld4(gamma + lane * 4), be = ld4(beta + lane * 4);
    st4(v + (size_t)pix * UB_WIDTH + lane * 4, make_float4(fmaf(g.x, r[0] * rstd, be.x), fmaf(g.y, r[1] * rstd, be.y),
                                                           fmaf(g.z, r[2] * rstd, be.z), fmaf(g.w, r[3] * rstd, be.w)));
}
// backward of the above down to the BatchNorm output: dyb = d(bn(m)) [rows][128]; bstats[b][c] += (sum dyb, sum dyb * m_hat);
// dgamma_o / dbeta_o += (sum dv * r_hat, sum dv)
__global__ void __launch_bounds__(256) ltaev_post_bwd_kernel(const float* __restrict__ m, const MeanRstd* __restrict__ mr,
                                                              const float* __restrict__ bn_g, const float* __restrict__ bn_b,
                                                              const float* __restrict__ gamma,
                                                              const unsigned char* __restrict__ keep_mask, unsigned long long seed,
                                                              unsigned long long offset, float drop_p, const float* __restrict__ dv,
                                                              float* __restrict__ dyb, double* bstats, float* dgamma, float* dbeta, int B,
                                                              float eps, int rows_per_warp) {
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int first = (blockIdx.x * 8 + warp) * rows_per_warp;      // rows_per_warp divides 1024: a warp stays inside one sample
    if (first >= B * LQ) return;
    const int b = first / LQ;
    const float4 g = ld4(gamma + lane * 4);
    const MeanRstd* ms = mr + (size_t)b * UB_WIDTH + lane * 4;
    const float4 bg4 = ld4(bn_g + lane * 4), bb4 = ld4(bn_b + lane * 4);
    const float bg[4] = {bg4.x, bg4.y, bg4.z, bg4.w}, bb[4] = {bb4.x, bb4.y, bb4.z, bb4.w};
    float dg[4] = {0, 0, 0, 0}, db[4] = {0, 0, 0, 0}, s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
    for (int pix = first; pix < first + rows_per_warp; ++pix) {
        const float4 mv4 = ld4(m + (size_t)pix * UB_WIDTH + lane * 4);
        const float mv[4] = {mv4.x, mv4.y, mv4.z, mv4.w};
        const float4 kp4 = v_keep4(keep_mask, seed, offset, drop_p, (size_t)pix * UB_WIDTH + lane * 4);
        const float kp[4] = {kp4.x, kp4.y, kp4.z, kp4.w};
        float pre[4], r[4];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { pre[i] = bn_relu_in(mv[i], ms[i], bg[i], bb[i]); r[i] = fmaxf(pre[i], 0.f) * kp[i]; sum += r[i]; }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        const float mean = sum * 0.125f;
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { r[i] -= mean; var = fmaf(r[i], r[i], var); }
        var += __shfl_xor_sync(0xffffffffu, var, 1);
        const float rstd = 1.0f / sqrtf(var * 0.125f + eps);
        const float4 d4 = ld4(dv + (size_t)pix * UB_WIDTH + lane * 4);
        const float d[4] = {d4.x, d4.y, d4.z, d4.w};
        const float gg[4] = {g.x, g.y, g.z, g.w};
        float dh[4], m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float rh = r[i] * rstd;
            dg[i] = fmaf(d[i], rh, dg[i]);
            db[i] += d[i];
            dh[i] = d[i] * gg[i];
            m1 += dh[i];
            m2 = fmaf(dh[i], rh, m2);
            r[i] = rh;
        }
        m1 += __shfl_xor_sync(0xffffffffu, m1, 1);
        m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
        m1 *= 0.125f; m2 *= 0.125f;
        float out[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float dr = rstd * (dh[i] - m1 - r[i] * m2);             // GroupNorm backward (per row, group of 8)
            out[i] = pre[i] > 0.f ? dr * kp[i] : 0.f;                     // dropout, ReLU
            s0[i] += out[i];
            s1[i] = fmaf(out[i], (mv[i] - ms[i].mean) * ms[i].rstd, s1[i]);
        }
        st4(dyb + (size_t)pix * UB_WIDTH + lane * 4, make_float4(out[0], out[1], out[2], out[3]));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        atomicAdd(&dgamma[lane * 4 + i], dg[i]);
        atomicAdd(&dbeta[lane * 4 + i], db[i]);
        atomicAdd(&bstats[((size_t)b * UB_WIDTH + lane * 4 + i) * 2 + 0], (double)s0[i]);
        atomicAdd(&bstats[((size_t)b * UB_WIDTH + lane * 4 + i) * 2 + 1], (double)s1[i]);
    }
}
// dm = a*dyb + b*m + c  (BatchNorm1d backward with the finalised coefficients; [B][1024][128])
__global__ void __launch_bounds__(256) normbwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const BCoef* __restrict__ bc,
                                                             float* __restrict__ dx, int rows_per_frame, size_t total4) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total4) return;
    const int c4 = (int)(i % (UB_WIDTH / 4));
    const size_t row = i / (UB_WIDTH / 4);
    const int n = (int)(row / rows_per_frame);
    const BCoef* k = bc + (size_t)n * UB_WIDTH + c4 * 4;
    const float4 d = ld4(dy + i * 4), xv = ld4(x + i * 4);
    st4(dx + i * 4, make_float4(fmaf(k[0].a, d.x, fmaf(k[0].b, xv.x, k[0].c)), fmaf(k[1].a, d.y, fmaf(k[1].b, xv.y, k[1].c)),
                                fmaf(k[2].a, d.z, fmaf(k[2].b, xv.z, k[2].c)), fmaf(k[3].a, d.w, fmaf(k[3].b, xv.w, k[3].c))));
}

// out[c] += sum_rows x[row][c]      C = blockDim.x (128 or 256); grid = row chunks
__global__ void colsum_kernel(const float* __restrict__ x, float* out, size_t rows, int chunk) {
    const int C = blockDim.x, c = threadIdx.x;
    const size_t r0 = (size_t)blockIdx.x * chunk, r1 = r0 + chunk < rows ? r0 + chunk : rows;
    float s = 0.f;
    for (size_t r = r0; r < r1; ++r) s += x[r * C + c];
    atomicAdd(&out[c], s);
}

// dlow[n][q][c] = sum over the full-resolution pixels p whose bilinear taps include q of weight * dfull[n][p][c]
// (adjoint of nn.Upsample(bilinear, align_corners=False), uncrtaints.py:416).  grid (32 low rows, N), 128 threads = channel.
__global__ void __launch_bounds__(128) upsample_adjoint128_kernel(const float* __restrict__ dfull, float* __restrict__ dlow, int H, int W) {
    const int n = blockIdx.y, Y = blockIdx.x, c = threadIdx.x;
    const int sy = H / UB_LOW, sx = W / UB_LOW;
    const float isy = (float)UB_LOW / (float)H, isx = (float)UB_LOW / (float)W;
    const float* src = dfull + (size_t)n * H * W * UB_WIDTH + c;
    // full-resolution rows that can touch low row Y: taps (y0, y0 + 1) with y0 in {Y - 1, Y}
    const int ylo = max(0, (Y - 1) * sy), yhi = min(H, (Y + 2) * sy);
    for (int X = 0; X < UB_LOW; ++X) {
        const int xlo = max(0, (X - 1) * sx), xhi = min(W, (X + 2) * sx);
        float acc = 0.f;
        for (int y = ylo; y < yhi; ++y) {
            int y0, y1; float ly;
            bilinear_tap(y, isy, UB_LOW, y0, y1, ly);
            const float wy = (y0 == Y ? 1.f - ly : 0.f) + (y1 == Y ? ly : 0.f);
            if (wy == 0.f) continue;
            for (int x = xlo; x < xhi; ++x) {
                int x0, x1; float lx;
                bilinear_tap(x, isx, UB_LOW, x0, x1, lx);
                const float wx = (x0 == X ? 1.f - lx : 0.f) + (x1 == X ? lx : 0.f);
                if (wx != 0.f) acc = fmaf(wy * wx, src[(size_t)(y * W + x) * UB_WIDTH], acc);
            }
        }
        dlow[((size_t)n * LQ + Y * UB_LOW + X) * UB_WIDTH + c] = acc;
    }
}

// Final backward of the temporal encoder with a value path: softmax backward of the summed attention gradient (aggregation +
// value), the folded score path (dAp, de), the value path's gradient dxn w.r.t. the normalised features, the in_norm affine
// gradients, and the GroupNorm backward -> dpooled.  One warp per low-res pixel, run-time T (ds staged in shared memory).
__global__ void __launch_bounds__(256) ltaev_final_bwd_kernel(const float* __restrict__ pooled, const float* __restrict__ Ap,
                                                               const float* __restrict__ attn, const float* __restrict__ dattn,
                                                               const float* __restrict__ dxn, const float* __restrict__ gamma,
                                                               float* __restrict__ dpooled, float* dAp, float* de, float* dgamma, float* dbeta,
                                                               int B, int T, float eps, int pix_per_warp) {
    __shared__ __align__(16) float sAp[UB_HEADS * UB_WIDTH];
    __shared__ float sds[8][UB_TLONG][UB_HEADS];
    for (int i = threadIdx.x; i < UB_HEADS * UB_WIDTH; i += 256) sAp[i] = Ap[i];
    __syncthreads();
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int npix = B * LQ;
    float4 gA[UB_HEADS];
#pragma unroll
    for (int h = 0; h < UB_HEADS; ++h) gA[h] = make_float4(0, 0, 0, 0);
    float4 dg = make_float4(0, 0, 0, 0), db = make_float4(0, 0, 0, 0);
    const float4 g = ld4(gamma + lane * 4);
    const int first = (blockIdx.x * 8 + warp) * pix_per_warp;
    for (int pix = first; pix < first + pix_per_warp && pix < npix; ++pix) {
        const int b = pix / LQ, q = pix % LQ;
        float mean, rstd;
        v_stats(pooled, b, q, lane, T, eps, mean, rstd);
        __syncwarp();
        if (lane < UB_HEADS) {          // softmax backward per head (masked entries have attn == 0 => zero gradient)
            float dot = 0.f;
            for (int t = 0; t < T; ++t) {
                const size_t o = (((size_t)lane * B + b) * T + t) * LQ + q;
                dot = fmaf(attn[o], dattn[o], dot);
            }
            for (int t = 0; t < T; ++t) {
                const size_t o = (((size_t)lane * B + b) * T + t) * LQ + q;
                const float ds = 0.5f * attn[o] * (dattn[o] - dot);
                sds[warp][t][lane] = ds;
                atomicAdd(&de[((size_t)b * T + t) * UB_HEADS + lane], ds);
            }
        }
        __syncwarp();
        float m1 = 0.f, m2 = 0.f;
        for (int pass = 0; pass < 2; ++pass) {      // pass 0: GroupNorm-backward means + parameter gradients; pass 1: dpooled
            for (int t = 0; t < T; ++t) {
                const float4 xh = v_xhat(pooled, b, t, q, lane, T, mean, rstd);
                const float4 dn = ld4(dxn + (((size_t)(b * T + t)) * LQ + q) * UB_WIDTH + lane * 4);
                float4 dx = make_float4(g.x * dn.x, g.y * dn.y, g.z * dn.z, g.w * dn.w);
#pragma unroll
                for (int h = 0; h < UB_HEADS; ++h) {
                    const float4 a = ld4(sAp + h * UB_WIDTH + lane * 4);
                    const float d = sds[warp][t][h];
                    if (pass == 0) {
                        gA[h].x = fmaf(d, xh.x, gA[h].x); gA[h].y = fmaf(d, xh.y, gA[h].y);
                        gA[h].z = fmaf(d, xh.z, gA[h].z); gA[h].w = fmaf(d, xh.w, gA[h].w);
                    }
                    dx.x = fmaf(a.x, d, dx.x); dx.y = fmaf(a.y, d, dx.y); dx.z = fmaf(a.z, d, dx.z); dx.w = fmaf(a.w, d, dx.w);
                }
                if (pass == 0) {
                    dg.x = fmaf(dn.x, xh.x, dg.x); dg.y = fmaf(dn.y, xh.y, dg.y); dg.z = fmaf(dn.z, xh.z, dg.z); dg.w = fmaf(dn.w, xh.w, dg.w);
                    db.x += dn.x; db.y += dn.y; db.z += dn.z; db.w += dn.w;
                    m1 += dx.x + dx.y + dx.z + dx.w;
                    m2 += dx.x * xh.x + dx.y * xh.y + dx.z * xh.z + dx.w * xh.w;
                } else {
                    st4(dpooled + (((size_t)(b * T + t)) * LQ + q) * UB_WIDTH + lane * 4,
                        make_float4(rstd * (dx.x - m1 - xh.x * m2), rstd * (dx.y - m1 - xh.y * m2), rstd * (dx.z - m1 - xh.z * m2),
                                    rstd * (dx.w - m1 - xh.w * m2)));
                }
            }
            if (pass == 0) {
                m1 += __shfl_xor_sync(0xffffffffu, m1, 1);
                m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
                m1 /= 8.f * T; m2 /= 8.f * T;
            }
        }
        __syncwarp();
    }
    atomicAdd(&dgamma[lane * 4 + 0], dg.x); atomicAdd(&dgamma[lane * 4 + 1], dg.y);
    atomicAdd(&dgamma[lane * 4 + 2], dg.z); atomicAdd(&dgamma[lane * 4 + 3], dg.w);
    atomicAdd(&dbeta[lane * 4 + 0], db.x); atomicAdd(&dbeta[lane * 4 + 1], db.y);
    atomicAdd(&dbeta[lane * 4 + 2], db.z); atomicAdd(&dbeta[lane * 4 + 3], db.w);
    __shared__ __align__(16) float red[8 * UB_WIDTH];
#pragma unroll 1
    for (int h = 0; h < UB_HEADS; ++h) {
        float4 v = gA[0];
#pragma unroll
        for (int j = 1; j < UB_HEADS; ++j) if (j == h) v = gA[j];
        __syncthreads();
        st4(red + warp * UB_WIDTH + lane * 4, v);
        __syncthreads();
        if (threadIdx.x < UB_WIDTH) {
            float t = 0.f;
#pragma unroll
            for (int r = 0; r < 8; ++r) t += red[r * UB_WIDTH + threadIdx.x];
            atomicAdd(&dAp[h * UB_WIDTH + threadIdx.x], t);
        }
    }
}

// dst[r][c] = src[r * ld + c0 + c]  (transpose == 0)   or   dst[c][r] = src[r * ld + c0 + c]  (transpose == 1)
__global__ void slice2d_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols, int ld, int c0, int transpose) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int r = i / cols, c = i % cols;
    const float v = src[(size_t)r * ld + c0 + c];
    if (transpose) dst[(size_t)c * rows + r] = v; else dst[i] = v;
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
int launch_ltaev_norm(const float* pooled, const float* gamma, const float* beta, float* xn, int B, int T, float eps, cudaStream_t st) {
    ltaev_norm_kernel<<<(B * LQ + 7) / 8, 256, 0, st>>>(pooled, gamma, beta, xn, B, T, eps);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_ltaev_attnv(const float* attn, const float* z, float* o, int B, int T, cudaStream_t st) {
    const size_t total = (size_t)B * LQ * (UB_HID / 4);
    ltaev_attnv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(attn, z, o, B, T);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_ltaev_attnv_bwd(const float* attn, const float* z, const float* d_o, float* dz, float* dattn, int B, int T, cudaStream_t st) {
    ltaev_attnv_bwd_kernel<<<(B * LQ + 7) / 8, 256, 0, st>>>(attn, z, d_o, dz, dattn, B, T);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_ltaev_post_fwd(const float* m, const MeanRstd* mr, const float* bn_g, const float* bn_b, const float* gamma, const float* beta,
                          const unsigned char* keep_mask, unsigned long long seed, unsigned long long offset, float drop_p, float* v, int B,
                          float eps, cudaStream_t st) {
    ltaev_post_fwd_kernel<<<(B * LQ + 7) / 8, 256, 0, st>>>(m, mr, bn_g, bn_b, gamma, beta, keep_mask, seed, offset, drop_p, v, B, eps);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_ltaev_post_bwd(const float* m, const MeanRstd* mr, const float* bn_g, const float* bn_b, const float* gamma,
                          const unsigned char* keep_mask, unsigned long long seed, unsigned long long offset, float drop_p, const float* dv,
                          float* dyb, double* bstats, float* dgamma, float* dbeta, int B, float eps, cudaStream_t st) {
    const int rpw = 16;                        // rows per warp (divides 1024)
    ltaev_post_bwd_kernel<<<(B * LQ / rpw + 7) / 8, 256, 0, st>>>(m, mr, bn_g, bn_b, gamma, keep_mask, seed, offset, drop_p, dv, dyb, bstats,
                                                                   dgamma, dbeta, B, eps, rpw);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_normbwd_apply(const float* dy, const float* x, const BCoef* bc, float* dx, int N, int rows_per_frame, cudaStream_t st) {
    const size_t total4 = (size_t)N * rows_per_frame * (UB_WIDTH / 4);
    normbwd_apply_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(dy, x, bc, dx, rows_per_frame, total4);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_colsum(const float* x, float* out, size_t rows, int C, cudaStream_t st) {
    if (C != 128 && C != 256) return UB_ERR_ARG;
    const int chunk = 256;
    colsum_kernel<<<(unsigned)((rows + chunk - 1) / chunk), C, 0, st>>>(x, out, rows, chunk);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_upsample_adjoint128(const float* dfull, float* dlow, int N, int H, int W, cudaStream_t st) {
    upsample_adjoint128_kernel<<<dim3(UB_LOW, N), 128, 0, st>>>(dfull, dlow, H, W);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_ltaev_final_bwd(const float* pooled, const float* Ap, const float* attn, const float* dattn, const float* dxn,
                           const float* gamma, float* dpooled, float* dAp, float* de, float* dgamma, float* dbeta, int B, int T, float eps,
                           cudaStream_t st) {
    const int ppw = 8;
    ltaev_final_bwd_kernel<<<(B * LQ / ppw + 7) / 8, 256, 0, st>>>(pooled, Ap, attn, dattn, dxn, gamma, dpooled, dAp, de, dgamma, dbeta, B, T,
                                                                    eps, ppw);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_slice2d(const float* src, float* dst, int rows, int cols, int ld, int c0, int transpose, cudaStream_t st) {
    slice2d_kernel<<<(rows * cols + 255) / 256, 256, 0, st>>>(src, dst, rows, cols, ld, c0, transpose);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

}  // namespace ub
