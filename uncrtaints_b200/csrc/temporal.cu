// Temporal path of UnCRtainTS: adaptive max-pool to 32x32, L-TAE-tiny attention, attention-weighted
// temporal aggregation -- forward and backward.
//
//   maxpool    : nn.AdaptiveMaxPool2d((32,32)) (model/src/backbones/uncrtaints.py:403-404), first-max argmax
//   ltae       : LTAE2dtiny.forward + MultiHeadAttentionSmall + ScaledDotProductAttentionSmall
//                (model/src/backbones/ltae.py:197-239, 341-385, 431-458)
//   aggregate  : Compact_Temporal_Aggregator, mode att_group (uncrtaints.py:156-210): bilinear x(H/32) upsample
//                (align_corners=False) -> Dropout(0.1) -> (~pad_mask) -> sum_t a[c//8] * x[t, c]
//
// L-TAE algebra.  The query is input independent (ltae.py:324,347), so with x_hat the GroupNorm(16)-normalised
// pooled features (groups = 8 channels x T time steps per low-res pixel, ltae.py:211):
//   score[h,t] = 1/2 * ( sum_c Ap[h,c] * x_hat[c,t] + e[b,t,h] )
//   Ap[h,c] = gamma_c * sum_k Ak[h,k] * W_in[k,c],  Ak[h,k] = sum_d Q[h,d] * W_k[4h+d,k]
//   e[b,t,h] = pe[b,t,:].Ak[h,:] + (terms constant in t, which cancel in the softmax)
// Ap and e are produced by the Python shim from the reference's parameters with differentiable torch ops
// (they are functions of ~100K weights, not of activations); the kernels here return dAp and de.
// The upsampled attention ([16*B, T, H, W], 201 MB at B=16 in the reference) is never materialised:
// the four bilinear taps are evaluated inside the aggregation kernel.
#include "common.cuh"
#include "kernels.h"

namespace ub {

// ------------------------------------------------------------------------------------------
// adaptive max pool (H, W multiples of 32): window s x s, s = H/32.  grid (1024, N), 128 threads (= channel)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ pooled,
                                                           int* __restrict__ idx, int H, int W) {
    constexpr int C = UB_WIDTH;
    const int n = blockIdx.y, win = blockIdx.x, wy = win / UB_LOW, wx = win % UB_LOW, c = threadIdx.x;
    const int sy = H / UB_LOW, sx = W / UB_LOW;
    const float* src = x + (size_t)n * H * W * C + c;
    float best = -INFINITY;
    int bi = (wy * sy) * W + wx * sx;
    for (int i = 0; i < sy; ++i)
        for (int j = 0; j < sx; ++j) {
            const int p = (wy * sy + i) * W + wx * sx + j;
            const float v = src[(size_t)p * C];
            if (v > best || v != v) { best = v; bi = p; }   // ATen rule: (val > max) || isnan(val)
        }
    pooled[((size_t)n * UB_LOW * UB_LOW + win) * C + c] = best;
    idx[((size_t)n * UB_LOW * UB_LOW + win) * C + c] = bi;
}

// dEnc[n][idx][c] += dpooled[n][win][c]   (windows are disjoint: plain read-modify-write)
__global__ void __launch_bounds__(128) maxpool_bwd_kernel(const float* __restrict__ dpooled, const int* __restrict__ idx,
                                                           float* __restrict__ denc, int HW) {
    constexpr int C = UB_WIDTH;
    const int n = blockIdx.y, win = blockIdx.x, c = threadIdx.x;
    const size_t o = ((size_t)n * UB_LOW * UB_LOW + win) * C + c;
    denc[((size_t)n * HW + idx[o]) * C + c] += dpooled[o];
}

// ------------------------------------------------------------------------------------------
// L-TAE tiny.  One warp per low-res pixel; lane l owns channels 4l..4l+3 (GroupNorm group = lanes 2g, 2g+1).
// ------------------------------------------------------------------------------------------
template <int T>
__device__ __forceinline__ void ltae_normalize(const float* __restrict__ pooled, int b, int q, int lane, float eps,
                                               float (&xh)[T][4], float& rstd) {
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const float4 v = ld4(pooled + (((size_t)(b * T + t)) * UB_LOW * UB_LOW + q) * UB_WIDTH + lane * 4);
        xh[t][0] = v.x; xh[t][1] = v.y; xh[t][2] = v.z; xh[t][3] = v.w;
        sum += v.x + v.y + v.z + v.w;
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    const float mean = sum / (8.f * T);
    float var = 0.f;
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) { xh[t][j] -= mean; var = fmaf(xh[t][j], xh[t][j], var); }
    var += __shfl_xor_sync(0xffffffffu, var, 1);
    rstd = 1.0f / sqrtf(var / (8.f * T) + eps);
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) xh[t][j] *= rstd;
}

template <int T>
__global__ void __launch_bounds__(256) ltae_fwd_kernel(const float* __restrict__ pooled, const float* __restrict__ Ap /* [16][128] */,
                                                        const float* __restrict__ e /* [B][T][16] */, const int* __restrict__ notpad /* [B*T] */,
                                                        float* __restrict__ attn /* [16][B][T][1024] */, int B, float eps) {
    __shared__ __align__(16) float sAp[UB_HEADS * UB_WIDTH];
    for (int i = threadIdx.x; i < UB_HEADS * UB_WIDTH; i += 256) sAp[i] = Ap[i];
    __syncthreads();
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int pix = blockIdx.x * 8 + warp;          // b * 1024 + q
    if (pix >= B * UB_LOW * UB_LOW) return;
    const int b = pix / (UB_LOW * UB_LOW), q = pix % (UB_LOW * UB_LOW);
    float xh[T][4], rstd;
    ltae_normalize<T>(pooled, b, q, lane, eps, xh, rstd);
    float sc[T];
#pragma unroll
    for (int t = 0; t < T; ++t) sc[t] = 0.f;
#pragma unroll 4
    for (int h = 0; h < UB_HEADS; ++h) {
        const float4 a = ld4(sAp + h * UB_WIDTH + lane * 4);
#pragma unroll
        for (int t = 0; t < T; ++t) {
            float part = a.x * xh[t][0] + a.y * xh[t][1] + a.z * xh[t][2] + a.w * xh[t][3];
            part = warp_sum(part);
            if (lane == h) sc[t] = part;
        }
    }
    if (lane < UB_HEADS) {
        float mx = -INFINITY;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            sc[t] = 0.5f * (sc[t] + e[((size_t)b * T + t) * UB_HEADS + lane]);    // / temperature sqrt(d_k)=2, ltae.py:339,433
            if (!notpad[b * T + t]) sc[t] = -1000.0f;                             // masked_fill(pad, -1e3), ltae.py:435
            mx = fmaxf(mx, sc[t]);
        }
        float den = 0.f;
#pragma unroll
        for (int t = 0; t < T; ++t) { sc[t] = expf(sc[t] - mx); den += sc[t]; }
        const float inv = 1.0f / den;
#pragma unroll
        for (int t = 0; t < T; ++t)
            attn[(((size_t)lane * B + b) * T + t) * (UB_LOW * UB_LOW) + q] = sc[t] * inv;
    }
}

template <int T>
__global__ void __launch_bounds__(256) ltae_bwd_kernel(const float* __restrict__ pooled, const float* __restrict__ Ap,
                                                        const float* __restrict__ attn, const float* __restrict__ dattn,
                                                        float* __restrict__ dpooled, float* dAp, float* de, int B, float eps,
                                                        int pix_per_warp) {
    __shared__ __align__(16) float sAp[UB_HEADS * UB_WIDTH];
    for (int i = threadIdx.x; i < UB_HEADS * UB_WIDTH; i += 256) sAp[i] = Ap[i];
    __syncthreads();
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int npix = B * UB_LOW * UB_LOW;
    float4 gA[UB_HEADS];
#pragma unroll
    for (int h = 0; h < UB_HEADS; ++h) gA[h] = make_float4(0, 0, 0, 0);
    const int first = (blockIdx.x * 8 + warp) * pix_per_warp;
    for (int pix = first; pix < first + pix_per_warp && pix < npix; ++pix) {
        const int b = pix / (UB_LOW * UB_LOW), q = pix % (UB_LOW * UB_LOW);
        float xh[T][4], rstd;
        ltae_normalize<T>(pooled, b, q, lane, eps, xh, rstd);
        // softmax backward on lanes < 16 (lane = head); masked entries have attn == 0 => zero gradient
        float ds[T];
#pragma unroll
        for (int t = 0; t < T; ++t) ds[t] = 0.f;
        if (lane < UB_HEADS) {
            float a[T], dot = 0.f;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const size_t o = (((size_t)lane * B + b) * T + t) * (UB_LOW * UB_LOW) + q;
                a[t] = attn[o];
                ds[t] = dattn[o];
                dot = fmaf(a[t], ds[t], dot);
            }
#pragma unroll
            for (int t = 0; t < T; ++t) {
                ds[t] = 0.5f * a[t] * (ds[t] - dot);
                atomicAdd(&de[((size_t)b * T + t) * UB_HEADS + lane], ds[t]);
            }
        }
        float dxh[T][4];
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
            for (int j = 0; j < 4; ++j) dxh[t][j] = 0.f;
#pragma unroll
        for (int h = 0; h < UB_HEADS; ++h) {
            const float4 a = ld4(sAp + h * UB_WIDTH + lane * 4);
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const float d = __shfl_sync(0xffffffffu, ds[t], h);
                gA[h].x = fmaf(d, xh[t][0], gA[h].x); gA[h].y = fmaf(d, xh[t][1], gA[h].y);
                gA[h].z = fmaf(d, xh[t][2], gA[h].z); gA[h].w = fmaf(d, xh[t][3], gA[h].w);
                dxh[t][0] = fmaf(a.x, d, dxh[t][0]); dxh[t][1] = fmaf(a.y, d, dxh[t][1]);
                dxh[t][2] = fmaf(a.z, d, dxh[t][2]); dxh[t][3] = fmaf(a.w, d, dxh[t][3]);
            }
        }
        // GroupNorm backward without affine (gamma/beta are folded into Ap / e)
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
            for (int j = 0; j < 4; ++j) { m1 += dxh[t][j]; m2 = fmaf(dxh[t][j], xh[t][j], m2); }
        m1 += __shfl_xor_sync(0xffffffffu, m1, 1);
        m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
        m1 /= 8.f * T; m2 /= 8.f * T;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            float4 o;
            o.x = rstd * (dxh[t][0] - m1 - xh[t][0] * m2);
            o.y = rstd * (dxh[t][1] - m1 - xh[t][1] * m2);
            o.z = rstd * (dxh[t][2] - m1 - xh[t][2] * m2);
            o.w = rstd * (dxh[t][3] - m1 - xh[t][3] * m2);
            st4(dpooled + (((size_t)(b * T + t)) * UB_LOW * UB_LOW + q) * UB_WIDTH + lane * 4, o);
        }
    }
    // dAp: reduce over the 8 warps, then one atomic per element per CTA
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

// ------------------------------------------------------------------------------------------
// aggregation.  One warp per group of 4 consecutive pixels (one Philox call per (head, t) covers the 4).
// lane l owns channels 4l..4l+3, head h = l/2.
// ------------------------------------------------------------------------------------------
struct AggArgs {
    const float* attn;        // [16][B][T][32][32]
    const int* notpad;        // [B*T]
    const unsigned char* keep_mask;   // optional explicit dropout keep mask [16][B][T][H][W] (tests), else Philox
    unsigned long long seed, offset;
    float drop_p;             // 0 => no dropout (eval)
    int B, T, H, W;
};

// Per-CTA shared state of the aggregation kernels.  A CTA owns ONE image row y of one sample, so the vertical half of the
// bilinear upsampling is the same for all its pixels: the row-interpolated attention v[h][t][X] (x notpad[b,t]) of all 16
// heads is staged once in shared memory (head stride padded by one float: the 16 heads of a warp hit 16 different banks)
// and every pixel only needs the horizontal lerp of two shared-memory values -- the kernels no longer issue 16-way
// scattered global loads per tap (ncu r01: 15 sectors per request, 56 % long-scoreboard stalls).
template <int T>
struct AggRow {
    float v[UB_HEADS][T * UB_LOW + 1];
};
template <int T>
__device__ __forceinline__ void agg_stage_row(const AggArgs& a, int b, int y, AggRow<T>& S) {
    const float inv_sy = (float)UB_LOW / (float)a.H;
    int y0, y1; float ly;
    bilinear_tap(y, inv_sy, UB_LOW, y0, y1, ly);
    for (int i = threadIdx.x; i < UB_HEADS * T * UB_LOW; i += blockDim.x) {
        const int h = i / (T * UB_LOW), r = i % (T * UB_LOW), t = r / UB_LOW, X = r % UB_LOW;
        const float* src = a.attn + (((size_t)h * a.B + b) * T + t) * (UB_LOW * UB_LOW);
        const float vv = src[y0 * UB_LOW + X] * (1.f - ly) + src[y1 * UB_LOW + X] * ly;
        S.v[h][r] = a.notpad[b * T + t] ? vv : 0.f;
    }
}
// dropout factors keep/(1-p) of head h (= lane/2) for frames t = 0..T-1 at the 4 pixels of group (y, x0): one Philox call
// covers the 4 pixels of one (h, t); the two lanes of a head split the frames (even lane: even t) and swap 4-bit masks.
template <int T>
__device__ __forceinline__ void agg_dropout(const AggArgs& a, int h, int b, int y, int x0, int lane, float (&k)[T][4]) {
    if (!(a.drop_p > 0.f)) {
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
            for (int i = 0; i < 4; ++i) k[t][i] = 1.f;
        return;
    }
    const float inv_keep = 1.f / (1.f - a.drop_p);
    uint32_t mine[(T + 1) / 2];
#pragma unroll
    for (int kk = 0; kk < (T + 1) / 2; ++kk) {
        const int t = 2 * kk + (lane & 1);
        uint32_t m = 0;
        if (t < T) {
            const size_t lin = ((((size_t)h * a.B + b) * T + t) * a.H + y) * a.W + x0;
            if (a.keep_mask) {
#pragma unroll
                for (int i = 0; i < 4; ++i) m |= (a.keep_mask[lin + i] ? 1u : 0u) << i;
            } else {
                const unsigned long long blk = (lin >> 2) + a.offset;
                const uint4 rnd = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), 0u, 0u),
                                                make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
                const uint32_t bits[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) m |= (((float)(bits[i] >> 8) * (1.0f / 16777216.0f)) >= a.drop_p ? 1u : 0u) << i;
            }
        }
        mine[kk] = m;
    }
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const uint32_t other = __shfl_xor_sync(0xffffffffu, mine[t >> 1], 1);
        const uint32_t m = ((lane & 1) == (t & 1)) ? mine[t >> 1] : other;
#pragma unroll
        for (int i = 0; i < 4; ++i) k[t][i] = ((m >> i) & 1u) ? inv_keep : 0.f;
    }
}

// forward: out[b,p,c] = sum_t w[h(c),b,t,p] * x[b,t,p,c].  grid (H, B), 256 threads; warp = groups of 4 pixels of the row,
// lane = 4 channels (head = lane / 2).  All T x 4 activation loads of a group are issued before the first FMA.
template <int T>
__global__ void __launch_bounds__(256) aggregate_fwd_kernel(AggArgs a, const float* __restrict__ x /* [B*T][P][128] */,
                                                             float* __restrict__ out /* [B][P][128] */, double* out_stats) {
    constexpr int C = UB_WIDTH;
    __shared__ AggRow<T> S;
    __shared__ __align__(16) float smem[2 * 8 * C];
    const int y = blockIdx.x, b = blockIdx.y, lane = threadIdx.x % 32, warp = threadIdx.x / 32, h = lane / 2;
    const int P = a.H * a.W;
    agg_stage_row<T>(a, b, y, S);
    __syncthreads();
    const float inv_sx = (float)UB_LOW / (float)a.W;
    float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
    for (int x0 = warp * 4; x0 < a.W; x0 += 32) {
        const int p = y * a.W + x0;
        float4 v[T][4];
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
            for (int i = 0; i < 4; ++i) v[t][i] = ld4_stream(x + (((size_t)(b * T + t)) * P + p + i) * C + lane * 4);
        float k[T][4];
        agg_dropout<T>(a, h, b, y, x0, lane, k);
        int xa[4], xb[4]; float lx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) bilinear_tap(x0 + i, inv_sx, UB_LOW, xa[i], xb[i], lx[i]);
        float4 acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float w = (S.v[h][t * UB_LOW + xa[i]] * (1.f - lx[i]) + S.v[h][t * UB_LOW + xb[i]] * lx[i]) * k[t][i];
                acc[i].x = fmaf(w, v[t][i].x, acc[i].x); acc[i].y = fmaf(w, v[t][i].y, acc[i].y);
                acc[i].z = fmaf(w, v[t][i].z, acc[i].z); acc[i].w = fmaf(w, v[t][i].w, acc[i].w);
            }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            st4(out + ((size_t)b * P + p + i) * C + lane * 4, acc[i]);
            s.x += acc[i].x; s.y += acc[i].y; s.z += acc[i].z; s.w += acc[i].w;
            q.x += acc[i].x * acc[i].x; q.y += acc[i].y * acc[i].y; q.z += acc[i].z * acc[i].z; q.w += acc[i].w * acc[i].w;
        }
    }
    // block reduce (c4 = lane, row = warp) -> per-(b, c) sums for the first decoder PreNorm
    float4* sa = reinterpret_cast<float4*>(smem);
    sa[warp * 32 + lane] = s;
    sa[256 + warp * 32 + lane] = q;
    __syncthreads();
    {
        const int which = threadIdx.x / C, ch = threadIdx.x % C;
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += (double)smem[which * 8 * C + r * C + ch];
        atomicAdd(&out_stats[((size_t)b * C + ch) * 2 + which], t);
    }
}

// backward A: dEnc[b,t,p,c] = w[h,b,t,p] * dAgg[b,p,c];  dwup[h,b,t,p] = d(upsampled attention) = keep/(1-p) * notpad * <dAgg, x>_head
template <int T>
__global__ void __launch_bounds__(256) aggregate_bwd_kernel(AggArgs a, const float* __restrict__ x, const float* __restrict__ dagg,
                                                             float* __restrict__ denc, float* __restrict__ dwup /* [16][B][T][P] */) {
    constexpr int C = UB_WIDTH;
    __shared__ AggRow<T> S;
    const int y = blockIdx.x, b = blockIdx.y, lane = threadIdx.x % 32, warp = threadIdx.x / 32, h = lane / 2;
    const int P = a.H * a.W;
    agg_stage_row<T>(a, b, y, S);
    __syncthreads();
    const float inv_sx = (float)UB_LOW / (float)a.W;
    for (int x0 = warp * 4; x0 < a.W; x0 += 32) {
        const int p = y * a.W + x0;
        float4 d[4], v[T][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = ld4_stream(dagg + ((size_t)b * P + p + i) * C + lane * 4);
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
            for (int i = 0; i < 4; ++i) v[t][i] = ld4_stream(x + (((size_t)(b * T + t)) * P + p + i) * C + lane * 4);
        float k[T][4];
        agg_dropout<T>(a, h, b, y, x0, lane, k);
        int xa[4], xb[4]; float lx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) bilinear_tap(x0 + i, inv_sx, UB_LOW, xa[i], xb[i], lx[i]);
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const size_t fr = ((size_t)(b * T + t)) * P + p;
            const float np = a.notpad[b * T + t] ? 1.f : 0.f;
            float dots[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float w = (S.v[h][t * UB_LOW + xa[i]] * (1.f - lx[i]) + S.v[h][t * UB_LOW + xb[i]] * lx[i]) * k[t][i];
                st4(denc + (fr + i) * C + lane * 4, make_float4(w * d[i].x, w * d[i].y, w * d[i].z, w * d[i].w));
                float dot = d[i].x * v[t][i].x + d[i].y * v[t][i].y + d[i].z * v[t][i].z + d[i].w * v[t][i].w;
                dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                dots[i] = dot * k[t][i] * np;
            }
            if ((lane & 1) == 0) {
                const size_t lin = ((((size_t)h * a.B + b) * T + t) * a.H + y) * a.W + x0;
                st4(dwup + lin, make_float4(dots[0], dots[1], dots[2], dots[3]));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Long sequences (T > UB_TMAX, up to UB_TLONG): the dataset yields up to 30 time points (data/dataLoader.py:200).  Same
// arithmetic with run-time loops over t instead of register arrays: the pooled features are re-read from the L1/L2 (they are
// B*T*1024*128 floats in total), scores / score gradients pass through shared memory.  These tensors are low resolution or
// touched once, so the simple form costs little; the T <= 8 templates above stay the fast path of the benchmark configurations.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ltae_stats_long(const float* __restrict__ pooled, int b, int q, int lane, int T, float eps,
                                                float& mean, float& rstd) {
    float sum = 0.f;
    for (int t = 0; t < T; ++t) {
        const float4 v = ld4(pooled + (((size_t)(b * T + t)) * UB_LOW * UB_LOW + q) * UB_WIDTH + lane * 4);
        sum += v.x + v.y + v.z + v.w;
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    mean = sum / (8.f * T);
    float var = 0.f;
    for (int t = 0; t < T; ++t) {
        const float4 v = ld4(pooled + (((size_t)(b * T + t)) * UB_LOW * UB_LOW + q) * UB_WIDTH + lane * 4);
        const float a = v.x - mean, bb = v.y - mean, c = v.z - mean, d = v.w - mean;
        var = fmaf(a, a, var); var = fmaf(bb, bb, var); var = fmaf(c, c, var); var = fmaf(d, d, var);
    }
    var += __shfl_xor_sync(0xffffffffu, var, 1);
    rstd = 1.0f / sqrtf(var / (8.f * T) + eps);
}
__device__ __forceinline__ float4 ltae_xhat(const float* __restrict__ pooled, int b, int t, int q, int lane, int T, float mean, float rstd) {
    const float4 v = ld4(pooled + (((size_t)(b * T + t)) * UB_LOW * UB_LOW + q) * UB_WIDTH + lane * 4);
    return make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd);
}

__global__ void __launch_bounds__(256) ltae_fwd_long_kernel(const float* __restrict__ pooled, const float* __restrict__ Ap,
                                                             const float* __restrict__ e, const int* __restrict__ notpad,
                                                             float* __restrict__ attn, int B, int T, float eps) {
    __shared__ __align__(16) float sAp[UB_HEADS * UB_WIDTH];
    __shared__ float ssc[8][UB_TLONG][UB_HEADS];
    for (int i = threadIdx.x; i < UB_HEADS * UB_WIDTH; i += 256) sAp[i] = Ap[i];
    __syncthreads();
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int pix = blockIdx.x * 8 + warp;
    if (pix >= B * UB_LOW * UB_LOW) return;
    const int b = pix / (UB_LOW * UB_LOW), q = pix % (UB_LOW * UB_LOW);
    float mean, rstd;
    ltae_stats_long(pooled, b, q, lane, T, eps, mean, rstd);
    for (int t = 0; t < T; ++t) {
        const float4 xh = ltae_xhat(pooled, b, t, q, lane, T, mean, rstd);
#pragma unroll 4
        for (int h = 0; h < UB_HEADS; ++h) {
            const float4 a = ld4(sAp + h * UB_WIDTH + lane * 4);
            const float part = warp_sum(a.x * xh.x + a.y * xh.y + a.z * xh.z + a.w * xh.w);
            if (lane == h) ssc[warp][t][h] = part;
        }
    }
    __syncwarp();
    if (lane < UB_HEADS) {
        float mx = -INFINITY;
        for (int t = 0; t < T; ++t) {
            float sc = 0.5f * (ssc[warp][t][lane] + e[((size_t)b * T + t) * UB_HEADS + lane]);
            if (!notpad[b * T + t]) sc = -1000.0f;
            ssc[warp][t][lane] = sc;
            mx = fmaxf(mx, sc);
        }
        float den = 0.f;
        for (int t = 0; t < T; ++t) { const float ex = expf(ssc[warp][t][lane] - mx); ssc[warp][t][lane] = ex; den += ex; }
        const float inv = 1.0f / den;
        for (int t = 0; t < T; ++t) attn[(((size_t)lane * B + b) * T + t) * (UB_LOW * UB_LOW) + q] = ssc[warp][t][lane] * inv;
    }
}

__global__ void __launch_bounds__(256) ltae_bwd_long_kernel(const float* __restrict__ pooled, const float* __restrict__ Ap,
                                                             const float* __restrict__ attn, const float* __restrict__ dattn,
                                                             float* __restrict__ dpooled, float* dAp, float* de, int B, int T, float eps,
                                                             int pix_per_warp) {
    __shared__ __align__(16) float sAp[UB_HEADS * UB_WIDTH];
    __shared__ float sds[8][UB_TLONG][UB_HEADS];
    for (int i = threadIdx.x; i < UB_HEADS * UB_WIDTH; i += 256) sAp[i] = Ap[i];
    __syncthreads();
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int npix = B * UB_LOW * UB_LOW;
    float4 gA[UB_HEADS];
#pragma unroll
    for (int h = 0; h < UB_HEADS; ++h) gA[h] = make_float4(0, 0, 0, 0);
    const int first = (blockIdx.x * 8 + warp) * pix_per_warp;
    for (int pix = first; pix < first + pix_per_warp && pix < npix; ++pix) {
        const int b = pix / (UB_LOW * UB_LOW), q = pix % (UB_LOW * UB_LOW);
        float mean, rstd;
        ltae_stats_long(pooled, b, q, lane, T, eps, mean, rstd);
        __syncwarp();
        if (lane < UB_HEADS) {          // softmax backward per head
            float dot = 0.f;
            for (int t = 0; t < T; ++t) {
                const size_t o = (((size_t)lane * B + b) * T + t) * (UB_LOW * UB_LOW) + q;
                dot = fmaf(attn[o], dattn[o], dot);
            }
            for (int t = 0; t < T; ++t) {
                const size_t o = (((size_t)lane * B + b) * T + t) * (UB_LOW * UB_LOW) + q;
                const float ds = 0.5f * attn[o] * (dattn[o] - dot);
                sds[warp][t][lane] = ds;
                atomicAdd(&de[((size_t)b * T + t) * UB_HEADS + lane], ds);
            }
        }
        __syncwarp();
        // pass A: GroupNorm-backward means and the dAp accumulation; pass B: the input gradient
        float m1 = 0.f, m2 = 0.f;
        for (int t = 0; t < T; ++t) {
            const float4 xh = ltae_xhat(pooled, b, t, q, lane, T, mean, rstd);
            float4 dx = make_float4(0, 0, 0, 0);
#pragma unroll
            for (int h = 0; h < UB_HEADS; ++h) {
                const float4 a = ld4(sAp + h * UB_WIDTH + lane * 4);
                const float d = sds[warp][t][h];
                gA[h].x = fmaf(d, xh.x, gA[h].x); gA[h].y = fmaf(d, xh.y, gA[h].y);
                gA[h].z = fmaf(d, xh.z, gA[h].z); gA[h].w = fmaf(d, xh.w, gA[h].w);
                dx.x = fmaf(a.x, d, dx.x); dx.y = fmaf(a.y, d, dx.y); dx.z = fmaf(a.z, d, dx.z); dx.w = fmaf(a.w, d, dx.w);
            }
            m1 += dx.x + dx.y + dx.z + dx.w;
            m2 += dx.x * xh.x + dx.y * xh.y + dx.z * xh.z + dx.w * xh.w;
        }
        m1 += __shfl_xor_sync(0xffffffffu, m1, 1);
        m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
        m1 /= 8.f * T; m2 /= 8.f * T;
        for (int t = 0; t < T; ++t) {
            const float4 xh = ltae_xhat(pooled, b, t, q, lane, T, mean, rstd);
            float4 dx = make_float4(0, 0, 0, 0);
#pragma unroll
            for (int h = 0; h < UB_HEADS; ++h) {
                const float4 a = ld4(sAp + h * UB_WIDTH + lane * 4);
                const float d = sds[warp][t][h];
                dx.x = fmaf(a.x, d, dx.x); dx.y = fmaf(a.y, d, dx.y); dx.z = fmaf(a.z, d, dx.z); dx.w = fmaf(a.w, d, dx.w);
            }
            st4(dpooled + (((size_t)(b * T + t)) * UB_LOW * UB_LOW + q) * UB_WIDTH + lane * 4,
                make_float4(rstd * (dx.x - m1 - xh.x * m2), rstd * (dx.y - m1 - xh.y * m2), rstd * (dx.z - m1 - xh.z * m2),
                            rstd * (dx.w - m1 - xh.w * m2)));
        }
        __syncwarp();
    }
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

// keep / (1 - p) of head h for frame t at the 4 pixels of group (y, x0) (same Philox stream / explicit mask as agg_dropout)
__device__ __forceinline__ void agg_dropout_long(const AggArgs& a, int h, int b, int t, int y, int x0, float (&k)[4]) {
    if (!(a.drop_p > 0.f)) { k[0] = k[1] = k[2] = k[3] = 1.f; return; }
    const float inv_keep = 1.f / (1.f - a.drop_p);
    const size_t lin = ((((size_t)h * a.B + b) * a.T + t) * a.H + y) * a.W + x0;
    if (a.keep_mask) {
#pragma unroll
        for (int i = 0; i < 4; ++i) k[i] = a.keep_mask[lin + i] ? inv_keep : 0.f;
        return;
    }
    const unsigned long long blk = (lin >> 2) + a.offset;
    const uint4 rnd = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), 0u, 0u), make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
    const uint32_t bits[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) k[i] = (((float)(bits[i] >> 8) * (1.0f / 16777216.0f)) >= a.drop_p) ? inv_keep : 0.f;
}
// stage the vertically interpolated attention row: sv[h * stride + t * 32 + X], stride = T * 32 + 1 (dynamic shared memory)
__device__ __forceinline__ void agg_stage_row_long(const AggArgs& a, int b, int y, float* sv, int stride) {
    const int T = a.T;
    const float inv_sy = (float)UB_LOW / (float)a.H;
    int y0, y1; float ly;
    bilinear_tap(y, inv_sy, UB_LOW, y0, y1, ly);
    for (int i = threadIdx.x; i < UB_HEADS * T * UB_LOW; i += blockDim.x) {
        const int h = i / (T * UB_LOW), r = i % (T * UB_LOW), t = r / UB_LOW, X = r % UB_LOW;
        const float* src = a.attn + (((size_t)h * a.B + b) * T + t) * (UB_LOW * UB_LOW);
        const float vv = src[y0 * UB_LOW + X] * (1.f - ly) + src[y1 * UB_LOW + X] * ly;
        sv[h * stride + r] = a.notpad[b * T + t] ? vv : 0.f;
    }
}

__global__ void __launch_bounds__(256) aggregate_fwd_long_kernel(AggArgs a, const float* __restrict__ x, float* __restrict__ out,
                                                                  double* out_stats) {
    constexpr int C = UB_WIDTH;
    extern __shared__ float sv[];
    __shared__ __align__(16) float smem[2 * 8 * C];
    const int T = a.T, stride = T * UB_LOW + 1;
    const int y = blockIdx.x, b = blockIdx.y, lane = threadIdx.x % 32, warp = threadIdx.x / 32, h = lane / 2;
    const int P = a.H * a.W;
    agg_stage_row_long(a, b, y, sv, stride);
    __syncthreads();
    const float inv_sx = (float)UB_LOW / (float)a.W;
    float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
    for (int x0 = warp * 4; x0 < a.W; x0 += 32) {
        const int p = y * a.W + x0;
        int xa[4], xb[4]; float lx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) bilinear_tap(x0 + i, inv_sx, UB_LOW, xa[i], xb[i], lx[i]);
        float4 acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = make_float4(0, 0, 0, 0);
        for (int t = 0; t < T; ++t) {
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = ld4_stream(x + (((size_t)(b * T + t)) * P + p + i) * C + lane * 4);
            float k[4];
            agg_dropout_long(a, h, b, t, y, x0, k);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float w = (sv[h * stride + t * UB_LOW + xa[i]] * (1.f - lx[i]) + sv[h * stride + t * UB_LOW + xb[i]] * lx[i]) * k[i];
                acc[i].x = fmaf(w, v[i].x, acc[i].x); acc[i].y = fmaf(w, v[i].y, acc[i].y);
                acc[i].z = fmaf(w, v[i].z, acc[i].z); acc[i].w = fmaf(w, v[i].w, acc[i].w);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            st4(out + ((size_t)b * P + p + i) * C + lane * 4, acc[i]);
            s.x += acc[i].x; s.y += acc[i].y; s.z += acc[i].z; s.w += acc[i].w;
            q.x += acc[i].x * acc[i].x; q.y += acc[i].y * acc[i].y; q.z += acc[i].z * acc[i].z; q.w += acc[i].w * acc[i].w;
        }
    }
    float4* sa = reinterpret_cast<float4*>(smem);
    sa[warp * 32 + lane] = s;
    sa[256 + warp * 32 + lane] = q;
    __syncthreads();
    {
        const int which = threadIdx.x / C, ch = threadIdx.x % C;
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += (double)smem[which * 8 * C + r * C + ch];
        atomicAdd(&out_stats[((size_t)b * C + ch) * 2 + which], t);
    }
}

__global__ void __launch_bounds__(256) aggregate_bwd_long_kernel(AggArgs a, const float* __restrict__ x, const float* __restrict__ dagg,
                                                                  float* __restrict__ denc, float* __restrict__ dwup) {
    constexpr int C = UB_WIDTH;
    extern __shared__ float sv[];
    const int T = a.T, stride = T * UB_LOW + 1;
    const int y = blockIdx.x, b = blockIdx.y, lane = threadIdx.x % 32, warp = threadIdx.x / 32, h = lane / 2;
    const int P = a.H * a.W;
    agg_stage_row_long(a, b, y, sv, stride);
    __syncthreads();
    const float inv_sx = (float)UB_LOW / (float)a.W;
    for (int x0 = warp * 4; x0 < a.W; x0 += 32) {
        const int p = y * a.W + x0;
        float4 d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = ld4_stream(dagg + ((size_t)b * P + p + i) * C + lane * 4);
        int xa[4], xb[4]; float lx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) bilinear_tap(x0 + i, inv_sx, UB_LOW, xa[i], xb[i], lx[i]);
        for (int t = 0; t < T; ++t) {
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = ld4_stream(x + (((size_t)(b * T + t)) * P + p + i) * C + lane * 4);
            float k[4];
            agg_dropout_long(a, h, b, t, y, x0, k);
            const size_t fr = ((size_t)(b * T + t)) * P + p;
            const float np = a.notpad[b * T + t] ? 1.f : 0.f;
            float dots[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float w = (sv[h * stride + t * UB_LOW + xa[i]] * (1.f - lx[i]) + sv[h * stride + t * UB_LOW + xb[i]] * lx[i]) * k[i];
                st4(denc + (fr + i) * C + lane * 4, make_float4(w * d[i].x, w * d[i].y, w * d[i].z, w * d[i].w));
                float dot = d[i].x * v[i].x + d[i].y * v[i].y + d[i].z * v[i].z + d[i].w * v[i].w;
                dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                dots[i] = dot * k[i] * np;
            }
            if ((lane & 1) == 0) {
                const size_t lin = ((((size_t)h * a.B + b) * T + t) * a.H + y) * a.W + x0;
                st4(dwup + lin, make_float4(dots[0], dots[1], dots[2], dots[3]));
            }
        }
    }
}

// backward B: adjoint of the bilinear upsampling, gather form (deterministic, no atomics):
// dattn[h,b,t,Y,X] = sum over the footprint of cell (Y,X) of tapweight * dwup[h,b,t,y,x]
__global__ void __launch_bounds__(256) upsample_adjoint_kernel(const float* __restrict__ dwup, float* __restrict__ dattn,
                                                                int H, int W, int total_cells) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= total_cells) return;
    const int X = cell % UB_LOW, Y = (cell / UB_LOW) % UB_LOW;
    const size_t img = (size_t)(cell / (UB_LOW * UB_LOW));
    const int sy = H / UB_LOW, sx = W / UB_LOW;
    const float inv_sy = (float)UB_LOW / (float)H, inv_sx = (float)UB_LOW / (float)W;
    const int ya = max(0, (Y - 1) * sy + sy / 2 - 1), yb = min(H - 1, (Y + 1) * sy + sy / 2);
    const int xa = max(0, (X - 1) * sx + sx / 2 - 1), xb = min(W - 1, (X + 1) * sx + sx / 2);
    const float* src = dwup + img * H * W;
    float acc = 0.f;
    for (int y = ya; y <= yb; ++y) {
        int y0, y1; float ly;
        bilinear_tap(y, inv_sy, UB_LOW, y0, y1, ly);
        const float wy = (y0 == Y ? 1.f - ly : 0.f) + (y1 == Y ? ly : 0.f);
        if (wy == 0.f) continue;
        float rowacc = 0.f;
        for (int x = xa; x <= xb; ++x) {
            int x0, x1; float lx;
            bilinear_tap(x, inv_sx, UB_LOW, x0, x1, lx);
            const float wx = (x0 == X ? 1.f - lx : 0.f) + (x1 == X ? lx : 0.f);
            rowacc = fmaf(wx, src[(size_t)y * W + x], rowacc);
        }
        acc = fmaf(wy, rowacc, acc);
    }
    dattn[cell] = acc;
}

// The same adjoint with coalesced reads (W = 32 * SX): a warp owns one low-resolution row Y of one (head, sample, frame) image,
// lane = cell X.  Per image row of the footprint every lane reads its own SX pixels (the warp reads the row contiguously) and
// forms three dot products: into its own cell and into the cells left / right of it (the footprint of a cell reaches SX/2+1
// pixels into each neighbour); the neighbour parts are exchanged by two shuffles at the very end.  The weight of every pixel
// comes from the same bilinear_tap as the forward pass, so the clamped borders need no special case.
// (The cell-per-thread form above reads with a 32-byte lane stride: 8 L1 wavefronts per load, 0.24 ms at B=16, T=3, 256x256.)
template <int SX>
__global__ void __launch_bounds__(256) upsample_adjoint_rows_kernel(const float* __restrict__ dwup, float* __restrict__ dattn,
                                                                     int H, int total_rows) {
    constexpr int W = UB_LOW * SX;
    const int row = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32), X = threadIdx.x % 32;
    if (row >= total_rows) return;
    const int Y = row % UB_LOW;
    const size_t img = (size_t)(row / UB_LOW);
    const int sy = H / UB_LOW;
    const float inv_sy = (float)UB_LOW / (float)H, inv_sx = 1.f / (float)SX;
    float wo[SX], wl[SX], wr[SX];                 // weight of own pixel i in cell X, X-1, X+1
#pragma unroll
    for (int i = 0; i < SX; ++i) {
        int x0, x1; float lx;
        bilinear_tap(X * SX + i, inv_sx, UB_LOW, x0, x1, lx);
        wo[i] = (x0 == X ? 1.f - lx : 0.f) + (x1 == X ? lx : 0.f);
        wl[i] = (x0 == X - 1 ? 1.f - lx : 0.f) + (x1 == X - 1 ? lx : 0.f);
        wr[i] = (x0 == X + 1 ? 1.f - lx : 0.f) + (x1 == X + 1 ? lx : 0.f);
    }
    const int ya = max(0, (Y - 1) * sy + sy / 2 - 1), yb = min(H - 1, (Y + 1) * sy + sy / 2);
    const float* src = dwup + img * H * W + X * SX;
    float ao = 0.f, al = 0.f, ar = 0.f;
    for (int y = ya; y <= yb; ++y) {
        int y0, y1; float ly;
        bilinear_tap(y, inv_sy, UB_LOW, y0, y1, ly);
        const float wy = (y0 == Y ? 1.f - ly : 0.f) + (y1 == Y ? ly : 0.f);
        if (wy == 0.f) continue;
        float v[SX];
        if constexpr (SX % 4 == 0) {
#pragma unroll
            for (int i = 0; i < SX; i += 4) {
                const float4 t = ld4(src + (size_t)y * W + i);
                v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < SX; ++i) v[i] = src[(size_t)y * W + i];
        }
        float ro = 0.f, rl = 0.f, rr = 0.f;
#pragma unroll
        for (int i = 0; i < SX; ++i) { ro = fmaf(wo[i], v[i], ro); rl = fmaf(wl[i], v[i], rl); rr = fmaf(wr[i], v[i], rr); }
        ao = fmaf(wy, ro, ao); al = fmaf(wy, rl, al); ar = fmaf(wy, rr, ar);
    }
    const float from_right = __shfl_down_sync(0xffffffffu, al, 1);      // lane X+1's part for its left neighbour = this cell
    const float from_left = __shfl_up_sync(0xffffffffu, ar, 1);         // lane X-1's part for its right neighbour = this cell
    dattn[(size_t)row * UB_LOW + X] = ao + (X < UB_LOW - 1 ? from_right : 0.f) + (X > 0 ? from_left : 0.f);
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static void launch_upsample_adjoint(const float* dwup, float* dattn, int H, int W, int images, cudaStream_t st) {
    const int rows = images * UB_LOW;
    if (W == UB_LOW * 8) upsample_adjoint_rows_kernel<8><<<(rows + 7) / 8, 256, 0, st>>>(dwup, dattn, H, rows);
    else if (W == UB_LOW * 4) upsample_adjoint_rows_kernel<4><<<(rows + 7) / 8, 256, 0, st>>>(dwup, dattn, H, rows);
    else {
        const int cells = rows * UB_LOW;
        upsample_adjoint_kernel<<<(cells + 255) / 256, 256, 0, st>>>(dwup, dattn, H, W, cells);
    }
}
int launch_maxpool_fwd(const float* x, float* pooled, int* idx, int N, int H, int W, cudaStream_t st) {
    if (H % UB_LOW || W % UB_LOW) return UB_ERR_ARG;
    maxpool_fwd_kernel<<<dim3(UB_LOW * UB_LOW, N), 128, 0, st>>>(x, pooled, idx, H, W);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_maxpool_bwd(const float* dpooled, const int* idx, float* denc, int N, int HW, cudaStream_t st) {
    maxpool_bwd_kernel<<<dim3(UB_LOW * UB_LOW, N), 128, 0, st>>>(dpooled, idx, denc, HW);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

#define UB_DISPATCH_T(T_, ...)                     \
    switch (T_) {                                  \
        case 1: { constexpr int TT = 1; __VA_ARGS__; } break; \
        case 2: { constexpr int TT = 2; __VA_ARGS__; } break; \
        case 3: { constexpr int TT = 3; __VA_ARGS__; } break; \
        case 4: { constexpr int TT = 4; __VA_ARGS__; } break; \
        case 5: { constexpr int TT = 5; __VA_ARGS__; } break; \
        case 6: { constexpr int TT = 6; __VA_ARGS__; } break; \
        case 7: { constexpr int TT = 7; __VA_ARGS__; } break; \
        case 8: { constexpr int TT = 8; __VA_ARGS__; } break; \
        default: return UB_ERR_ARG;                \
    }

int launch_ltae_fwd(const float* pooled, const float* Ap, const float* e, const int* notpad, float* attn, int B, int T,
                    float eps, cudaStream_t st) {
    const int npix = B * UB_LOW * UB_LOW;
    if (T > UB_TMAX && T <= UB_TLONG) {
        ltae_fwd_long_kernel<<<(npix + 7) / 8, 256, 0, st>>>(pooled, Ap, e, notpad, attn, B, T, eps);
        UB_CHECK_LAUNCH();
        return UB_OK;
    }
    UB_DISPATCH_T(T, (ltae_fwd_kernel<TT><<<(npix + 7) / 8, 256, 0, st>>>(pooled, Ap, e, notpad, attn, B, eps)));
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_ltae_bwd(const float* pooled, const float* Ap, const float* attn, const float* dattn, float* dpooled, float* dAp,
                    float* de, int B, int T, float eps, cudaStream_t st) {
    const int npix = B * UB_LOW * UB_LOW;
    const int ppw = 8;
    const int blocks = (npix + 8 * ppw - 1) / (8 * ppw);
    if (T > UB_TMAX && T <= UB_TLONG) {
        ltae_bwd_long_kernel<<<blocks, 256, 0, st>>>(pooled, Ap, attn, dattn, dpooled, dAp, de, B, T, eps, ppw);
        UB_CHECK_LAUNCH();
        return UB_OK;
    }
    UB_DISPATCH_T(T, (ltae_bwd_kernel<TT><<<blocks, 256, 0, st>>>(pooled, Ap, attn, dattn, dpooled, dAp, de, B, eps, ppw)));
    UB_CHECK_LAUNCH();
    return UB_OK;
}

static AggArgs make_agg(const float* attn, const int* notpad, const unsigned char* keep_mask, unsigned long long seed,
                        unsigned long long offset, float drop_p, int B, int T, int H, int W) {
    AggArgs a;
    a.attn = attn; a.notpad = notpad; a.keep_mask = keep_mask; a.seed = seed; a.offset = offset; a.drop_p = drop_p;
    a.B = B; a.T = T; a.H = H; a.W = W;
    return a;
}
int launch_aggregate_fwd(const float* attn, const int* notpad, const unsigned char* keep_mask, unsigned long long seed,
                         unsigned long long offset, float drop_p, const float* x, float* out, double* out_stats, int B,
                         int T, int H, int W, cudaStream_t st) {
    if (W % 32) return UB_ERR_ARG;
    const AggArgs a = make_agg(attn, notpad, keep_mask, seed, offset, drop_p, B, T, H, W);
    if (T > UB_TMAX && T <= UB_TLONG) {
        const size_t smem = (size_t)UB_HEADS * (T * UB_LOW + 1) * sizeof(float);
        UB_SET_SMEM(aggregate_fwd_long_kernel, (size_t)UB_HEADS * (UB_TLONG * UB_LOW + 1) * sizeof(float));
        aggregate_fwd_long_kernel<<<dim3(H, B), 256, smem, st>>>(a, x, out, out_stats);
        UB_CHECK_LAUNCH();
        return UB_OK;
    }
    UB_DISPATCH_T(T, (aggregate_fwd_kernel<TT><<<dim3(H, B), 256, 0, st>>>(a, x, out, out_stats)));
    UB_CHECK_LAUNCH();
    return UB_OK;
}
int launch_aggregate_bwd(const float* attn, const int* notpad, const unsigned char* keep_mask, unsigned long long seed,
                         unsigned long long offset, float drop_p, const float* x, const float* dagg, float* denc,
                         float* dwup, float* dattn, int B, int T, int H, int W, cudaStream_t st) {
    if (W % 32) return UB_ERR_ARG;
    const AggArgs a = make_agg(attn, notpad, keep_mask, seed, offset, drop_p, B, T, H, W);
    if (T > UB_TMAX && T <= UB_TLONG) {
        const size_t smem = (size_t)UB_HEADS * (T * UB_LOW + 1) * sizeof(float);
        UB_SET_SMEM(aggregate_bwd_long_kernel, (size_t)UB_HEADS * (UB_TLONG * UB_LOW + 1) * sizeof(float));
        aggregate_bwd_long_kernel<<<dim3(H, B), 256, smem, st>>>(a, x, dagg, denc, dwup);
        UB_CHECK_LAUNCH();
    } else {
        UB_DISPATCH_T(T, (aggregate_bwd_kernel<TT><<<dim3(H, B), 256, 0, st>>>(a, x, dagg, denc, dwup)));
        UB_CHECK_LAUNCH();
    }
    launch_upsample_adjoint(dwup, dattn, H, W, UB_HEADS * B * T, st);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

}  // namespace ub
