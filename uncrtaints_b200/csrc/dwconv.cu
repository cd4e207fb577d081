// Depthwise 3x3 convolution with reflect padding (groups = 256, no bias) of the MBConv block
// (model/src/backbones/uncrtaints.py:130-131), forward (K2) and backward (B3), pixel-major layout.
//
// K2:  h2 = DW3x3_reflect( gelu(h1*scale1 + shift1) );  epilogue: column sums of h2 (Norm2 statistics)
// B3:  dz2 = (du*s + dpool/P) * gelu'(z2);  dh2 = a2*dz2 + b2*h2 + c2      (SE gate + Norm2 backward)
//      dg1 = DW3x3^T(dh2) with the adjoint of reflect padding;  dz1 = dg1 * gelu'(z1)
//      dWdw[c][i][j] += sum_p dh2[p] * g1[reflect(p + (i-1, j-1))];  bstats1 += (sum dz1, sum dz1*h1_hat)
//
// A CTA owns a 32-column strip of one frame for a 32-channel chunk (one 128-byte line per pixel) and walks down the
// image in tiles of 8 rows; the (8+2) x (32+2) x 32 halo tile of the transformed operand lives in shared memory
// (43.5 KB), so the normalisation + GELU is evaluated once per element (1.33x with halo) instead of 9x, and 4-5
// (forward) / 2 (backward) CTAs are resident per SM to overlap one CTA's load phase with another's stencil phase.
// Reflection is index arithmetic on the load (no padded copy as in the reference's F.pad path).
#include "common.cuh"
#include "kernels.h"

namespace ub {

constexpr int DW_TH = 8, DW_TW = 32, DW_CC = 32;            // tile rows, cols, channel chunk
constexpr int DW_Q = DW_CC / 4;                              // channel quads per chunk (8)
constexpr int DW_PT = 256 / DW_Q;                            // pixel threads (32): one column each
constexpr int DW_HR = DW_TH + 2, DW_HC = DW_TW + 2;         // halo tile
constexpr int DW_TILE_FLOATS = DW_HR * DW_HC * DW_CC;       // 10880 floats = 43.5 KB

__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// reduce per-thread float4 partials over the 32 pixel-threads (tid / 8) for each of the 8 channel quads and add the
// result to dst[ch][0..1] (fp64)
__device__ __forceinline__ void reduce_pt_atomic2(float4 a, float4 b, double* dst /* &stats[n][cbase][0] */, float* smem) {
    const int cq = threadIdx.x % DW_Q, pt = threadIdx.x / DW_Q;
    float4* sa = reinterpret_cast<float4*>(smem);
    float4* sb = sa + 256;
    __syncthreads();
    sa[pt * DW_Q + cq] = a;
    sb[pt * DW_Q + cq] = b;
    __syncthreads();
    if (threadIdx.x < 2 * DW_CC) {
        const int which = threadIdx.x / DW_CC, ch = threadIdx.x % DW_CC;
        const float* src = reinterpret_cast<const float*>(which ? sb : sa);
        double t = 0.0;
#pragma unroll 8
        for (int r = 0; r < DW_PT; ++r) t += (double)src[r * DW_CC + ch];
        atomicAdd(&dst[ch * 2 + which], t);
    }
    __syncthreads();
}

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = valid ? 16 : 0;                    // src-size 0 => the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}

// K2.  The raw h1 halo tile of the NEXT 8-row tile is fetched by cp.async into the second shared-memory buffer while
// the current tile is normalised + GELU'd in place and convolved, so no thread ever waits on HBM with work undone.
__global__ void __launch_bounds__(256, 2)
dwconv_fwd_kernel(const float* __restrict__ h1, const Coef* __restrict__ coef1, const float* __restrict__ wdw /* [256][9] */,
                  float* __restrict__ h2, double* stats2, int H, int W) {
    extern __shared__ __align__(16) float smem[];
    constexpr int C = UB_HID;
    constexpr int NITEM = DW_HR * DW_HC * DW_Q;
    const int n = blockIdx.z, cbase = blockIdx.y * DW_CC, x0 = blockIdx.x * DW_TW;
    const int cq = threadIdx.x % DW_Q, col = threadIdx.x / DW_Q;
    const int c0 = cbase + cq * 4;
    Coef k[4];
    float w[9][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        k[i] = coef1[(size_t)n * C + c0 + i];
#pragma unroll
        for (int j = 0; j < 9; ++j) w[j][i] = wdw[(size_t)(c0 + i) * 9 + j];
    }
    const float* src = h1 + (size_t)n * H * W * C;
    float* dst = h2 + (size_t)n * H * W * C;
    float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);

    auto fetch = [&](int y0, float* buf) {          // raw halo tile (reflected source rows / columns) -> shared memory
        for (int e = threadIdx.x; e < NITEM; e += 256) {
            const int pix = e / DW_Q, ry = pix / DW_HC, rx = pix - ry * DW_HC;
            const int sy = reflect_idx(y0 - 1 + ry, H), sx = reflect_idx(x0 - 1 + rx, W);
            cp_async16(buf + pix * DW_CC + cq * 4, src + ((size_t)sy * W + sx) * C + c0, true);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    fetch(0, smem);
    for (int y0 = 0, it = 0; y0 < H; y0 += DW_TH, ++it) {
        float* tile = smem + (it & 1) * DW_TILE_FLOATS;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                               // tile landed; everybody is done with the other buffer
        if (y0 + DW_TH < H) fetch(y0 + DW_TH, smem + ((it + 1) & 1) * DW_TILE_FLOATS);
        for (int e = threadIdx.x; e < NITEM; e += 256) {   // in-place Norm1 + GELU (e % 8 == cq)
            float* ptr = tile + (e / DW_Q) * DW_CC + cq * 4;
            const float4 v = ld4(ptr);
            float4 g;
            g.x = gelu_f(fmaf(v.x, k[0].scale, k[0].shift));
            g.y = gelu_f(fmaf(v.y, k[1].scale, k[1].shift));
            g.z = gelu_f(fmaf(v.z, k[2].scale, k[2].shift));
            g.w = gelu_f(fmaf(v.w, k[3].scale, k[3].shift));
            st4(ptr, g);
        }
        __syncthreads();
        // one column per thread, 8 output rows
#pragma unroll 2
        for (int r = 0; r < DW_TH; ++r) {
            float4 o = make_float4(0, 0, 0, 0);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float4 t = ld4(tile + ((r + i) * DW_HC + col + j) * DW_CC + cq * 4);
                    o.x = fmaf(w[i * 3 + j][0], t.x, o.x);
                    o.y = fmaf(w[i * 3 + j][1], t.y, o.y);
                    o.z = fmaf(w[i * 3 + j][2], t.z, o.z);
                    o.w = fmaf(w[i * 3 + j][3], t.w, o.w);
                }
            st4(dst + ((size_t)(y0 + r) * W + x0 + col) * C + c0, o);
            s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
            q.x += o.x * o.x; q.y += o.y * o.y; q.z += o.z * o.z; q.w += o.w * o.w;
        }
    }
    reduce_pt_atomic2(s, q, stats2 + ((size_t)n * C + cbase) * 2, smem);
}

// B3a (pointwise, in place): du <- dh2 = a2*dz2 + b2*h2 + c2 with dz2 = (du*s + dpool/P) * gelu'(h2*scale2 + shift2):
// the SE-gate backward and the Norm2 backward of one element.  Splitting this off the stencil kernel keeps the stencil
// kernel's halo tile free of arithmetic (it is filled by cp.async) and both kernels out of register spills.
__global__ void __launch_bounds__(256) dh2_kernel(float* __restrict__ du, const float* __restrict__ h2, const float* __restrict__ gate,
                                                   const float* __restrict__ dmp, const Coef* __restrict__ coef2,
                                                   const BCoef* __restrict__ bc2, int P, int chunk) {
    constexpr int C = UB_HID, Q = C / 4, ROWS = 256 / Q;
    const int n = blockIdx.y, c4 = threadIdx.x % Q, r = threadIdx.x / Q;
    Coef k[4]; BCoef b[4]; float sg[4], dm[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const size_t ci = (size_t)n * C + c4 * 4 + i;
        k[i] = coef2[ci]; b[i] = bc2[ci]; sg[i] = gate[ci]; dm[i] = dmp[ci];
    }
    const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, P);
    const size_t base = (size_t)n * P * C + c4 * 4;
#pragma unroll 4
    for (int p = p0 + r; p < p1; p += ROWS) {
        const float4 d = ld4(du + base + (size_t)p * C);
        const float4 hv = ld4_stream(h2 + base + (size_t)p * C);
        const float dv[4] = {d.x, d.y, d.z, d.w}, hh[4] = {hv.x, hv.y, hv.z, hv.w};
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float z = fmaf(hh[i], k[i].scale, k[i].shift);
            const float dz = fmaf(dv[i], sg[i], dm[i]) * gelu_grad_f(z);
            o[i] = fmaf(b[i].a, dz, fmaf(b[i].b, hh[i], b[i].c));
        }
        st4(du + base + (size_t)p * C, make_float4(o[0], o[1], o[2], o[3]));
    }
}

// per-chunk coefficient block staged in shared memory for the backward kernel (one float4 per channel and kind)
struct DwBwdCoef {
    float4 k1m1[DW_CC];   // scale1, shift1, mean1, rstd1
    float4 w[9][DW_Q];    // depthwise taps [tap][channel quad]: read per use (LDS.128), keeps the stencil loop at <= 85 registers
};

// B3b (stencil).  Both the input gradient and the weight gradient are written per INPUT pixel p of the convolution:
//   D[tap] = sum of dh2[q] over all outputs q that read p through `tap`   (normally the single q = p - off(tap);
//            on rows/cols 1 and H-2/W-2 also the border output that reaches p through reflect padding)
//   dg1[p] = sum_tap w[tap] * D[tap]         dW[tap] += g1[p] * D[tap]         dz1[p] = dg1[p] * gelu'(z1[p])
// so gelu(z1) and gelu'(z1) are evaluated once per pixel (no halo arithmetic at all) and the only shared-memory tile
// is the raw dh2 halo tile, filled by cp.async with zero-fill outside the image.
__global__ void __launch_bounds__(256, 2)
dwconv_bwd_kernel(const float* __restrict__ dh2, const float* __restrict__ h1, const Coef* __restrict__ coef1,
                  const MeanRstd* __restrict__ mr1, const float* __restrict__ wdw, float* __restrict__ dz1,
                  double* bstats1, float* dwdw /* [256][9] accumulated */, int H, int W) {
    extern __shared__ __align__(16) float smem[];
    float* tdh = smem;                       // dh2 halo tile, zero outside the image
    DwBwdCoef* cf = reinterpret_cast<DwBwdCoef*>(smem + 2 * DW_TILE_FLOATS);
    constexpr int C = UB_HID;
    constexpr int ROW = DW_HC * DW_CC;       // floats per halo row
    const int n = blockIdx.z, cbase = blockIdx.y * DW_CC, x0 = blockIdx.x * DW_TW;
    const int cq = threadIdx.x % DW_Q, col = threadIdx.x / DW_Q;
    const int c0 = cbase + cq * 4;
    if (threadIdx.x < DW_CC) {
        const size_t ci = (size_t)n * C + cbase + threadIdx.x;
        const Coef b = coef1[ci];
        const MeanRstd m = mr1[ci];
        cf->k1m1[threadIdx.x] = make_float4(b.scale, b.shift, m.mean, m.rstd);
    }
    for (int e = threadIdx.x; e < 9 * DW_CC; e += 256) {
        const int j = e / DW_CC, ch = e % DW_CC;
        reinterpret_cast<float*>(&cf->w[j][0])[ch] = wdw[(size_t)(cbase + ch) * 9 + j];
    }
    const size_t fbase = (size_t)n * H * W * C;
    float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
    float4 gw[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) gw[j] = make_float4(0, 0, 0, 0);
    const int qx = x0 + col;
    const bool x_lo = (qx == 1), x_hi = (qx == W - 2);
    auto fetch = [&](int y0, float* buf) {          // raw dh2 halo tile, zero-filled outside the image
        for (int e0 = threadIdx.x; e0 < DW_HR * DW_HC * DW_Q; e0 += 256) {
            const int pix = e0 / DW_Q, ry = pix / DW_HC, rx = pix - ry * DW_HC;
            const int yy = y0 - 1 + ry, xx = x0 - 1 + rx;
            const bool inside = (yy >= 0 && yy < H && xx >= 0 && xx < W);
            const size_t off = fbase + ((size_t)(inside ? yy : 0) * W + (inside ? xx : 0)) * C + c0;
            cp_async16(buf + pix * DW_CC + cq * 4, dh2 + off, inside);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    fetch(0, tdh);
    for (int y0 = 0, it = 0; y0 < H; y0 += DW_TH, ++it) {
        const float* tcol = tdh + (it & 1) * DW_TILE_FLOATS + (col + 1) * DW_CC + cq * 4;   // this thread's column, halo row 0
        // h1 of this thread's column, software-prefetched one row ahead (first row in flight across the barrier)
        float4 h1next = ld4(h1 + fbase + ((size_t)y0 * W + qx) * C + c0);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                               // tile landed; everybody is done with the other buffer
        if (y0 + DW_TH < H) fetch(y0 + DW_TH, tdh + ((it + 1) & 1) * DW_TILE_FLOATS);
        const float4 km0 = cf->k1m1[cq * 4 + 0], km1 = cf->k1m1[cq * 4 + 1], km2 = cf->k1m1[cq * 4 + 2], km3 = cf->k1m1[cq * 4 + 3];
#pragma unroll 1
        for (int r = 0; r < DW_TH; ++r) {
            const int qy = y0 + r;
            const float4 hv = h1next;
            if (r + 1 < DW_TH) h1next = ld4(h1 + fbase + ((size_t)(qy + 1) * W + qx) * C + c0);
            float4 g, gp;
            gelu_both(fmaf(hv.x, km0.x, km0.y), g.x, gp.x);
            gelu_both(fmaf(hv.y, km1.x, km1.y), g.y, gp.y);
            gelu_both(fmaf(hv.z, km2.x, km2.y), g.z, gp.z);
            gelu_both(fmaf(hv.w, km3.x, km3.y), g.w, gp.w);
            const float* tc = tcol + (r + 1) * ROW;            // centre of the 3x3 neighbourhood of p
            const bool y_lo = (qy == 1), y_hi = (qy == H - 2);
            const bool border = y_lo | y_hi | x_lo | x_hi;
            float4 o = make_float4(0, 0, 0, 0);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    // tap (i, j) has offset (i-1, j-1); the regular reader of p through it is q = p - off
                    float4 d = ld4(tc + (1 - i) * ROW + (1 - j) * DW_CC);
                    if (border) {   // rare: adjoint of reflect padding, the mirrored reader q = p + off
                        const bool ya = (i == 0 && y_lo) || (i == 2 && y_hi);
                        const bool xa = (j == 0 && x_lo) || (j == 2 && x_hi);
                        if (ya) { const float4 e = ld4(tc + (i - 1) * ROW + (1 - j) * DW_CC); d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w; }
                        if (xa) { const float4 e = ld4(tc + (1 - i) * ROW + (j - 1) * DW_CC); d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w; }
                        if (ya && xa) { const float4 e = ld4(tc + (i - 1) * ROW + (j - 1) * DW_CC); d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w; }
                    }
                    const float4 ww = cf->w[i * 3 + j][cq];
                    o.x = fmaf(ww.x, d.x, o.x); o.y = fmaf(ww.y, d.y, o.y);
                    o.z = fmaf(ww.z, d.z, o.z); o.w = fmaf(ww.w, d.w, o.w);
                    float4& a = gw[i * 3 + j];
                    a.x = fmaf(g.x, d.x, a.x); a.y = fmaf(g.y, d.y, a.y);
                    a.z = fmaf(g.z, d.z, a.z); a.w = fmaf(g.w, d.w, a.w);
                }
            const float4 dz = make_float4(o.x * gp.x, o.y * gp.y, o.z * gp.z, o.w * gp.w);
            st4(dz1 + fbase + ((size_t)qy * W + qx) * C + c0, dz);
            s.x += dz.x; s.y += dz.y; s.z += dz.z; s.w += dz.w;
            q.x = fmaf(dz.x, (hv.x - km0.z) * km0.w, q.x);
            q.y = fmaf(dz.y, (hv.y - km1.z) * km1.w, q.y);
            q.z = fmaf(dz.z, (hv.z - km2.z) * km2.w, q.z);
            q.w = fmaf(dz.w, (hv.w - km3.z) * km3.w, q.w);
        }
    }
    reduce_pt_atomic2(s, q, bstats1 + ((size_t)n * C + cbase) * 2, smem);
    // depthwise weight gradient: reduce over the 32 pixel-threads, one atomic per (channel, tap) per CTA
    float* red = smem;  // [9][32 pt][32 ch]
#pragma unroll
    for (int j = 0; j < 9; ++j) st4(red + (j * DW_PT + col) * DW_CC + cq * 4, gw[j]);
    __syncthreads();
    for (int e = threadIdx.x; e < 9 * DW_CC; e += 256) {
        const int j = e / DW_CC, ch = e % DW_CC;
        float t = 0.f;
#pragma unroll 8
        for (int r = 0; r < DW_PT; ++r) t += red[(j * DW_PT + r) * DW_CC + ch];
        atomicAdd(&dwdw[(size_t)(cbase + ch) * 9 + j], t);
    }
}

// B3 fused (optional, ub200_dwconv_set_bwd_split(0); currently slower than the two-kernel path): one kernel, 4 Hh of HBM traffic instead of 6.  512 threads (16 warps), 1 CTA per SM.
// The raw du and h2 halo tiles of the NEXT 8-row tile arrive by cp.async (zero-filled outside the image) into the second
// pair of shared-memory buffers while the current pair is turned into dh2 in place (SE-gate backward, gelu'(z2), Norm2
// backward; forced to 0 outside the image) and consumed by the input-stationary stencil described above.
struct DwFusedCoef {
    float4 k2sg[DW_CC];   // scale2, shift2, gate, dpool/P
    float4 b2[DW_CC];     // a2, b2, c2, -
    float4 k1m1[DW_CC];   // scale1, shift1, mean1, rstd1
    float4 w[9][DW_Q];    // depthwise taps [tap][channel quad]
};

__global__ void __launch_bounds__(512, 1)
dwconv_bwd_fused_kernel(const float* __restrict__ du, const float* __restrict__ h2, const float* __restrict__ h1,
                        const float* __restrict__ gate, const float* __restrict__ dmp, const Coef* __restrict__ coef2,
                        const BCoef* __restrict__ bc2, const Coef* __restrict__ coef1, const MeanRstd* __restrict__ mr1,
                        const float* __restrict__ wdw, float* __restrict__ dz1, double* bstats1, float* dwdw, int H, int W) {
    extern __shared__ __align__(16) float smem[];
    // buffers: [2 stages][du tile | h2 tile]
    DwFusedCoef* cf = reinterpret_cast<DwFusedCoef*>(smem + 4 * DW_TILE_FLOATS);
    constexpr int C = UB_HID;
    constexpr int ROW = DW_HC * DW_CC;
    constexpr int NITEM = DW_HR * DW_HC * DW_Q;
    const int n = blockIdx.z, cbase = blockIdx.y * DW_CC, x0 = blockIdx.x * DW_TW;
    const int tid = threadIdx.x;
    const int cq = tid % DW_Q, col = (tid / DW_Q) % DW_TW, rh = tid / 256;     // channel quad, column, row half (4 rows)
    const int c0 = cbase + cq * 4;
    if (tid < DW_CC) {
        const size_t ci = (size_t)n * C + cbase + tid;
        const Coef a = coef2[ci], b = coef1[ci];
        const BCoef bb = bc2[ci];
        const MeanRstd m = mr1[ci];
        cf->k2sg[tid] = make_float4(a.scale, a.shift, gate[ci], dmp[ci]);
        cf->b2[tid] = make_float4(bb.a, bb.b, bb.c, 0.f);
        cf->k1m1[tid] = make_float4(b.scale, b.shift, m.mean, m.rstd);
    }
    for (int e = tid; e < 9 * DW_CC; e += 512) {
        const int j = e / DW_CC, ch = e % DW_CC;
        reinterpret_cast<float*>(&cf->w[j][0])[ch] = wdw[(size_t)(cbase + ch) * 9 + j];
    }
    const size_t fbase = (size_t)n * H * W * C;
    float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
    float4 gw[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) gw[j] = make_float4(0, 0, 0, 0);
    const int qx = x0 + col;
    const bool x_lo = (qx == 1), x_hi = (qx == W - 2);

    auto fetch = [&](int y0, float* bufd, float* bufh) {
        for (int e = tid; e < NITEM; e += 512) {
            const int pix = e / DW_Q, ry = pix / DW_HC, rx = pix - ry * DW_HC;
            const int yy = y0 - 1 + ry, xx = x0 - 1 + rx;
            const bool inside = (yy >= 0 && yy < H && xx >= 0 && xx < W);
            const size_t off = fbase + ((size_t)(inside ? yy : 0) * W + (inside ? xx : 0)) * C + c0;
            cp_async16(bufd + pix * DW_CC + cq * 4, du + off, inside);
            cp_async16(bufh + pix * DW_CC + cq * 4, h2 + off, inside);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    fetch(0, smem, smem + DW_TILE_FLOATS);
    for (int y0 = 0, it = 0; y0 < H; y0 += DW_TH, ++it) {
        float* td = smem + (it & 1) * 2 * DW_TILE_FLOATS;      // du tile -> dh2 tile (in place)
        const float* th = td + DW_TILE_FLOATS;                  // h2 tile
        // h1 of this thread's column / row half, software-prefetched one row ahead
        float4 h1next = ld4(h1 + fbase + ((size_t)(y0 + rh * 4) * W + qx) * C + c0);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                        // tiles landed; everybody is done with the other stage
        if (y0 + DW_TH < H) fetch(y0 + DW_TH, smem + ((it + 1) & 1) * 2 * DW_TILE_FLOATS, smem + ((it + 1) & 1) * 2 * DW_TILE_FLOATS + DW_TILE_FLOATS);
        // ---- dh2 in place (e % 8 == cq): dz2 = (du*s + dpool/P) * gelu'(z2); dh2 = a2*dz2 + b2*h2 + c2; 0 outside the image ----
        {
            const float4 k2a = cf->k2sg[cq * 4 + 0], k2b = cf->k2sg[cq * 4 + 1], k2c = cf->k2sg[cq * 4 + 2], k2d = cf->k2sg[cq * 4 + 3];
            const float4 b2a = cf->b2[cq * 4 + 0], b2b = cf->b2[cq * 4 + 1], b2c = cf->b2[cq * 4 + 2], b2d = cf->b2[cq * 4 + 3];
            for (int e = tid; e < NITEM; e += 512) {
                const int pix = e / DW_Q, ry = pix / DW_HC, rx = pix - ry * DW_HC;
                const int yy = y0 - 1 + ry, xx = x0 - 1 + rx;
                float* ptr = td + pix * DW_CC + cq * 4;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                    const float4 d = ld4(ptr);
                    const float4 hv = ld4(th + pix * DW_CC + cq * 4);
                    float4 o;
                    o.x = fmaf(b2a.x, fmaf(d.x, k2a.z, k2a.w) * gelu_grad_f(fmaf(hv.x, k2a.x, k2a.y)), fmaf(b2a.y, hv.x, b2a.z));
                    o.y = fmaf(b2b.x, fmaf(d.y, k2b.z, k2b.w) * gelu_grad_f(fmaf(hv.y, k2b.x, k2b.y)), fmaf(b2b.y, hv.y, b2b.z));
                    o.z = fmaf(b2c.x, fmaf(d.z, k2c.z, k2c.w) * gelu_grad_f(fmaf(hv.z, k2c.x, k2c.y)), fmaf(b2c.y, hv.z, b2c.z));
                    o.w = fmaf(b2d.x, fmaf(d.w, k2d.z, k2d.w) * gelu_grad_f(fmaf(hv.w, k2d.x, k2d.y)), fmaf(b2d.y, hv.w, b2d.z));
                    st4(ptr, o);
                }   // else: already zero-filled by cp.async
            }
        }
        __syncthreads();
        // ---- input-stationary stencil: 4 rows of this thread's column ----
        const float4 km0 = cf->k1m1[cq * 4 + 0], km1 = cf->k1m1[cq * 4 + 1], km2 = cf->k1m1[cq * 4 + 2], km3 = cf->k1m1[cq * 4 + 3];
        const float* tcol = td + (col + 1) * DW_CC + cq * 4;
#pragma unroll 1
        for (int rr = 0; rr < 4; ++rr) {
            const int r = rh * 4 + rr, qy = y0 + r;
            const float4 hv = h1next;
            if (rr + 1 < 4) h1next = ld4(h1 + fbase + ((size_t)(qy + 1) * W + qx) * C + c0);
            float4 g, gp;
            gelu_both(fmaf(hv.x, km0.x, km0.y), g.x, gp.x);
            gelu_both(fmaf(hv.y, km1.x, km1.y), g.y, gp.y);
            gelu_both(fmaf(hv.z, km2.x, km2.y), g.z, gp.z);
            gelu_both(fmaf(hv.w, km3.x, km3.y), g.w, gp.w);
            const float* tc = tcol + (r + 1) * ROW;
            const bool y_lo = (qy == 1), y_hi = (qy == H - 2);
            const bool border = y_lo | y_hi | x_lo | x_hi;
            float4 o = make_float4(0, 0, 0, 0);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    float4 d = ld4(tc + (1 - i) * ROW + (1 - j) * DW_CC);
                    if (border) {
                        const bool ya = (i == 0 && y_lo) || (i == 2 && y_hi);
                        const bool xa = (j == 0 && x_lo) || (j == 2 && x_hi);
                        if (ya) { const float4 e = ld4(tc + (i - 1) * ROW + (1 - j) * DW_CC); d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w; }
                        if (xa) { const float4 e = ld4(tc + (1 - i) * ROW + (j - 1) * DW_CC); d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w; }
                        if (ya && xa) { const float4 e = ld4(tc + (i - 1) * ROW + (j - 1) * DW_CC); d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w; }
                    }
                    const float4 ww = cf->w[i * 3 + j][cq];
                    o.x = fmaf(ww.x, d.x, o.x); o.y = fmaf(ww.y, d.y, o.y);
                    o.z = fmaf(ww.z, d.z, o.z); o.w = fmaf(ww.w, d.w, o.w);
                    float4& a = gw[i * 3 + j];
                    a.x = fmaf(g.x, d.x, a.x); a.y = fmaf(g.y, d.y, a.y);
                    a.z = fmaf(g.z, d.z, a.z); a.w = fmaf(g.w, d.w, a.w);
                }
            const float4 dz = make_float4(o.x * gp.x, o.y * gp.y, o.z * gp.z, o.w * gp.w);
            st4(dz1 + fbase + ((size_t)qy * W + qx) * C + c0, dz);
            s.x += dz.x; s.y += dz.y; s.z += dz.z; s.w += dz.w;
            q.x = fmaf(dz.x, (hv.x - km0.z) * km0.w, q.x);
            q.y = fmaf(dz.y, (hv.y - km1.z) * km1.w, q.y);
            q.z = fmaf(dz.z, (hv.z - km2.z) * km2.w, q.z);
            q.w = fmaf(dz.w, (hv.w - km3.z) * km3.w, q.w);
        }
    }
    // statistics and weight gradient: reduce over the 64 (column, row-half) threads of each channel quad
    __syncthreads();
    {
        float4* sa = reinterpret_cast<float4*>(smem);
        float4* sb = sa + 512;
        const int pt = tid / DW_Q;                       // 0..63
        sa[pt * DW_Q + cq] = s;
        sb[pt * DW_Q + cq] = q;
        __syncthreads();
        if (tid < 2 * DW_CC) {
            const int which = tid / DW_CC, ch = tid % DW_CC;
            const float* src = reinterpret_cast<const float*>(which ? sb : sa);
            double t = 0.0;
#pragma unroll 8
            for (int r = 0; r < 64; ++r) t += (double)src[r * DW_CC + ch];
            atomicAdd(&bstats1[((size_t)n * C + cbase + ch) * 2 + which], t);
        }
        __syncthreads();
        float* red = smem;  // [9][64 pt][32 ch] = 73.7 KB
#pragma unroll
        for (int j = 0; j < 9; ++j) st4(red + (j * 64 + pt) * DW_CC + cq * 4, gw[j]);
        __syncthreads();
        for (int e = tid; e < 9 * DW_CC; e += 512) {
            const int j = e / DW_CC, ch = e % DW_CC;
            float t = 0.f;
#pragma unroll 8
            for (int r = 0; r < 64; ++r) t += red[(j * 64 + r) * DW_CC + ch];
            atomicAdd(&dwdw[(size_t)(cbase + ch) * 9 + j], t);
        }
    }
}

// dwconv_set_mode(): bit 0 = row-streaming forward (dwconv_rows.cu), bit 1 = row-streaming fused backward,
// bit 2 = packed FFMA2 arithmetic in those kernels, bit 3 = packed f32x2 GELU, bit 5 = 512-thread channel-pair backward;
// 0 = the tile kernels of this file.  Default 47 (packed GELU: another 3-4 %, channel-pair backward: another 2 %): measured at B=16,
// 256x256 forward 4.44 -> 3.38 ms, backward 14.0 -> 8.1 ms per step against the tile kernels.
static int g_dw_mode = 47;
int dwconv_set_mode(int mode) { g_dw_mode = mode; return UB_OK; }

int launch_dwconv_fwd(const float* h1, const Coef* coef1, const float* wdw, float* h2, double* stats2, int N, int H, int W,
                      cudaStream_t st) {
    if (g_dw_mode & 1) return launch_dwrows_fwd(h1, coef1, wdw, h2, stats2, N, H, W, (g_dw_mode >> 2) & 15, st);
    if (W % DW_TW != 0 || H % DW_TH != 0) return UB_ERR_ARG;
    constexpr size_t smem = (size_t)2 * DW_TILE_FLOATS * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(dwconv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return UB_ERR_CUDA;
        attr_set = true;
    }
    dwconv_fwd_kernel<<<dim3(W / DW_TW, UB_HID / DW_CC, N), 256, smem, st>>>(h1, coef1, wdw, h2, stats2, H, W);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

static int g_dw_bwd_split = 1;     // dwconv_set_bwd_split(): 1 = pointwise dh2 kernel + stencil kernel (default; measured 13.9 ms vs
                                   // 19.6 ms per step at B=16), 0 = fused kernel
int dwconv_set_bwd_split(int on) { g_dw_bwd_split = on ? 1 : 0; return UB_OK; }

int launch_dwconv_bwd(float* du, const float* h2, const float* h1, const float* gate, const float* dmp, const Coef* coef2,
                      const BCoef* bc2, const Coef* coef1, const MeanRstd* mr1, const float* wdw, float* dz1, double* bstats1,
                      float* dwdw, int N, int H, int W, cudaStream_t st) {
    if (g_dw_mode & 2)
        return launch_dwrows_bwd(du, h2, h1, gate, dmp, coef2, bc2, coef1, mr1, wdw, dz1, bstats1, dwdw, N, H, W, (g_dw_mode >> 2) & 15, st);
    if (W % DW_TW != 0 || H % DW_TH != 0) return UB_ERR_ARG;
    if (!g_dw_bwd_split) {
        constexpr size_t smem = (size_t)4 * DW_TILE_FLOATS * sizeof(float) + sizeof(DwFusedCoef);
        static bool attr_set = false;
        if (!attr_set) {
            if (cudaFuncSetAttribute(dwconv_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return UB_ERR_CUDA;
            attr_set = true;
        }
        dwconv_bwd_fused_kernel<<<dim3(W / DW_TW, UB_HID / DW_CC, N), 512, smem, st>>>(du, h2, h1, gate, dmp, coef2, bc2, coef1, mr1, wdw,
                                                                                     dz1, bstats1, dwdw, H, W);
        UB_CHECK_LAUNCH();
        return UB_OK;
    }
    const int P = H * W;
    const int chunk = P >= 4096 ? 1024 : (P >= 1024 ? 256 : 64);
    dh2_kernel<<<dim3((P + chunk - 1) / chunk, N), 256, 0, st>>>(du, h2, gate, dmp, coef2, bc2, P, chunk);   // du <- dh2 in place
    UB_CHECK_LAUNCH();
    constexpr size_t smem = (size_t)2 * DW_TILE_FLOATS * sizeof(float) + sizeof(DwBwdCoef);
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(dwconv_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return UB_ERR_CUDA;
        attr_set = true;
    }
    dwconv_bwd_kernel<<<dim3(W / DW_TW, UB_HID / DW_CC, N), 256, smem, st>>>(du, h1, coef1, mr1, wdw, dz1, bstats1, dwdw, H, W);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

}  // namespace ub
