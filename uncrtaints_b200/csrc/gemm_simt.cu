// FP32 CUDA-core streaming GEMMs for the MBConv 1x1 convolutions (model/src/backbones/uncrtaints.py:126,136)
// and their backward passes, with the normalisation / activation / SE-gate applied in the operand
// prologue and the next normalisation's statistics gathered in the epilogue.
//
//   C[64 px x NOUT] = f(A)[64 px x K] * Wt[K x NOUT]        (weights resident in shared memory)
//
// One CTA (256 threads, 1 per SM: 160-195 KB shared memory) keeps the whole K x NOUT weight matrix in
// shared memory and streams 64-pixel tiles of one frame through it.  Thread (tx, ty) = (tid%16, tid/16)
// owns pixels ty*4..ty*4+3 and the float4 column groups tx*4 + 64*j.
//
// The weight-gradient kernel reduces over pixels instead: dW[128 x 256] = sum_p fa(p)[128] (x) fb(p)[256],
// one partial per CTA, summed by reduce_partials_kernel.
//
// This is the exact-fp32 path (and the reference for the tcgen05 bf16x3 path in gemm_tc.cu).
#include "common.cuh"
#include "kernels.h"

namespace ub {

// ------------------------------------------------------------------------------------------
// operand loaders: produce the transformed A element quad for (frame n, pixel p, channel quad kq)
// ------------------------------------------------------------------------------------------
struct LoadNormed {            // K1: a = x*scale0 + shift0
    const float* x; const Coef* coef; int C;
    Coef k[4];
    __device__ void init(int n, int kq) {
#pragma unroll
        for (int i = 0; i < 4; ++i) k[i] = coef[(size_t)n * C + kq * 4 + i];
    }
    __device__ float4 load(size_t row, int kq) const {
        const float4 v = ld4_stream(x + row * C + kq * 4);
        return make_float4(fmaf(v.x, k[0].scale, k[0].shift), fmaf(v.y, k[1].scale, k[1].shift),
                           fmaf(v.z, k[2].scale, k[2].shift), fmaf(v.w, k[3].scale, k[3].shift));
    }
};
struct LoadGeluGate {          // K4: a = gelu(h2*scale2 + shift2) * s[n][k]
    const float* h2; const Coef* coef; const float* gate; int C;
    Coef k[4]; float g[4];
    __device__ void init(int n, int kq) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { k[i] = coef[(size_t)n * C + kq * 4 + i]; g[i] = gate[(size_t)n * C + kq * 4 + i]; }
    }
    __device__ float4 load(size_t row, int kq) const {
        const float4 v = ld4_stream(h2 + row * C + kq * 4);
        return make_float4(gelu_f(fmaf(v.x, k[0].scale, k[0].shift)) * g[0], gelu_f(fmaf(v.y, k[1].scale, k[1].shift)) * g[1],
                           gelu_f(fmaf(v.z, k[2].scale, k[2].shift)) * g[2], gelu_f(fmaf(v.w, k[3].scale, k[3].shift)) * g[3]);
    }
};
struct LoadNormBwd {           // B5b / B2b: a = ca*dy + cb*v + cc   (normalisation backward applied on the fly)
    const float* dy; const float* v; const BCoef* bc; int C;
    BCoef k[4];
    __device__ void init(int n, int kq) {
#pragma unroll
        for (int i = 0; i < 4; ++i) k[i] = bc[(size_t)n * C + kq * 4 + i];
    }
    __device__ float4 load(size_t row, int kq) const {
        const float4 d = ld4_stream(dy + row * C + kq * 4);
        const float4 w = ld4(v + row * C + kq * 4);
        return make_float4(fmaf(k[0].a, d.x, fmaf(k[0].b, w.x, k[0].c)), fmaf(k[1].a, d.y, fmaf(k[1].b, w.y, k[1].c)),
                           fmaf(k[2].a, d.z, fmaf(k[2].b, w.z, k[2].c)), fmaf(k[3].a, d.w, fmaf(k[3].b, w.w, k[3].c)));
    }
};

// ------------------------------------------------------------------------------------------
// epilogues: store the output quad and accumulate NS per-column statistics
// ------------------------------------------------------------------------------------------
struct EpiStoreStats {         // K1 / K4: raw output + (sum, sumsq)
    static constexpr int NS = 2;
    float* out; double* stats; int NOUT;
    __device__ void apply(int n, size_t row, int col, float4 v, float4* s) const {
        st4(out + row * NOUT + col, v);
        s[0].x += v.x; s[0].y += v.y; s[0].z += v.z; s[0].w += v.w;
        s[1].x += v.x * v.x; s[1].y += v.y * v.y; s[1].z += v.z * v.z; s[1].w += v.w * v.w;
    }
    __device__ double* dst(int n) const { return stats + (size_t)n * NOUT * NS; }
};
struct EpiGemm2Bwd {           // B5b: du = acc; sums (du*g2, du*gp2, du*gp2*h2hat) with g2 = gelu(z2), gp2 = gelu'(z2)
    static constexpr int NS = 3;
    float* du; const float* h2; const Coef* coef2; const MeanRstd* mr2; double* sums;
    __device__ void apply(int n, size_t row, int col, float4 v, float4* s) const {
        st4(du + row * UB_HID + col, v);
        const float4 h = ld4(h2 + row * UB_HID + col);
        const float hv[4] = {h.x, h.y, h.z, h.w};
        const float dv[4] = {v.x, v.y, v.z, v.w};
        float r0[4], r1[4], r2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const Coef k = ldg(&coef2[(size_t)n * UB_HID + col + i]);
            const MeanRstd m = ldg(&mr2[(size_t)n * UB_HID + col + i]);
            const float z = fmaf(hv[i], k.scale, k.shift);
            float gz, gp;
            gelu_both(z, gz, gp);
            r0[i] = dv[i] * gz;
            r1[i] = dv[i] * gp;
            r2[i] = dv[i] * gp * (hv[i] - m.mean) * m.rstd;
        }
        s[0].x += r0[0]; s[0].y += r0[1]; s[0].z += r0[2]; s[0].w += r0[3];
        s[1].x += r1[0]; s[1].y += r1[1]; s[1].z += r1[2]; s[1].w += r1[3];
        s[2].x += r2[0]; s[2].y += r2[1]; s[2].z += r2[2]; s[2].w += r2[3];
    }
    __device__ double* dst(int n) const { return sums + (size_t)n * UB_HID * NS; }
};
struct EpiGemm1Bwd {           // B2b: dn0 = acc; sums (dn0, dn0 * x_hat)
    static constexpr int NS = 2;
    float* dn0; const float* x; const MeanRstd* mr0; double* bstats;
    __device__ void apply(int n, size_t row, int col, float4 v, float4* s) const {
        st4(dn0 + row * UB_WIDTH + col, v);
        const float4 xv = ld4(x + row * UB_WIDTH + col);
        const MeanRstd m0 = ldg(&mr0[(size_t)n * UB_WIDTH + col + 0]), m1 = ldg(&mr0[(size_t)n * UB_WIDTH + col + 1]),
                       m2 = ldg(&mr0[(size_t)n * UB_WIDTH + col + 2]), m3 = ldg(&mr0[(size_t)n * UB_WIDTH + col + 3]);
        s[0].x += v.x; s[0].y += v.y; s[0].z += v.z; s[0].w += v.w;
        s[1].x += v.x * (xv.x - m0.mean) * m0.rstd;
        s[1].y += v.y * (xv.y - m1.mean) * m1.rstd;
        s[1].z += v.z * (xv.z - m2.mean) * m2.rstd;
        s[1].w += v.w * (xv.w - m3.mean) * m3.rstd;
    }
    __device__ double* dst(int n) const { return bstats + (size_t)n * UB_WIDTH * NS; }
};

// ------------------------------------------------------------------------------------------
// streaming GEMM kernel
// ------------------------------------------------------------------------------------------
template <int K, int NOUT, class ALoad, class Epi>
__global__ void __launch_bounds__(256, 1)
gemm_px_kernel(ALoad al, const float* __restrict__ Wt /* [K][NOUT] */, Epi ep, int P, int tiles_per_block) {
    constexpr int TM = 64, PITCH = K + 4, J = NOUT / 64, Q = K / 4, RPP = 256 / Q;
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;                 // K*NOUT
    float* As = smem + K * NOUT;      // TM*PITCH
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16, n = blockIdx.y;

    for (int i = tid; i < K * NOUT / 4; i += 256) reinterpret_cast<float4*>(Ws)[i] = ld4(Wt + (size_t)i * 4);
    const int kq = tid % Q, r0 = tid / Q;
    al.init(n, kq);

    float4 st[Epi::NS][J];
#pragma unroll
    for (int s = 0; s < Epi::NS; ++s)
#pragma unroll
        for (int j = 0; j < J; ++j) st[s][j] = make_float4(0, 0, 0, 0);

    const int tiles_per_frame = P / TM;
    const int t0 = blockIdx.x * tiles_per_block;
    const int t1 = min(t0 + tiles_per_block, tiles_per_frame);
    for (int t = t0; t < t1; ++t) {
        const size_t row0 = (size_t)n * P + (size_t)t * TM;
        __syncthreads();  // previous tile fully consumed (and weights visible on the first pass)
#pragma unroll 4
        for (int r = r0; r < TM; r += RPP) st4(As + r * PITCH + kq * 4, al.load(row0 + r, kq));
        __syncthreads();

        float4 acc[4][J];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < J; ++j) acc[i][j] = make_float4(0, 0, 0, 0);

#pragma unroll 2
        for (int k4 = 0; k4 < K; k4 += 4) {
            float a[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = ld4(As + (ty * 4 + i) * PITCH + k4);
                a[i][0] = v.x; a[i][1] = v.y; a[i][2] = v.z; a[i][3] = v.w;
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const float4 w = ld4(Ws + (k4 + kk) * NOUT + tx * 4 + 64 * j);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[i][j].x = fmaf(a[i][kk], w.x, acc[i][j].x);
                        acc[i][j].y = fmaf(a[i][kk], w.y, acc[i][j].y);
                        acc[i][j].z = fmaf(a[i][kk], w.z, acc[i][j].z);
                        acc[i][j].w = fmaf(a[i][kk], w.w, acc[i][j].w);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < J; ++j) {
            float4 s[Epi::NS];
#pragma unroll
            for (int q = 0; q < Epi::NS; ++q) s[q] = st[q][j];
#pragma unroll
            for (int i = 0; i < 4; ++i) ep.apply(n, row0 + ty * 4 + i, tx * 4 + 64 * j, acc[i][j], s);
#pragma unroll
            for (int q = 0; q < Epi::NS; ++q) st[q][j] = s[q];
        }
    }

    // column statistics: reduce the per-thread partials over ty through shared memory (weights are dead)
    double* dst = ep.dst(n);
    float* red = smem;  // [16][NOUT]
#pragma unroll
    for (int s = 0; s < Epi::NS; ++s) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < J; ++j) st4(red + ty * NOUT + tx * 4 + 64 * j, st[s][j]);
        __syncthreads();
        for (int c = tid; c < NOUT; c += 256) {
            double tsum = 0.0;
#pragma unroll
            for (int r = 0; r < 16; ++r) tsum += (double)red[r * NOUT + c];
            atomicAdd(&dst[(size_t)c * Epi::NS + s], tsum);
        }
    }
}

template <int K, int NOUT, class ALoad, class Epi>
static int launch_gemm(ALoad al, const float* Wt, Epi ep, int N, int P, cudaStream_t st) {
    if (P % 64 != 0) return UB_ERR_ARG;
    constexpr size_t smem = (size_t)(K * NOUT + 64 * (K + 4)) * sizeof(float);
    auto kern = gemm_px_kernel<K, NOUT, ALoad, Epi>;
    UB_SET_SMEM(kern, smem);
    const int tiles = P / 64;
    const int tpb = tiles >= 16 ? 16 : tiles;
    kern<<<dim3((tiles + tpb - 1) / tpb, N), 256, smem, st>>>(al, Wt, ep, P, tpb);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

// ------------------------------------------------------------------------------------------
// weight gradient: partial[blk][ia*sa + ib*sb] = sum over the block's pixels of fa(p)[ia] * fb(p)[ib]
// fa has 128 columns, fb has 256 columns.
// ------------------------------------------------------------------------------------------
template <class LA, class LB, int CB = 256>
__global__ void __launch_bounds__(256, 1)
wgrad_kernel(LA la, LB lb, float* __restrict__ partial, int P, long long total_tiles, int sa, int sb) {
    constexpr int TM = 64, CA = 128, JB = CB / 64, QB = CB / 4, RB = 256 / QB;
    extern __shared__ __align__(16) float smem[];
    float* As = smem;             // TM*CA
    float* Bs = smem + TM * CA;   // TM*CB
    const int tid = threadIdx.x, tb = tid % 16, ta = tid / 16;
    float4 acc[8][JB];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < JB; ++j) acc[i][j] = make_float4(0, 0, 0, 0);

    const int tiles_per_frame = P / TM;
    const long long per = (total_tiles + gridDim.x - 1) / gridDim.x;
    const long long t0 = (long long)blockIdx.x * per, t1 = min(t0 + per, total_tiles);
    int cur_n = -1;
    const int qa = tid % 32, ra = tid / 32;   // A tile: 32 quads per row, 8 rows per pass
    const int qb = tid % QB, rb = tid / QB;   // B tile: CB/4 quads per row, 256/(CB/4) rows per pass
    for (long long t = t0; t < t1; ++t) {
        const int n = (int)(t / tiles_per_frame);
        if (n != cur_n) { la.init(n, qa); lb.init(n, qb); cur_n = n; }
        const size_t row0 = (size_t)t * TM;   // == n*P + (t % tiles_per_frame)*TM
        __syncthreads();
#pragma unroll 4
        for (int r = ra; r < TM; r += 8) st4(As + r * CA + qa * 4, la.load(row0 + r, qa));
#pragma unroll 4
        for (int r = rb; r < TM; r += RB) st4(Bs + r * CB + qb * 4, lb.load(row0 + r, qb));
        __syncthreads();
#pragma unroll 4
        for (int p = 0; p < TM; ++p) {
            const float4 a0 = ld4(As + p * CA + ta * 8), a1 = ld4(As + p * CA + ta * 8 + 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int j = 0; j < JB; ++j) {
                const float4 b = ld4(Bs + p * CB + tb * 4 + 64 * j);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    acc[i][j].x = fmaf(av[i], b.x, acc[i][j].x);
                    acc[i][j].y = fmaf(av[i], b.y, acc[i][j].y);
                    acc[i][j].z = fmaf(av[i], b.z, acc[i][j].z);
                    acc[i][j].w = fmaf(av[i], b.w, acc[i][j].w);
                }
            }
        }
    }
    float* dst = partial + (size_t)blockIdx.x * CA * CB;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < JB; ++j) {
            const int ia = ta * 8 + i, ib = tb * 4 + 64 * j;
            const float v[4] = {acc[i][j].x, acc[i][j].y, acc[i][j].z, acc[i][j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) dst[(size_t)ia * sa + (size_t)(ib + e) * sb] = v[e];
        }
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ grad, int count, int nparts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float s = 0.f;
    for (int b = 0; b < nparts; ++b) s += partial[(size_t)b * count + i];
    grad[i] += s;
}

template <class LA, class LB>
static int launch_wgrad(LA la, LB lb, float* partial, int max_parts, float* grad, int N, int P, int sa, int sb,
                        cudaStream_t st) {
    if (P % 64 != 0) return UB_ERR_ARG;
    constexpr size_t smem = (size_t)64 * (128 + 256) * sizeof(float);
    auto kern = wgrad_kernel<LA, LB>;
    UB_SET_SMEM(kern, smem);
    const long long total = (long long)N * (P / 64);
    int blocks = (int)(total < max_parts ? total : max_parts);
    kern<<<blocks, 256, smem, st>>>(la, lb, partial, P, total, sa, sb);
    UB_CHECK_LAUNCH();
    reduce_partials_kernel<<<(128 * 256 + 255) / 256, 256, 0, st>>>(partial, grad, 128 * 256, blocks);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

int launch_reduce_partials(const float* partial, float* grad, int count, int nparts, cudaStream_t st) {
    reduce_partials_kernel<<<(count + 255) / 256, 256, 0, st>>>(partial, grad, count, nparts);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

// [rows][cols] -> [cols][rows]; used once per step for the forward weight layouts
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int r = i / cols, c = i % cols;
    out[(size_t)c * rows + r] = in[i];
}

// ------------------------------------------------------------------------------------------
// launchers (kernels.h)
// ------------------------------------------------------------------------------------------
int launch_transpose(const float* in, float* out, int rows, int cols, cudaStream_t st) {
    transpose_kernel<<<(rows * cols + 255) / 256, 256, 0, st>>>(in, out, rows, cols);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

int simt_gemm1_fwd(const float* x, const Coef* coef0, const float* w1t, float* h1, double* stats1, int N, int P, cudaStream_t st) {
    LoadNormed al{x, coef0, UB_WIDTH};
    EpiStoreStats ep{h1, stats1, UB_HID};
    return launch_gemm<UB_WIDTH, UB_HID>(al, w1t, ep, N, P, st);
}
int simt_gemm2_fwd(const float* h2, const Coef* coef2, const float* gate, const float* w2t, float* y, double* stats3, int N, int P, cudaStream_t st) {
    LoadGeluGate al{h2, coef2, gate, UB_HID};
    EpiStoreStats ep{y, stats3, UB_WIDTH};
    return launch_gemm<UB_HID, UB_WIDTH>(al, w2t, ep, N, P, st);
}
int simt_gemm2_bwd(const float* dout, const float* y, const BCoef* bc3, const float* w2, float* du, const float* h2,
                   const Coef* coef2, const MeanRstd* mr2, double* sums3, int N, int P, cudaStream_t st) {
    LoadNormBwd al{dout, y, bc3, UB_WIDTH};
    EpiGemm2Bwd ep{du, h2, coef2, mr2, sums3};
    return launch_gemm<UB_WIDTH, UB_HID>(al, w2, ep, N, P, st);
}
int simt_gemm1_bwd(const float* dz1, const float* h1, const BCoef* bc1, const float* w1, float* dn0, const float* x,
                   const MeanRstd* mr0, double* bstats0, int N, int P, cudaStream_t st) {
    LoadNormBwd al{dz1, h1, bc1, UB_HID};
    EpiGemm1Bwd ep{dn0, x, mr0, bstats0};
    return launch_gemm<UB_HID, UB_WIDTH>(al, w1, ep, N, P, st);
}
// dW2[o][k] += sum_p dy[p][o] * u[p][k],  dy = norm3-bwd(dOut, y), u = gelu(norm2(h2)) * s
int simt_wgrad2(const float* dout, const float* y, const BCoef* bc3, const float* h2, const Coef* coef2, const float* gate,
                float* partial, int max_parts, float* dw2, int N, int P, cudaStream_t st) {
    LoadNormBwd la{dout, y, bc3, UB_WIDTH};
    LoadGeluGate lb{h2, coef2, gate, UB_HID};
    return launch_wgrad(la, lb, partial, max_parts, dw2, N, P, UB_HID, 1, st);
}
// dW1[o][k] += sum_p dh1[p][o] * n0[p][k],  dh1 = norm1-bwd(dz1, h1), n0 = norm0(x)
int simt_wgrad1(const float* x, const Coef* coef0, const float* dz1, const float* h1, const BCoef* bc1, float* partial,
                int max_parts, float* dw1, int N, int P, cudaStream_t st) {
    LoadNormed la{x, coef0, UB_WIDTH};
    LoadNormBwd lb{dz1, h1, bc1, UB_HID};
    return launch_wgrad(la, lb, partial, max_parts, dw1, N, P, 1, UB_WIDTH, st);
}

// ------------------------------------------------------------------------------------------
// Generic fp32 linear layers of the `use_v` L-TAE value path (ltae.py:10-141, uncrtaints.py:324-338,414-417): row-major
// [rows][K] x [K][NOUT] products on the low-resolution grid (B*1024 rows) and the include_v 1x1 convolution at full resolution.
// They reuse the streaming GEMM / weight-gradient kernels above with plain loaders; all operands are fp32 (exact).
// ------------------------------------------------------------------------------------------
struct LoadPlain {
    const float* x; int C;
    __device__ void init(int, int) {}
    __device__ float4 load(size_t row, int kq) const { return ld4(x + row * C + kq * 4); }
};
struct EpiLinear {             // out = acc + bias[col] + rowbias[n][col]; column (sum, sumsq) into stats
    static constexpr int NS = 2;
    float* out; const float* bias; const float* rowbias; double* stats; int NOUT;
    __device__ void apply(int n, size_t row, int col, float4 v, float4* s) const {
        if (bias) { const float4 b = ld4(bias + col); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
        if (rowbias) { const float4 b = ld4(rowbias + (size_t)n * NOUT + col); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
        st4(out + row * NOUT + col, v);
        s[0].x += v.x; s[0].y += v.y; s[0].z += v.z; s[0].w += v.w;
        s[1].x += v.x * v.x; s[1].y += v.y * v.y; s[1].z += v.z * v.z; s[1].w += v.w * v.w;
    }
    __device__ double* dst(int n) const { return stats + (size_t)n * NOUT * NS; }
};
struct EpiUpAdd {              // out = acc + bilinear_up(low)[n][p][col] (align_corners=False, uncrtaints.py:416); low: [N][32*32][128]
    static constexpr int NS = 2;
    float* out; const float* low; double* stats; int H, W;
    __device__ void apply(int n, size_t row, int col, float4 v, float4* s) const {
        const int p = (int)(row - (size_t)n * H * W), y = p / W, x = p % W;
        int y0, y1, x0, x1; float ly, lx;
        bilinear_tap(y, (float)UB_LOW / (float)H, UB_LOW, y0, y1, ly);
        bilinear_tap(x, (float)UB_LOW / (float)W, UB_LOW, x0, x1, lx);
        const float* base = low + (size_t)n * UB_LOW * UB_LOW * UB_WIDTH + col;
        const float4 a = ld4(base + (size_t)(y0 * UB_LOW + x0) * UB_WIDTH), b = ld4(base + (size_t)(y0 * UB_LOW + x1) * UB_WIDTH);
        const float4 c = ld4(base + (size_t)(y1 * UB_LOW + x0) * UB_WIDTH), d = ld4(base + (size_t)(y1 * UB_LOW + x1) * UB_WIDTH);
        const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
        v.x += w00 * a.x + w01 * b.x + w10 * c.x + w11 * d.x;
        v.y += w00 * a.y + w01 * b.y + w10 * c.y + w11 * d.y;
        v.z += w00 * a.z + w01 * b.z + w10 * c.z + w11 * d.z;
        v.w += w00 * a.w + w01 * b.w + w10 * c.w + w11 * d.w;
        st4(out + row * UB_WIDTH + col, v);
        s[0].x += v.x; s[0].y += v.y; s[0].z += v.z; s[0].w += v.w;
        s[1].x += v.x * v.x; s[1].y += v.y * v.y; s[1].z += v.z * v.z; s[1].w += v.w * v.w;
    }
    __device__ double* dst(int n) const { return stats + (size_t)n * UB_WIDTH * NS; }
};

int simt_linear(const float* a, int K, const float* wt, int NOUT, const float* bias, const float* rowbias, float* out, double* stats,
                int N, int P, cudaStream_t st) {
    if (!a || !wt || !out || !stats) return UB_ERR_ARG;
    LoadPlain al{a, K};
    EpiLinear ep{out, bias, rowbias, stats, NOUT};
    if (K == 128 && NOUT == 256) return launch_gemm<128, 256>(al, wt, ep, N, P, st);
    if (K == 256 && NOUT == 128) return launch_gemm<256, 128>(al, wt, ep, N, P, st);
    if (K == 128 && NOUT == 128) return launch_gemm<128, 128>(al, wt, ep, N, P, st);
    return UB_ERR_ARG;
}
int simt_linear_upadd(const float* a, const float* wt, const float* low, float* out, double* stats, int N, int H, int W, cudaStream_t st) {
    LoadPlain al{a, UB_WIDTH};
    EpiUpAdd ep{out, low, stats, H, W};
    return launch_gemm<128, 128>(al, wt, ep, N, H * W, st);
}

// grad[r * ld + c] += sum_b partial[b][r][c]   (compact [rows][cols] partials into a wider matrix)
__global__ void reduce_partials_2d_kernel(const float* __restrict__ partial, float* __restrict__ grad, int rows, int cols, int ld, int nparts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    float s = 0.f;
    for (int b = 0; b < nparts; ++b) s += partial[(size_t)b * rows * cols + i];
    grad[(size_t)(i / cols) * ld + i % cols] += s;
}
// grad[ia*sa + ib*sb] (+ld handling for cb == 128) += sum_rows a[row][ia] * b[row][ib];  a: [N*P][128], b: [N*P][cb], cb = 128 | 256
int simt_wgrad_plain(const float* a, const float* b, int cb, float* partial, int max_parts, float* grad, int sa, int sb, int ld, int N, int P,
                     cudaStream_t st) {
    if (!a || !b || !partial || !grad || P % 64 != 0) return UB_ERR_ARG;
    LoadPlain la{a, 128}, lb{b, cb};
    if (cb == 256) return launch_wgrad(la, lb, partial, max_parts, grad, N, P, sa, sb, st);
    if (cb != 128) return UB_ERR_ARG;
    constexpr size_t smem = (size_t)64 * (128 + 128) * sizeof(float);
    auto kern = wgrad_kernel<LoadPlain, LoadPlain, 128>;
    UB_SET_SMEM(kern, smem);
    const long long total = (long long)N * (P / 64);
    const int blocks = (int)(total < max_parts ? total : max_parts);
    kern<<<blocks, 256, smem, st>>>(la, lb, partial, P, total, 128, 1);
    UB_CHECK_LAUNCH();
    reduce_partials_2d_kernel<<<(128 * 128 + 255) / 256, 256, 0, st>>>(partial, grad, 128, 128, ld, blocks);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

}  // namespace ub
