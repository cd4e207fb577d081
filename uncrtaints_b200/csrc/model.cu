// Host-side orchestration of the UnCRtainTS hot path and the C ABI (include/uncrtaints_b200.h).
//
// Forward  = UNCRTAINTS.forward   (model/src/backbones/uncrtaints.py:391-446)
// Backward = what autograd derives from it in the reference (triggered at base_model.py:77)
//
// Kernel sequence per MBConv block (uncrtaints.py:100-146), "one HBM materialisation per normalisation barrier":
//   fwd  K1 gemm1 (PreNorm apply -> 1x1 expand -> h1 + stats)      K2 dwconv (Norm1+GELU -> 3x3 reflect -> h2 + stats)
//        K3 se_pool (Norm2+GELU -> pooled sums)  se_fwd            K4 gemm2 (Norm2+GELU+gate -> 1x1 project -> y + stats)
//        K5 residual (x + Norm3(y) -> out + stats for the next PreNorm)
//   bwd  B5a norm_bwd_stats  B5b gemm2_bwd (+ wgrad2)  se_bwd  B3 dwconv_bwd  B2b gemm1_bwd (+ wgrad1)  B1 residual_bwd
// with tiny finalize kernels turning column sums into per-(frame, channel) coefficients in between.
#include <string.h>
#include <stdio.h>
#include "common.cuh"
#include "kernels.h"
#include "../../include/uncrtaints_b200.h"

namespace ub {

unsigned long long g_launch_count = 0;

// ---- optional per-kernel timing (bench.py): CUDA events recorded on the launch stream around selected kernels ----
enum KernelId {
    KID_GEMM1_FWD = 0, KID_DWCONV_FWD, KID_SE_POOL, KID_GEMM2_FWD, KID_RESIDUAL_FWD, KID_NORM_BWD_STATS, KID_GEMM2_BWD,
    KID_WGRAD2, KID_DWCONV_BWD, KID_GEMM1_BWD, KID_WGRAD1, KID_RESIDUAL_BWD, KID_INCONV, KID_MAXPOOL, KID_LTAE, KID_AGGREGATE,
    KID_HEAD, KID_INCONV_BWD, KID_TEMPORAL_BWD, KID_HEAD_BWD, KID_COUNT
};
struct ProfRec { int kid; cudaEvent_t a, b; };
static unsigned long long g_prof_mask = 0;
static ProfRec g_prof[8192];
static int g_prof_n = 0, g_prof_created = 0;
static inline void prof_begin(int kid, cudaStream_t st) {
    if (!((g_prof_mask >> kid) & 1ull) || g_prof_n >= 8192) return;
    ProfRec& r = g_prof[g_prof_n];
    if (g_prof_n >= g_prof_created) { cudaEventCreate(&r.a); cudaEventCreate(&r.b); g_prof_created = g_prof_n + 1; }
    r.kid = kid;
    cudaEventRecord(r.a, st);
}
static inline void prof_end(int kid, cudaStream_t st) {
    if (!((g_prof_mask >> kid) & 1ull) || g_prof_n >= 8192) return;
    cudaEventRecord(g_prof[g_prof_n].b, st);
    ++g_prof_n;
}
#define UB_PROF(kid, st, expr)          \
    do {                                \
        prof_begin(kid, st);            \
        int rc__ = (expr);              \
        prof_end(kid, st);              \
        if (rc__ != UB_OK) return rc__; \
    } while (0)

#define UB_TRY(expr)                    \
    do {                                \
        int rc__ = (expr);              \
        if (rc__ != UB_OK) return rc__; \
    } while (0)

struct Bump {
    size_t off = 0;
    size_t take(size_t bytes) {
        const size_t o = off;
        off += (bytes + 255) & ~(size_t)255;
        return o;
    }
};

// Workspace slices of one MBConv block (offsets in bytes).
struct BlockWs {
    // saved activations
    size_t h1, h2, y, out;
    // forward statistics (fp64, inside the forward zero arena)
    size_t stats0, stats1, stats2, stats3, pool, gp;
    // coefficients
    size_t coef0, coef1, coef2, coef3, mr0, mr1, mr2, mr3, se_save, gate, w1t, w2t;
    size_t w1img, w2img, w2timg, w1timg;   // tcgen05 weight images (bf16 hi/lo, swizzled), 128 KB each
    // backward statistics (inside the backward zero arena) and coefficients
    size_t bstats0, bstats1, bstats2, bstats3, sums3, bc0, bc1, bc2, bc3, dmp;
};

// Workspace slices of one residual block (block_type = 'residual'): three ConvLayers conv3x3 -> norm -> ReLU (uncrtaints.py:24-69)
struct ResWs {
    size_t c[3], out;                       // saved activations: the three convolution outputs (pre-norm) and the block output
    size_t stats[3], coef[3], mr[3];        // forward statistics (zero arena), coefficients
    size_t bstats[3], bc[3];                // backward statistics (zero arena), coefficients
    size_t wimg[3], wimgT[3];               // tcgen05 weight images (forward: [co][tap, ci]; input gradient: [ci][tap, co], taps mirrored), 576 KB each
};

struct Layout {
    int Ne, B, P, nblk, Nmax, max_parts;
    int residual;               // block_type: 0 = MBConv, 1 = ResidualConvBlock
    ResWs rblk[1 + 16];
    size_t res_scratch;         // [Nmax][128] doubles: statistics sink of the input-gradient GEMM epilogue
    size_t res_split;           // pre-split (hi / lo) image of the current convolution's input activations (common.cuh: split_image_bytes)
    size_t res_dcsplit;         // pre-split image of the current convolution's output gradient (backward)
    size_t res_wfold;           // [9][128][128] fp32: tap-major weights of the current layer for the reflection-adjoint kernel
    size_t hes;                 // bytes per element of the 256-channel hidden tensors h1, h2, du, dz1 (4, or 2 with gemm_backend bit 5)
    size_t fwd_zero_begin, fwd_zero_end, bwd_zero_begin, bwd_zero_end;
    size_t notpad, stats_c0, coef_in, mr_in, x0, pooled, pool_idx, attn, agg;
    size_t bstats_in, bc_in, mom_in, gram_in;
    size_t gA, gB, dn0, du, dz1, partial, dwup, dattn, dpooled;
    BlockWs blk[1 + 16];
    // use_v (ltae_v.cu): low-resolution value path + include_v
    struct {
        size_t xn, z, o, m, v, vv, mix, stats_m, coef_m, mr_m, scratch_stats, winT, wmT, wvT, waT, wv, wa, dump;     // forward
        size_t bstats_m, bc_m, dvv, dv, dyb, dm, d_o, dz, dxn;                                                   // backward
    } V;
    size_t total;
};

constexpr int MAX_PARTS = 148;     // persistent weight-gradient CTAs: one [128][256] partial each

static void block_fwd_stats(Bump& b, BlockWs& w, int N) {
    w.stats0 = b.take((size_t)N * UB_WIDTH * 2 * sizeof(double));
    w.stats1 = b.take((size_t)N * UB_HID * 2 * sizeof(double));
    w.stats2 = b.take((size_t)N * UB_HID * 2 * sizeof(double));
    w.stats3 = b.take((size_t)N * UB_WIDTH * 2 * sizeof(double));
    w.pool = b.take((size_t)N * UB_HID * 2 * sizeof(double));
    w.gp = b.take((size_t)N * UB_HID * 2 * sizeof(double));
}
static void block_bwd_stats(Bump& b, BlockWs& w, int N) {
    w.bstats0 = b.take((size_t)N * UB_WIDTH * 2 * sizeof(double));
    w.bstats1 = b.take((size_t)N * UB_HID * 2 * sizeof(double));
    w.bstats2 = b.take((size_t)N * UB_HID * 2 * sizeof(double));
    w.bstats3 = b.take((size_t)N * UB_WIDTH * 2 * sizeof(double));
    w.sums3 = b.take((size_t)N * UB_HID * 3 * sizeof(double));
}
// acts == false: the caller assigns h1 / h2 / y / out itself (forward-only layout: blocks share them, see make_layout)
static void block_rest(Bump& b, BlockWs& w, int N, size_t P, bool acts = true, size_t hes = sizeof(float)) {
    if (acts) {
        w.h1 = b.take((size_t)N * P * UB_HID * hes);
        w.h2 = b.take((size_t)N * P * UB_HID * hes);
        w.y = b.take((size_t)N * P * UB_WIDTH * sizeof(float));
        w.out = b.take((size_t)N * P * UB_WIDTH * sizeof(float));
    }
    w.coef0 = b.take((size_t)N * UB_WIDTH * sizeof(Coef));
    w.coef1 = b.take((size_t)N * UB_HID * sizeof(Coef));
    w.coef2 = b.take((size_t)N * UB_HID * sizeof(Coef));
    w.coef3 = b.take((size_t)N * UB_WIDTH * sizeof(Coef));
    w.mr0 = b.take((size_t)N * UB_WIDTH * sizeof(MeanRstd));
    w.mr1 = b.take((size_t)N * UB_HID * sizeof(MeanRstd));
    w.mr2 = b.take((size_t)N * UB_HID * sizeof(MeanRstd));
    w.mr3 = b.take((size_t)N * UB_WIDTH * sizeof(MeanRstd));
    w.se_save = b.take((size_t)N * (UB_HID + UB_SE + UB_HID) * sizeof(float));
    w.gate = b.take((size_t)N * UB_HID * sizeof(float));
    w.w1t = b.take((size_t)UB_WIDTH * UB_HID * sizeof(float));
    w.w2t = b.take((size_t)UB_WIDTH * UB_HID * sizeof(float));
    w.w1img = b.take((size_t)UB_WIDTH * UB_HID * 4);
    w.w2img = b.take((size_t)UB_WIDTH * UB_HID * 4);
    w.w2timg = b.take((size_t)UB_WIDTH * UB_HID * 4);
    w.w1timg = b.take((size_t)UB_WIDTH * UB_HID * 4);
    w.bc0 = b.take((size_t)N * UB_WIDTH * sizeof(BCoef));
    w.bc1 = b.take((size_t)N * UB_HID * sizeof(BCoef));
    w.bc2 = b.take((size_t)N * UB_HID * sizeof(BCoef));
    w.bc3 = b.take((size_t)N * UB_WIDTH * sizeof(BCoef));
    w.dmp = b.take((size_t)N * UB_HID * sizeof(float));
}

static int make_layout(const ub200_desc* d, Layout& L) {
    if (!d || d->B < 1 || d->T < 1 || d->T > UB_TLONG || d->C_in < 1 || d->C_in > 16) return UB_ERR_ARG;
    if (d->H < UB_LOW || d->W < UB_LOW || d->H % UB_LOW || d->W % UB_LOW) return UB_ERR_ARG;
    if (d->n_dec_blocks < 1 || d->n_dec_blocks > 16) return UB_ERR_ARG;
    if (d->out_dim < UB_S2 || d->out_dim > 26) return UB_ERR_ARG;
    if (d->is_mono && d->use_v) return UB_ERR_ARG;
    if (d->is_mono && d->T != 1) return UB_ERR_ARG;             // out.squeeze(dim=1) (uncrtaints.py:418) needs a single frame
    L.B = d->B; L.Ne = d->B * d->T; L.P = d->H * d->W; L.nblk = 1 + d->n_dec_blocks;
    L.Nmax = L.Ne;
    // bf16 hidden storage needs the tcgen05 GEMMs (the CUDA-core comparators read fp32) and excludes the fused project-conv backward
    if ((d->gemm_backend & 3) != 0 && (d->gemm_backend & 3) != 3) return UB_ERR_ARG;
    if ((d->gemm_backend & 32) && (d->gemm_backend & 3) != 3) return UB_ERR_ARG;
    L.hes = (d->gemm_backend & 32) ? 2 : 4;
    L.residual = d->block_type == 1;
    if (d->block_type != 0 && d->block_type != 1) return UB_ERR_ARG;
    // the residual blocks exist on the tcgen05 path only, with fp32 storage
    if (L.residual && ((d->gemm_backend & 3) != 3 || (d->gemm_backend & 32))) return UB_ERR_ARG;
    const size_t P = (size_t)L.P;
    Bump b;
    L.fwd_zero_begin = b.off;
    L.notpad = b.take((size_t)L.Ne * sizeof(int));
    L.stats_c0 = b.take((size_t)L.Ne * UB_WIDTH * 2 * sizeof(double));
    L.mom_in = b.take(inconv_moments_bytes(L.Ne));
    for (int i = 0; i < L.nblk; ++i) {
        const int N = i == 0 ? L.Ne : L.B;
        if (L.residual) {
            L.blk[i].stats0 = b.take((size_t)N * UB_WIDTH * 2 * sizeof(double));       // sink of the producers' "next PreNorm" sums
            for (int l = 0; l < 3; ++l) L.rblk[i].stats[l] = b.take((size_t)N * UB_WIDTH * 2 * sizeof(double));
        } else {
            block_fwd_stats(b, L.blk[i], N);
        }
    }
    if (d->use_v) L.V.stats_m = b.take((size_t)L.B * UB_WIDTH * 2 * sizeof(double));
    L.fwd_zero_end = b.off;
    L.bwd_zero_begin = b.off;
    L.bstats_in = b.take((size_t)L.Ne * UB_WIDTH * 2 * sizeof(double));
    L.gram_in = b.take(inconv_gram_bytes(L.Ne));
    for (int i = 0; i < L.nblk; ++i) {
        const int N = i == 0 ? L.Ne : L.B;
        if (L.residual) {
            for (int l = 0; l < 3; ++l) L.rblk[i].bstats[l] = b.take((size_t)N * UB_WIDTH * 2 * sizeof(double));
        } else {
            block_bwd_stats(b, L.blk[i], N);
        }
    }
    if (d->use_v) L.V.bstats_m = b.take((size_t)L.B * UB_WIDTH * 2 * sizeof(double));
    L.bwd_zero_end = b.off;
    L.coef_in = b.take((size_t)L.Ne * UB_WIDTH * sizeof(Coef));
    L.mr_in = b.take((size_t)L.Ne * UB_WIDTH * sizeof(MeanRstd));
    L.bc_in = b.take((size_t)L.Ne * UB_WIDTH * sizeof(BCoef));
    L.x0 = b.take((size_t)L.Ne * P * UB_WIDTH * sizeof(float));
    L.pooled = b.take((size_t)L.Ne * UB_LOW * UB_LOW * UB_WIDTH * sizeof(float));
    L.pool_idx = b.take((size_t)L.Ne * UB_LOW * UB_LOW * UB_WIDTH * sizeof(int));
    L.attn = b.take((size_t)UB_HEADS * L.Ne * UB_LOW * UB_LOW * sizeof(float));
    L.agg = b.take((size_t)L.B * P * UB_WIDTH * sizeof(float));
    if (d->use_v) {
        const size_t Q = (size_t)L.B * UB_LOW * UB_LOW, NQ = (size_t)L.Ne * UB_LOW * UB_LOW;
        L.V.xn = b.take(NQ * UB_WIDTH * 4); L.V.z = b.take(NQ * UB_HID * 4); L.V.o = b.take(Q * UB_HID * 4);
        L.V.m = b.take(Q * UB_WIDTH * 4); L.V.v = b.take(Q * UB_WIDTH * 4); L.V.vv = b.take(Q * UB_WIDTH * 4);
        L.V.mix = b.take((size_t)L.B * P * UB_WIDTH * 4);
        L.V.coef_m = b.take((size_t)L.B * UB_WIDTH * sizeof(Coef)); L.V.mr_m = b.take((size_t)L.B * UB_WIDTH * sizeof(MeanRstd));
        L.V.scratch_stats = b.take((size_t)L.Ne * UB_HID * 2 * sizeof(double));
        L.V.winT = b.take((size_t)UB_WIDTH * UB_HID * 4); L.V.wmT = b.take((size_t)UB_WIDTH * UB_HID * 4);
        L.V.wvT = b.take((size_t)UB_WIDTH * UB_WIDTH * 4); L.V.waT = b.take((size_t)UB_WIDTH * UB_WIDTH * 4);
        L.V.wv = b.take((size_t)UB_WIDTH * UB_WIDTH * 4); L.V.wa = b.take((size_t)UB_WIDTH * UB_WIDTH * 4);
        L.V.dump = b.take(2 * UB_HID * 4);
        if (d->need_grad) {
            L.V.bc_m = b.take((size_t)L.B * UB_WIDTH * sizeof(BCoef));
            L.V.dvv = b.take(Q * UB_WIDTH * 4); L.V.dv = b.take(Q * UB_WIDTH * 4); L.V.dyb = b.take(Q * UB_WIDTH * 4);
            L.V.dm = b.take(Q * UB_WIDTH * 4); L.V.d_o = b.take(Q * UB_HID * 4); L.V.dz = b.take(NQ * UB_HID * 4);
            L.V.dxn = b.take(NQ * UB_WIDTH * 4);
        }
    }
    if (L.residual) {
        for (int i = 0; i < L.nblk; ++i) {
            const size_t N = i == 0 ? L.Ne : L.B;
            ResWs& r = L.rblk[i];
            for (int l = 0; l < 3; ++l) {
                r.c[l] = b.take(N * P * UB_WIDTH * sizeof(float));
                r.coef[l] = b.take(N * UB_WIDTH * sizeof(Coef));
                r.mr[l] = b.take(N * UB_WIDTH * sizeof(MeanRstd));
                r.bc[l] = b.take(N * UB_WIDTH * sizeof(BCoef));
                r.wimg[l] = b.take((size_t)9 * UB_WIDTH * UB_WIDTH * 4);
                r.wimgT[l] = b.take((size_t)9 * UB_WIDTH * UB_WIDTH * 4);
            }
            r.out = b.take(N * P * UB_WIDTH * sizeof(float));
            L.blk[i].out = r.out;
        }
        L.res_scratch = b.take((size_t)L.Nmax * UB_WIDTH * sizeof(double));
        L.res_split = b.take(split_image_bytes((size_t)L.Nmax, d->H, d->W));
        L.res_dcsplit = d->need_grad ? b.take(split_image_bytes((size_t)L.Nmax, d->H, d->W)) : 0;
        L.res_wfold = d->need_grad ? b.take((size_t)9 * UB_WIDTH * UB_WIDTH * sizeof(float)) : 0;
    } else if (d->need_grad) {
        for (int i = 0; i < L.nblk; ++i) block_rest(b, L.blk[i], i == 0 ? L.Ne : L.B, P, true, L.hes);
    } else {
        // forward only (validation / inference): nothing is saved for a backward, so all blocks share one set of hidden
        // buffers (kernels run in stream order) and the decoder outputs ping-pong between two buffers: 1 x (2 Hh + A) + 3 A
        // instead of 6 x (2 Hh + 2 A) -- at B=32, T=5 about 40 GB instead of 107 GB.
        const size_t h1 = b.take((size_t)L.Ne * P * UB_HID * L.hes), h2 = b.take((size_t)L.Ne * P * UB_HID * L.hes);
        const size_t y = b.take((size_t)L.Ne * P * UB_WIDTH * sizeof(float));
        const size_t out0 = b.take((size_t)L.Ne * P * UB_WIDTH * sizeof(float));
        const size_t pp[2] = {b.take((size_t)L.B * P * UB_WIDTH * sizeof(float)), b.take((size_t)L.B * P * UB_WIDTH * sizeof(float))};
        for (int i = 0; i < L.nblk; ++i) {
            L.blk[i].h1 = h1; L.blk[i].h2 = h2; L.blk[i].y = y;
            L.blk[i].out = i == 0 ? out0 : pp[i & 1];
            block_rest(b, L.blk[i], i == 0 ? L.Ne : L.B, P, false);
        }
    }
    if (d->need_grad) {
        L.gA = b.take((size_t)L.Nmax * P * UB_WIDTH * sizeof(float));
        L.gB = b.take((size_t)L.Nmax * P * UB_WIDTH * sizeof(float));
        L.dn0 = b.take((size_t)L.Nmax * P * UB_WIDTH * sizeof(float));
        L.du = b.take((size_t)L.Nmax * P * UB_HID * L.hes);
        L.dz1 = b.take((size_t)L.Nmax * P * UB_HID * L.hes);
        L.max_parts = MAX_PARTS;
        L.partial = b.take((size_t)L.max_parts * UB_WIDTH * UB_HID * sizeof(float));
        L.dwup = b.take((size_t)UB_HEADS * L.Ne * P * sizeof(float));
        L.dattn = b.take((size_t)UB_HEADS * L.Ne * UB_LOW * UB_LOW * sizeof(float));
        L.dpooled = b.take((size_t)L.Ne * UB_LOW * UB_LOW * UB_WIDTH * sizeof(float));
    } else {
        L.gA = L.gB = L.dn0 = L.du = L.dz1 = L.partial = L.dwup = L.dattn = L.dpooled = 0;
        L.max_parts = 0;
    }
    L.total = b.off;
    return UB_OK;
}

template <class T>
static inline T* at(void* ws, size_t off) { return reinterpret_cast<T*>(static_cast<char*>(ws) + off); }
static inline const float* pf(const void* const* tab, int i) { return static_cast<const float*>(tab[i]); }
static inline float* pfm(const void* const* tab, int i) { return const_cast<float*>(static_cast<const float*>(tab[i])); }
static inline float* gf(void* const* tab, int i) { return static_cast<float*>(tab[i]); }

struct BlockCtx {
    const void* const* p;   // UB200_BLOCK_STRIDE parameter pointers
    void* const* g;         // gradient pointers (backward) or null
    const BlockWs* w;
    void* ws;
    int N, H, W, groups, training, backend;
    int max_parts = MAX_PARTS;      // [128][256] slots available in the weight-gradient partial buffer
    int relu_mask_dx = 0;           // backward: multiply dX by [x > 0] (the block input is in_conv's ReLU output)
    float eps, momentum;
    cudaStream_t st;
};

static int finalize(const BlockCtx& c, size_t stats, int wi, size_t coef, size_t mr, int C) {
    return launch_norm_finalize(at<double>(c.ws, stats), pf(c.p, wi), pf(c.p, wi + 1), pfm(c.p, wi + 2), pfm(c.p, wi + 3),
                                at<Coef>(c.ws, coef), at<MeanRstd>(c.ws, mr), c.N, C, c.groups, (double)c.H * c.W, c.eps,
                                c.momentum, c.training, c.st);
}
static int finalize_bwd(const BlockCtx& c, size_t bstats, int wi, size_t mr, size_t bc, int C) {
    return launch_norm_finalize_bwd(at<double>(c.ws, bstats), pf(c.p, wi), at<MeanRstd>(c.ws, mr), at<BCoef>(c.ws, bc),
                                    gf(c.g, wi), gf(c.g, wi + 1), c.N, C, c.groups, (double)c.H * c.W, c.training, c.st);
}

// x: block input (stats0 already accumulated); next_stats: where K5 adds the column sums of `out` (or null)
static int mbconv_forward(const BlockCtx& c, const float* x, double* next_stats, bool need_gp) {
    const BlockWs& w = *c.w;
    const int P = c.H * c.W;
    void* ws = c.ws;
    const bool tcb = (c.backend & 3) == 3;
    if ((c.backend & 3) != 0 && !tcb) return UB_ERR_ARG;
    // forward operands (normalised activations, weights) take the fp16 hi/lo split (2^-22), see gemm_tc.cu: SPLIT_F16X3
    const int single = (c.backend & 4) ? 1 : 2;
    const int hbf = (c.backend & 32) != 0;
    if (hbf && !tcb) return UB_ERR_ARG;
    if (tcb) {
        // M-operand images: W1 [256][128] and W2 [128][256] as stored (forward); W2^T / W1^T for the input-gradient GEMMs (bf16)
        UB_TRY(tc_prep_weights(pf(c.p, UB200_B_W1), at<char>(ws, w.w1img), UB_HID, UB_WIDTH, 0, single == 2, c.st));
        UB_TRY(tc_prep_weights(pf(c.p, UB200_B_W2), at<char>(ws, w.w2img), UB_WIDTH, UB_HID, 0, single == 2, c.st));
        if (need_gp) {
            UB_TRY(tc_prep_weights(pf(c.p, UB200_B_W2), at<char>(ws, w.w2timg), UB_HID, UB_WIDTH, 1, 0, c.st));
            UB_TRY(tc_prep_weights(pf(c.p, UB200_B_W1), at<char>(ws, w.w1timg), UB_WIDTH, UB_HID, 1, 0, c.st));
        }
    } else {
        UB_TRY(launch_transpose(pf(c.p, UB200_B_W1), at<float>(ws, w.w1t), UB_HID, UB_WIDTH, c.st));
        UB_TRY(launch_transpose(pf(c.p, UB200_B_W2), at<float>(ws, w.w2t), UB_WIDTH, UB_HID, c.st));
    }
    UB_TRY(finalize(c, w.stats0, UB200_B_N0_W, w.coef0, w.mr0, UB_WIDTH));
    if (tcb)
        UB_PROF(KID_GEMM1_FWD, c.st, tc_gemm1_fwd(x, at<Coef>(ws, w.coef0), at<char>(ws, w.w1img), at<char>(ws, w.h1), at<double>(ws, w.stats1), c.N, P, single, hbf, c.st));
    else
        UB_PROF(KID_GEMM1_FWD, c.st, simt_gemm1_fwd(x, at<Coef>(ws, w.coef0), at<float>(ws, w.w1t), at<float>(ws, w.h1), at<double>(ws, w.stats1), c.N, P, c.st));
    UB_TRY(finalize(c, w.stats1, UB200_B_N1_W, w.coef1, w.mr1, UB_HID));
    UB_PROF(KID_DWCONV_FWD, c.st, launch_dwconv_fwd(at<char>(ws, w.h1), at<Coef>(ws, w.coef1), pf(c.p, UB200_B_WDW), at<char>(ws, w.h2),
                             at<double>(ws, w.stats2), c.N, c.H, c.W, hbf, c.st));
    UB_TRY(finalize(c, w.stats2, UB200_B_N2_W, w.coef2, w.mr2, UB_HID));
    UB_PROF(KID_SE_POOL, c.st, launch_se_pool(at<char>(ws, w.h2), at<Coef>(ws, w.coef2), at<MeanRstd>(ws, w.mr2), at<double>(ws, w.pool),
                          need_gp ? at<double>(ws, w.gp) : nullptr, c.N, P, hbf, c.st));
    UB_TRY(launch_se_fwd(at<double>(ws, w.pool), pf(c.p, UB200_B_F1), pf(c.p, UB200_B_F2), at<float>(ws, w.se_save),
                         at<float>(ws, w.gate), c.N, P, c.st));
    if (tcb)
        UB_PROF(KID_GEMM2_FWD, c.st, tc_gemm2_fwd(at<char>(ws, w.h2), at<Coef>(ws, w.coef2), at<float>(ws, w.gate), at<char>(ws, w.w2img),
                              at<float>(ws, w.y), at<double>(ws, w.stats3), c.N, P, single, hbf, c.st));
    else
        UB_PROF(KID_GEMM2_FWD, c.st, simt_gemm2_fwd(at<float>(ws, w.h2), at<Coef>(ws, w.coef2), at<float>(ws, w.gate), at<float>(ws, w.w2t),
                              at<float>(ws, w.y), at<double>(ws, w.stats3), c.N, P, c.st));
    UB_TRY(finalize(c, w.stats3, UB200_B_N3_W, w.coef3, w.mr3, UB_WIDTH));
    UB_PROF(KID_RESIDUAL_FWD, c.st, launch_residual_fwd(x, at<float>(ws, w.y), at<Coef>(ws, w.coef3), at<float>(ws, w.out), next_stats, c.N, P, c.st));
    return UB_OK;
}

// Eval-mode specialisation of a BatchNorm block without a backward (validation / test_reconstruct.py:104-109): all four
// normalisations use running statistics, so every coefficient is known before the first kernel runs and two of the five passes
// disappear -- the squeeze-excite pooling rides in the depthwise kernel's epilogue and the residual add in the project GEMM's:
//   K1 gemm1 (A + Hh)   K2' dwconv + pool (2 Hh)   se_fwd   K4' gemm2 + Norm3 + residual (Hh + 2 A)      = 3 A + 4 Hh per frame
// instead of 5 A + 5 Hh.  (GroupNorm blocks normalise with instance statistics in eval mode too and keep the general path.)
static int mbconv_forward_eval_bn(const BlockCtx& c, const float* x, double* next_stats) {
    const BlockWs& w = *c.w;
    const int P = c.H * c.W;
    void* ws = c.ws;
    const int single = (c.backend & 4) ? 1 : 2;
    const int hbf = (c.backend & 32) != 0;
    UB_TRY(tc_prep_weights(pf(c.p, UB200_B_W1), at<char>(ws, w.w1img), UB_HID, UB_WIDTH, 0, single == 2, c.st));
    UB_TRY(tc_prep_weights(pf(c.p, UB200_B_W2), at<char>(ws, w.w2img), UB_WIDTH, UB_HID, 0, single == 2, c.st));
    UB_TRY(finalize(c, w.stats0, UB200_B_N0_W, w.coef0, w.mr0, UB_WIDTH));
    UB_TRY(finalize(c, w.stats1, UB200_B_N1_W, w.coef1, w.mr1, UB_HID));
    UB_TRY(finalize(c, w.stats2, UB200_B_N2_W, w.coef2, w.mr2, UB_HID));
    UB_TRY(finalize(c, w.stats3, UB200_B_N3_W, w.coef3, w.mr3, UB_WIDTH));
    UB_PROF(KID_GEMM1_FWD, c.st, tc_gemm1_fwd(x, at<Coef>(ws, w.coef0), at<char>(ws, w.w1img), at<char>(ws, w.h1), at<double>(ws, w.stats1), c.N, P, single, hbf, c.st));
    UB_PROF(KID_DWCONV_FWD, c.st, launch_dwconv_fwd_pool(at<char>(ws, w.h1), at<Coef>(ws, w.coef1), pf(c.p, UB200_B_WDW), at<char>(ws, w.h2),
                             at<Coef>(ws, w.coef2), at<double>(ws, w.pool), c.N, c.H, c.W, hbf, c.st));
    UB_TRY(launch_se_fwd(at<double>(ws, w.pool), pf(c.p, UB200_B_F1), pf(c.p, UB200_B_F2), at<float>(ws, w.se_save),
                         at<float>(ws, w.gate), c.N, P, c.st));
    // the column sums of `out` go to the next block's PreNorm accumulator (unused by an eval-mode BatchNorm, harmless) or stats3
    UB_PROF(KID_GEMM2_FWD, c.st, tc_gemm2_fwd_residual(at<char>(ws, w.h2), at<Coef>(ws, w.coef2), at<float>(ws, w.gate), at<char>(ws, w.w2img), x,
                          at<Coef>(ws, w.coef3), at<float>(ws, w.out), next_stats ? next_stats : at<double>(ws, w.stats3), c.N, P, single, hbf, c.st));
    return UB_OK;
}

// dout: gradient w.r.t. the block output; dx: gradient w.r.t. the block input (may not alias dout)
// stats_ready: the Norm3-backward statistics of (dout, y) were gathered by the residual pass of the block above (which produced dout);
// below: the block that produced x -- dx is its output gradient, and this block's residual pass gathers ITS Norm3-backward statistics.
static int mbconv_backward(const BlockCtx& c, const float* x, const float* dout, float* dx, float* dn0, void* du, void* dz1,
                           float* partial, bool stats_ready = false, const BlockWs* below = nullptr) {
    const BlockWs& w = *c.w;
    const int P = c.H * c.W;
    void* ws = c.ws;
    if (!stats_ready)
        UB_PROF(KID_NORM_BWD_STATS, c.st, launch_norm_bwd_stats(dout, at<float>(ws, w.y), at<MeanRstd>(ws, w.mr3), at<double>(ws, w.bstats3), c.N, P, c.st));
    UB_TRY(finalize_bwd(c, w.bstats3, UB200_B_N3_W, w.mr3, w.bc3, UB_WIDTH));
    const bool tcb = (c.backend & 3) == 3;                      // 3 = tcgen05 product path, 0 = fp32 CUDA-core comparator
    if ((c.backend & 3) != 0 && !tcb) return UB_ERR_ARG;
    const int single = (c.backend & 4) != 0;
    const int hbf = (c.backend & 32) != 0;
    if (hbf && !tcb) return UB_ERR_ARG;
    float* du_f = static_cast<float*>(du);                     // fp32 view for the CUDA-core comparators / the fused project kernel
    float* dz1_f = static_cast<float*>(dz1);
    if (tcb)
        UB_PROF(KID_GEMM2_BWD, c.st, tc_gemm2_bwd(dout, at<float>(ws, w.y), at<BCoef>(ws, w.bc3), at<char>(ws, w.w2timg), du, at<char>(ws, w.h2),
                              at<Coef>(ws, w.coef2), at<MeanRstd>(ws, w.mr2), at<double>(ws, w.sums3), c.N, P, single, hbf, c.st));
    else
        UB_PROF(KID_GEMM2_BWD, c.st, simt_gemm2_bwd(dout, at<float>(ws, w.y), at<BCoef>(ws, w.bc3), pf(c.p, UB200_B_W2), du_f, at<float>(ws, w.h2),
                              at<Coef>(ws, w.coef2), at<MeanRstd>(ws, w.mr2), at<double>(ws, w.sums3), c.N, P, c.st));
    if (tcb)
        UB_PROF(KID_WGRAD2, c.st, tc_wgrad2(dout, at<float>(ws, w.y), at<BCoef>(ws, w.bc3), at<char>(ws, w.h2), at<Coef>(ws, w.coef2),
                           at<float>(ws, w.gate), partial, MAX_PARTS, gf(c.g, UB200_B_W2), c.N, P, single, hbf, c.st));
    else
        UB_PROF(KID_WGRAD2, c.st, simt_wgrad2(dout, at<float>(ws, w.y), at<BCoef>(ws, w.bc3), at<float>(ws, w.h2), at<Coef>(ws, w.coef2),
                           at<float>(ws, w.gate), partial, MAX_PARTS, gf(c.g, UB200_B_W2), c.N, P, c.st));
    UB_TRY(launch_se_bwd(at<double>(ws, w.sums3), at<double>(ws, w.gp), pf(c.p, UB200_B_F1), pf(c.p, UB200_B_F2),
                         at<float>(ws, w.se_save), gf(c.g, UB200_B_F1), gf(c.g, UB200_B_F2), at<float>(ws, w.dmp),
                         at<double>(ws, w.bstats2), c.N, P, c.st));
    UB_TRY(finalize_bwd(c, w.bstats2, UB200_B_N2_W, w.mr2, w.bc2, UB_HID));
    UB_PROF(KID_DWCONV_BWD, c.st, launch_dwconv_bwd(du, at<char>(ws, w.h2), at<char>(ws, w.h1), at<float>(ws, w.gate), at<float>(ws, w.dmp),
                             at<Coef>(ws, w.coef2), at<BCoef>(ws, w.bc2), at<Coef>(ws, w.coef1), at<MeanRstd>(ws, w.mr1),
                             pf(c.p, UB200_B_WDW), dz1, at<double>(ws, w.bstats1), gf(c.g, UB200_B_WDW), c.N, c.H, c.W, hbf, c.st));
    UB_TRY(finalize_bwd(c, w.bstats1, UB200_B_N1_W, w.mr1, w.bc1, UB_HID));
    if (tcb) {     // input gradient AND weight gradient of the expand convolution in one kernel (one read of dz1, h1, x)
        UB_PROF(KID_GEMM1_BWD, c.st, tc_gemm1_bwd(dz1, at<char>(ws, w.h1), at<BCoef>(ws, w.bc1), at<char>(ws, w.w1timg), dn0, x, at<MeanRstd>(ws, w.mr0),
                              at<double>(ws, w.bstats0), at<Coef>(ws, w.coef0), partial, MAX_PARTS, gf(c.g, UB200_B_W1), c.N, P, single, hbf, c.st));
    } else {
        UB_PROF(KID_GEMM1_BWD, c.st, simt_gemm1_bwd(dz1_f, at<float>(ws, w.h1), at<BCoef>(ws, w.bc1), pf(c.p, UB200_B_W1), dn0, x, at<MeanRstd>(ws, w.mr0),
                              at<double>(ws, w.bstats0), c.N, P, c.st));
        UB_PROF(KID_WGRAD1, c.st, simt_wgrad1(x, at<Coef>(ws, w.coef0), dz1_f, at<float>(ws, w.h1), at<BCoef>(ws, w.bc1), partial, MAX_PARTS,
                           gf(c.g, UB200_B_W1), c.N, P, c.st));
    }
    UB_TRY(finalize_bwd(c, w.bstats0, UB200_B_N0_W, w.mr0, w.bc0, UB_WIDTH));
    UB_PROF(KID_RESIDUAL_BWD, c.st, launch_residual_bwd(dout, dn0, x, at<BCoef>(ws, w.bc0), dx, c.N, P, c.relu_mask_dx,
                               below ? at<float>(ws, below->y) : nullptr, below ? at<MeanRstd>(ws, below->mr3) : nullptr,
                               below ? at<double>(ws, below->bstats3) : nullptr, c.st));
    return UB_OK;
}

// column sums of a [N][P][128] tensor (standalone MBConv entry point only)
__global__ void __launch_bounds__(256) colstats_kernel(const float* __restrict__ x, double* stats, int P, int chunk) {
    constexpr int C = UB_WIDTH;
    __shared__ float red[2 * 8 * C];
    const int n = blockIdx.y, lane = threadIdx.x % 32, row = threadIdx.x / 32;
    const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, P);
    float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
    for (int p = p0 + row; p < p1; p += 8) {
        const float4 v = ld4(x + ((size_t)n * P + p) * C + lane * 4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w;
    }
    float4* r4 = reinterpret_cast<float4*>(red);
    r4[row * 32 + lane] = s;
    r4[256 + row * 32 + lane] = q;
    __syncthreads();
    const int which = threadIdx.x / C, ch = threadIdx.x % C;
    double t = 0.0;
    for (int r = 0; r < 8; ++r) t += (double)red[which * 8 * C + r * C + ch];
    atomicAdd(&stats[((size_t)n * C + ch) * 2 + which], t);
}

// ---- block_type = 'residual' (ResidualConvBlock, uncrtaints.py:24-69): x + L3(L2(L1(x))), L = conv3x3(reflect, bias) -> norm -> ReLU ----
// Per ConvLayer ONE tensor is materialised (the convolution output c, with its statistics from the GEMM epilogue); the norm + ReLU
// is applied in the next convolution's operand loader, the last one in the residual pass.  The convolutions are implicit GEMMs
// over (tap, channel) on the tcgen05 path (gemm_tc.cu: TLoadConvSplit on pre-split operand images, streamed weight slabs).
static int residual_forward(const ub200_desc* d, const Layout& L, int i, const void* const* params, const float* x, double* next_stats,
                            void* ws, cudaStream_t st) {
    const ResWs& w = L.rblk[i];
    const int N = i == 0 ? L.Ne : L.B, groups = i == 0 ? d->enc_groups : d->dec_groups, P = L.P;
    const void* const* p = params + UB200_P_BLOCK0 + i * UB200_BLOCK_STRIDE;
    const int single = (d->gemm_backend & 4) ? 1 : 2;
    const float* in = x;
    const Coef* incoef = nullptr;
    for (int l = 0; l < 3; ++l) {
        const void* const* q = p + l * UB200_R_STRIDE;
        if (!pf(q, UB200_R_W) || !pf(q, UB200_R_B) || !pf(q, UB200_R_N_W)) return UB_ERR_ARG;
        UB_TRY(tc_prep_conv_weights(pf(q, UB200_R_W), at<char>(ws, w.wimg[l]), 0, single == 2, st));
        if (d->need_grad) UB_TRY(tc_prep_conv_weights(pf(q, UB200_R_W), at<char>(ws, w.wimgT[l]), 1, 0, st));
        // the convolution reads every input element nine times (once per tap): the previous layer's norm + ReLU and the hi / lo operand
        // split are applied ONCE into a pre-split image (2 A of extra traffic), the GEMM's loader then only copies 16-byte chunks
        // (measured at B=16 against converting in the loader: conv forward 40.5 -> 35.1, input gradient 46.6 -> 41.6, weight
        // gradient 45.2 -> 37.8 ms per step, for 12.3 ms of split passes)
        UB_PROF(KID_SE_POOL, st, tc_split_act(in, incoef, l > 0, at<char>(ws, L.res_split), N, d->H, d->W, single, st));
        UB_PROF(KID_GEMM1_FWD, st, tc_conv3x3_fwd(at<char>(ws, L.res_split), at<char>(ws, w.wimg[l]), pf(q, UB200_R_B), at<float>(ws, w.c[l]),
                                                        at<double>(ws, w.stats[l]), N, d->H, d->W, single, st));
        UB_TRY(launch_norm_finalize(at<double>(ws, w.stats[l]), pf(q, UB200_R_N_W), pf(q, UB200_R_N_B), pfm(q, UB200_R_N_RM), pfm(q, UB200_R_N_RV),
                                    at<Coef>(ws, w.coef[l]), at<MeanRstd>(ws, w.mr[l]), N, UB_WIDTH, groups, (double)P, d->norm_eps,
                                    d->bn_momentum, d->training, st));
        in = at<float>(ws, w.c[l]);
        incoef = at<Coef>(ws, w.coef[l]);
    }
    UB_PROF(KID_RESIDUAL_FWD, st, launch_residual_relu_fwd(x, at<float>(ws, w.c[2]), at<Coef>(ws, w.coef[2]), at<float>(ws, w.out), next_stats, N, P, st));
    return UB_OK;
}
// dout -> dx (may not alias); t0, t1: two [N][P][128] scratch buffers for the gradients between the layers
static int residual_backward(const ub200_desc* d, const Layout& L, int i, const void* const* params, void* const* grads, const float* x,
                             const float* dout, float* dx, float* t0, float* t1, float* partial, void* ws, cudaStream_t st) {
    const ResWs& w = L.rblk[i];
    char* dc = at<char>(ws, L.res_dcsplit);                // the convolution's output gradient as a pre-split image
    const int N = i == 0 ? L.Ne : L.B, groups = i == 0 ? d->enc_groups : d->dec_groups, P = L.P;
    const void* const* p = params + UB200_P_BLOCK0 + i * UB200_BLOCK_STRIDE;
    void* const* g = grads + UB200_P_BLOCK0 + i * UB200_BLOCK_STRIDE;
    const int single = (d->gemm_backend & 4) != 0;
    const float* dy = dout;
    for (int l = 2; l >= 0; --l) {
        const void* const* q = p + l * UB200_R_STRIDE;
        void* const* gq = g + l * UB200_R_STRIDE;
        UB_PROF(KID_NORM_BWD_STATS, st, launch_relu_norm_bwd_stats(dy, at<float>(ws, w.c[l]), at<Coef>(ws, w.coef[l]), at<MeanRstd>(ws, w.mr[l]),
                                                                   at<double>(ws, w.bstats[l]), N, P, st));
        UB_TRY(launch_norm_finalize_bwd(at<double>(ws, w.bstats[l]), pf(q, UB200_R_N_W), at<MeanRstd>(ws, w.mr[l]), at<BCoef>(ws, w.bc[l]),
                                        gf(gq, UB200_R_N_W), gf(gq, UB200_R_N_B), N, UB_WIDTH, groups, (double)P, d->training, st));
        UB_PROF(KID_RESIDUAL_BWD, st, launch_relu_norm_bwd_apply(dy, at<float>(ws, w.c[l]), at<Coef>(ws, w.coef[l]), at<BCoef>(ws, w.bc[l]), dc,
                                                                 gf(gq, UB200_R_B), N, d->H, d->W, st));
        const float* in = l == 0 ? x : at<float>(ws, w.c[l - 1]);
        const Coef* incoef = l == 0 ? nullptr : at<Coef>(ws, w.coef[l - 1]);
        if (gf(gq, UB200_R_W)) {
            UB_PROF(KID_SE_POOL, st, tc_split_act(in, incoef, l > 0, at<char>(ws, L.res_split), N, d->H, d->W, single, st));      // bf16 hi/lo (gradient GEMM)
            UB_PROF(KID_WGRAD1, st, tc_conv3x3_wgrad(dc, at<char>(ws, L.res_split), partial, L.max_parts, gf(gq, UB200_R_W), N, d->H, d->W,
                                                           single, st));
        }
        float* din = l == 0 ? dx : (l == 2 ? t0 : t1);
        UB_PROF(KID_GEMM1_BWD, st, tc_conv3x3_dgrad(dc, at<char>(ws, w.wimgT[l]), l == 0 ? dout : nullptr, din, at<double>(ws, L.res_scratch), N,
                                                          d->H, d->W, single, st));
        UB_PROF(KID_WGRAD2, st, launch_conv_fold(dc, pf(q, UB200_R_W), at<float>(ws, L.res_wfold), din, N, d->H, d->W, st));
        dy = din;
    }
    return UB_OK;
}

// ---- use_v: full LTAE2d value path + include_v (ltae.py:96-141, uncrtaints.py:414-417); kernels in ltae_v.cu ----
static int value_path_forward(const ub200_desc* d, const Layout& L, const void* const* params, const unsigned char* v_keep_mask, void* ws,
                              cudaStream_t st) {
    const int B = d->B, T = d->T, Ne = L.Ne;
    const float* w_in = pf(params, UB200_P_LTAE_WIN);
    const float* w_m = pf(params, UB200_P_LTAE_MLP_W);
    const float* w_cv = pf(params, UB200_P_INCV_W);
    if (!w_in || !w_m || !w_cv || !pf(params, UB200_P_LTAE_GN_W) || !pf(params, UB200_P_LTAE_BN_W) || !pf(params, UB200_P_LTAE_ON_W)) return UB_ERR_ARG;
    double* scratch = at<double>(ws, L.V.scratch_stats);
    // [K][NOUT] forms of the forward weights; compact copies of the two halves of include_v.weight for the backward GEMMs
    UB_TRY(launch_slice2d(w_in, at<float>(ws, L.V.winT), UB_HID, UB_WIDTH, UB_WIDTH, 0, 1, st));
    UB_TRY(launch_slice2d(w_m, at<float>(ws, L.V.wmT), UB_WIDTH, UB_HID, UB_HID, 0, 1, st));
    UB_TRY(launch_slice2d(w_cv, at<float>(ws, L.V.waT), UB_WIDTH, UB_WIDTH, UB_HID, 0, 1, st));
    UB_TRY(launch_slice2d(w_cv, at<float>(ws, L.V.wvT), UB_WIDTH, UB_WIDTH, UB_HID, UB_WIDTH, 1, st));
    UB_TRY(launch_slice2d(w_cv, at<float>(ws, L.V.wa), UB_WIDTH, UB_WIDTH, UB_HID, 0, 0, st));
    UB_TRY(launch_slice2d(w_cv, at<float>(ws, L.V.wv), UB_WIDTH, UB_WIDTH, UB_HID, UB_WIDTH, 0, st));
    UB_TRY(launch_ltaev_norm(at<float>(ws, L.pooled), pf(params, UB200_P_LTAE_GN_W), pf(params, UB200_P_LTAE_GN_B), at<float>(ws, L.V.xn), B, T,
                             d->norm_eps, st));
    UB_TRY(simt_linear(at<float>(ws, L.V.xn), UB_WIDTH, at<float>(ws, L.V.winT), UB_HID, pf(params, UB200_P_LTAE_BIN), pf(params, UB200_P_LTAE_PE),
                       at<float>(ws, L.V.z), scratch, Ne, UB_LOW * UB_LOW, st));
    UB_TRY(launch_ltaev_attnv(at<float>(ws, L.attn), at<float>(ws, L.V.z), at<float>(ws, L.V.o), B, T, st));
    UB_TRY(simt_linear(at<float>(ws, L.V.o), UB_HID, at<float>(ws, L.V.wmT), UB_WIDTH, pf(params, UB200_P_LTAE_MLP_B), nullptr, at<float>(ws, L.V.m),
                       at<double>(ws, L.V.stats_m), B, UB_LOW * UB_LOW, st));
    // BatchNorm1d over the B*1024 rows (ltae.py:79): per-sample partial sums, finalised like a BatchNorm2d over B frames of 1024 pixels
    UB_TRY(launch_norm_finalize(at<double>(ws, L.V.stats_m), pf(params, UB200_P_LTAE_BN_W), pf(params, UB200_P_LTAE_BN_B),
                                pfm(params, UB200_P_LTAE_BN_RM), pfm(params, UB200_P_LTAE_BN_RV), at<Coef>(ws, L.V.coef_m), at<MeanRstd>(ws, L.V.mr_m),
                                B, UB_WIDTH, 0, (double)(UB_LOW * UB_LOW), d->norm_eps, d->bn_momentum, d->training, st));
    const float vp = d->training ? d->v_dropout_p : 0.f;
    UB_TRY(launch_ltaev_post_fwd(at<float>(ws, L.V.m), at<MeanRstd>(ws, L.V.mr_m), pf(params, UB200_P_LTAE_BN_W), pf(params, UB200_P_LTAE_BN_B),
                                 pf(params, UB200_P_LTAE_ON_W), pf(params, UB200_P_LTAE_ON_B),
                                 v_keep_mask, d->seed, d->offset, vp, at<float>(ws, L.V.v), B, d->norm_eps, st));
    UB_TRY(simt_linear(at<float>(ws, L.V.v), UB_WIDTH, at<float>(ws, L.V.wvT), UB_WIDTH, pf(params, UB200_P_INCV_B), nullptr, at<float>(ws, L.V.vv),
                       scratch, B, UB_LOW * UB_LOW, st));
    // include_v(cat(agg, up(v))) = W_a agg + up(W_v v + b): one full-resolution 128 -> 128 GEMM; its column sums feed the first decoder PreNorm
    UB_TRY(simt_linear_upadd(at<float>(ws, L.agg), at<float>(ws, L.V.waT), at<float>(ws, L.V.vv), at<float>(ws, L.V.mix),
                             at<double>(ws, L.blk[1].stats0), B, d->H, d->W, st));
    return UB_OK;
}
// dmix -> dagg (into `dagg`), the include_v gradients and dvv (low resolution)
static int value_path_backward_head(const ub200_desc* d, const Layout& L, const void* const* params, void* const* grads, const float* dmix,
                                    float* dagg, float* partial, void* ws, cudaStream_t st) {
    const int B = d->B, P = L.P;
    double* scratch = at<double>(ws, L.V.scratch_stats);
    float* g_cv = gf(grads, UB200_P_INCV_W);
    UB_TRY(launch_upsample_adjoint128(dmix, at<float>(ws, L.V.dvv), B, d->H, d->W, st));
    UB_TRY(simt_linear(dmix, UB_WIDTH, at<float>(ws, L.V.wa), UB_WIDTH, nullptr, nullptr, dagg, scratch, B, P, st));
    if (g_cv) {
        UB_TRY(simt_wgrad_plain(dmix, at<float>(ws, L.agg), UB_WIDTH, partial, L.max_parts, g_cv, 0, 0, UB_HID, B, P, st));
        UB_TRY(simt_wgrad_plain(at<float>(ws, L.V.dvv), at<float>(ws, L.V.v), UB_WIDTH, partial, L.max_parts, g_cv + UB_WIDTH, 0, 0, UB_HID, B,
                                UB_LOW * UB_LOW, st));
    }
    if (gf(grads, UB200_P_INCV_B)) UB_TRY(launch_colsum(at<float>(ws, L.V.dvv), gf(grads, UB200_P_INCV_B), (size_t)B * UB_LOW * UB_LOW, UB_WIDTH, st));
    return UB_OK;
}
// after aggregate_bwd has written dattn: the low-resolution chain back to dpooled and every temporal-encoder gradient
static int value_path_backward_tail(const ub200_desc* d, const Layout& L, const void* const* params, void* const* grads,
                                    const unsigned char* v_keep_mask, float* partial, void* ws, cudaStream_t st) {
    const int B = d->B, T = d->T, Ne = L.Ne, LQ = UB_LOW * UB_LOW;
    double* scratch = at<double>(ws, L.V.scratch_stats);
    const float vp = d->training ? d->v_dropout_p : 0.f;
    float* dump = at<float>(ws, L.V.dump);     // sink for gradient slots the caller left NULL
    auto G = [&](int slot) { float* g = gf(grads, slot); return g ? g : dump; };
    UB_TRY(simt_linear(at<float>(ws, L.V.dvv), UB_WIDTH, at<float>(ws, L.V.wv), UB_WIDTH, nullptr, nullptr, at<float>(ws, L.V.dv), scratch, B, LQ, st));
    UB_TRY(launch_ltaev_post_bwd(at<float>(ws, L.V.m), at<MeanRstd>(ws, L.V.mr_m), pf(params, UB200_P_LTAE_BN_W), pf(params, UB200_P_LTAE_BN_B),
                                 pf(params, UB200_P_LTAE_ON_W), v_keep_mask,
                                 d->seed, d->offset, vp, at<float>(ws, L.V.dv), at<float>(ws, L.V.dyb), at<double>(ws, L.V.bstats_m),
                                 G(UB200_P_LTAE_ON_W), G(UB200_P_LTAE_ON_B), B, d->norm_eps, st));
    UB_TRY(launch_norm_finalize_bwd(at<double>(ws, L.V.bstats_m), pf(params, UB200_P_LTAE_BN_W), at<MeanRstd>(ws, L.V.mr_m), at<BCoef>(ws, L.V.bc_m),
                                    gf(grads, UB200_P_LTAE_BN_W), gf(grads, UB200_P_LTAE_BN_B), B, UB_WIDTH, 0, (double)LQ, d->training, st));
    UB_TRY(launch_normbwd_apply(at<float>(ws, L.V.dyb), at<float>(ws, L.V.m), at<BCoef>(ws, L.V.bc_m), at<float>(ws, L.V.dm), B, LQ, st));
    if (gf(grads, UB200_P_LTAE_MLP_B)) UB_TRY(launch_colsum(at<float>(ws, L.V.dm), gf(grads, UB200_P_LTAE_MLP_B), (size_t)B * LQ, UB_WIDTH, st));
    if (gf(grads, UB200_P_LTAE_MLP_W))
        UB_TRY(simt_wgrad_plain(at<float>(ws, L.V.dm), at<float>(ws, L.V.o), UB_HID, partial, L.max_parts, gf(grads, UB200_P_LTAE_MLP_W), UB_HID, 1, 0, B, LQ, st));
    UB_TRY(simt_linear(at<float>(ws, L.V.dm), UB_WIDTH, pf(params, UB200_P_LTAE_MLP_W), UB_HID, nullptr, nullptr, at<float>(ws, L.V.d_o), scratch, B, LQ, st));
    UB_TRY(launch_ltaev_attnv_bwd(at<float>(ws, L.attn), at<float>(ws, L.V.z), at<float>(ws, L.V.d_o), at<float>(ws, L.V.dz), at<float>(ws, L.dattn), B, T, st));
    if (gf(grads, UB200_P_LTAE_BIN)) UB_TRY(launch_colsum(at<float>(ws, L.V.dz), gf(grads, UB200_P_LTAE_BIN), (size_t)Ne * LQ, UB_HID, st));
    if (gf(grads, UB200_P_LTAE_WIN))
        UB_TRY(simt_wgrad_plain(at<float>(ws, L.V.xn), at<float>(ws, L.V.dz), UB_HID, partial, L.max_parts, gf(grads, UB200_P_LTAE_WIN), 1, UB_WIDTH, 0, Ne, LQ, st));
    UB_TRY(simt_linear(at<float>(ws, L.V.dz), UB_HID, pf(params, UB200_P_LTAE_WIN), UB_WIDTH, nullptr, nullptr, at<float>(ws, L.V.dxn), scratch, Ne, LQ, st));
    UB_TRY(launch_ltaev_final_bwd(at<float>(ws, L.pooled), pf(params, UB200_P_LTAE_AP), at<float>(ws, L.attn), at<float>(ws, L.dattn),
                                  at<float>(ws, L.V.dxn), pf(params, UB200_P_LTAE_GN_W), at<float>(ws, L.dpooled), gf(grads, UB200_P_LTAE_AP),
                                  gf(grads, UB200_P_LTAE_E), G(UB200_P_LTAE_GN_W), G(UB200_P_LTAE_GN_B), B, T, d->norm_eps, st));
    return UB_OK;
}

static int num_sms() { return device_sm_count(); }

}  // namespace ub

using namespace ub;

extern "C" {

int ub200_version(void) { return 100; }

unsigned long long ub200_launch_count(void) { return g_launch_count; }

// h1[N*P][256] = (x[N*P][128] * scale + shift) . W1^T, stats[N][256][2] += column (sum, sumsq): the 1x1 expand GEMM alone
// (unit tests and the roofline micro-benchmark).  coef: [N][128] (scale, shift) pairs; scratch: 256 KB.
int ub200_gemm1_forward(int backend, const float* x, const float* coef, const float* w1, float* h1, double* stats, int N, int P,
                        void* scratch, void* stream) {
    if (!x || !coef || !w1 || !h1 || !stats || !scratch || P % 128) return UB_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(stats, 0, (size_t)N * UB_HID * 2 * sizeof(double), st) != cudaSuccess) return UB_ERR_CUDA;
    if ((backend & 3) == 3) {
        const int mode = (backend & 4) ? 1 : 2;
        UB_TRY(tc_prep_weights(w1, scratch, UB_HID, UB_WIDTH, 0, mode == 2, st));
        return tc_gemm1_fwd(x, reinterpret_cast<const Coef*>(coef), scratch, h1, stats, N, P, mode, 0, st);
    }
    UB_TRY(launch_transpose(w1, static_cast<float*>(scratch), UB_HID, UB_WIDTH, st));
    return simt_gemm1_fwd(x, reinterpret_cast<const Coef*>(coef), static_cast<const float*>(scratch), h1, stats, N, P, st);
}

int ub200_prof_enable(unsigned long long mask) {
    g_prof_mask = mask;
    g_prof_n = 0;
    return UB_OK;
}
int ub200_prof_num_kernels(void) { return KID_COUNT; }
const char* ub200_prof_kernel_name(int kid) {
    static const char* names[KID_COUNT] = {"gemm1_fwd", "dwconv_fwd", "se_pool", "gemm2_fwd", "residual_fwd", "norm_bwd_stats",
        "gemm2_bwd", "wgrad2", "dwconv_bwd", "gemm1_bwd", "wgrad1", "residual_bwd", "inconv", "maxpool", "ltae", "aggregate",
        "head", "inconv_bwd", "temporal_bwd", "head_bwd"};
    return (kid >= 0 && kid < KID_COUNT) ? names[kid] : "";
}
int ub200_prof_read(int kid, double* total_ms, int* launches) {
    if (!total_ms || !launches) return UB_ERR_ARG;
    double t = 0.0;
    int n = 0;
    for (int i = 0; i < g_prof_n; ++i) {
        if (g_prof[i].kid != kid) continue;
        if (cudaEventSynchronize(g_prof[i].b) != cudaSuccess) return UB_ERR_CUDA;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_prof[i].a, g_prof[i].b) != cudaSuccess) return UB_ERR_CUDA;
        t += ms;
        ++n;
    }
    *total_ms = t;
    *launches = n;
    return UB_OK;
}

int ub200_num_param_slots(const ub200_desc* d) {
    if (!d) return UB_ERR_ARG;
    return UB200_P_BLOCK0 + (1 + d->n_dec_blocks) * UB200_BLOCK_STRIDE;
}

size_t ub200_workspace_bytes(const ub200_desc* d) {
    Layout L;
    if (make_layout(d, L) != UB_OK) return 0;
    return L.total;
}

int ub200_workspace_tap(const ub200_desc* d, const char* name, size_t* offset, size_t* bytes) {
    Layout L;
    if (!name || !offset || !bytes || make_layout(d, L) != UB_OK) return UB_ERR_ARG;
    const size_t P = (size_t)L.P;
    struct { const char* n; size_t off, sz; } fixed[] = {
        {"x0", L.x0, (size_t)L.Ne * P * UB_WIDTH * 4},
        {"pooled", L.pooled, (size_t)L.Ne * UB_LOW * UB_LOW * UB_WIDTH * 4},
        {"pool_idx", L.pool_idx, (size_t)L.Ne * UB_LOW * UB_LOW * UB_WIDTH * 4},
        {"attn", L.attn, (size_t)UB_HEADS * L.Ne * UB_LOW * UB_LOW * 4},
        {"agg", L.agg, (size_t)L.B * P * UB_WIDTH * 4},
        {"notpad", L.notpad, (size_t)L.Ne * 4},
    };
    for (auto& f : fixed)
        if (!strcmp(f.n, name)) { *offset = f.off; *bytes = f.sz; return UB_OK; }
    if (d->use_v) {           // low-resolution value path: MLP output before the BatchNorm1d, its (mean, rstd), the value output
        const size_t Q = (size_t)L.B * UB_LOW * UB_LOW;
        if (!strcmp(name, "v.m")) { *offset = L.V.m; *bytes = Q * UB_WIDTH * 4; return UB_OK; }
        if (!strcmp(name, "v.mr")) { *offset = L.V.mr_m; *bytes = (size_t)L.B * UB_WIDTH * sizeof(MeanRstd); return UB_OK; }
        if (!strcmp(name, "v.v")) { *offset = L.V.v; *bytes = Q * UB_WIDTH * 4; return UB_OK; }
    }
    int bi = -1;
    char what[16];
    if (sscanf(name, "blk%d.%15s", &bi, what) == 2 && bi >= 0 && bi < L.nblk) {
        const size_t N = bi == 0 ? L.Ne : L.B;
        if (L.residual) {           // residual blocks: the block output and the three convolution outputs c1..c3
            if (!strcmp(what, "out")) { *offset = L.rblk[bi].out; *bytes = N * P * UB_WIDTH * 4; return UB_OK; }
            if (what[0] == 'c' && what[1] >= '1' && what[1] <= '3' && !what[2]) { *offset = L.rblk[bi].c[what[1] - '1']; *bytes = N * P * UB_WIDTH * 4; return UB_OK; }
            if (what[0] == 'k' && what[1] >= '1' && what[1] <= '3' && !what[2]) { *offset = L.rblk[bi].coef[what[1] - '1']; *bytes = N * UB_WIDTH * sizeof(Coef); return UB_OK; }
            return UB_ERR_ARG;
        }
        const BlockWs& w = L.blk[bi];
        // forward-only layout: hidden buffers are shared by all blocks and decoder outputs ping-pong; only the encoder output and
        // the last two decoder outputs still hold what their name says after the call
        if (!d->need_grad && (strcmp(what, "out") != 0 || (bi > 0 && bi < L.nblk - 2))) return UB_ERR_ARG;
        if (!strcmp(what, "h1")) { *offset = w.h1; *bytes = N * P * UB_HID * L.hes; return UB_OK; }
        if (!strcmp(what, "h2")) { *offset = w.h2; *bytes = N * P * UB_HID * L.hes; return UB_OK; }
        if (!strcmp(what, "y")) { *offset = w.y; *bytes = N * P * UB_WIDTH * 4; return UB_OK; }
        if (!strcmp(what, "out")) { *offset = w.out; *bytes = N * P * UB_WIDTH * 4; return UB_OK; }
    }
    return UB_ERR_ARG;
}

static BlockCtx make_ctx(const ub200_desc* d, const Layout& L, int i, const void* const* params, void* const* grads, void* ws,
                         cudaStream_t st) {
    BlockCtx c;
    c.p = params + UB200_P_BLOCK0 + i * UB200_BLOCK_STRIDE;
    c.g = grads ? grads + UB200_P_BLOCK0 + i * UB200_BLOCK_STRIDE : nullptr;
    c.w = &L.blk[i];
    c.ws = ws;
    c.N = i == 0 ? L.Ne : L.B;
    c.H = d->H; c.W = d->W;
    c.groups = i == 0 ? d->enc_groups : d->dec_groups;
    c.training = d->training;
    c.backend = d->gemm_backend;
    c.eps = d->norm_eps; c.momentum = d->bn_momentum;
    c.max_parts = L.max_parts > 0 ? L.max_parts : MAX_PARTS;
    c.st = st;
    return c;
}

int ub200_forward(const ub200_desc* d, const float* input, const void* const* params, const unsigned char* keep_mask,
                  float* output, void* ws, size_t ws_bytes, void* stream) {
    return ub200_forward_v(d, input, params, keep_mask, nullptr, output, ws, ws_bytes, stream);
}
int ub200_forward_v(const ub200_desc* d, const float* input, const void* const* params, const unsigned char* keep_mask,
                    const unsigned char* v_keep_mask, float* output, void* ws, size_t ws_bytes, void* stream) {
    Layout L;
    UB_TRY(make_layout(d, L));
    if (!input || !params || !output || !ws) return UB_ERR_ARG;
    if (ws_bytes < L.total) return UB_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int P = L.P;
    if (cudaMemsetAsync(at<char>(ws, L.fwd_zero_begin), 0, L.fwd_zero_end - L.fwd_zero_begin, st) != cudaSuccess) return UB_ERR_CUDA;

    // in_conv: conv1x1 + norm + ReLU (+ pad-mask test), NCHW -> pixel-major
    UB_PROF(KID_INCONV, st, launch_inconv_stats_moments(input, pf(params, UB200_P_IN_W), pf(params, UB200_P_IN_B), at<double>(ws, L.mom_in),
                               at<double>(ws, L.stats_c0), at<int>(ws, L.notpad), d->pad_value, L.Ne, d->C_in, P, st));
    UB_TRY(launch_norm_finalize(at<double>(ws, L.stats_c0), pf(params, UB200_P_IN_NORM_W), pf(params, UB200_P_IN_NORM_B),
                                pfm(params, UB200_P_IN_NORM_RM), pfm(params, UB200_P_IN_NORM_RV), at<Coef>(ws, L.coef_in),
                                at<MeanRstd>(ws, L.mr_in), L.Ne, UB_WIDTH, d->enc_groups, (double)P, d->norm_eps, d->bn_momentum,
                                d->training, st));
    UB_PROF(KID_INCONV, st, launch_inconv_apply(input, pf(params, UB200_P_IN_W), pf(params, UB200_P_IN_B), at<Coef>(ws, L.coef_in),
                               at<float>(ws, L.x0), at<double>(ws, L.blk[0].stats0), L.Ne, d->C_in, P, st));
    // encoder block
    BlockCtx enc = make_ctx(d, L, 0, params, nullptr, ws, st);
    // is_mono (uncrtaints.py:296,418: `--pretrain`, single-date input): no temporal encoder / aggregator, the encoder output IS the
    // decoder input, so the encoder's residual pass gathers the first decoder PreNorm's statistics
    if (L.residual)
        UB_TRY(residual_forward(d, L, 0, params, at<float>(ws, L.x0), nullptr, ws, st));
    else
    UB_TRY(mbconv_forward(enc, at<float>(ws, L.x0), d->is_mono ? at<double>(ws, L.blk[1].stats0) : nullptr, d->need_grad != 0));
    const float* enc_out = at<float>(ws, L.blk[0].out);
    const float* x = enc_out;
    if (!d->is_mono) {
    // temporal path
    UB_PROF(KID_MAXPOOL, st, launch_maxpool_fwd(enc_out, at<float>(ws, L.pooled), at<int>(ws, L.pool_idx), L.Ne, d->H, d->W, st));
    UB_PROF(KID_LTAE, st, launch_ltae_fwd(at<float>(ws, L.pooled), pf(params, UB200_P_LTAE_AP), pf(params, UB200_P_LTAE_E),
                           at<int>(ws, L.notpad), at<float>(ws, L.attn), d->B, d->T, d->norm_eps, st));
    const float drop_p = d->training ? d->dropout_p : 0.f;
    UB_PROF(KID_AGGREGATE, st, launch_aggregate_fwd(at<float>(ws, L.attn), at<int>(ws, L.notpad), keep_mask, d->seed, d->offset, drop_p, enc_out,
                                at<float>(ws, L.agg), at<double>(ws, d->use_v ? L.V.scratch_stats : L.blk[1].stats0), d->B, d->T, d->H, d->W, st));
    x = at<float>(ws, L.agg);
    if (d->use_v) {
        UB_TRY(value_path_forward(d, L, params, v_keep_mask, ws, st));
        x = at<float>(ws, L.V.mix);
    }
    }
    // decoder blocks
    for (int i = 1; i < L.nblk; ++i) {
        BlockCtx c = make_ctx(d, L, i, params, nullptr, ws, st);
        double* next_stats = i + 1 < L.nblk ? at<double>(ws, L.blk[i + 1].stats0) : nullptr;
        if (L.residual)
            UB_TRY(residual_forward(d, L, i, params, x, nullptr, ws, st));
        else if (!d->training && !d->need_grad && d->dec_groups == 0 && (d->gemm_backend & 3) == 3)
            UB_TRY(mbconv_forward_eval_bn(c, x, next_stats));
        else
            UB_TRY(mbconv_forward(c, x, next_stats, d->need_grad != 0));
        x = at<float>(ws, L.blk[i].out);
    }
    UB_PROF(KID_HEAD, st, launch_head_fwd(x, pf(params, UB200_P_OUT_W), pf(params, UB200_P_OUT_B), output, d->B, d->out_dim, P, d->scale_by,
                           d->mean_sigmoid, d->var_eps, st));
    return UB_OK;
}

int ub200_backward(const ub200_desc* d, const float* input, const void* const* params, const unsigned char* keep_mask,
                   const float* output, const float* grad_output, void* const* grads, void* ws, size_t ws_bytes, void* stream) {
    return ub200_backward_v(d, input, params, keep_mask, nullptr, output, grad_output, grads, ws, ws_bytes, stream);
}
int ub200_backward_v(const ub200_desc* d, const float* input, const void* const* params, const unsigned char* keep_mask,
                     const unsigned char* v_keep_mask, const float* output, const float* grad_output, void* const* grads, void* ws,
                     size_t ws_bytes, void* stream) {
    Layout L;
    UB_TRY(make_layout(d, L));
    if (!input || !params || !output || !grad_output || !grads || !ws || !d->need_grad) return UB_ERR_ARG;
    if (ws_bytes < L.total) return UB_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int P = L.P;
    if (cudaMemsetAsync(at<char>(ws, L.bwd_zero_begin), 0, L.bwd_zero_end - L.bwd_zero_begin, st) != cudaSuccess) return UB_ERR_CUDA;
    float* gA = at<float>(ws, L.gA);
    float* gB = at<float>(ws, L.gB);
    float* dn0 = at<float>(ws, L.dn0);
    void* du = at<char>(ws, L.du);
    void* dz1 = at<char>(ws, L.dz1);
    float* partial = at<float>(ws, L.partial);

    const float* dec_out = at<float>(ws, L.blk[L.nblk - 1].out);
    UB_PROF(KID_HEAD_BWD, st, launch_head_bwd(grad_output, output, dec_out, pf(params, UB200_P_OUT_W), gA, gf(grads, UB200_P_OUT_W),
                           gf(grads, UB200_P_OUT_B), d->B, d->out_dim, P, d->scale_by, d->mean_sigmoid, d->var_eps, num_sms(), st));
    for (int i = L.nblk - 1; i >= 1; --i) {
        BlockCtx c = make_ctx(d, L, i, params, grads, ws, st);
        const float* x = i == 1 ? (d->is_mono ? at<float>(ws, L.blk[0].out) : at<float>(ws, d->use_v ? L.V.mix : L.agg)) : at<float>(ws, L.blk[i - 1].out);
        if (L.residual)
            UB_TRY(residual_backward(d, L, i, params, grads, x, gA, gB, static_cast<float*>(du), static_cast<float*>(du) + (size_t)L.Nmax * P * UB_WIDTH,
                                     partial, ws, st));
        else
        // decoder blocks i >= 2: dx is the output gradient of block i-1 (same frame count), whose statistics pass rides in this block's
        // residual pass
        UB_TRY(mbconv_backward(c, x, gA, gB, dn0, du, dz1, partial, /*stats_ready=*/i < L.nblk - 1, i >= 2 ? &L.blk[i - 1] : nullptr));
        float* t = gA; gA = gB; gB = t;
    }
    if (d->is_mono) {         // gA = dEnc directly: make it the encoder block's incoming gradient
        float* t = gA; gA = gB; gB = t;
    } else {
    // gA = dAgg (use_v: the gradient of the include_v output; dAgg = gA . W_a goes to dn0, free at this point).
    // Temporal path backward; dEnc accumulates in gB.
    const float* enc_out = at<float>(ws, L.blk[0].out);
    const float drop_p = d->training ? d->dropout_p : 0.f;
    const float* dagg = gA;
    if (d->use_v) {
        UB_TRY(value_path_backward_head(d, L, params, grads, gA, dn0, partial, ws, st));
        dagg = dn0;
    }
    UB_PROF(KID_TEMPORAL_BWD, st, launch_aggregate_bwd(at<float>(ws, L.attn), at<int>(ws, L.notpad), keep_mask, d->seed, d->offset, drop_p, enc_out, dagg,
                                gB, at<float>(ws, L.dwup), at<float>(ws, L.dattn), d->B, d->T, d->H, d->W, st));
    if (d->use_v)
        UB_TRY(value_path_backward_tail(d, L, params, grads, v_keep_mask, partial, ws, st));
    else
    UB_PROF(KID_TEMPORAL_BWD, st, launch_ltae_bwd(at<float>(ws, L.pooled), pf(params, UB200_P_LTAE_AP), at<float>(ws, L.attn), at<float>(ws, L.dattn),
                           at<float>(ws, L.dpooled), gf(grads, UB200_P_LTAE_AP), gf(grads, UB200_P_LTAE_E), d->B, d->T,
                           d->norm_eps, st));
    UB_PROF(KID_TEMPORAL_BWD, st, launch_maxpool_bwd(at<float>(ws, L.dpooled), at<int>(ws, L.pool_idx), gB, L.Ne, P, st));
    }
    BlockCtx enc = make_ctx(d, L, 0, params, grads, ws, st);
    enc.relu_mask_dx = 1;                     // the gram pass then consumes dgn = dX0 * [x0 > 0] directly
    if (L.residual)
        UB_TRY(residual_backward(d, L, 0, params, grads, at<float>(ws, L.x0), gB, gA, static_cast<float*>(du),
                                 static_cast<float*>(du) + (size_t)L.Nmax * P * UB_WIDTH, partial, ws, st));
    else
    UB_TRY(mbconv_backward(enc, at<float>(ws, L.x0), gB, gA, dn0, du, dz1, partial));
    // in_conv backward (no input gradient); MBConv: the ReLU mask of in_conv was applied by the encoder block's residual_bwd
    UB_PROF(KID_INCONV_BWD, st, launch_inconv_bwd_gram(input, L.residual ? at<float>(ws, L.x0) : nullptr, gA, pf(params, UB200_P_IN_W), pf(params, UB200_P_IN_B),
                               at<MeanRstd>(ws, L.mr_in), at<double>(ws, L.gram_in), at<double>(ws, L.bstats_in), L.Ne, d->C_in, P, st));
    UB_TRY(launch_norm_finalize_bwd(at<double>(ws, L.bstats_in), pf(params, UB200_P_IN_NORM_W), at<MeanRstd>(ws, L.mr_in),
                                    at<BCoef>(ws, L.bc_in), gf(grads, UB200_P_IN_NORM_W), gf(grads, UB200_P_IN_NORM_B), L.Ne,
                                    UB_WIDTH, d->enc_groups, (double)P, d->training, st));
    UB_PROF(KID_INCONV_BWD, st, launch_inconv_bwd_finish(at<double>(ws, L.gram_in), at<double>(ws, L.mom_in), pf(params, UB200_P_IN_W),
                               pf(params, UB200_P_IN_B), at<BCoef>(ws, L.bc_in), gf(grads, UB200_P_IN_W), gf(grads, UB200_P_IN_B),
                               L.Ne, d->C_in, P, st));
    return UB_OK;
}

int ub200_mgnll_forward(const float* pred, long long pred_sb, const float* target, long long targ_sb, const float* var,
                        long long var_sb, int var_ch, int B, int P, float eps, float* loss, float* dpred, float* dvar,
                        int* neg_flag, void* scratch, void* stream) {
    if (!pred || !target || !var || !loss || !neg_flag || !scratch || (var_ch != 1 && var_ch != UB_S2) || B < 1 || P < 1)
        return UB_ERR_ARG;
    if ((dpred == nullptr) != (dvar == nullptr)) return UB_ERR_ARG;
    return launch_mgnll(pred, pred_sb, target, targ_sb, var, var_sb, var_ch, dpred, dvar, static_cast<double*>(scratch),
                        neg_flag, loss, B, P, eps, static_cast<cudaStream_t>(stream));
}

int ub200_gnll_forward(const float* pred, long long pred_sb, const float* target, long long targ_sb, const float* var,
                       long long var_sb, int B, int P, float eps, int full, float* loss, float* dpred, float* dvar, float* var_out,
                       int* neg_flag, void* scratch, void* stream) {
    if (!pred || !target || !var || !loss || !neg_flag || !scratch || B < 1 || P < 1) return UB_ERR_ARG;
    if ((dpred == nullptr) != (dvar == nullptr)) return UB_ERR_ARG;
    return launch_gnll(pred, pred_sb, target, targ_sb, var, var_sb, dpred, dvar, var_out, static_cast<double*>(scratch), neg_flag, loss,
                       B, P, eps, full, static_cast<cudaStream_t>(stream));
}

int ub200_mgnll_none(const float* pred, long long pred_sb, const float* target, long long targ_sb, const float* var, long long var_sb,
                     int var_ch, int B, int P, float eps, const float* grad_loss, float* loss, float* dpred, float* dvar, int* neg_flag,
                     void* stream) {
    if (!pred || !target || !var || (var_ch != 1 && var_ch != UB_S2) || B < 1 || P < 1) return UB_ERR_ARG;
    if (grad_loss ? (!dpred || !dvar) : (!loss || !neg_flag)) return UB_ERR_ARG;
    return launch_mgnll_none(pred, pred_sb, target, targ_sb, var, var_sb, var_ch, grad_loss, loss, dpred, dvar, neg_flag, B, P, eps,
                             static_cast<cudaStream_t>(stream));
}

int ub200_gnll_none(const float* pred, long long pred_sb, const float* target, long long targ_sb, const float* var, long long var_sb,
                    int B, int P, float eps, int full, const float* grad_loss, float* loss, float* var_out, float* dpred, float* dvar,
                    int* neg_flag, void* stream) {
    if (!pred || !target || !var || B < 1 || P < 1) return UB_ERR_ARG;
    if (grad_loss ? (!dpred || !dvar) : (!loss || !neg_flag)) return UB_ERR_ARG;
    return launch_gnll_none(pred, pred_sb, target, targ_sb, var, var_sb, grad_loss, loss, var_out, dpred, dvar, neg_flag, B, P, eps, full,
                            static_cast<cudaStream_t>(stream));
}

int ub200_scale_by_scalar(const float* in, const float* grad_loss, float* out, size_t n, void* stream) {
    if (!in || !grad_loss || !out) return UB_ERR_ARG;
    return launch_scale_by_scalar(in, grad_loss, out, n, static_cast<cudaStream_t>(stream));
}

int ub200_covariance(const float* var, long long var_sb, int var_ch, int B, int P, float eps, float* cov, void* stream) {
    if (!var || !cov || (var_ch != 1 && var_ch != UB_S2)) return UB_ERR_ARG;
    return launch_covariance(var, var_sb, var_ch, cov, B, P, eps, static_cast<cudaStream_t>(stream));
}

// ---- fused Adam over flat buffers (torch.optim.Adam semantics; base_model.py:48-51,122) ----
int ub200_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, size_t n, int step, float lr, float beta1,
                    float beta2, float eps, float weight_decay, float grad_scale, int zero_grad, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || step < 1) return UB_ERR_ARG;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    return launch_adam_step(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, (float)(1.0 / bc1),
                            (float)(1.0 / sqrt(bc2)), grad_scale, zero_grad, static_cast<cudaStream_t>(stream));
}

// ---- device-side batch assembly (prepare_data_multi, train_reconstruct.py:161-179) ----
int ub200_assemble_input(const void* const* src_table, float* x, int B, int T, int c_s1, int c_s2, int P, void* stream) {
    if (!src_table || !x) return UB_ERR_ARG;
    return launch_assemble_input(reinterpret_cast<const float* const*>(src_table), x, B, T, c_s1, c_s2, P, static_cast<cudaStream_t>(stream));
}

// ---- image metrics of the validation loop (metrics.py:20-57, pytorch_ssim) for a whole batch ----
int ub200_img_metrics(const float* target, const float* pred, const float* var, int B, int H, int W, double* acc, float* pixelwise,
                      void* stream) {
    if (!target || !pred || !acc || B < 1 || H < 1 || W < 1) return UB_ERR_ARG;
    return launch_img_metrics(target, pred, var, acc, pixelwise, B, H, W, static_cast<cudaStream_t>(stream));
}

// ---- standalone out_conv + head (tests, calibration sweep) ----
int ub200_head_forward(const float* dec, const float* w, const float* bias, float* out, int B, int out_dim, int P, float scale_by,
                       int mean_sigmoid, float var_eps, void* stream) {
    if (!dec || !w || !bias || !out || B < 1) return UB_ERR_ARG;
    return launch_head_fwd(dec, w, bias, out, B, out_dim, P, scale_by, mean_sigmoid, var_eps, static_cast<cudaStream_t>(stream));
}
int ub200_head_backward(const float* grad_out, const float* out, const float* dec, const float* w, float* ddec, float* dw, float* db,
                        int B, int out_dim, int P, float scale_by, int mean_sigmoid, float var_eps, void* stream) {
    if (!grad_out || !out || !dec || !w || !ddec || !dw || !db || B < 1) return UB_ERR_ARG;
    return launch_head_bwd(grad_out, out, dec, w, ddec, dw, db, B, out_dim, P, scale_by, mean_sigmoid, var_eps, num_sms(),
                           static_cast<cudaStream_t>(stream));
}

// ---- standalone MBConv block (tests) -------------------------------------------------------------------
struct MbLayout { BlockWs w; size_t zero_begin, zero_end, bzero_begin, bzero_end, dn0, du, dz1, partial, total; int max_parts; };
static void mb_layout(int N, int H, int W, MbLayout& M, size_t hes = 4) {
    Bump b;
    M.zero_begin = b.off;
    block_fwd_stats(b, M.w, N);
    M.zero_end = b.off;
    M.bzero_begin = b.off;
    block_bwd_stats(b, M.w, N);
    M.bzero_end = b.off;
    block_rest(b, M.w, N, (size_t)H * W, true, hes);
    M.dn0 = b.take((size_t)N * H * W * UB_WIDTH * 4);
    M.du = b.take((size_t)N * H * W * UB_HID * hes);
    M.dz1 = b.take((size_t)N * H * W * UB_HID * hes);
    M.max_parts = MAX_PARTS;
    M.partial = b.take((size_t)M.max_parts * UB_WIDTH * UB_HID * 4);
    M.total = b.off;
}
size_t ub200_mbconv_workspace_bytes(int N, int H, int W) {
    MbLayout M;
    mb_layout(N, H, W, M);
    return M.total;
}
static BlockCtx mb_ctx(const MbLayout& M, const void* const* p, void* const* g, void* ws, int N, int H, int W, int groups,
                       int training, float eps, float momentum, int backend, cudaStream_t st) {
    BlockCtx c;
    c.p = p; c.g = g; c.w = &M.w; c.ws = ws; c.N = N; c.H = H; c.W = W; c.groups = groups; c.training = training;
    c.backend = backend; c.eps = eps; c.momentum = momentum; c.st = st;
    c.max_parts = M.max_parts;
    return c;
}
int ub200_mbconv_forward(const float* x, const void* const* block_params, int N, int H, int W, int groups, int training,
                         float eps, float momentum, int gemm_backend, float* out, void* ws, size_t ws_bytes, void* stream) {
    if (!x || !block_params || !out || !ws || N < 1 || H % 8 || W % 16) return UB_ERR_ARG;
    MbLayout M;
    mb_layout(N, H, W, M, (gemm_backend & 32) ? 2 : 4);      // the workspace size query assumes fp32 storage (an upper bound)
    if (ws_bytes < M.total) return UB_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int P = H * W;
    if (cudaMemsetAsync(at<char>(ws, M.zero_begin), 0, M.zero_end - M.zero_begin, st) != cudaSuccess) return UB_ERR_CUDA;
    const int chunk = 256;
    colstats_kernel<<<dim3((P + chunk - 1) / chunk, N), 256, 0, st>>>(x, at<double>(ws, M.w.stats0), P, chunk);
    UB_CHECK_LAUNCH();
    BlockCtx c = mb_ctx(M, block_params, nullptr, ws, N, H, W, groups, training, eps, momentum, gemm_backend, st);
    UB_TRY(mbconv_forward(c, x, nullptr, true));
    if (cudaMemcpyAsync(out, at<float>(ws, M.w.out), (size_t)N * P * UB_WIDTH * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return UB_ERR_CUDA;
    return UB_OK;
}
int ub200_mbconv_backward(const float* x, const void* const* block_params, const float* dout, void* const* block_grads, int N,
                          int H, int W, int groups, int training, int gemm_backend, float* dx, void* ws, size_t ws_bytes,
                          void* stream) {
    if (!x || !block_params || !dout || !block_grads || !dx || !ws) return UB_ERR_ARG;
    MbLayout M;
    mb_layout(N, H, W, M, (gemm_backend & 32) ? 2 : 4);
    if (ws_bytes < M.total) return UB_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(at<char>(ws, M.bzero_begin), 0, M.bzero_end - M.bzero_begin, st) != cudaSuccess) return UB_ERR_CUDA;
    BlockCtx c = mb_ctx(M, block_params, block_grads, ws, N, H, W, groups, training, 1e-5f, 0.1f, gemm_backend, st);
    return mbconv_backward(c, x, dout, dx, at<float>(ws, M.dn0), at<char>(ws, M.du), at<char>(ws, M.dz1),
                           at<float>(ws, M.partial));
}

}  // extern "C"
