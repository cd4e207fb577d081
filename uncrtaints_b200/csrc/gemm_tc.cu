// tcgen05 (5th-gen tensor core) streaming GEMMs for the MBConv 1x1 convolutions and their input-gradient passes
// (model/src/backbones/uncrtaints.py:126,136 and their autograd backward).
//
//   D[ch_out (128 TMEM lanes per M-block), px (128 TMEM columns)] = sum_k W[ch_out][k] * f(act)[px][k]
//
// * fp32 parity on bf16 tensor cores: every operand is split x = hi + lo (two bf16), and three MMAs accumulate
//   W_hi*A_hi + W_hi*A_lo + W_lo*A_hi in the fp32 TMEM accumulator (relative error ~1e-5; single-pass TF32
//   measured 1e-3..8e-3 on this network's gradients and fails the 1e-3 parity target).
// * the activation operand cannot come straight from HBM by TMA: the normalisation / GELU / SE gate / norm-backward
//   must be applied first (the statistics barrier forces the apply into the consumer).  All 512 threads load
//   fp32 rows with coalesced 32-byte pieces, transform in registers, split, and write the bf16 hi/lo tiles into
//   shared memory in the canonical K-major SWIZZLE_128B layout (8 rows x 128 B atoms, 16-byte chunk index XOR
//   row%8), then fence.proxy.async and hand them to the single MMA-issuing thread.
// * weights are the M operand (prepared once per step as a swizzled bf16 hi/lo image, 128 KB, resident in shared
//   memory); pixels are the N operand.  With channels on TMEM lanes the epilogue is trivial: tcgen05.ld 32x32b
//   gives each thread ONE channel x 32 pixels, so global stores are 128-byte coalesced across the warp, the
//   per-channel statistics are plain per-thread sums, and per-channel coefficients are thread constants.
// * pipeline: 2-slot shared-memory ring of 64-channel K-blocks (tcgen05.commit -> mbarrier frees a slot) and a
//   double-buffered TMEM accumulator (2 x 256 columns): the MMAs of tile t run while the threads do the epilogue
//   of tile t-1 and load tile t+1.
#include <cuda.h>            // CUtensorMap (types only: cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint, no libcuda link)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "kernels.h"

#ifndef UB_TRY
#define UB_TRY(expr)                    \
    do {                                \
        int rc__ = (expr);              \
        if (rc__ != UB_OK) return rc__; \
    } while (0)
#endif

namespace ub {
namespace tc {

constexpr int THREADS = 512;
constexpr int TILE_PX = 128;          // UMMA N
constexpr int KBLK = 64;              // channels per K-block (one 128-byte swizzle row of bf16)
constexpr int NSTAGE = 2;
constexpr int STAGE_BYTES = 2 * TILE_PX * KBLK * 2;   // hi + lo tiles, 32 KB

// shared-memory matrix descriptor constants (layout of cute/arch/mma_sm100_desc.hpp)
constexpr uint32_t c_desc_hi = (64u) | (1u << 14) | (2u << 29);   // SBO=1024B>>4, version=1, SWIZZLE_128B
constexpr uint32_t c_desc_lbo = 1u;                                // LBO field (ignored for SW128 K-major)
// `single` kernel argument = 1: single-pass bf16 (gemm_backend bit 2; BASELINE config #3's "bf16 tensor-core path"): operands are
// rounded to bf16 once, one MMA per k-step instead of three, the lo tiles are neither written nor read.  ~3e-3 relative
// accuracy instead of ~1e-5.
constexpr uint32_t c_idesc = (1u << 4) | (1u << 7) | (1u << 10) | (16u << 17) | (8u << 24);  // F32 acc, BF16 x BF16, N=128, M=128

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(c_desc_lbo & 0x3FFFu) << 16) | ((uint64_t)c_desc_hi << 32);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// N = 8 / 16 / 32 consecutive TMEM columns of this thread's lane (tcgen05.ld.32x32b.xN)
template <int N>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, float* v) {
    static_assert(N == 8 || N == 16 || N == 32, "unsupported tcgen05.ld width");
    if constexpr (N == 32) {
        float t[32];
        tmem_ld32(taddr, t);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = t[i];
    } else if constexpr (N == 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
    } else {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
    }
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Operand split modes ("single" argument of the kernels): 0 = bf16 hi + bf16 lo (three MMAs; ~2^-17: the input-/weight-gradient
// GEMMs, whose operands are gradients of unbounded range), 1 = one bf16 (single-pass, reduced precision), 2 = fp16 hi + fp16 lo
// (three MMAs; ~2^-22, i.e. fp32-grade: the FORWARD GEMMs, whose operands are normalised activations and weights, well inside
// fp16's range -- conversions saturate instead of overflowing).  Mode 2 exists because the max-pool argmax downstream of the
// encoder is discrete: a forward error of 6e-6 flips a handful of near-tie windows, which alone costs 5-9e-4 on the encoder
// gradients (tests/test_gpu_parity.py::test_headline_resolution_default_backend); with 2^-22 operands the forward is as accurate as
// an fp32 FMA chain and the flips all but disappear.
constexpr int SPLIT_BF16X3 = 0, SPLIT_BF16X1 = 1, SPLIT_F16X3 = 2;
// split 8 fp32 values into hi / lo halves (x ~= hi + lo) and store both 16-byte chunks
__device__ __forceinline__ void split_store8(const float (&v)[8], char* hi_chunk, char* lo_chunk, int single = 0) {
    uint32_t h[4], l[4];
    if (single == SPLIT_F16X3) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float a = v[2 * i], b = v[2 * i + 1];
            uint32_t hp, lp;
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hp) : "f"(b), "f"(a));   // upper half <- b, lower half <- a
            float ah, bh;
            asm("{\n\t.reg .b16 lo16, hi16;\n\tmov.b32 {lo16, hi16}, %2;\n\tcvt.f32.f16 %0, lo16;\n\tcvt.f32.f16 %1, hi16;\n\t}"
                : "=f"(ah), "=f"(bh) : "r"(hp));
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lp) : "f"(b - bh), "f"(a - ah));
            h[i] = hp;
            l[i] = lp;
        }
        *reinterpret_cast<uint4*>(hi_chunk) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(lo_chunk) = make_uint4(l[0], l[1], l[2], l[3]);
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        uint32_t hp;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hp) : "f"(b), "f"(a));     // upper half <- b, lower half <- a
        h[i] = hp;
        if (single != SPLIT_BF16X1) {
            const float ah = __uint_as_float(hp << 16), bh = __uint_as_float(hp & 0xFFFF0000u);
            uint32_t lp;
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lp) : "f"(b - bh), "f"(a - ah));
            l[i] = lp;
        }
    }
    *reinterpret_cast<uint4*>(hi_chunk) = make_uint4(h[0], h[1], h[2], h[3]);
    if (single != SPLIT_BF16X1) *reinterpret_cast<uint4*>(lo_chunk) = make_uint4(l[0], l[1], l[2], l[3]);
}

// ------------------------------------------------------------------------------------------
// operand loaders.  issue(): raw global loads only (kept in flight across the previous K-block's convert / MMA
// issue / epilogue).  finish(): apply the transform with the per-frame coefficients, which live in shared memory as
// three structure-of-arrays vectors cf[0..K), cf[K..2K), cf[2K..3K) (conflict-free 32-byte reads per lane).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld4_nc(const float* p) {      // read-only path WITH L1 allocation: the two 16-byte halves of
    float4 r;                                                     // a lane's 32-byte piece share sectors
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
    const float4 a = ld4_nc(p), b = ld4_nc(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// Coefficient vectors are stored "quad-split": the first float4 of every 8-channel chunk in the lower half of the vector,
// the second in the upper half, so that the two LDS.128 of a lane are 16-byte contiguous across lanes (no bank conflicts).
__device__ __forceinline__ int cf_pos(int k, int K) { return ((k >> 2) & 1) * (K / 2) + (k >> 3) * 4 + (k & 3); }
__device__ __forceinline__ void lds8(const float* vec, int K, int ch0, float (&v)[8]) {      // ch0 % 8 == 0
    const float4 a = *reinterpret_cast<const float4*>(vec + (ch0 >> 1)), b = *reinterpret_cast<const float4*>(vec + K / 2 + (ch0 >> 1));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// Storage type of the 256-channel hidden tensors (h1, h2 and the gradients du, dz1): float, or __nv_bfloat16 with gemm_backend bit 5
// (BASELINE config #3: "bf16 tensor-core path" -- these four tensors are 16 of the 34 tensor passes of an MBConv frame and two
// thirds of its bytes).  Raw8 = 8 consecutive elements as loaded (32 or 16 bytes in flight per item).
typedef __nv_bfloat16 bf16_t;
template <class T> struct Raw8;
template <> struct Raw8<float> { float a[8]; };
template <> struct Raw8<bf16_t> { uint4 p; };
__device__ __forceinline__ void ld8_raw(const float* p, Raw8<float>& r) { ld8(p, r.a); }
__device__ __forceinline__ void ld8_raw(const bf16_t* p, Raw8<bf16_t>& r) {
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.p.x), "=r"(r.p.y), "=r"(r.p.z), "=r"(r.p.w) : "l"(p));
}
__device__ __forceinline__ void unpack8(const Raw8<float>& r, float (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = r.a[i];
}
__device__ __forceinline__ void unpack8(const Raw8<bf16_t>& r, float (&v)[8]) {
    const uint32_t w[4] = {r.p.x, r.p.y, r.p.z, r.p.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
}
__device__ __forceinline__ float ldh(const float* p) { return *p; }
__device__ __forceinline__ float ldh(const bf16_t* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void sth(float* p, float v) { *p = v; }
__device__ __forceinline__ void sth(bf16_t* p, float v) { *p = __float2bfloat16_rn(v); }

// Loaders with `static constexpr bool PRESPLIT = true` read operands that were split into hi / lo halves beforehand (split_act_kernel:
// a [rows][128] image of 512-byte rows, 128 hi halves then 128 lo halves): their Raw is the two 16-byte chunks as loaded and put()
// copies them straight into the swizzled tiles -- no conversion in the GEMM.  Used where an element is read many times (the nine taps
// of a 3x3 convolution).
template <class L, class = void> struct IsPresplit { static constexpr bool v = false; };
template <class L> struct IsPresplit<L, decltype((void)L::PRESPLIT)> { static constexpr bool v = L::PRESPLIT; };
// Per-row position state for loaders whose source address needs integer division (the convolution loaders: frame / y / x of the
// output pixel).  It changes once per tile while issue() runs once per K-block (18 times per tile for the 3x3 convolution), and
// the producer warps are bound by the latency of their own instruction stream (ncu r02, conv kernel: 250 instructions per thread
// and K-block, 10.7 cycles per issued instruction at four warps per scheduler), so the divisions are hoisted out of the K loop.
// loaders of a plain row-major [rows][K] tensor offer issue_off(row * K + ch0, raw): the persistent kernels advance that offset by
// constants instead of rebuilding the address from the tile index in every step
template <class L, class = void> struct HasIssueOff { static constexpr bool v = false; };
template <class L> struct HasIssueOff<L, decltype((void)&L::issue_off)> { static constexpr bool v = true; };
// gemm_tc_body spreads a step's prefetch over the step (second row's loads after the first row's convert) for loaders that set
// SPREAD_ROWS: measured gemm2_fwd 3.47 -> 3.20 ms per step (GELU loader), gemm1_fwd neutral, gemm2_bwd 4.82 -> 4.95 (two tensors per
// row: its loads want the whole step in flight)
template <class L, class = void> struct SpreadRows { static constexpr bool v = false; };
template <class L> struct SpreadRows<L, decltype((void)L::SPREAD_ROWS)> { static constexpr bool v = L::SPREAD_ROWS; };
template <class L, class = void> struct PosOf { struct type { size_t row; }; };
template <class L> struct PosOf<L, decltype((void)sizeof(typename L::Pos))> { typedef typename L::Pos type; };
template <class L>
__device__ __forceinline__ void locate_row(const L& l, size_t row, typename PosOf<L>::type& pos) {
    if constexpr (IsPresplit<L>::v) l.locate(row, pos); else pos.row = row;
}
template <class L>
__device__ __forceinline__ void issue_row(const L& l, const typename PosOf<L>::type& pos, int K, int ch0, typename L::Raw& raw) {
    if constexpr (IsPresplit<L>::v) l.issue_at(pos, K, ch0, raw); else l.issue(pos.row, K, ch0, raw);
}
template <class L>
__device__ __forceinline__ void convert_store(const L& l, const typename L::Raw& raw, const typename L::Cf& cf, char* hi_chunk, char* lo_chunk,
                                              int single) {
    if constexpr (IsPresplit<L>::v) {
        l.put(raw, hi_chunk, lo_chunk, single);
    } else {
        float v[8];
        l.finish(raw, cf, v);
        split_store8(v, hi_chunk, lo_chunk, single);
    }
}

struct TLoadNormed {           // a = x*scale + shift
    const float* x; const Coef* coef;
    struct Raw { float a[8]; };
    __device__ void fill(int n, int K, float* cf) const {
        for (int k = threadIdx.x; k < K; k += THREADS) { const Coef c = coef[(size_t)n * K + k]; const int q = cf_pos(k, K); cf[q] = c.scale; cf[K + q] = c.shift; }
    }
    __device__ void issue(size_t row, int K, int ch0, Raw& r) const { ld8(x + row * K + ch0, r.a); }
    __device__ void issue_off(size_t off, Raw& r) const { ld8(x + off, r.a); }                  // off = row * K + ch0
    struct Cf { float sc[8], sh[8]; };
    __device__ void coefs(int K, int ch0, const float* cf, Cf& c) const { lds8(cf, K, ch0, c.sc); lds8(cf + K, K, ch0, c.sh); }
    __device__ void finish(const Raw& r, const Cf& c, float (&v)[8]) const {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(r.a[i], c.sc[i], c.sh[i]);
    }
};
template <class HT>
struct TLoadGeluGateT {        // a = gelu(h2*scale + shift) * gate
    static constexpr bool SPREAD_ROWS = true;
    const HT* h2; const Coef* coef; const float* gate;
    typedef Raw8<HT> Raw;
    __device__ void fill(int n, int K, float* cf) const {
        for (int k = threadIdx.x; k < K; k += THREADS) {
            const Coef c = coef[(size_t)n * K + k];
            const int q = cf_pos(k, K);
            cf[q] = c.scale; cf[K + q] = c.shift; cf[2 * K + q] = gate[(size_t)n * K + k];
        }
    }
    __device__ void issue(size_t row, int K, int ch0, Raw& r) const { ld8_raw(h2 + row * K + ch0, r); }
    __device__ void issue_off(size_t off, Raw& r) const { ld8_raw(h2 + off, r); }                 // off = row * K + ch0
    struct Cf { float sc[8], sh[8], g[8]; };
    __device__ void coefs(int K, int ch0, const float* cf, Cf& c) const {
        lds8(cf, K, ch0, c.sc); lds8(cf + K, K, ch0, c.sh); lds8(cf + 2 * K, K, ch0, c.g);
    }
    __device__ void finish(const Raw& r, const Cf& c, float (&v)[8]) const {
        float a[8];
        unpack8(r, a);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {          // packed f32x2 single-MUFU GELU: 4 pairs
            float g0, g1;
            gelu_val_pair(fmaf(a[i], c.sc[i], c.sh[i]), fmaf(a[i + 1], c.sc[i + 1], c.sh[i + 1]), g0, g1);
            v[i] = g0 * c.g[i];
            v[i + 1] = g1 * c.g[i + 1];
        }
    }
};
typedef TLoadGeluGateT<float> TLoadGeluGate;
template <class T>
struct TLoadNormBwdT {         // a = ca*dy + cb*v + cc
    const T* dy; const T* vv; const BCoef* bc;
    struct Raw { Raw8<T> a, b; };
    __device__ void fill(int n, int K, float* cf) const {
        for (int k = threadIdx.x; k < K; k += THREADS) { const BCoef c = bc[(size_t)n * K + k]; const int q = cf_pos(k, K); cf[q] = c.a; cf[K + q] = c.b; cf[2 * K + q] = c.c; }
    }
    __device__ void issue(size_t row, int K, int ch0, Raw& r) const { ld8_raw(dy + row * K + ch0, r.a); ld8_raw(vv + row * K + ch0, r.b); }
    __device__ void issue_off(size_t off, Raw& r) const { ld8_raw(dy + off, r.a); ld8_raw(vv + off, r.b); }      // off = row * K + ch0
    struct Cf { float ca[8], cb[8], cc[8]; };
    __device__ void coefs(int K, int ch0, const float* cf, Cf& c) const {
        lds8(cf, K, ch0, c.ca); lds8(cf + K, K, ch0, c.cb); lds8(cf + 2 * K, K, ch0, c.cc);
    }
    __device__ void finish(const Raw& r, const Cf& c, float (&v)[8]) const {
        float a[8], b[8];
        unpack8(r.a, a);
        unpack8(r.b, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(c.ca[i], a[i], fmaf(c.cb[i], b[i], c.cc[i]));
    }
};
typedef TLoadNormBwdT<float> TLoadNormBwd;

// 3x3 convolution operands (ResidualConvBlock, uncrtaints.py:24-69; nn.Conv2d(k=3, padding=1, padding_mode='reflect'), utae.py:478-487).
// The GEMM's K index is (tap, channel), k = tap*128 + c, and the operand row of output pixel p for tap (dy, dx) is the input pixel
// p + (dy, dx) -- reflected at the border (forward, weight gradient) or ZERO outside the image (input gradient: the transposed
// convolution of the zero-extended output gradient; the adjoint of the reflection is added by conv_fold_kernel).
// Every input element is read nine times, so the previous layer's norm + ReLU and the hi / lo split are applied ONCE into a
// pre-split image (split_act_kernel / relu_norm_bwd_apply_kernel) and the loaders only copy 16-byte chunks.
// paired: the operand has 256 "channels" = taps (tap0, tap1) side by side (the N operand of the weight-gradient GEMM); tap1 < 0 -> zeros.
constexpr int SPLIT_ROW = UB_SPLIT_ROW; // bytes per pixel of a pre-split image (layout: common.cuh, split_pixel_offset)
struct TLoadConvSplit {
    static constexpr bool PRESPLIT = true;
    const char* img; int H, W, zero_pad, paired, tap0, tap1;
    struct Raw { uint4 h, l; };
    struct Pos { const char* frame; int y, x; };          // frame base of the image, pixel coordinates of the operand row
    __device__ void fill(int, int, float*) const {}
    __device__ void locate(size_t row, Pos& pos) const {
        const uint32_t P = (uint32_t)(H * W), row32 = (uint32_t)row;
        const uint32_t n = row32 / P, p = row32 - n * P;
        pos.y = (int)(p / (uint32_t)W);
        pos.x = (int)(p - (uint32_t)pos.y * (uint32_t)W);
        pos.frame = img + split_pixel_offset(n, 0, -1, H, W);          // first (halo) pixel of the frame
    }
    __device__ void issue(size_t row, int K, int ch0, Raw& r) const {
        Pos pos;
        locate(row, pos);
        issue_at(pos, K, ch0, r);
    }
    __device__ void issue_at(const Pos& pos, int /*K*/, int ch0, Raw& r) const {
        const int tap = paired ? (ch0 < 128 ? tap0 : tap1) : (ch0 >> 7);
        const int t3 = tap >= 6 ? 2 : (tap >= 3 ? 1 : 0);              // tap / 3 without a division (tap may be -1: absent)
        int yy = pos.y + t3 - 1, xs = pos.x + (tap - 3 * t3) - 1;
        bool ok = tap >= 0;
        if (zero_pad) {
            ok = ok && yy >= 0 && yy < H && xs >= 0 && xs < W;
        } else {
            yy = yy < 0 ? 1 : (yy >= H ? H - 2 : yy);
            xs = xs < 0 ? 1 : (xs >= W ? W - 2 : xs);
        }
        if (ok) {
            const char* src = pos.frame + (size_t)(yy * (W + 2) + xs + 1) * SPLIT_ROW + (ch0 & 127) * 2;
            asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.h.x), "=r"(r.h.y), "=r"(r.h.z), "=r"(r.h.w) : "l"(src));
            asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.l.x), "=r"(r.l.y), "=r"(r.l.z), "=r"(r.l.w) : "l"(src + 256));
        } else {
            r.h = make_uint4(0, 0, 0, 0);
            r.l = make_uint4(0, 0, 0, 0);
        }
    }
    struct Cf {};
    __device__ void coefs(int, int, const float*, Cf&) const {}
    __device__ void put(const Raw& r, char* hi_chunk, char* lo_chunk, int single) const {
        *reinterpret_cast<uint4*>(hi_chunk) = r.h;
        if (single != SPLIT_BF16X1) *reinterpret_cast<uint4*>(lo_chunk) = r.l;
    }
};
struct TLoadPlainSplit {       // pixels of a pre-split image as they are (A operand of the convolution weight gradient)
    static constexpr bool PRESPLIT = true;
    const char* img; int H, W;
    struct Raw { uint4 h, l; };
    struct Pos { size_t row; };
    __device__ void locate(size_t row, Pos& pos) const { pos.row = row; }
    __device__ void issue_at(const Pos& pos, int K, int ch0, Raw& r) const { issue(pos.row, K, ch0, r); }
    __device__ void fill(int, int, float*) const {}
    __device__ void issue(size_t row, int /*K*/, int ch0, Raw& r) const {
        const uint32_t P = (uint32_t)(H * W), row32 = (uint32_t)row;
        const uint32_t n = row32 / P, p = row32 - n * P;
        const int y = (int)(p / (uint32_t)W), x = (int)(p - (uint32_t)y * (uint32_t)W);
        const char* src = img + split_pixel_offset(n, y, x, H, W) + ch0 * 2;
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.h.x), "=r"(r.h.y), "=r"(r.h.z), "=r"(r.h.w) : "l"(src));
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.l.x), "=r"(r.l.y), "=r"(r.l.z), "=r"(r.l.w) : "l"(src + 256));
    }
    struct Cf {};
    __device__ void coefs(int, int, const float*, Cf&) const {}
    __device__ void put(const Raw& r, char* hi_chunk, char* lo_chunk, int single) const {
        *reinterpret_cast<uint4*>(hi_chunk) = r.h;
        if (single != SPLIT_BF16X1) *reinterpret_cast<uint4*>(lo_chunk) = r.l;
    }
};
// out = split(relu?(x*scale + shift)): the pre-split image of a [N][P][128] fp32 tensor.  thread = 8 channels of one pixel.
__global__ void __launch_bounds__(256) split_act_kernel(const float* __restrict__ x, const Coef* __restrict__ coef, char* __restrict__ out,
                                                        int H, int W, size_t rows, int relu, int mode) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= rows * 16) return;
    const size_t row = i / 16;
    const int c0 = (int)(i % 16) * 8;
    const uint32_t P = (uint32_t)(H * W), n = (uint32_t)row / P, p = (uint32_t)row - n * P;
    const int y = (int)(p / (uint32_t)W), xx = (int)(p - (uint32_t)y * (uint32_t)W);
    float a[8], v[8];
    ld8(x + row * 128 + c0, a);
    if (coef) {
        const Coef* k = coef + (size_t)n * 128 + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], k[j].scale, k[j].shift);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = relu ? fmaxf(a[j], 0.f) : a[j];
    char* dst = out + split_pixel_offset(n, y, xx, H, W) + c0 * 2;
    split_store8(v, dst, dst + 256, mode);
    // halo pixels hold the reflected neighbour: x' = 0 <- x = 1, x' = W + 1 <- x = W - 2
    if (xx == 1) { char* h = out + split_pixel_offset(n, y, -1, H, W) + c0 * 2; split_store8(v, h, h + 256, mode); }
    if (xx == W - 2) { char* h = out + split_pixel_offset(n, y, W, H, W) + c0 * 2; split_store8(v, h, h + 256, mode); }
}

// ------------------------------------------------------------------------------------------
// epilogues: thread = one output channel; v[32] = 32 consecutive pixels of that channel
// ------------------------------------------------------------------------------------------
struct TEpiBiasStoreStats {    // out = acc + bias[ch]; (sum, sumsq) of out: the 3x3 convolutions of the residual blocks
    static constexpr int NS = 2;
    static constexpr int PARTS = 2;
    float* out; const float* bias; double* stats;
    struct State { float b; };
    __device__ void init(int n, int NOUT, int ch, State& st) const { st.b = bias ? bias[ch] : 0.f; }
    template <int NPX>
    __device__ void apply(const State& st, size_t row0, int NOUT, int ch, const float* v, float* s) const {
#pragma unroll
        for (int i = 0; i < NPX; ++i) {
            const float o = v[i] + st.b;
            out[(row0 + i) * NOUT + ch] = o;
            s[0] += o;
            s[1] = fmaf(o, o, s[1]);
        }
    }
    __device__ double* dst(int n, int NOUT) const { return stats + (size_t)n * NOUT * NS; }
};
struct TEpiStoreAdd {          // out = acc (+ add): input gradient of a 3x3 convolution (+ the residual branch's gradient)
    static constexpr int NS = 1;
    static constexpr int PARTS = 2;
    float* out; const float* add; double* scratch;
    struct State {};
    __device__ void init(int, int, int, State&) const {}
    template <int NPX>
    __device__ void apply(const State&, size_t row0, int NOUT, int ch, const float* v, float* s) const {
#pragma unroll
        for (int i = 0; i < NPX; ++i) {
            const size_t o = (row0 + i) * NOUT + ch;
            out[o] = add ? v[i] + add[o] : v[i];
        }
    }
    __device__ double* dst(int n, int NOUT) const { return scratch + (size_t)n * NOUT * NS; }
};
template <class OT>
struct TEpiStoreStatsT {       // raw output + (sum, sumsq)
    static constexpr int NS = 2;
    static constexpr int PARTS = 2;      // split epilogue (see gemm_tc_kernel): measured -9 % (128 -> 256) and -5 % (256 -> 128)
    OT* out; double* stats;
    struct State {};
    __device__ void init(int n, int NOUT, int ch, State&) const {}
    template <int NPX>
    __device__ void apply(const State&, size_t row0, int NOUT, int ch, const float* v, float* s) const {
#pragma unroll
        for (int i = 0; i < NPX; ++i) {
            sth(out + (row0 + i) * NOUT + ch, v[i]);
            s[0] += v[i];
            s[1] = fmaf(v[i], v[i], s[1]);
        }
    }
    __device__ double* dst(int n, int NOUT) const { return stats + (size_t)n * NOUT * NS; }
};
typedef TEpiStoreStatsT<float> TEpiStoreStats;
struct TEpiResidual {          // out = x + acc*scale3 + shift3 (+ column sums of out): eval-mode BatchNorm blocks, Norm3 known up front
    static constexpr int NS = 2;
    static constexpr int PARTS = 1;
    float* out; const float* x; const Coef* coef3; double* stats;
    struct State { Coef k; };
    __device__ void init(int n, int NOUT, int ch, State& st) const { st.k = coef3[(size_t)n * NOUT + ch]; }
    template <int NPX>
    __device__ void apply(const State& st, size_t row0, int NOUT, int ch, const float* v, float* s) const {
#pragma unroll
        for (int i = 0; i < NPX; ++i) {
            const size_t o = (row0 + i) * NOUT + ch;
            const float r = x[o] + fmaf(v[i], st.k.scale, st.k.shift);
            out[o] = r;
            s[0] += r;
            s[1] = fmaf(r, r, s[1]);
        }
    }
    __device__ double* dst(int n, int NOUT) const { return stats + (size_t)n * NOUT * NS; }
};
template <class HT>
struct TEpiGemm2BwdT {         // du = acc; sums (du*g2, du*gp2, du*gp2*h2hat)
    static constexpr int NS = 3;
    static constexpr int PARTS = 1;      // these epilogues also READ a tensor per element: splitting measured neutral to +4 %
    HT* du; const HT* h2; const Coef* coef2; const MeanRstd* mr2; double* sums;
    struct State { Coef k; MeanRstd m; };
    __device__ void init(int n, int NOUT, int ch, State& st) const { st.k = coef2[(size_t)n * NOUT + ch]; st.m = mr2[(size_t)n * NOUT + ch]; }
    template <int NPX>
    __device__ void apply(const State& st, size_t row0, int NOUT, int ch, const float* v, float* s) const {
#pragma unroll
        for (int i = 0; i < NPX; i += 2) {        // two pixels per step: packed f32x2 GELU + derivative
            const size_t o0 = (row0 + i) * NOUT + ch, o1 = o0 + NOUT;
            sth(du + o0, v[i]);
            sth(du + o1, v[i + 1]);
            const float h0 = ldh(h2 + o0), h1v = ldh(h2 + o1);
            float g0, g1, p0, p1;
            gelu_pair<true, true>(fmaf(h0, st.k.scale, st.k.shift), fmaf(h1v, st.k.scale, st.k.shift), g0, g1, p0, p1);
            s[0] = fmaf(v[i], g0, s[0]);
            s[1] = fmaf(v[i], p0, s[1]);
            s[2] = fmaf(v[i] * p0, (h0 - st.m.mean) * st.m.rstd, s[2]);
            s[0] = fmaf(v[i + 1], g1, s[0]);
            s[1] = fmaf(v[i + 1], p1, s[1]);
            s[2] = fmaf(v[i + 1] * p1, (h1v - st.m.mean) * st.m.rstd, s[2]);
        }
    }
    __device__ double* dst(int n, int NOUT) const { return sums + (size_t)n * NOUT * NS; }
};
typedef TEpiGemm2BwdT<float> TEpiGemm2Bwd;
struct TEpiGemm1Bwd {          // dn0 = acc; sums (dn0, dn0*x_hat)
    static constexpr int NS = 2;
    static constexpr int PARTS = 1;
    float* dn0; const float* x; const MeanRstd* mr0; double* bstats;
    struct State { MeanRstd m; };
    __device__ void init(int n, int NOUT, int ch, State& st) const { st.m = mr0[(size_t)n * NOUT + ch]; }
    template <int NPX>
    __device__ void apply(const State& st, size_t row0, int NOUT, int ch, const float* v, float* s) const {
        float xv[NPX];
        preload<NPX>(row0, NOUT, ch, xv);
        apply_pre<NPX>(st, row0, NOUT, ch, v, xv, s);
    }
    // the slice's loads of x, issued by the fused kernel BEFORE it waits for the accumulator and reads TMEM (their latency overlaps
    // both), and in any case before the slice's stores instead of alternating with them
    template <int NPX>
    __device__ void preload(size_t row0, int NOUT, int ch, float* xv) const {
#pragma unroll
        for (int i = 0; i < NPX; ++i) xv[i] = __ldg(x + (row0 + i) * NOUT + ch);
    }
    template <int NPX>
    __device__ void apply_pre(const State& st, size_t row0, int NOUT, int ch, const float* v, const float* xv, float* s) const {
#pragma unroll
        for (int i = 0; i < NPX; ++i) {
            dn0[(row0 + i) * NOUT + ch] = v[i];
            s[0] += v[i];
            s[1] = fmaf(v[i], (xv[i] - st.m.mean) * st.m.rstd, s[1]);
        }
    }
    __device__ double* dst(int n, int NOUT) const { return bstats + (size_t)n * NOUT * NS; }
};

// ------------------------------------------------------------------------------------------
// the kernel.  grid = (G, N frames); CTA b of a frame owns tiles [b*T/G, (b+1)*T/G) of that frame.
// The (tile, K-block) sequence is software pipelined: the raw loads of step q+1 are issued before step q is
// converted, so HBM requests stay in flight across the convert / fence / barrier / MMA issue / epilogue of step q.
// ------------------------------------------------------------------------------------------
// WSTREAM (the 3x3 convolutions of the residual blocks, K = 9 taps x 128 channels): the weight image (576 KB) does not fit in
// shared memory, so the [NOUT][64] hi and lo slabs of each K-block are streamed from the (L2-resident) global image into a
// four-slot ring by the TMA engine: thread 0 issues two cp.async.bulk (16 KB each, mbarrier complete_tx) two pipeline steps ahead
// and waits for the slab's barrier right before it issues the step's MMAs -- no registers, no shared-memory stores by the threads
// (the first version copied the slabs through registers, 4 x LDG.128 + 4 x STS.128 per thread and step; same speed, 40 more registers).
template <int K, int NOUT, class ALoad, class Epi, int PARTS, bool WSTREAM>
__device__ __forceinline__ void gemm_tc_body(const ALoad& al, const uint4* __restrict__ wimg, const Epi& ep, int P, int single,
                                             const int bx /* CTA index within the frame */, const int gx /* CTAs per frame */, const int n) {
    constexpr int KB = K / KBLK, MH = NOUT / 128;
    constexpr int W_HALF = K * NOUT * 2;                  // bytes of the hi image (the lo image follows it)
    constexpr int SLAB = NOUT * 128;                      // one K-block of one image: [NOUT rows][128 B]
    constexpr int WRING = 4, WAHEAD = 2;                  // streamed weights: ring slots, prefetch distance in pipeline steps
    constexpr int W_BYTES = WSTREAM ? WRING * 2 * SLAB : K * NOUT * 4;    // resident: hi image + lo image; streamed: ring of slab pairs
    constexpr int CFK = WSTREAM ? 128 : K;                // coefficient entries per vector
    constexpr int ACC_COLS = MH * TILE_PX;                // TMEM columns per accumulator stage
    extern __shared__ __align__(1024) char smem_raw[];
    char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B atoms need 1024-byte alignment
    char* sW = smem;                                      // [hi|lo][KB][NOUT rows][128 B], or NSTAGE x {hi slab, lo slab}
    char* sA = smem + W_BYTES;                            // NSTAGE x {hi tile 16 KB, lo tile 16 KB}
    float* sCf = reinterpret_cast<float*>(sA + NSTAGE * STAGE_BYTES);     // 3 x CFK coefficients (SoA)
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sCf + 3 * CFK);          // free[NSTAGE], accfull[2], wfull[WRING]
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + NSTAGE + 2 + WRING);
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;

    const int tiles_per_frame = P / TILE_PX;
    const int t0 = (int)(((long long)bx * tiles_per_frame) / gx);
    const int t1 = (int)(((long long)(bx + 1) * tiles_per_frame) / gx);
    const int Q = (t1 - t0) * KB;                         // pipeline steps of this CTA
    // producer role: 16-byte chunk pc8 of rows pr and pr + 64
    const int pc8 = tid % 8, pr = tid / 8;
    typename ALoad::Raw raw[2];
    typename PosOf<ALoad>::type pos[2];                   // the two operand rows of this thread in the tile being prefetched
    if (Q > 0) {
        const size_t row0 = (size_t)n * P + (size_t)t0 * TILE_PX;
        if constexpr (IsPresplit<ALoad>::v) {
            locate_row(al, row0 + pr, pos[0]);
            locate_row(al, row0 + pr + 64, pos[1]);
            issue_row(al, pos[0], K, pc8 * 8, raw[0]);
            issue_row(al, pos[1], K, pc8 * 8, raw[1]);
        } else if constexpr (HasIssueOff<ALoad>::v) {
            al.issue_off((row0 + pr) * K + pc8 * 8, raw[0]);
            al.issue_off((row0 + pr + 64) * K + pc8 * 8, raw[1]);
        } else {
            al.issue(row0 + pr, K, pc8 * 8, raw[0]);
            al.issue(row0 + pr + 64, K, pc8 * 8, raw[1]);
        }
    }
    // element offset of this thread's next prefetch (step 1), advanced by constants: + one K-block, or to block 0 of the next tile
    size_t poff = ((size_t)n * P + (size_t)t0 * TILE_PX + pr) * K + pc8 * 8 + KBLK;

    // ---- one-time setup: weights, coefficients, barriers, TMEM ----
    // (thread 0 only) slab pair of pipeline step qq -> weight ring slot qq % WRING.  The slot was last read by the MMAs of step
    // qq - WRING, which are complete whenever this is called (see the call sites).
    auto w_issue = [&](int qq) {
        const int kb = qq % KB;
        const uint32_t bar = smem_u32(&sBar[NSTAGE + 2 + qq % WRING]), dst = smem_u32(sW) + (uint32_t)(qq % WRING) * 2 * SLAB;
        mbar_expect_tx(bar, 2 * SLAB);
        bulk_g2s(dst, reinterpret_cast<const char*>(wimg) + (size_t)kb * SLAB, SLAB, bar);
        bulk_g2s(dst + SLAB, reinterpret_cast<const char*>(wimg) + (size_t)W_HALF + (size_t)kb * SLAB, SLAB, bar);
    };
    al.fill(n, K, sCf);
    if (tid == 0) {
        for (int i = 0; i < NSTAGE + 2 + WRING; ++i) mbar_init(smem_u32(&sBar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if constexpr (WSTREAM) {
            for (int a = 0; a < WAHEAD && a < Q; ++a) w_issue(a);
        } else if (Q > 0) {
            // resident weight image (128 KB): ONE thread hands it to the TMA engine (eight 16 KB bulk copies onto the first weight
            // barrier) instead of 16 LDG.128 + 16 STS.128 per thread; it lands while the CTA sets up and the first operand loads are in
            // flight, and only the MMA issuer waits for it.  With 4 ... 12 CTAs per SM and launch the prologue is paid that often.
            const uint32_t bar = smem_u32(&sBar[NSTAGE + 2]);
            mbar_expect_tx(bar, W_BYTES);
#pragma unroll
            for (int c = 0; c < W_BYTES / 16384; ++c)
                bulk_g2s(smem_u32(sW) + c * 16384, reinterpret_cast<const char*>(wimg) + (size_t)c * 16384, 16384, bar);
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sTmem;

    // epilogue role of this thread: TMEM lane quarter = warp % 4, pixel block = warp / 4, M-blocks j = 0..MH-1
    const int lq = warp % 4, pc = warp / 4;
    typename Epi::State est[MH];
    float stat[MH][Epi::NS];
#pragma unroll
    for (int j = 0; j < MH; ++j) {
        ep.init(n, NOUT, j * 128 + lq * 32 + lane, est[j]);
#pragma unroll
        for (int s = 0; s < Epi::NS; ++s) stat[j][s] = 0.f;
    }

    auto epilogue = [&](int pit) {        // local tile index pit: TMEM -> registers -> global (+ statistics)
        const size_t prow0 = (size_t)n * P + (size_t)(t0 + pit) * TILE_PX + pc * 32;
        mbar_wait(smem_u32(&sBar[NSTAGE + (pit & 1)]), (uint32_t)(pit >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < MH; ++j) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(pit & 1) * ACC_COLS + j * TILE_PX + pc * 32, v);
            ep.template apply<32>(est[j], prow0, NOUT, j * 128 + lq * 32 + lane, v, stat[j]);
        }
        tc_fence_before();     // order the TMEM reads before the barrier that precedes the next overwrite of this stage
    };
    // PARTS > 1: the epilogue of tile t-1 is spread over the pipeline steps of tile t (32/PARTS pixels after every
    // KB/PARTS-th K-block) instead of running as one burst of 64+ global stores / loads per thread after the last K-block:
    // the operand prefetch of the next step is never more than a partial epilogue away, and the LSU queue does not fill
    // up (ncu r01: lg_throttle 26 % in the write-heavy 128 -> 256 kernel, which gains 9 %; the 256 -> 128 kernels do not).
    static_assert(PARTS == 1 || (KB % PARTS == 0 && 32 % PARTS == 0), "PARTS must divide the K-block count");
    auto epilogue_part = [&](int pit, int part) {
        constexpr int CH = 32 / PARTS;
        const size_t prow0 = (size_t)n * P + (size_t)(t0 + pit) * TILE_PX + pc * 32 + part * CH;
        if (part == 0) mbar_wait(smem_u32(&sBar[NSTAGE + (pit & 1)]), (uint32_t)(pit >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < MH; ++j) {
            float v[CH];
            tmem_ldn<CH>(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(pit & 1) * ACC_COLS + j * TILE_PX + pc * 32 + part * CH, v);
            ep.template apply<CH>(est[j], prow0, NOUT, j * 128 + lq * 32 + lane, v, stat[j]);
        }
        tc_fence_before();
    };

    // (Alternating two raw register sets statically, as wgrad_tc_body does, was measured here too: it removes the 16 ... 32 register
    // moves of `cur = raw` per step, but the twice-unrolled step body ran slower -- gemm2_fwd 3.41 -> 3.65, gemm1_fwd 2.36 -> 2.77,
    // gemm2_bwd 4.90 -> 5.92 ms per step with 108 bytes of spills at its 128-register cap.)
    for (int q = 0; q < Q; ++q) {
        const int it = q / KB, kb = q % KB;               // local tile, K-block
        typename ALoad::Raw cur[2] = {raw[0], raw[1]};
        if (q + 1 < Q) {                                  // prefetch the next step's operands
            const int nit = (q + 1) / KB, nkb = (q + 1) % KB;
            const size_t nrow0 = (size_t)n * P + (size_t)(t0 + nit) * TILE_PX;
            if constexpr (IsPresplit<ALoad>::v) {
                if (nkb == 0) {                           // the prefetch enters the next tile
                    locate_row(al, nrow0 + pr, pos[0]);
                    locate_row(al, nrow0 + pr + 64, pos[1]);
                }
                issue_row(al, pos[0], K, nkb * KBLK + pc8 * 8, raw[0]);
                issue_row(al, pos[1], K, nkb * KBLK + pc8 * 8, raw[1]);
            } else if constexpr (HasIssueOff<ALoad>::v) {
                al.issue_off(poff, raw[0]);
                if constexpr (!SpreadRows<ALoad>::v) {        // else: the second row's loads follow the first row's convert (below)
                    al.issue_off(poff + (size_t)64 * K, raw[1]);
                    poff += nkb == KB - 1 ? (size_t)TILE_PX * K - (KB - 1) * KBLK : KBLK;
                }
            } else {
                al.issue(nrow0 + pr, K, nkb * KBLK + pc8 * 8, raw[0]);
                al.issue(nrow0 + pr + 64, K, nkb * KBLK + pc8 * 8, raw[1]);
            }
        }
        const uint32_t slot = (uint32_t)q % NSTAGE, u = (uint32_t)q / NSTAGE;
        mbar_wait(smem_u32(&sBar[slot]), (u & 1) ^ 1);    // MMAs that read this ring slot are done
        char* hi = sA + slot * STAGE_BYTES;
        char* lo = hi + STAGE_BYTES / 2;
        if constexpr (WSTREAM) {
            // the wait above proves the MMAs of step q - NSTAGE (and all earlier ones) complete: weight slot (q + WAHEAD) % WRING,
            // last read by step q + WAHEAD - WRING = q - NSTAGE, is free
            static_assert(WRING - WAHEAD == NSTAGE, "weight ring reuse is tied to the operand ring's barrier");
            if (tid == 0 && q + WAHEAD < Q) w_issue(q + WAHEAD);
        }
        {
            typename ALoad::Cf cfr;                        // this thread's 8 channels of the K-block: one coefficient read for both rows
            al.coefs(K, kb * KBLK + pc8 * 8, sCf, cfr);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int r = pr + 64 * j;
                const int off = r * 128 + ((pc8 ^ (r & 7)) << 4);
                convert_store(al, cur[j], cfr, hi + off, lo + off, single);
                if constexpr (!IsPresplit<ALoad>::v && HasIssueOff<ALoad>::v && SpreadRows<ALoad>::v) {
                    // the step's prefetch is spread over it: the LDG.128 issued back to back filled the LSU queue (lg_throttle) and
                    // held up the converts behind them (the same move in wgrad_tc_kernel: 4.34 -> 4.20 ms per step)
                    if (j == 0 && q + 1 < Q) {
                        al.issue_off(poff + (size_t)64 * K, raw[1]);
                        poff += (q + 1) % KB == KB - 1 ? (size_t)TILE_PX * K - (KB - 1) * KBLK : KBLK;
                    }
                }
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t acc = (uint32_t)(it & 1) * ACC_COLS;
            if constexpr (WSTREAM) mbar_wait(smem_u32(&sBar[NSTAGE + 2 + q % WRING]), (uint32_t)(q / WRING) & 1);   // slab pair landed
            else if (q == 0) mbar_wait(smem_u32(&sBar[NSTAGE + 2]), 0);                                              // weight image landed
            const uint32_t a_hi = WSTREAM ? smem_u32(sW) + (uint32_t)(q % WRING) * 2 * SLAB : smem_u32(sW) + kb * SLAB;
            const uint32_t a_lo = a_hi + (WSTREAM ? SLAB : W_HALF);
            const uint32_t b_hi = smem_u32(hi), b_lo = smem_u32(lo);
            // fp16 operands: A/B format fields (bits 7-9, 10-12) = 0 (F16) instead of 1 (BF16)
            const uint32_t idesc = single == SPLIT_F16X3 ? (c_idesc & ~((7u << 7) | (7u << 10))) : c_idesc;
#pragma unroll
            for (int j = 0; j < MH; ++j) {
                const uint32_t d = tmem_base + acc + j * TILE_PX;
#pragma unroll
                for (int k16 = 0; k16 < KBLK / 16; ++k16) {
                    const uint64_t wa = make_desc(a_hi + j * (128 * 128) + k16 * 32);
                    const uint64_t wl = make_desc(a_lo + j * (128 * 128) + k16 * 32);
                    const uint64_t xa = make_desc(b_hi + k16 * 32);
                    const uint64_t xl = make_desc(b_lo + k16 * 32);
                    tc_mma(d, wa, xa, idesc, (kb | k16) != 0);
                    if (single != SPLIT_BF16X1) {
                        tc_mma(d, wa, xl, idesc, 1);
                        tc_mma(d, wl, xa, idesc, 1);
                    }
                }
            }
            tc_commit(smem_u32(&sBar[slot]));                                  // frees the ring slot
            if (kb == KB - 1) tc_commit(smem_u32(&sBar[NSTAGE + (it & 1)]));   // accumulator of this tile complete
        }
        if constexpr (PARTS > 1) {
            constexpr int EVERY = KB / PARTS;
            if (it > 0 && (kb + 1) % EVERY == 0) epilogue_part(it - 1, (kb + 1) / EVERY - 1);   // a slice of the previous tile, while this tile's MMAs run
        } else {
            if (kb == KB - 1 && it > 0) epilogue(it - 1);     // previous tile, while this tile's MMAs run
        }
    }
    if (Q > 0) {
        if constexpr (PARTS > 1) {
#pragma unroll
            for (int part = 0; part < PARTS; ++part) epilogue_part(t1 - t0 - 1, part);
        } else {
            epilogue(t1 - t0 - 1);
        }
    }
    // per-channel statistics: 4 warps (pixel blocks) share a channel -> 4 atomics per channel per CTA
    double* dst = ep.dst(n, NOUT);
#pragma unroll
    for (int j = 0; j < MH; ++j)
#pragma unroll
        for (int s = 0; s < Epi::NS; ++s)
            atomicAdd(&dst[(size_t)(j * 128 + lq * 32 + lane) * Epi::NS + s], (double)stat[j][s]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

template <int K, int NOUT, class ALoad, class Epi, int PARTS, bool WSTREAM>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(ALoad al, const uint4* __restrict__ wimg /* prepared bf16 hi/lo image, K*NOUT*4 bytes */, Epi ep, int P, int single) {
    gemm_tc_body<K, NOUT, ALoad, Epi, PARTS, WSTREAM>(al, wimg, ep, P, single, (int)blockIdx.x, (int)gridDim.x, (int)blockIdx.y);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
static int gcd_int(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }
static int sm_count() { return device_sm_count(); }
// CTAs per frame: G*N a multiple of the SM count (equal waves), about 8-16 tiles per CTA
static int blocks_per_frame(int N, int tiles) {
    const int sms = sm_count();
    const int unit = sms / gcd_int(sms, N);
    int g = unit;
    while (g * 2 <= tiles / 8) g *= 2;
    if (g > tiles) g = tiles;
    return g < 1 ? 1 : g;
}

template <int K, int NOUT, class ALoad, class Epi, bool WSTREAM = false>
static int launch(ALoad al, const void* wimg, Epi ep, int N, int P, int single, cudaStream_t st) {
    if (P % TILE_PX != 0) return UB_ERR_ARG;
    constexpr size_t wbytes = WSTREAM ? (size_t)4 * 2 * NOUT * 128 : (size_t)K * NOUT * 4;       // streamed: four-slot ring of slab pairs
    constexpr size_t smem = wbytes + NSTAGE * STAGE_BYTES + 3 * (WSTREAM ? 128 : K) * sizeof(float) + 8 * 8 + 16 + 1024;
    auto kern = gemm_tc_kernel<K, NOUT, ALoad, Epi, Epi::PARTS, WSTREAM>;
    UB_SET_SMEM(kern, smem);
    const int tiles = P / TILE_PX;
    const dim3 grid(blocks_per_frame(N, tiles), N);
    kern<<<grid, THREADS, smem, st>>>(al, static_cast<const uint4*>(wimg), ep, P, single);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

// ------------------------------------------------------------------------------------------
// weight gradient on tensor cores:  dW[m][n] = sum_px fa(px)[m] * fb(px)[n],  m < 128, n < 256
//
// The contraction runs over pixels, i.e. over the ROWS of the pixel-major tiles, so both operands are
// "MN-major" for the MMA.  The canonical MN-major SWIZZLE_128B layout ((8,n),(8,k)):((1,LBO),(8,SBO)) [uint128
// units, cute/atom/mma_traits_sm100.hpp] is byte-for-byte the layout the producers already write for K-major
// tiles: 128-byte rows (one pixel, 64 channels), 8-row 1024-byte atoms with the 16-byte chunk index XOR row%8.
// Column blocks of 64 channels are LBO apart, 8-pixel groups SBO = 1024 B apart; one MMA consumes 16 pixels.
// One persistent CTA per SM walks a contiguous range of 64-pixel tiles (2-stage ring), accumulating the whole
// 128 x 256 gradient in TMEM (256 columns); partials are summed by reduce_partials_kernel.
// ------------------------------------------------------------------------------------------
constexpr int WG_PX = 64;
constexpr int WG_A_BYTES = 2 * WG_PX * 128 * 2;      // hi+lo, 128 channels: 32 KB
constexpr int WG_B_BYTES = 2 * WG_PX * 256 * 2;      // hi+lo, 256 channels: 64 KB
constexpr int WG_STAGE = WG_A_BYTES + WG_B_BYTES;    // 96 KB
constexpr int WG_BLK = WG_PX * 128;                  // one [64 px][64 ch] bf16 block: 8 KB

constexpr uint32_t c_wg_desc_hi = (64u) | (1u << 14) | (2u << 29);                  // SBO = 1024 B
constexpr uint32_t c_wg_desc_lbo = (uint32_t)(WG_BLK >> 4);                           // LBO = 8 KB between 64-channel blocks
constexpr uint32_t c_wg_idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (32u << 17) | (8u << 24);

__device__ __forceinline__ uint64_t make_wg_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(c_wg_desc_lbo & 0x3FFFu) << 16) | ((uint64_t)c_wg_desc_hi << 32);
}

template <class LA, class LB>
__device__ __forceinline__ void wgrad_tc_body(const LA& la, const LB& lb, float* __restrict__ dst /* this CTA's [128][256] partial */, int P,
                                              const long long t0, const long long t1 /* 64-pixel tiles [t0, t1) of the whole tensor */,
                                              int sa, int sb, int single) {
    extern __shared__ __align__(1024) char smem_raw[];
    char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    char* sStage = smem;                                                       // 2 x {A hi, A lo, B hi, B lo}
    float* sCfA = reinterpret_cast<float*>(smem + 2 * WG_STAGE);              // 3 x 128 coefficients (SoA)
    float* sCfB = sCfA + 3 * 128;                                              // 3 x 256 coefficients
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sCfB + 3 * 256);             // free[2], done
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 3);
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;

    const int tiles_per_frame = P / WG_PX;
    // 32-bit step counter and an incrementally tracked frame index: `frame = tile / tiles_per_frame` on 64-bit indices was an emulated
    // division in every half step (see bwd_tc_kernel)
    const int H = (int)(t1 > t0 ? (t1 - t0) : 0) * 2;      // pipeline steps: half tiles of 32 pixel rows
    const int frame0 = (int)(t0 / tiles_per_frame), first_boundary = 2 * (int)((long long)(frame0 + 1) * tiles_per_frame - t0);
    int ld_next = frame0, ld_boundary = 0;                 // the frame changes at half-step index ld_boundary
    // producer role per half tile: one A item (row ra, chunk ca) and two B items (rows rb, rb+16; chunk cb)
    const int ra = tid / 16, ca = tid % 16, rb = tid / 32, cb = tid % 32;
    // Two raw register sets, alternating STATICALLY (the half-step loop is unrolled by two): while one set is converted the other receives
    // the next half tile's loads.  One set plus a per-step copy cost 33 register moves per thread and tile across the loop back-edge
    // (7 % of the instruction stream in the ncu source view); measured 5.06 -> 4.46 ms per step.
    typename LA::Raw rawa, rawa2;
    typename LB::Raw rawb[2], rawb2[2];
    constexpr bool LINEAR = HasIssueOff<LA>::v && HasIssueOff<LB>::v;
    size_t aoff = ((size_t)t0 * WG_PX + ra) * 128 + ca * 8, boff = ((size_t)t0 * WG_PX + rb) * 256 + cb * 8;      // next prefetch (LINEAR)
    if (H > 0) {
        const size_t row0 = (size_t)t0 * WG_PX;
        if constexpr (LINEAR) {
            la.issue_off(aoff, rawa);
            lb.issue_off(boff, rawb[0]);
            lb.issue_off(boff + 16 * 256, rawb[1]);
            aoff += 32 * 128;                           // every half step is the next 32 pixel rows
            boff += 32 * 256;
        } else {
            la.issue(row0 + ra, 128, ca * 8, rawa);
            lb.issue(row0 + rb, 256, cb * 8, rawb[0]);
            lb.issue(row0 + rb + 16, 256, cb * 8, rawb[1]);
        }
    }

    if (tid == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&sBar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sTmem;

    auto hstep = [&](const int h, const typename LA::Raw& ca_raw, const typename LB::Raw (&cb_raw)[2], typename LA::Raw& na_raw,
                     typename LB::Raw (&nb_raw)[2]) {
        const int half = h & 1;
        const uint32_t use = (uint32_t)(h >> 1);
        if (h == ld_boundary) {                 // block-uniform: refill the per-frame coefficients
            const int n = ld_next;
            ld_next = n + 1;
            ld_boundary = n == frame0 ? first_boundary : ld_boundary + 2 * tiles_per_frame;
            __syncthreads();
            la.fill(n, 128, sCfA);
            lb.fill(n, 256, sCfB);
            __syncthreads();
        }
        if (h + 1 < H) {                        // prefetch the next half tile into the other register set
            if constexpr (LINEAR) {
                // the eight LDG.128 of a half step are spread over it (A here, B[0] after the A convert, B[1] after the first B convert):
                // issued back to back they filled the LSU queue (lg_throttle 20 % of the stalls) and held up the converts behind them
                la.issue_off(aoff, na_raw);
                aoff += 32 * 128;
            } else {
                const size_t nrow0 = (size_t)(t0 + ((h + 1) >> 1)) * WG_PX + (size_t)((h + 1) & 1) * 32;
                la.issue(nrow0 + ra, 128, ca * 8, na_raw);
                lb.issue(nrow0 + rb, 256, cb * 8, nb_raw[0]);
                lb.issue(nrow0 + rb + 16, 256, cb * 8, nb_raw[1]);
            }
        }
        const uint32_t slot = use & 1, u = use >> 1;
        if (half == 0) mbar_wait(smem_u32(&sBar[slot]), (u & 1) ^ 1);
        char* a_hi = sStage + slot * WG_STAGE;
        char* a_lo = a_hi + WG_A_BYTES / 2;
        char* b_hi = a_hi + WG_A_BYTES;
        char* b_lo = b_hi + WG_B_BYTES / 2;
        {
            const int r = half * 32 + ra;
            typename LA::Cf cfa;
            la.coefs(128, ca * 8, sCfA, cfa);
            const int off = (ca / 8) * WG_BLK + r * 128 + (((ca % 8) ^ (r & 7)) << 4);
            convert_store(la, ca_raw, cfa, a_hi + off, a_lo + off, single);
        }
        if constexpr (LINEAR) {
            if (h + 1 < H) lb.issue_off(boff, nb_raw[0]);
        }
        typename LB::Cf cfb;                              // both B rows of this thread share the channel chunk
        lb.coefs(256, cb * 8, sCfB, cfb);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int r = half * 32 + rb + 16 * j;
            const int off = (cb / 8) * WG_BLK + r * 128 + (((cb % 8) ^ (r & 7)) << 4);
            convert_store(lb, cb_raw[j], cfb, b_hi + off, b_lo + off, single);
            if constexpr (LINEAR) {
                if (j == 0 && h + 1 < H) { lb.issue_off(boff + 16 * 256, nb_raw[1]); boff += 32 * 256; }
            }
        }
        if (half == 1) {
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t idesc = c_wg_idesc;
#pragma unroll
                for (int p16 = 0; p16 < WG_PX / 16; ++p16) {
                    const uint64_t ah = make_wg_desc(smem_u32(a_hi) + p16 * 2048), al = make_wg_desc(smem_u32(a_lo) + p16 * 2048);
                    const uint64_t bh = make_wg_desc(smem_u32(b_hi) + p16 * 2048), bl = make_wg_desc(smem_u32(b_lo) + p16 * 2048);
                    tc_mma(tmem_base, ah, bh, idesc, (use | (uint32_t)p16) != 0);
                    if (!single) {
                        tc_mma(tmem_base, ah, bl, idesc, 1);
                        tc_mma(tmem_base, al, bh, idesc, 1);
                    }
                }
                tc_commit(smem_u32(&sBar[slot]));
                if (h == H - 1) tc_commit(smem_u32(&sBar[2]));
            }
        }
    };
    for (int h = 0; h < H; h += 2) {                // H = 2 x tiles: half 0 consumes set 1, half 1 set 2
        hstep(h, rawa, rawb, rawa2, rawb2);
        hstep(h + 1, rawa2, rawb2, rawa, rawb);
    }
    if (H > 0) {
        mbar_wait(smem_u32(&sBar[2]), 0);
        tc_fence_after();
        const int lq = warp % 4, cbk = warp / 4;          // lanes lq*32.., columns cbk*64..
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + cbk * 64 + hh * 32, v);
            const int m = lq * 32 + lane;
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[(size_t)m * sa + (size_t)(cbk * 64 + hh * 32 + i) * sb] = v[i];
        }
    } else {
        for (int i = tid; i < 128 * 256; i += THREADS) dst[i] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
}

template <class LA, class LB>
__global__ void __launch_bounds__(THREADS, 1)
wgrad_tc_kernel(LA la, LB lb, float* __restrict__ partial, int P, long long total_tiles, int sa, int sb, int single) {
    const long long per = (total_tiles + gridDim.x - 1) / gridDim.x;
    const long long t0 = (long long)blockIdx.x * per, t1 = min(t0 + per, total_tiles);
    wgrad_tc_body<LA, LB>(la, lb, partial + (size_t)blockIdx.x * 128 * 256, P, t0, t1, sa, sb, single);
}

template <class LA, class LB>
static int launch_wgrad_tc(LA la, LB lb, float* partial, int max_parts, int N, int P, int sa, int sb, int single, int* nparts,
                           cudaStream_t st) {
    if (P % WG_PX != 0) return UB_ERR_ARG;
    constexpr size_t smem = (size_t)2 * WG_STAGE + 3 * 384 * sizeof(float) + 3 * 8 + 16 + 1024;
    auto kern = wgrad_tc_kernel<LA, LB>;
    UB_SET_SMEM(kern, smem);
    const long long total = (long long)N * (P / WG_PX);
    const int blocks = (int)(total < max_parts ? total : max_parts);
    kern<<<blocks, THREADS, smem, st>>>(la, lb, partial, P, total, sa, sb, single);
    UB_CHECK_LAUNCH();
    *nparts = blocks;
    return UB_OK;
}

// ------------------------------------------------------------------------------------------
// Fused input-gradient + weight-gradient GEMM of the 1x1 EXPAND convolution (the backward of uncrtaints.py:126).
//
// The two GEMMs of a convolution's backward read the same activation tensors: launched separately they stream those tensors from
// HBM twice (for the expand convolution dz1, h1: 2 Hh, and x: A per frame).  Here ONE persistent kernel builds every operand tile
// once and feeds it to both GEMMs:
//     S = dh1 = normbwd1(dz1, h1) (256 channels, streamed in four 64-channel blocks),   R = n0 = x*scale0 + shift0 (128 channels)
//     dn0[128 ch][px] += W1^T[:, blk] . S_blk   (K-major operands, per block)           dW1^T[k][o in blk] += R^T . S_blk   (MN-major)
// The bytes of a [64 px][64 ch] SWIZZLE_128B block are at the same time a K-major tile (pixels = rows) and an MN-major tile
// (pixels = the contraction), so the same shared-memory block serves both MMAs.
//
// Shared memory cannot hold the 128 KB weight image, a 128-pixel operand ring AND a second 128-pixel operand, so the tile is
// 64 pixels (UMMA M=128, N=64, K=16): 128 KB weights + 32 KB resident operand R + 3 x 16 KB ring of S blocks + coefficients.
// TMEM: double-buffered input-gradient accumulator (2 x 64 columns) + the whole 128 x 256 weight gradient (256 columns),
// accumulated over all tiles of the CTA and written once as a partial (reduce_partials_kernel sums the <= 148 partials).
// One persistent CTA per SM walks a contiguous range of tiles across frames; per-frame coefficient tables are refilled at frame
// boundaries, the per-channel statistics are flushed by the epilogue when ITS tile (one behind) changes frame.
// Measured (B=16, ms per step): gemm1_bwd 5.4 + wgrad1 4.7 = 10.1 separate -> 7.6 fused; ncu r02: 204.6 MB of DRAM traffic per frame
// against 201.3 MB algorithmic, tensor pipe 11 % busy, stalls long_scoreboard 29 % / short_scoreboard 15 % / wait 12 % / barrier 11 %.
// Tried and rejected (measured): the same fusion for the PROJECT convolution (11.0 ms against 5.0 + 5.0: its GELU-heavy loader and
// epilogue do not shrink with the bytes, and 64-pixel tiles double its per-step overhead); N=128 weight-gradient batches with the
// coefficient tables read from global memory (2x slower: the tables thrash the ~26 KB of L1 left beside 218 KB of shared memory);
// prefetching the epilogue's x values one step ahead (+9 %: the extra loads compete with the operand prefetch).
// ------------------------------------------------------------------------------------------
constexpr int FPX = 64;                       // pixels per tile
constexpr int FBLK = FPX * 128;               // one [64 px][64 ch] bf16 block: 8 KB
constexpr int FRING = 3;                      // ring stages (hi block + lo block = 16 KB each)
constexpr int F_W_BYTES = UB_WIDTH * UB_HID * 4;      // hi + lo weight image: 128 KB
constexpr uint32_t F_IDESC_K = (1u << 4) | (1u << 7) | (1u << 10) | (8u << 17) | (8u << 24);      // F32 acc, BF16 x BF16, K-major A/B, N=64, M=128
constexpr uint32_t F_IDESC_MN = F_IDESC_K | (1u << 15) | (1u << 16);                               // both operands MN-major

__device__ __forceinline__ void mbar_wait_guard(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    const long long t0 = clock64();
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 4000000000ll) __trap();      // ~2 s: a protocol error must fault, not hang the GPU
    } while (!ok);
}

template <class LS, class LR, class Epi>
__global__ void __launch_bounds__(THREADS, 1)
bwd_tc_kernel(LS ls, LR lr, const uint4* __restrict__ wimg, Epi ep, float* __restrict__ partial, int P, long long total_tiles,
              int sa, int sb, int single) {
    constexpr int NOUT = UB_WIDTH;                            // channels of the input gradient (TMEM lanes x MH)
    constexpr int MH = NOUT / 128;
    constexpr int ACC = MH * FPX;                             // TMEM columns of one input-gradient stage
    constexpr int DW_COL = 2 * ACC;                           // weight-gradient accumulator: 256 columns behind the two stages
    constexpr int W_HALF = F_W_BYTES / 2;
    constexpr int PARTS = 2, CH = 16 / PARTS;                 // the epilogue of tile t-1 runs in two slices during tile t
    extern __shared__ __align__(1024) char smem_raw[];
    char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    char* sW = smem;                                          // weight image [hi|lo][K/64][NOUT rows][128 B]
    char* sR = sW + F_W_BYTES;                                // resident operand: hi {blk0, blk1}, lo {blk0, blk1}: 32 KB
    char* sS = sR + 4 * FBLK;                                 // ring: FRING x {hi block, lo block}
    float* sCfS = reinterpret_cast<float*>(sS + FRING * 2 * FBLK);      // 3 x 256 coefficients of S (SoA, quad-split)
    float* sCfR = sCfS + 3 * UB_HID;                                     // 3 x 128 coefficients of R
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sCfR + 3 * UB_WIDTH);   // ringfree[FRING], accfull[2], rfree, done
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + FRING + 4);
    const uint32_t bRing = smem_u32(&sBar[0]), bAcc = smem_u32(&sBar[FRING]), bRfree = smem_u32(&sBar[FRING + 2]),
                   bDone = smem_u32(&sBar[FRING + 3]);
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;

    // Tile indices are 32-bit and the frame of a tile is tracked INCREMENTALLY (one counter pair for the loader, one for the epilogue, which
    // runs a tile behind): `frame = tile / tiles_per_frame` on 64-bit indices was an emulated division (~80 instructions, I2F + loop)
    // three times per tile -- a fifth of this kernel's instruction stream in the ncu source view.
    const int per = (int)((total_tiles + gridDim.x - 1) / gridDim.x);
    const int t0 = (int)blockIdx.x * per, t1 = (int)min((long long)t0 + per, total_tiles);
    const int ntiles = t1 > t0 ? t1 - t0 : 0;
    const int tiles_per_frame = P / FPX;
    const int Q = ntiles * 4;                                 // pipeline steps: (tile, 64-channel block of S)
    const int frame0 = t0 / tiles_per_frame, first_boundary = (frame0 + 1) * tiles_per_frame - t0;      // local index of the first tile of frame0 + 1
    float* dst = partial + (size_t)blockIdx.x * UB_WIDTH * UB_HID;
    if (Q == 0) {
        for (int i = tid; i < UB_WIDTH * UB_HID; i += THREADS) dst[i] = 0.f;
        return;
    }
    // producer roles: S item = 8 channels (chunk sc) of row sr of the current block; R items = chunk rc of rows rr, rr + 32.
    // Raw operand loads are issued TWO pipeline steps ahead (two S register sets, alternating; at the END of a step, see below): a
    // 64-pixel step moves only 32 KB per SM, and one step of work does not cover the loaded HBM latency (measured: 0.51 of the HBM
    // roof with one step ahead -- Little's law wants ~64 KB in flight per SM).
    const int sr = tid / 8, sc = tid % 8, rr = tid / 16, rc = tid % 16;
    typename LS::Raw raws[2];
    typename LR::Raw rawr[2];
    // element offsets of this thread's NEXT prefetch, advanced by constants (the addresses used to be rebuilt from the tile index in
    // every step: ~30 integer instructions of 64-bit arithmetic per thread and step)
    size_t soff = ((size_t)t0 * FPX + sr) * UB_HID + sc * 8, roff = ((size_t)t0 * FPX + rr) * UB_WIDTH + rc * 8;
    {
        ls.issue_off(soff, raws[0]);
        ls.issue_off(soff + 64, raws[1]);
        lr.issue_off(roff, rawr[0]);
        lr.issue_off(roff + 32 * UB_WIDTH, rawr[1]);
        soff += 128;                                          // step 2: block 2 of the first tile
        roff += (size_t)FPX * UB_WIDTH;                       // next tile
    }
    // ---- one-time setup: weights, barriers, TMEM ----
    for (int i = tid; i < F_W_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(sW)[i] = wimg[i];
    if (tid == 0) {
        for (int i = 0; i < FRING + 4; ++i) mbar_init(smem_u32(&sBar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sTmem;

    // ---- epilogue state (thread = one output channel per M-block; 16 pixels of a tile, in PARTS slices) ----
    const int lq = warp % 4, pc = warp / 4;
    typename Epi::State est[MH];
    float stat[MH][Epi::NS];
    int epi_n = -1, epi_next = frame0, epi_boundary = 0;      // frame of the next epilogue tile changes at local tile index epi_boundary
    auto flush_stats = [&]() {
        if (epi_n < 0) return;
        double* d = ep.dst(epi_n, NOUT);
#pragma unroll
        for (int j = 0; j < MH; ++j)
#pragma unroll
            for (int s = 0; s < Epi::NS; ++s) atomicAdd(&d[(size_t)(j * 128 + lq * 32 + lane) * Epi::NS + s], (double)stat[j][s]);
    };
    auto epilogue_part = [&](int e, int part) {              // local tile e, slice `part`: TMEM -> registers -> global (+ statistics)
        const int t = t0 + e;
        if (part == 0 && e == epi_boundary) {                 // the epilogue's tile entered a new frame (thread-local bookkeeping)
            const int n = epi_next;
            epi_next = n + 1;
            epi_boundary = n == frame0 ? first_boundary : epi_boundary + tiles_per_frame;
            flush_stats();
            epi_n = n;
#pragma unroll
            for (int j = 0; j < MH; ++j) {
                ep.init(n, NOUT, j * 128 + lq * 32 + lane, est[j]);
#pragma unroll
                for (int s = 0; s < Epi::NS; ++s) stat[j][s] = 0.f;
            }
        }
        const size_t prow0 = (size_t)t * FPX + pc * 16 + part * CH;
        float xv[MH][CH];
#pragma unroll
        for (int j = 0; j < MH; ++j) ep.template preload<CH>(prow0, NOUT, j * 128 + lq * 32 + lane, xv[j]);
        if (part == 0) mbar_wait_guard(bAcc + (e & 1) * 8, (uint32_t)(e >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < MH; ++j) {
            float v[CH];
            tmem_ldn<CH>(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(e & 1) * ACC + j * FPX + pc * 16 + part * CH, v);
            ep.template apply_pre<CH>(est[j], prow0, NOUT, j * 128 + lq * 32 + lane, v, xv[j], stat[j]);
        }
        tc_fence_before();
    };

    int ld_next = frame0, ld_boundary = 0;                    // frame of the loader's tile changes at local tile index ld_boundary
    auto step = [&](const int q, typename LS::Raw& sraw) {
        const int it = q >> 2, blk = q & 3;
        if (blk == 0 && it == ld_boundary) {                  // block-uniform: per-frame coefficient tables of the loaders
            const int n = ld_next;
            ld_next = n + 1;
            ld_boundary = n == frame0 ? first_boundary : ld_boundary + tiles_per_frame;
            __syncthreads();
            ls.fill(n, UB_HID, sCfS);
            lr.fill(n, UB_WIDTH, sCfR);
            __syncthreads();
        }
        const typename LS::Raw curs = sraw;
        const typename LR::Raw curr[2] = {rawr[0], rawr[1]};
        const uint32_t slot = (uint32_t)q % FRING, use = (uint32_t)q / FRING;
        mbar_wait_guard(bRing + slot * 8, (use & 1) ^ 1);              // MMAs that read this ring slot are done
        char* s_hi = sS + slot * 2 * FBLK;
        char* s_lo = s_hi + FBLK;
        {
            typename LS::Cf cfs;
            ls.coefs(UB_HID, blk * 64 + sc * 8, sCfS, cfs);
            float v[8];
            ls.finish(curs, cfs, v);
            const int off = sr * 128 + ((sc ^ (sr & 7)) << 4);
            split_store8(v, s_hi + off, s_lo + off, single);
        }
        if (blk == 0) {
            // R after S: the previous tile's last MMAs (issued a moment ago) still read R; converting the S block first gives them time
            // (the wait used to be polled ~11 times per tile by every warp: ncu source view)
            mbar_wait_guard(bRfree, ((uint32_t)it & 1) ^ 1);           // the previous tile's MMAs no longer read R
            typename LR::Cf cfr;
            lr.coefs(UB_WIDTH, rc * 8, sCfR, cfr);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int r = rr + 32 * j;
                float v[8];
                lr.finish(curr[j], cfr, v);
                const int off = (rc / 8) * FBLK + r * 128 + (((rc % 8) ^ (r & 7)) << 4);
                split_store8(v, sR + off, sR + 2 * FBLK + off, single);
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t r_hi = smem_u32(sR), r_lo = r_hi + 2 * FBLK, b_hi = smem_u32(s_hi), b_lo = smem_u32(s_lo);
            {
                // dn0[128][px] += W1^T[:, blk] . dh1_blk   and   dW1^T[128 (channels of R)][blk*64 .. +64] += R^T . S_blk (contraction over
                // the 64 pixels), interleaved: two independent accumulation chains
                const uint32_t a_hi = smem_u32(sW) + blk * (NOUT * 128), a_lo = a_hi + W_HALF;
                const uint32_t d = tmem_base + (uint32_t)(it & 1) * ACC, dw = tmem_base + DW_COL + blk * 64;
#pragma unroll
                for (int k16 = 0; k16 < 4; ++k16) {
                    const uint64_t wa = make_desc(a_hi + k16 * 32), wl = make_desc(a_lo + k16 * 32);
                    const uint64_t xa = make_desc(b_hi + k16 * 32), xl = make_desc(b_lo + k16 * 32);
                    const uint64_t ah = make_wg_desc(r_hi + k16 * 2048), al = make_wg_desc(r_lo + k16 * 2048);
                    const uint64_t bh = make_wg_desc(b_hi + k16 * 2048), bl = make_wg_desc(b_lo + k16 * 2048);
                    tc_mma(d, wa, xa, F_IDESC_K, (blk | k16) != 0);
                    tc_mma(dw, ah, bh, F_IDESC_MN, (it | k16) != 0);
                    if (!single) {
                        tc_mma(d, wa, xl, F_IDESC_K, 1);
                        tc_mma(dw, ah, bl, F_IDESC_MN, 1);
                        tc_mma(d, wl, xa, F_IDESC_K, 1);
                        tc_mma(dw, al, bh, F_IDESC_MN, 1);
                    }
                }
            }
            tc_commit(bRing + slot * 8);                                   // frees the ring slot
            if (blk == 3) {
                tc_commit(bAcc + (it & 1) * 8);                            // input-gradient accumulator of this tile complete
                tc_commit(bRfree);                                         // R may be overwritten by the next tile
                if (q == Q - 1) tc_commit(bDone);
            }
        }
        // Prefetch the operands of step q + 2 -- AFTER this step's proxy fence.  fence.proxy.async is MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC in
        // SASS and waits for every outstanding global load of the thread (ncu source view: long_scoreboard on the FENCE / BAR.SYNC), so
        // loads issued at the top of the step were complete when the step ended -- in flight for the convert only; issued here they stay
        // in flight across the epilogue slice and the next step (measured 7.6 -> 7.3 ms per step at B=16).
        if (q + 2 < Q) {
            const int nblk = (q + 2) & 3;
            ls.issue_off(soff, sraw);
            soff += nblk == 3 ? (size_t)FPX * UB_HID - 192 : 64;      // next block, or block 0 of the next tile
            if (nblk == 0) {                                  // blk == 2: this tile's R was consumed two steps ago
                lr.issue_off(roff, rawr[0]);
                lr.issue_off(roff + 32 * UB_WIDTH, rawr[1]);
                roff += (size_t)FPX * UB_WIDTH;
            }
        }
        // a slice of the previous tile's epilogue while this step's MMAs run; the slice after the last block also covers the
        // MMA latency before the next tile overwrites R
        if (it > 0 && (blk & 1)) epilogue_part(it - 1, blk >> 1);
    };
    for (int q = 0; q < Q; q += 2) {                          // Q is a multiple of 4: the two register sets alternate statically
        step(q, raws[0]);
        step(q + 1, raws[1]);
    }
#pragma unroll
    for (int part = 0; part < PARTS; ++part) epilogue_part(ntiles - 1, part);
    flush_stats();
    // ---- weight-gradient partial of this CTA ----
    mbar_wait_guard(bDone, 0);
    tc_fence_after();
    {
        const int cbk = warp / 4;                             // lanes lq*32.., columns cbk*64..
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + DW_COL + cbk * 64 + hh * 32, v);
            const int m = lq * 32 + lane;
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[(size_t)m * sa + (size_t)(cbk * 64 + hh * 32 + i) * sb] = v[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

template <class LS, class LR, class Epi>
static int launch_bwd_fused(LS ls, LR lr, const void* wimg, Epi ep, float* partial, int max_parts, int N, int P, int sa, int sb,
                            int single, int* nparts, cudaStream_t st) {
    if (P % FPX != 0) return UB_ERR_ARG;
    constexpr size_t smem = (size_t)F_W_BYTES + 4 * FBLK + FRING * 2 * FBLK + 3 * (UB_HID + UB_WIDTH) * sizeof(float) + (FRING + 4) * 8 + 16 + 1024;
    auto kern = bwd_tc_kernel<LS, LR, Epi>;
    UB_SET_SMEM(kern, smem);
    const long long total = (long long)N * (P / FPX);
    const int blocks = (int)(total < max_parts ? total : max_parts);
    kern<<<blocks, THREADS, smem, st>>>(ls, lr, static_cast<const uint4*>(wimg), ep, partial, P, total, sa, sb, single);
    UB_CHECK_LAUNCH();
    *nparts = blocks;
    return UB_OK;
}

// ------------------------------------------------------------------------------------------
// 3x3 convolution (forward and input gradient) fed entirely by the TMA engine -- no producer warps.
//
// The pre-split image (common.cuh: 512-byte pixels, image rows padded by one halo pixel per side) is a 4-D tensor
// [n][y][x'][256 halves] for the TMA; the operand tile of output pixels (y, x0 .. x0+127) for tap (dy, dx) and channel block h
// is the box {64 halves, 128 pixels, 1, 1} at coordinates (h*64 [+128 for the lo halves], x0 + dx + 1, y + dy, n), written with the
// hardware's 128-byte swizzle -- byte for byte the K-major SWIZZLE_128B tile the thread loaders wrote.  The halo pixels make the
// shifted box always lie inside the padded row (reflected neighbour for the activations, zeros for the output gradient); rows
// above / below the image are the reflected coordinate (forward) or out of bounds = zero fill (input gradient).
// Roles: warp 0 one lane: per K-block two cp.async.bulk for the weight slabs + two cp.async.bulk.tensor for the operand halves into
// a 3-stage ring (64 KB per stage, mbarrier complete_tx); warp 1 one lane: tcgen05.mma + commit; warps 2..9: epilogue out of the
// double-buffered TMEM accumulator.  Requires W % 128 == 0 (a tile inside one image row); other widths use the thread-loader kernel.
// ncu of the thread-loader version (profiles/r02_ncu_conv3x3.md): the 16 producer warps were bound by the latency of their own
// instruction stream (250 instructions per thread and K-block), tensor pipe 29 % active.
// ------------------------------------------------------------------------------------------
constexpr int CT_STAGES = 3;
constexpr int CT_STAGE_BYTES = 4 * 16384;      // weight hi slab, weight lo slab, operand hi tile, operand lo tile
constexpr int CT_EPI_WARPS = 8;
constexpr int CT_THREADS = 32 * (2 + CT_EPI_WARPS);

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <class Epi>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv_tma_kernel(const __grid_constant__ CUtensorMap tmap, const char* __restrict__ wimg, Epi ep, int H, int W, int reflect, int single) {
    constexpr int KB = 18, NOUT = UB_WIDTH, SLAB = NOUT * 128, W_HALF = 9 * UB_WIDTH * NOUT * 2;
    extern __shared__ __align__(1024) char smem_raw[];
    char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* sBar = reinterpret_cast<uint64_t*>(smem + CT_STAGES * CT_STAGE_BYTES);      // full[S], empty[S], accfull[2], accempty[2]
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 2 * CT_STAGES + 4);
    const uint32_t bFull = smem_u32(&sBar[0]), bEmpty = smem_u32(&sBar[CT_STAGES]), bAccFull = smem_u32(&sBar[2 * CT_STAGES]),
                   bAccEmpty = smem_u32(&sBar[2 * CT_STAGES + 2]);
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, n = blockIdx.y;
    const int P = H * W, tiles_per_frame = P / TILE_PX, tiles_per_row = W / TILE_PX;
    const int t0 = (int)(((long long)blockIdx.x * tiles_per_frame) / gridDim.x);
    const int t1 = (int)(((long long)(blockIdx.x + 1) * tiles_per_frame) / gridDim.x);
    const int ntiles = t1 - t0, Q = ntiles * KB;
    if (tid == 0) {
        for (int i = 0; i < 2 * CT_STAGES + 2; ++i) mbar_init(smem_u32(&sBar[i]), 1);
        for (int i = 0; i < 2; ++i) mbar_init(bAccEmpty + i * 8, CT_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sTmem;

    if (warp == 0) {
        if (lane == 0) {                       // ---- TMA producer ----
            for (int q = 0; q < Q; ++q) {
                const int it = q / KB, kb = q - it * KB, s = q % CT_STAGES, use = q / CT_STAGES;
                mbar_wait_guard(bEmpty + s * 8, ((uint32_t)use & 1) ^ 1);
                const int t = t0 + it, y = t / tiles_per_row, x0 = (t - y * tiles_per_row) * TILE_PX;
                const int tap = kb >> 1, half = kb & 1, t3 = tap >= 6 ? 2 : (tap >= 3 ? 1 : 0);
                int yy = y + t3 - 1;
                const int xc = x0 + (tap - 3 * t3) - 1 + 1;                     // + 1: halo pixel in front of every image row
                if (reflect) yy = yy < 0 ? 1 : (yy >= H ? H - 2 : yy);            // else: rows outside the image are zero-filled by the TMA
                const uint32_t stage = smem_u32(smem) + (uint32_t)s * CT_STAGE_BYTES, bar = bFull + s * 8;
                mbar_expect_tx(bar, CT_STAGE_BYTES);
                bulk_g2s(stage, wimg + (size_t)kb * SLAB, SLAB, bar);
                bulk_g2s(stage + 16384, wimg + (size_t)W_HALF + (size_t)kb * SLAB, SLAB, bar);
                tma_load_4d(stage + 32768, &tmap, half * 64, xc, yy, n, bar);
                tma_load_4d(stage + 49152, &tmap, 128 + half * 64, xc, yy, n, bar);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                       // ---- MMA issuer ----
            const uint32_t idesc = single == SPLIT_F16X3 ? (c_idesc & ~((7u << 7) | (7u << 10))) : c_idesc;
            for (int q = 0; q < Q; ++q) {
                const int it = q / KB, kb = q - it * KB, s = q % CT_STAGES, use = q / CT_STAGES;
                if (kb == 0) {                 // the accumulator stage must have been drained by the epilogue of tile it - 2
                    mbar_wait_guard(bAccEmpty + (it & 1) * 8, (((uint32_t)it >> 1) & 1) ^ 1);
                }
                mbar_wait_guard(bFull + s * 8, (uint32_t)use & 1);
                tc_fence_after();
                const uint32_t stage = smem_u32(smem) + (uint32_t)s * CT_STAGE_BYTES;
                const uint32_t a_hi = stage, a_lo = stage + 16384, b_hi = stage + 32768, b_lo = stage + 49152;
                const uint32_t d = tmem_base + (uint32_t)(it & 1) * TILE_PX;
#pragma unroll
                for (int k16 = 0; k16 < KBLK / 16; ++k16) {
                    const uint64_t wa = make_desc(a_hi + k16 * 32), wl = make_desc(a_lo + k16 * 32);
                    const uint64_t xa = make_desc(b_hi + k16 * 32), xl = make_desc(b_lo + k16 * 32);
                    tc_mma(d, wa, xa, idesc, (kb | k16) != 0);
                    if (single != SPLIT_BF16X1) {
                        tc_mma(d, wa, xl, idesc, 1);
                        tc_mma(d, wl, xa, idesc, 1);
                    }
                }
                tc_commit(bEmpty + s * 8);
                if (kb == KB - 1) tc_commit(bAccFull + (it & 1) * 8);
            }
        }
    } else {                                   // ---- epilogue: TMEM lane quarter = warp % 4, 64-pixel half = (warp - 2) / 4 ----
        const int lq = warp % 4, ph = (warp - 2) / 4, ch = lq * 32 + lane;
        typename Epi::State est;
        float stat[Epi::NS];
        ep.init(n, NOUT, ch, est);
#pragma unroll
        for (int i = 0; i < Epi::NS; ++i) stat[i] = 0.f;
        for (int it = 0; it < ntiles; ++it) {
            mbar_wait_guard(bAccFull + (it & 1) * 8, ((uint32_t)it >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(it & 1) * TILE_PX + ph * 64 + c * 32, v);
                ep.template apply<32>(est, (size_t)n * P + (size_t)(t0 + it) * TILE_PX + ph * 64 + c * 32, NOUT, ch, v, stat);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bAccEmpty + (it & 1) * 8);
        }
        double* dst = ep.dst(n, NOUT);
#pragma unroll
        for (int i = 0; i < Epi::NS; ++i) atomicAdd(&dst[(size_t)ch * Epi::NS + i], (double)stat[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
}

// 4-D tensor map of a pre-split image [N][H][W + 2][256 halves] with box {64, 128, 1, 1} and the 128-byte swizzle
static int make_split_tmap(CUtensorMap* tm, const void* img, int N, int H, int W, int box_px = TILE_PX) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return UB_ERR_CUDA;
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const cuuint64_t dims[4] = {256, (cuuint64_t)(W + 2), (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {UB_SPLIT_ROW, (cuuint64_t)(W + 2) * UB_SPLIT_ROW, (cuuint64_t)H * (W + 2) * UB_SPLIT_ROW};
    const cuuint32_t box[4] = {64, (cuuint32_t)box_px, 1, 1}, estr[4] = {1, 1, 1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(img), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? UB_OK : UB_ERR_CUDA;
}

template <class Epi>
static int launch_conv_tma(const void* img, const void* wimg, Epi ep, int N, int H, int W, int reflect, int single, cudaStream_t st) {
    CUtensorMap tm;
    UB_TRY(make_split_tmap(&tm, img, N, H, W));
    constexpr size_t smem = (size_t)CT_STAGES * CT_STAGE_BYTES + (2 * CT_STAGES + 4) * 8 + 16 + 1024;
    auto kern = conv_tma_kernel<Epi>;
    UB_SET_SMEM(kern, smem);
    const dim3 grid(blocks_per_frame(N, H * W / TILE_PX), N);
    kern<<<grid, CT_THREADS, smem, st>>>(tm, static_cast<const char*>(wimg), ep, H, W, reflect, single);
    UB_CHECK_LAUNCH();
    return UB_OK;
}

// ------------------------------------------------------------------------------------------
// Weight gradient of a 3x3 convolution for one tap pair, TMA-fed:  dW[co][j*128 + ci] += sum_p dc[p][co] * act[reflect(p + tap_j)][ci].
// Same MN-major tcgen05 GEMM as wgrad_tc_body (64-pixel tiles, the whole 128 x 256 gradient in TMEM), but both operands arrive as
// TMA boxes of the pre-split images: A = four [64 px][64 halves] boxes of the output gradient (hi / lo x two channel blocks), B = eight
// boxes of the activations (two taps x hi / lo x two channel blocks), each tap's box shifted by (dy, dx) (halo pixel / reflected row
// coordinate).  Warp 0 lane 0: TMA producer (2-stage ring, 96 KB per stage); warp 1 lane 0: MMA issuer; warps 2..5: final TMEM read-out.
// Requires W % 64 == 0 (a tile inside one image row); persistent CTAs walk contiguous tile ranges across frames.
// ------------------------------------------------------------------------------------------
constexpr int CW_STAGE_BYTES = WG_STAGE;       // 32 KB A + 64 KB B
constexpr int CW_THREADS = 32 * 6;
__global__ void __launch_bounds__(CW_THREADS, 1)
conv_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmap_dc, const __grid_constant__ CUtensorMap tmap_act, float* __restrict__ partial,
                      int H, int W, long long total_tiles, int tapA, int tapB, int single) {
    extern __shared__ __align__(1024) char smem_raw[];
    char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* sBar = reinterpret_cast<uint64_t*>(smem + 2 * CW_STAGE_BYTES);     // full[2], empty[2], done
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 5);
    const uint32_t bFull = smem_u32(&sBar[0]), bEmpty = smem_u32(&sBar[2]), bDone = smem_u32(&sBar[4]);
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const long long per = (total_tiles + gridDim.x - 1) / gridDim.x;
    const long long t0 = (long long)blockIdx.x * per, t1 = min(t0 + per, total_tiles);
    const int ntiles = t1 > t0 ? (int)(t1 - t0) : 0;
    float* dst = partial + (size_t)blockIdx.x * 128 * 256;
    if (ntiles == 0) {
        for (int i = tid; i < 128 * 256; i += CW_THREADS) dst[i] = 0.f;
        return;
    }
    if (tid == 0) {
        for (int i = 0; i < 5; ++i) mbar_init(smem_u32(&sBar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sTmem;
    const int tiles_per_row = W / WG_PX, tiles_per_frame = H * tiles_per_row;
    if (warp == 0) {
        if (lane == 0) {                       // ---- TMA producer ----
            const int tb = tapB >= 0 ? tapB : tapA;                // absent second tap: load the first one again (its columns are ignored)
            const int taps[2] = {tapA, tb};
            for (int i = 0; i < ntiles; ++i) {
                const int s = i & 1, use = i >> 1;
                mbar_wait_guard(bEmpty + s * 8, ((uint32_t)use & 1) ^ 1);
                const long long t = t0 + i;
                const int n = (int)(t / tiles_per_frame), r = (int)(t - (long long)n * tiles_per_frame);
                const int y = r / tiles_per_row, x0 = (r - y * tiles_per_row) * WG_PX;
                const uint32_t stage = smem_u32(smem) + (uint32_t)s * CW_STAGE_BYTES, bar = bFull + s * 8;
                mbar_expect_tx(bar, CW_STAGE_BYTES);
                // A: output gradient, hi blocks 0..1 then lo blocks 0..1 (8 KB each)
#pragma unroll
                for (int b = 0; b < 4; ++b) tma_load_4d(stage + b * WG_BLK, &tmap_dc, (b >> 1) * 128 + (b & 1) * 64, x0 + 1, y, n, bar);
                // B: activations, hi blocks {tapA c0, tapA c64, tapB c0, tapB c64} then the same four lo blocks
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int tap = taps[(b >> 1) & 1], t3 = tap >= 6 ? 2 : (tap >= 3 ? 1 : 0);
                    int yy = y + t3 - 1;
                    yy = yy < 0 ? 1 : (yy >= H ? H - 2 : yy);
                    tma_load_4d(stage + WG_A_BYTES + b * WG_BLK, &tmap_act, (b >> 2) * 128 + (b & 1) * 64, x0 + (tap - 3 * t3) - 1 + 1, yy, n, bar);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                       // ---- MMA issuer ----
            for (int i = 0; i < ntiles; ++i) {
                const int s = i & 1, use = i >> 1;
                mbar_wait_guard(bFull + s * 8, (uint32_t)use & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem) + (uint32_t)s * CW_STAGE_BYTES, a_lo = a_hi + WG_A_BYTES / 2;
                const uint32_t b_hi = a_hi + WG_A_BYTES, b_lo = b_hi + WG_B_BYTES / 2;
#pragma unroll
                for (int p16 = 0; p16 < WG_PX / 16; ++p16) {
                    const uint64_t ah = make_wg_desc(a_hi + p16 * 2048), al = make_wg_desc(a_lo + p16 * 2048);
                    const uint64_t bh = make_wg_desc(b_hi + p16 * 2048), bl = make_wg_desc(b_lo + p16 * 2048);
                    tc_mma(tmem_base, ah, bh, c_wg_idesc, (i | p16) != 0);
                    if (!single) {
                        tc_mma(tmem_base, ah, bl, c_wg_idesc, 1);
                        tc_mma(tmem_base, al, bh, c_wg_idesc, 1);
                    }
                }
                tc_commit(bEmpty + s * 8);
                if (i == ntiles - 1) tc_commit(bDone);
            }
        }
    } else {                                   // ---- read-out of the 128 x 256 gradient (warp's TMEM lane quarter = warp % 4) ----
        mbar_wait_guard(bDone, 0);
        tc_fence_after();
        const int lq = warp % 4, m = lq * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + c * 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[(size_t)m * 256 + c * 32 + i] = v[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
}

static int launch_conv_wgrad_tma(const void* dcs, const void* xs, float* partial, int max_parts, int N, int H, int W, int tapA, int tapB,
                                 int single, int* nparts, cudaStream_t st) {
    CUtensorMap tm_dc, tm_act;
    UB_TRY(make_split_tmap(&tm_dc, dcs, N, H, W, WG_PX));
    UB_TRY(make_split_tmap(&tm_act, xs, N, H, W, WG_PX));
    constexpr size_t smem = (size_t)2 * CW_STAGE_BYTES + 5 * 8 + 16 + 1024;
    UB_SET_SMEM(conv_wgrad_tma_kernel, smem);
    const long long total = (long long)N * (H * W / WG_PX);
    const int blocks = (int)(total < max_parts ? total : max_parts);
    conv_wgrad_tma_kernel<<<blocks, CW_THREADS, smem, st>>>(tm_dc, tm_act, partial, H, W, total, tapA, tapB, single);
    UB_CHECK_LAUNCH();
    *nparts = blocks;
    return UB_OK;
}

// Weight image: src is fp32 [rows][K] (transpose == 0) or [K][rows] (transpose == 1, i.e. the M operand is src^T).
// Output: bf16 hi image then lo image, each [K/64][rows][64] in the K-major SWIZZLE_128B layout.
__global__ void prep_weights_kernel(const float* __restrict__ src, uint16_t* __restrict__ img, int rows, int K, int transpose, int f16) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * K) return;
    const int r = i / K, k = i % K;
    const float w = transpose ? src[(size_t)k * rows + r] : src[i];
    uint16_t hbits, lbits;
    if (f16) {            // fp16 hi / lo (forward images, SPLIT_F16X3); saturating like the activation split
        const float ws = fminf(fmaxf(w, -65504.f), 65504.f);
        const __half hh = __float2half_rn(ws);
        const __half lh = __float2half_rn(fminf(fmaxf(w - __half2float(hh), -65504.f), 65504.f));
        hbits = __half_as_ushort(hh);
        lbits = __half_as_ushort(lh);
    } else {
        const __nv_bfloat16 hb = __float2bfloat16_rn(w);
        const __nv_bfloat16 lb = __float2bfloat16_rn(w - __bfloat162float(hb));
        hbits = __bfloat16_as_ushort(hb);
        lbits = __bfloat16_as_ushort(lb);
    }
    const int kb = k / KBLK, kk = k % KBLK;
    const size_t off = ((size_t)kb * rows + r) * 128 + (((kk / 8) ^ (r & 7)) << 4) + (kk % 8) * 2;   // bytes
    img[off / 2] = hbits;
    img[(off + (size_t)rows * K * 2) / 2] = lbits;
}

// Weight image of a 3x3 convolution, src = nn.Conv2d weight [co][ci][3][3] (128 x 128 x 9).  The GEMM's K index is k = tap*128 + c.
//   dgrad == 0 (forward):        M[r = co][k = tap*128 + ci] = W[co][ci][tap]                       out[p] = sum_tap W_tap . in[p + tap]
//   dgrad == 1 (input gradient): M[r = ci][k = tap*128 + co] = W[co][ci][8 - tap]                   din[q] = sum_tap W_(-tap)^T . dout[q + tap]
// Same hi/lo split and K-major SWIZZLE_128B layout as prep_weights_kernel, rows = 128, K = 1152.
__global__ void prep_conv_weights_kernel(const float* __restrict__ src, uint16_t* __restrict__ img, int dgrad, int f16) {
    constexpr int rows = 128, K = 9 * 128;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * K) return;
    const int r = i / K, k = i % K, tap = k / 128, c = k % 128;
    const float w = dgrad ? src[((size_t)c * 128 + r) * 9 + (8 - tap)] : src[((size_t)r * 128 + c) * 9 + tap];
    uint16_t hbits, lbits;
    if (f16) {
        const float ws = fminf(fmaxf(w, -65504.f), 65504.f);
        const __half hh = __float2half_rn(ws);
        const __half lh = __float2half_rn(fminf(fmaxf(w - __half2float(hh), -65504.f), 65504.f));
        hbits = __half_as_ushort(hh);
        lbits = __half_as_ushort(lh);
    } else {
        const __nv_bfloat16 hb = __float2bfloat16_rn(w);
        const __nv_bfloat16 lb = __float2bfloat16_rn(w - __bfloat162float(hb));
        hbits = __bfloat16_as_ushort(hb);
        lbits = __bfloat16_as_ushort(lb);
    }
    const int kb = k / KBLK, kk = k % KBLK;
    const size_t off = ((size_t)kb * rows + r) * 128 + (((kk / 8) ^ (r & 7)) << 4) + (kk % 8) * 2;   // bytes
    img[off / 2] = hbits;
    img[(off + (size_t)rows * K * 2) / 2] = lbits;
}

// dW[co][ci][tapA | tapB] += sum over the CTA partials [co][j*128 + ci] (j = 0: tapA, 1: tapB; tapB < 0: absent)
__global__ void reduce_conv_partials_kernel(const float* __restrict__ partial, float* __restrict__ grad, int tapA, int tapB, int nparts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 128 * 256) return;
    const int co = i / 256, n = i % 256, tap = n < 128 ? tapA : tapB;
    if (tap < 0) return;
    float s = 0.f;
    for (int b = 0; b < nparts; ++b) s += partial[(size_t)b * 128 * 256 + i];
    grad[((size_t)co * 128 + (n & 127)) * 9 + tap] += s;
}

}  // namespace tc

// ---- 3x3 convolutions of the residual blocks (uncrtaints.py:24-69) on the tcgen05 path ----
int tc_prep_conv_weights(const float* w, void* img, int dgrad, int f16, cudaStream_t st) {
    tc::prep_conv_weights_kernel<<<(128 * 1152 + 255) / 256, 256, 0, st>>>(w, static_cast<uint16_t*>(img), dgrad, f16);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
// split[N*P][512 B] = hi / lo halves of relu?(x*scale + shift) (coef may be null); mode: 2 = fp16 hi/lo (forward operands),
// 0 = bf16 hi/lo (gradient GEMMs), 1 = bf16 hi only
int tc_split_act(const float* x, const Coef* coef, int relu, void* split, int N, int H, int W, int mode, cudaStream_t st) {
    const size_t rows = (size_t)N * H * W;
    tc::split_act_kernel<<<(unsigned)((rows * 16 + 255) / 256), 256, 0, st>>>(x, coef, static_cast<char*>(split), H, W, rows, relu, mode);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
// c[N][P][128] = conv3x3_reflect(xs) + bias; stats[N][128][2] += column (sum, sumsq).  xs: pre-split input activations
int tc_conv3x3_fwd(const void* xs, const void* wimg, const float* bias, float* c, double* stats, int N, int H, int W, int single,
                   cudaStream_t st) {
    if (W % tc::TILE_PX == 0) return tc::launch_conv_tma(xs, wimg, tc::TEpiBiasStoreStats{c, bias, stats}, N, H, W, 1, single, st);
    tc::TLoadConvSplit al{static_cast<const char*>(xs), H, W, 0, 0, 0, 0};
    return tc::launch<9 * UB_WIDTH, UB_WIDTH, tc::TLoadConvSplit, tc::TEpiBiasStoreStats, true>(al, wimg, tc::TEpiBiasStoreStats{c, bias, stats}, N, H * W, single, st);
}
// din[N][P][128] = transposed 3x3 convolution of the zero-extended dc (+ add); the reflection's adjoint is added by launch_conv_fold.
// dcs: pre-split output gradient
int tc_conv3x3_dgrad(const void* dcs, const void* wimg_t, const float* add, float* din, double* scratch, int N, int H, int W, int single,
                     cudaStream_t st) {
    if (W % tc::TILE_PX == 0) return tc::launch_conv_tma(dcs, wimg_t, tc::TEpiStoreAdd{din, add, scratch}, N, H, W, 0, single, st);
    tc::TLoadConvSplit al{static_cast<const char*>(dcs), H, W, 1, 0, 0, 0};
    return tc::launch<9 * UB_WIDTH, UB_WIDTH, tc::TLoadConvSplit, tc::TEpiStoreAdd, true>(al, wimg_t, tc::TEpiStoreAdd{din, add, scratch}, N, H * W, single, st);
}
// dW[co][ci][tap] += sum_p dc[p][co] * act[reflect(p + tap)][ci] for all 9 taps (five launches of tap pairs); both operands pre-split
int tc_conv3x3_wgrad(const void* dcs, const void* xs, float* partial, int max_parts, float* dw, int N, int H, int W, int single,
                     cudaStream_t st) {
    for (int t = 0; t < 9; t += 2) {
        const int tapB = t + 1 < 9 ? t + 1 : -1;
        tc::TLoadConvSplit lb{static_cast<const char*>(xs), H, W, 0, 1, t, tapB};
        int nparts = 0;
        if (W % tc::WG_PX == 0)
            UB_TRY(tc::launch_conv_wgrad_tma(dcs, xs, partial, max_parts, N, H, W, t, tapB, single, &nparts, st));
        else
        UB_TRY(tc::launch_wgrad_tc(tc::TLoadPlainSplit{static_cast<const char*>(dcs), H, W}, lb, partial, max_parts, N, H * W, UB_HID, 1, single, &nparts, st));
        tc::reduce_conv_partials_kernel<<<(128 * 256 + 255) / 256, 256, 0, st>>>(partial, dw, t, tapB, nparts);
        UB_CHECK_LAUNCH();
    }
    return UB_OK;
}

int tc_prep_weights(const float* src, void* img, int rows, int K, int transpose, int f16, cudaStream_t st) {
    tc::prep_weights_kernel<<<(rows * K + 255) / 256, 256, 0, st>>>(src, static_cast<uint16_t*>(img), rows, K, transpose, f16);
    UB_CHECK_LAUNCH();
    return UB_OK;
}
// hbf != 0: the 256-channel hidden tensors (h1, h2, du, dz1) are stored as bf16 (gemm_backend bit 5)
int tc_gemm1_fwd(const float* x, const Coef* coef0, const void* w1img, void* h1, double* stats1, int N, int P, int single, int hbf,
                 cudaStream_t st) {
    tc::TLoadNormed al{x, coef0};
    if (hbf) return tc::launch<UB_WIDTH, UB_HID>(al, w1img, tc::TEpiStoreStatsT<tc::bf16_t>{static_cast<tc::bf16_t*>(h1), stats1}, N, P, single, st);
    return tc::launch<UB_WIDTH, UB_HID>(al, w1img, tc::TEpiStoreStats{static_cast<float*>(h1), stats1}, N, P, single, st);
}
int tc_gemm2_fwd(const void* h2, const Coef* coef2, const float* gate, const void* w2img, float* y, double* stats3, int N, int P,
                 int single, int hbf, cudaStream_t st) {
    tc::TEpiStoreStats ep{y, stats3};
    if (hbf) return tc::launch<UB_HID, UB_WIDTH>(tc::TLoadGeluGateT<tc::bf16_t>{static_cast<const tc::bf16_t*>(h2), coef2, gate}, w2img, ep, N, P, single, st);
    return tc::launch<UB_HID, UB_WIDTH>(tc::TLoadGeluGate{static_cast<const float*>(h2), coef2, gate}, w2img, ep, N, P, single, st);
}
// eval-mode BatchNorm block: out = x + Norm3(W2 . u) in the GEMM epilogue (no y tensor, no residual pass); stats = column sums of out
int tc_gemm2_fwd_residual(const void* h2, const Coef* coef2, const float* gate, const void* w2img, const float* x, const Coef* coef3,
                          float* out, double* stats, int N, int P, int single, int hbf, cudaStream_t st) {
    tc::TEpiResidual ep{out, x, coef3, stats};
    if (hbf) return tc::launch<UB_HID, UB_WIDTH>(tc::TLoadGeluGateT<tc::bf16_t>{static_cast<const tc::bf16_t*>(h2), coef2, gate}, w2img, ep, N, P, single, st);
    return tc::launch<UB_HID, UB_WIDTH>(tc::TLoadGeluGate{static_cast<const float*>(h2), coef2, gate}, w2img, ep, N, P, single, st);
}
int tc_gemm2_bwd(const float* dout, const float* y, const BCoef* bc3, const void* w2timg, void* du, const void* h2,
                 const Coef* coef2, const MeanRstd* mr2, double* sums3, int N, int P, int single, int hbf, cudaStream_t st) {
    tc::TLoadNormBwd al{dout, y, bc3};
    if (hbf)
        return tc::launch<UB_WIDTH, UB_HID>(al, w2timg, tc::TEpiGemm2BwdT<tc::bf16_t>{static_cast<tc::bf16_t*>(du), static_cast<const tc::bf16_t*>(h2), coef2, mr2, sums3},
                                            N, P, single, st);
    return tc::launch<UB_WIDTH, UB_HID>(al, w2timg, tc::TEpiGemm2Bwd{static_cast<float*>(du), static_cast<const float*>(h2), coef2, mr2, sums3}, N, P, single, st);
}
// dW2[o][k] += sum_p dy[p][o] * u[p][k]
int tc_wgrad2(const float* dout, const float* y, const BCoef* bc3, const void* h2, const Coef* coef2, const float* gate,
              float* partial, int max_parts, float* dw2, int N, int P, int single, int hbf, cudaStream_t st) {
    tc::TLoadNormBwd la{dout, y, bc3};
    int nparts = 0;
    int rc = hbf ? tc::launch_wgrad_tc(la, tc::TLoadGeluGateT<tc::bf16_t>{static_cast<const tc::bf16_t*>(h2), coef2, gate}, partial, max_parts, N, P, UB_HID, 1, single, &nparts, st)
                 : tc::launch_wgrad_tc(la, tc::TLoadGeluGate{static_cast<const float*>(h2), coef2, gate}, partial, max_parts, N, P, UB_HID, 1, single, &nparts, st);
    if (rc != UB_OK) return rc;
    return launch_reduce_partials(partial, dw2, UB_WIDTH * UB_HID, nparts, st);
}
// Fused backward of the expand convolution: dn0 + PreNorm-backward sums (as tc_gemm1_bwd) AND dW1 += dh1^T n0 (as tc_wgrad1)
int tc_gemm1_bwd(const void* dz1, const void* h1, const BCoef* bc1, const void* w1timg, float* dn0, const float* x,
                       const MeanRstd* mr0, double* bstats0, const Coef* coef0, float* partial, int max_parts, float* dw1, int N,
                       int P, int single, int hbf, cudaStream_t st) {
    tc::TLoadNormed lr{x, coef0};
    tc::TEpiGemm1Bwd ep{dn0, x, mr0, bstats0};
    int nparts = 0;
    int rc = hbf ? tc::launch_bwd_fused(tc::TLoadNormBwdT<tc::bf16_t>{static_cast<const tc::bf16_t*>(dz1), static_cast<const tc::bf16_t*>(h1), bc1}, lr, w1timg, ep, partial,
                                   max_parts, N, P, 1, UB_WIDTH, single, &nparts, st)
                 : tc::launch_bwd_fused(tc::TLoadNormBwd{static_cast<const float*>(dz1), static_cast<const float*>(h1), bc1}, lr, w1timg, ep, partial,
                                   max_parts, N, P, 1, UB_WIDTH, single, &nparts, st);
    if (rc != UB_OK) return rc;
    return launch_reduce_partials(partial, dw1, UB_WIDTH * UB_HID, nparts, st);
}

}  // namespace ub
