// Internal launcher declarations shared by the translation units of libuncrtaints_b200.so.
// Every launcher enqueues on the given stream only, allocates nothing, never synchronises, and
// returns UB_OK or a negative UB_ERR_* code.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace ub {

struct Coef;
struct MeanRstd;
struct BCoef;

// norm.cu
int launch_norm_finalize(const double* stats, const float* gamma, const float* beta, float* rm, float* rv, Coef* coef,
                         MeanRstd* mr, int N, int C, int groups, double count, float eps, float momentum, int training,
                         cudaStream_t st);
int launch_norm_finalize_bwd(const double* bstats, const float* gamma, const MeanRstd* mr, BCoef* bcoef, float* dgamma,
                             float* dbeta, int N, int C, int groups, double count, int training, cudaStream_t st);
int launch_residual_fwd(const float* x, const float* y, const Coef* coef3, float* out, double* out_stats, int N, int P,
                        cudaStream_t st);
int launch_se_pool(const void* h2, const Coef* coef2, const MeanRstd* mr2, double* pool_stats, double* gp_stats, int N,
                   int P, int hbf, cudaStream_t st);
int launch_norm_bwd_stats(const float* dy, const float* v, const MeanRstd* mr, double* bstats, int N, int P, cudaStream_t st);
int launch_residual_bwd(const float* dout, const float* dn0, const float* x, const BCoef* bc0, float* dx, int N, int P,
                        int relu_mask, const float* y_below, const MeanRstd* mr_below, double* bstats_below, cudaStream_t st);

// norm.cu: residual blocks (conv3x3 -> norm -> ReLU layers)
int launch_residual_relu_fwd(const float* x, const float* c3, const Coef* coef3, float* out, double* out_stats, int N, int P, cudaStream_t st);
int launch_relu_norm_bwd_stats(const float* dy, const float* c, const Coef* coef, const MeanRstd* mr, double* bstats, int N, int P, cudaStream_t st);
int launch_relu_norm_bwd_apply(const float* dy, const float* c, const Coef* coef, const BCoef* bc, void* dc_split, float* dbias, int N, int H,
                               int W, cudaStream_t st);
int launch_conv_fold(const void* dc_split, const float* w, float* wt_scratch /* 9*128*128 floats */, float* din, int N, int H, int W,
                     cudaStream_t st);

// gemm_simt.cu
int launch_transpose(const float* in, float* out, int rows, int cols, cudaStream_t st);
int launch_reduce_partials(const float* partial, float* grad, int count, int nparts, cudaStream_t st);
int simt_gemm1_fwd(const float* x, const Coef* coef0, const float* w1t, float* h1, double* stats1, int N, int P, cudaStream_t st);
int simt_gemm2_fwd(const float* h2, const Coef* coef2, const float* gate, const float* w2t, float* y, double* stats3, int N,
                   int P, cudaStream_t st);
int simt_gemm2_bwd(const float* dout, const float* y, const BCoef* bc3, const float* w2, float* du, const float* h2,
                   const Coef* coef2, const MeanRstd* mr2, double* sums3, int N, int P, cudaStream_t st);
int simt_gemm1_bwd(const float* dz1, const float* h1, const BCoef* bc1, const float* w1, float* dn0, const float* x,
                   const MeanRstd* mr0, double* bstats0, int N, int P, cudaStream_t st);
int simt_wgrad2(const float* dout, const float* y, const BCoef* bc3, const float* h2, const Coef* coef2, const float* gate,
                float* partial, int max_parts, float* dw2, int N, int P, cudaStream_t st);
int simt_wgrad1(const float* x, const Coef* coef0, const float* dz1, const float* h1, const BCoef* bc1, float* partial,
                int max_parts, float* dw1, int N, int P, cudaStream_t st);

// gemm_tc.cu (tcgen05 bf16x3 tensor-core versions of the four streaming GEMMs and the two weight-gradient GEMMs).
// single: operand split mode -- 0 = bf16 hi/lo (three MMAs), 1 = one bf16 (single pass, gemm_backend bit 2, reduced precision),
// 2 = fp16 hi/lo (three MMAs, fp32-grade; forward GEMMs only: needs a weight image prepared with f16 = 1).
int tc_prep_weights(const float* src, void* img, int rows, int K, int transpose, int f16, cudaStream_t st);
// hbf != 0: the 256-channel hidden tensors (h1, h2, du, dz1; passed as void*) are stored as bf16 (gemm_backend bit 5)
int tc_gemm1_fwd(const float* x, const Coef* coef0, const void* w1img, void* h1, double* stats1, int N, int P, int single, int hbf,
                 cudaStream_t st);
int tc_gemm2_fwd(const void* h2, const Coef* coef2, const float* gate, const void* w2img, float* y, double* stats3, int N,
                 int P, int single, int hbf, cudaStream_t st);
int tc_gemm2_fwd_residual(const void* h2, const Coef* coef2, const float* gate, const void* w2img, const float* x, const Coef* coef3,
                          float* out, double* stats, int N, int P, int single, int hbf, cudaStream_t st);
int tc_gemm2_bwd(const float* dout, const float* y, const BCoef* bc3, const void* w2timg, void* du, const void* h2,
                 const Coef* coef2, const MeanRstd* mr2, double* sums3, int N, int P, int single, int hbf, cudaStream_t st);
int tc_wgrad2(const float* dout, const float* y, const BCoef* bc3, const void* h2, const Coef* coef2, const float* gate,
              float* partial, int max_parts, float* dw2, int N, int P, int single, int hbf, cudaStream_t st);
// expand convolution backward: dn0 + PreNorm-backward sums AND dW1 += dh1^T n0 in ONE kernel (one read of dz1, h1, x)
int tc_gemm1_bwd(const void* dz1, const void* h1, const BCoef* bc1, const void* w1timg, float* dn0, const float* x,
                 const MeanRstd* mr0, double* bstats0, const Coef* coef0, float* partial, int max_parts, float* dw1, int N,
                 int P, int single, int hbf, cudaStream_t st);

// gemm_tc.cu: dense 3x3 convolutions (reflect padding) of the residual blocks as implicit GEMMs over (tap, channel), weights streamed,
// operands read from pre-split (hi / lo) images
int tc_prep_conv_weights(const float* w, void* img, int dgrad, int f16, cudaStream_t st);
int tc_split_act(const float* x, const Coef* coef, int relu, void* split, int N, int H, int W, int mode, cudaStream_t st);
int tc_conv3x3_fwd(const void* xs, const void* wimg, const float* bias, float* c, double* stats, int N, int H, int W, int single,
                   cudaStream_t st);
int tc_conv3x3_dgrad(const void* dcs, const void* wimg_t, const float* add, float* din, double* scratch, int N, int H, int W, int single,
                     cudaStream_t st);
int tc_conv3x3_wgrad(const void* dcs, const void* xs, float* partial, int max_parts, float* dw, int N, int H, int W, int single,
                     cudaStream_t st);

// dwconv_rows.cu (row-streaming depthwise kernels fed by TMA bulk copies; the backward one is fused: du, h2, h1 -> dz1 in one pass)
int launch_dwconv_fwd(const void* h1, const Coef* coef1, const float* wdw, void* h2, double* stats2, int N, int H, int W, int hbf,
                      cudaStream_t st);
int launch_dwconv_fwd_pool(const void* h1, const Coef* coef1, const float* wdw, void* h2, const Coef* coef2, double* pool, int N,
                           int H, int W, int hbf, cudaStream_t st);
int launch_dwconv_bwd(const void* du, const void* h2, const void* h1, const float* gate, const float* dmp,
                      const Coef* coef2, const BCoef* bc2, const Coef* coef1, const MeanRstd* mr1, const float* wdw,
                      void* dz1, double* bstats1, float* dwdw, int N, int H, int W, int hbf, cudaStream_t st);

// se.cu
int launch_se_fwd(const double* pool_stats, const float* f1, const float* f2, float* save, float* gate, int N, int P,
                  cudaStream_t st);
int launch_se_bwd(const double* sums3, const double* gp_stats, const float* f1, const float* f2, const float* save,
                  float* df1, float* df2, float* dmp, double* bstats2, int N, int P, cudaStream_t st);

// inconv.cu
int launch_inconv_apply(const float* x, const float* w, const float* b, const Coef* coef, float* x0, double* stats_x0, int N,
                        int Cin, int P, cudaStream_t st);

size_t inconv_moments_bytes(int N);
size_t inconv_gram_bytes(int N);
int launch_inconv_stats_moments(const float* x, const float* w, const float* b, double* mom, double* stats, int* notpad,
                                float pad_value, int N, int Cin, int P, cudaStream_t st);
int launch_inconv_bwd_gram(const float* x, const float* x0, const float* dx0, const float* w, const float* b, const MeanRstd* mr,
                           double* gacc, double* bstats, int N, int Cin, int P, cudaStream_t st);
int launch_inconv_bwd_finish(const double* gacc, const double* mom, const float* w, const float* b, const BCoef* bc, float* dw,
                             float* db, int N, int Cin, int P, cudaStream_t st);

// temporal.cu
int launch_maxpool_fwd(const float* x, float* pooled, int* idx, int N, int H, int W, cudaStream_t st);
int launch_maxpool_bwd(const float* dpooled, const int* idx, float* denc, int N, int HW, cudaStream_t st);
int launch_ltae_fwd(const float* pooled, const float* Ap, const float* e, const int* notpad, float* attn, int B, int T,
                    float eps, cudaStream_t st);
int launch_ltae_bwd(const float* pooled, const float* Ap, const float* attn, const float* dattn, float* dpooled, float* dAp,
                    float* de, int B, int T, float eps, cudaStream_t st);
int launch_aggregate_fwd(const float* attn, const int* notpad, const unsigned char* keep_mask, unsigned long long seed,
                         unsigned long long offset, float drop_p, const float* x, float* out, double* out_stats, int B,
                         int T, int H, int W, cudaStream_t st);
int launch_aggregate_bwd(const float* attn, const int* notpad, const unsigned char* keep_mask, unsigned long long seed,
                         unsigned long long offset, float drop_p, const float* x, const float* dagg, float* denc,
                         float* dwup, float* dattn, int B, int T, int H, int W, cudaStream_t st);

// gemm_simt.cu: generic fp32 linear layers (the `use_v` value path, ltae_v.cu).  a: [N*P][K], wt: [K][NOUT], out: [N*P][NOUT] =
// a . wt + bias[NOUT] + rowbias[N][NOUT] (either may be null); stats[N][NOUT][2] += column (sum, sumsq) (always written: pass scratch)
int simt_linear(const float* a, int K, const float* wt, int NOUT, const float* bias, const float* rowbias, float* out, double* stats,
                int N, int P, cudaStream_t st);
// out[N][H*W][128] = a . wt + bilinear_up(low [N][32*32][128]); stats as above
int simt_linear_upadd(const float* a, const float* wt, const float* low, float* out, double* stats, int N, int H, int W, cudaStream_t st);
// grad[ia*sa + ib*sb] += sum_rows a[row][ia] * b[row][ib] (cb = 256), or grad[ia*ld + ib] += ... (cb = 128); a: [N*P][128], b: [N*P][cb]
int simt_wgrad_plain(const float* a, const float* b, int cb, float* partial, int max_parts, float* grad, int sa, int sb, int ld, int N, int P,
                     cudaStream_t st);

// ltae_v.cu (`use_v`: full LTAE2d value path + include_v)
int launch_ltaev_norm(const float* pooled, const float* gamma, const float* beta, float* xn, int B, int T, float eps, cudaStream_t st);
int launch_ltaev_attnv(const float* attn, const float* z, float* o, int B, int T, cudaStream_t st);
int launch_ltaev_attnv_bwd(const float* attn, const float* z, const float* d_o, float* dz, float* dattn, int B, int T, cudaStream_t st);
int launch_ltaev_post_fwd(const float* m, const MeanRstd* mr, const float* bn_g, const float* bn_b, const float* gamma, const float* beta,
                          const unsigned char* keep_mask, unsigned long long seed, unsigned long long offset, float drop_p, float* v, int B,
                          float eps, cudaStream_t st);
int launch_ltaev_post_bwd(const float* m, const MeanRstd* mr, const float* bn_g, const float* bn_b, const float* gamma,
                          const unsigned char* keep_mask, unsigned long long seed, unsigned long long offset, float drop_p, const float* dv,
                          float* dyb, double* bstats, float* dgamma, float* dbeta, int B, float eps, cudaStream_t st);
int launch_normbwd_apply(const float* dy, const float* x, const BCoef* bc, float* dx, int N, int rows_per_frame, cudaStream_t st);
int launch_colsum(const float* x, float* out, size_t rows, int C, cudaStream_t st);
int launch_upsample_adjoint128(const float* dfull, float* dlow, int N, int H, int W, cudaStream_t st);
int launch_ltaev_final_bwd(const float* pooled, const float* Ap, const float* attn, const float* dattn, const float* dxn,
                           const float* gamma, float* dpooled, float* dAp, float* de, float* dgamma, float* dbeta, int B, int T, float eps,
                           cudaStream_t st);
int launch_slice2d(const float* src, float* dst, int rows, int cols, int ld, int c0, int transpose, cudaStream_t st);

// head_loss.cu
int launch_head_fwd(const float* dec, const float* w, const float* bias, float* out, int B, int O, int P, float scale_by,
                    int mean_sigmoid, float var_eps, cudaStream_t st);
int launch_head_bwd(const float* dout, const float* out, const float* dec, const float* w, float* ddec, float* dw, float* db,
                    int B, int O, int P, float scale_by, int mean_sigmoid, float var_eps, int num_sms, cudaStream_t st);
int launch_mgnll(const float* pred, long long pred_sb, const float* target, long long targ_sb, const float* var,
                 long long var_sb, int var_ch, float* dpred, float* dvar, double* acc, int* neg_flag, float* loss, int B,
                 int P, float eps, cudaStream_t st);
int launch_gnll(const float* pred, long long pred_sb, const float* target, long long targ_sb, const float* var, long long var_sb,
                float* dpred, float* dvar, float* var_out, double* acc, int* neg_flag, float* loss, int B, int P, float eps, int full,
                cudaStream_t st);
int launch_mgnll_none(const float* pred, long long pred_sb, const float* target, long long targ_sb, const float* var, long long var_sb,
                      int var_ch, const float* g, float* loss, float* dpred, float* dvar, int* neg_flag, int B, int P, float eps,
                      cudaStream_t st);
int launch_gnll_none(const float* pred, long long pred_sb, const float* target, long long targ_sb, const float* var, long long var_sb,
                     const float* g, float* loss, float* var_out, float* dpred, float* dvar, int* neg_flag, int B, int P, float eps,
                     int full, cudaStream_t st);
int launch_scale_by_scalar(const float* in, const float* g, float* out, size_t n, cudaStream_t st);
int launch_covariance(const float* var, long long var_sb, int var_ch, float* cov, int B, int P, float eps, cudaStream_t st);

// metrics.cu
int launch_img_metrics(const float* target, const float* pred, const float* var, double* acc, float* pix, int B, int H, int W,
                       cudaStream_t st);

// optim.cu
int launch_adam_step(float* p, float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                     float weight_decay, float inv_bc1, float inv_sqrt_bc2, float grad_scale, int zero_grad, cudaStream_t st);

int launch_assemble_input(const float* const* src, float* x, int B, int T, int c1, int c2, int P, cudaStream_t st);

}  // namespace ub
