/* libuncrtaints_b200.so -- C ABI of the B200 (sm_100a) UnCRtainTS forward/backward hot path.
 *
 * The reference (PatrickTUM/UnCRtainTS @ 5e1f1b5) is pure Python and has no FFI; the boundary this
 * library sits under is the Python operator surface listed below.  Each entry point names the
 * reference interface (file:line under /root/reference) whose computation it replaces; the Python shim
 * in uncrtaints_b200/ binds these with ctypes and re-exposes the reference's classes unchanged
 * (INTEGRATION.md shows the binding).
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *  - the caller owns every buffer (inputs, outputs, gradients, workspace); the library allocates nothing, keeps no
 *    configuration between calls (every option is an argument or a field of ub200_desc) and never synchronises the
 *    device.  The only process-level state are two diagnostics: the launch counter and the optional profiling
 *    events of ub200_prof_* (bench.py); they are not thread-safe and do not influence any result;
 *  - every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns immediately;
 *  - return value: 0 on success, negative UB200_ERR_* otherwise;
 *  - API tensors are float32, contiguous, in the reference's layouts (NCHW); internal activations in the
 *    workspace are pixel-major.
 */
#ifndef UNCRTAINTS_B200_H
#define UNCRTAINTS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UB200_OK 0
#define UB200_ERR_ARG -1        /* unsupported shape / configuration */
#define UB200_ERR_CUDA -2       /* a CUDA launch or runtime call failed */
#define UB200_ERR_WORKSPACE -3  /* workspace too small */

/* Parameter table.  `params` / `grads` arguments are arrays of device pointers in this order; entries that
 * do not exist for the configuration (BatchNorm running statistics of a GroupNorm layer, and the
 * corresponding gradient slots) are NULL.  Names are the reference's state-dict keys. */
enum {
    UB200_P_IN_W = 0,      /* in_conv.conv.conv.0.weight [128][C_in]                 (utae.py:478)  */
    UB200_P_IN_B,          /* in_conv.conv.conv.0.bias   [128]                                      */
    UB200_P_IN_NORM_W,     /* in_conv.conv.conv.1.weight [128]                       (utae.py:470)  */
    UB200_P_IN_NORM_B,     /* in_conv.conv.conv.1.bias   [128]                                      */
    UB200_P_IN_NORM_RM,    /* running_mean (BatchNorm encoder only)                                 */
    UB200_P_IN_NORM_RV,    /* running_var                                                           */
    UB200_P_LTAE_AP,       /* folded L-TAE score matrix Ap [16][128]   (uncrtaints_b200/backbone.py: fold_ltae) */
    UB200_P_LTAE_E,        /* folded additive score term e [B][T][16]                               */
    UB200_P_OUT_W,         /* out_conv.conv.conv.0.weight [out_dim][128]             (uncrtaints.py:381) */
    UB200_P_OUT_B,         /* out_conv.conv.conv.0.bias   [out_dim]                                 */
    /* use_v only (full LTAE2d value path + include_v; ltae.py:10-141, uncrtaints.py:324-338,414-417) */
    UB200_P_LTAE_GN_W,     /* temporal_encoder.in_norm.weight [128]   (value path; the score path has it folded into Ap) */
    UB200_P_LTAE_GN_B,     /* temporal_encoder.in_norm.bias   [128]                                 */
    UB200_P_LTAE_WIN,      /* temporal_encoder.inconv.weight [256][128]                  (ltae.py:48) */
    UB200_P_LTAE_BIN,      /* temporal_encoder.inconv.bias   [256]                                  */
    UB200_P_LTAE_PE,       /* positional table [B][T][256] (positional_encoding.py:16-31; no gradient; NULL = none) */
    UB200_P_LTAE_MLP_W,    /* temporal_encoder.mlp.0.weight [128][256]                   (ltae.py:78) */
    UB200_P_LTAE_MLP_B,    /* temporal_encoder.mlp.0.bias   [128]                                   */
    UB200_P_LTAE_BN_W,     /* temporal_encoder.mlp.1.* BatchNorm1d(128)                  (ltae.py:79) */
    UB200_P_LTAE_BN_B,
    UB200_P_LTAE_BN_RM,
    UB200_P_LTAE_BN_RV,
    UB200_P_LTAE_ON_W,     /* temporal_encoder.out_norm.* GroupNorm(16, 128)             (ltae.py:68-71) */
    UB200_P_LTAE_ON_B,
    UB200_P_INCV_W,        /* include_v.weight [128][256]: columns 0..127 act on the aggregated features, 128..255 on up(v) (uncrtaints.py:338,417) */
    UB200_P_INCV_B,        /* include_v.bias   [128]                                                */
    UB200_P_BLOCK0         /* first MBConv block (in_block.0); block i starts at UB200_P_BLOCK0 + i*UB200_BLOCK_STRIDE;
                              blocks 1..n_dec_blocks are out_block.0 .. out_block.(n-1)              */
};
/* Per-MBConv-block parameter offsets (uncrtaints.py:100-146; X = in_block.0 / out_block.i). */
enum {
    UB200_B_N0_W = 0,      /* X.conv.norm.weight [128]    PreNorm (uncrtaints.py:72-79)              */
    UB200_B_N0_B,          /* X.conv.norm.bias                                                       */
    UB200_B_N0_RM,         /* X.conv.norm.running_mean (BatchNorm only, updated in training)         */
    UB200_B_N0_RV,
    UB200_B_W1,            /* X.conv.fn.0.weight [256][128]   1x1 expand (uncrtaints.py:126)         */
    UB200_B_N1_W,          /* X.conv.fn.1.* [256]                                                    */
    UB200_B_N1_B,
    UB200_B_N1_RM,
    UB200_B_N1_RV,
    UB200_B_WDW,           /* X.conv.fn.3.weight [256][1][3][3] depthwise (uncrtaints.py:130-131)    */
    UB200_B_N2_W,          /* X.conv.fn.4.* [256]                                                    */
    UB200_B_N2_B,
    UB200_B_N2_RM,
    UB200_B_N2_RV,
    UB200_B_F1,            /* X.conv.fn.6.fc.0.weight [32][256]  SE (uncrtaints.py:82-97)            */
    UB200_B_F2,            /* X.conv.fn.6.fc.2.weight [256][32]                                      */
    UB200_B_W2,            /* X.conv.fn.7.weight [128][256]   1x1 project (uncrtaints.py:136)        */
    UB200_B_N3_W,          /* X.conv.fn.8.* [128]                                                    */
    UB200_B_N3_B,
    UB200_B_N3_RM,
    UB200_B_N3_RV,
    UB200_BLOCK_STRIDE
};

/* Per-block parameter offsets with block_type = 'residual' (ResidualConvBlock, uncrtaints.py:24-69; X = in_block.0 / out_block.i): the
 * same UB200_BLOCK_STRIDE slots hold three ConvLayers l = 0..2 (X.conv1, X.conv2, X.conv3) at l*UB200_R_STRIDE + ... */
enum {
    UB200_R_W = 0,         /* X.conv<l+1>.conv.0.weight [128][128][3][3]  3x3, reflect padding (utae.py:478-487) */
    UB200_R_B,             /* X.conv<l+1>.conv.0.bias   [128]                                                   */
    UB200_R_N_W,           /* X.conv<l+1>.conv.1.*      [128]  norm (BatchNorm2d / GroupNorm(4))                */
    UB200_R_N_B,
    UB200_R_N_RM,
    UB200_R_N_RV,
    UB200_R_STRIDE
};

typedef struct ub200_desc {
    int B, T, C_in, H, W;       /* input [B][T][C_in][H][W]; H, W multiples of 32; T <= 64 (T <= 8: register-resident temporal kernels); C_in <= 16 */
    int n_dec_blocks;           /* len(decoder_widths), default 5 (uncrtaints.py:236); one encoder block */
    int out_dim;                /* 13 + covar_dim: 26 (diag/uni), 14 (iso), 13 (none) (uncrtaints.py:357-368) */
    int enc_groups;             /* encoder_norm: 4 = GroupNorm(4) (default), 0 = BatchNorm2d */
    int dec_groups;             /* decoder_norm: 0 = BatchNorm2d (default), 4 = GroupNorm(4) */
    int training;               /* nn.Module.training: BatchNorm batch statistics + running update, dropout on */
    int need_grad;              /* keep what ub200_backward needs */
    int mean_sigmoid;           /* out_nonlin_mean (uncrtaints.py:384) */
    int gemm_backend;           /* 3 (default) = tcgen05 tensor-core path: forward GEMMs with fp16 hi/lo operands (three MMAs, 2^-22), input- and
                                   weight-gradient GEMMs with bf16 hi/lo operands, the expand convolution's two backward GEMMs fused into one kernel;
                                   0 = fp32 CUDA-core GEMMs (test comparator).  Flags on top of 3: +4 single-pass bf16 MMAs (reduced precision),
                                   +32 the 256-channel hidden tensors h1, h2, du, dz1 stored as bf16 (3 + 4 + 32 = 39: BASELINE config #3) */
    float scale_by;             /* uncrtaints.py:250,384 */
    float var_eps;              /* 1e-9 if scale_by == 1 else 1e-3 (uncrtaints.py:374) */
    float pad_value;            /* uncrtaints.py:245,392 */
    float norm_eps;             /* 1e-5 */
    float bn_momentum;          /* 0.1 */
    float dropout_p;            /* 0.1 in training (uncrtaints.py:154), forced to 0 when !training */
    unsigned long long seed;    /* Philox seed / offset for the attention dropout (ignored when keep_mask != NULL) */
    unsigned long long offset;
    int block_type;             /* 0 = 'mbconv' (default), 1 = 'residual' (uncrtaints.py:253,291-294,324-327; tcgen05 path, fp32 storage only) */
    int use_v;                  /* use_v (uncrtaints.py:252,300-314,414-417): full LTAE2d with a value output, include_v 1x1 convolution */
    float v_dropout_p;          /* nn.Dropout on the MLP-processed values (ltae.py:17,85,129: 0.2), training only */
    int is_mono;                /* is_mono (uncrtaints.py:253,296,418; `--pretrain`): T == 1, no temporal encoder / aggregator -- the encoder
                                   output feeds the decoder; UB200_P_LTAE_* are unused */
} ub200_desc;

int ub200_version(void);

/* Number of CUDA kernels this library has launched so far in this process (bench.py's gpu_launches). */
unsigned long long ub200_launch_count(void);

/* Optional per-kernel timing used by bench.py for the roofline line: while bit k of `mask` is set, every launch of
 * kernel class k inside ub200_forward / ub200_backward is bracketed by CUDA events on the launch stream.
 * ub200_prof_enable resets the records; ub200_prof_read synchronises the recorded events (host-blocking) and returns
 * the summed device time and the number of launches of class `kid`. */
int ub200_prof_enable(unsigned long long mask);
int ub200_prof_num_kernels(void);
const char* ub200_prof_kernel_name(int kid);
int ub200_prof_read(int kid, double* total_ms, int* launches);

/* The 1x1 expand GEMM of an MBConv block alone (uncrtaints.py:126 with the PreNorm apply fused in front and the
 * Norm1 statistics behind): h1[N*P][256] = (x[N*P][128]*scale + shift) . W1^T; stats[N][256][2] = column (sum, sumsq).
 * coef: [N][128] (scale, shift); scratch: 256 KB device memory; backend as in ub200_desc.gemm_backend. */
int ub200_gemm1_forward(int backend, const float* x, const float* coef, const float* w1, float* h1, double* stats, int N, int P,
                        void* scratch, void* stream);

/* Number of pointer slots in a params / grads table for this configuration. */
int ub200_num_param_slots(const ub200_desc* d);

/* Bytes of workspace ub200_forward / ub200_backward need.  The same workspace must be passed, untouched,
 * from a forward call to its backward call (it holds the saved activations). */
size_t ub200_workspace_bytes(const ub200_desc* d);

/* Locate a named intermediate inside the workspace (tests / debugging): "x0", "pooled", "pool_idx", "attn",
 * "agg", "notpad", "blk<i>.h1|h2|y|out" (residual blocks: "blk<i>.out|c1|c2|c3" and the (scale, shift) pairs "blk<i>.k1|k2|k3"), and with use_v "v.m" (MLP output [B*1024][128]), "v.mr" ((mean, rstd) of its BatchNorm1d, [B][128][2])
 * and "v.v" (value output [B*1024][128]).  Returns 0 and fills offset/bytes, or UB200_ERR_ARG. */
int ub200_workspace_tap(const ub200_desc* d, const char* name, size_t* offset, size_t* bytes);

/* UNCRTAINTS.forward (model/src/backbones/uncrtaints.py:391-446).
 *   input  [B][T][C_in][H][W], output [B][1][out_dim][H][W];
 *   keep_mask: optional explicit Bernoulli keep mask of the attention dropout, uint8 [16][B][T][H][W]
 *   (uncrtaints.py:202); NULL -> in-kernel Philox(seed, offset).
 * BatchNorm running_mean / running_var in `params` are updated in place when d->training. */
int ub200_forward(const ub200_desc* d, const float* input, const void* const* params, const unsigned char* keep_mask,
                  float* output, void* workspace, size_t workspace_bytes, void* stream);
/* Same with an explicit keep mask for the value dropout of the use_v path as well (uint8 [B*1024][128], ltae.py:129; tests). */
int ub200_forward_v(const ub200_desc* d, const float* input, const void* const* params, const unsigned char* keep_mask,
                    const unsigned char* v_keep_mask, float* output, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of UNCRTAINTS.forward (autograd of the reference, triggered at base_model.py:77).
 *   grad_output [B][1][out_dim][H][W]; output = the tensor ub200_forward produced;
 *   grads: table like params; every non-NULL slot is ACCUMULATED into (zero it first for plain gradients).
 * The input has no gradient (in_conv is the first layer, the reference never asks for it). */
int ub200_backward(const ub200_desc* d, const float* input, const void* const* params, const unsigned char* keep_mask,
                   const float* output, const float* grad_output, void* const* grads, void* workspace,
                   size_t workspace_bytes, void* stream);
int ub200_backward_v(const ub200_desc* d, const float* input, const void* const* params, const unsigned char* keep_mask,
                     const unsigned char* v_keep_mask, const float* output, const float* grad_output, void* const* grads, void* workspace,
                     size_t workspace_bytes, void* stream);

/* MultiGaussianNLLLoss.forward -> multi_gaussian_nll_loss (model/src/losses.py:353,149-218), reduction='mean'.
 *   pred/target/var: base pointers of [B][1][13 | var_ch][H][W] tensors whose batch stride (in elements) is
 *   *_sb and whose [c][h][w] block is contiguous (slices of the network output qualify); var_ch = 13 (diag)
 *   or 1 (iso).  loss: device scalar.  dpred [B][13][P] / dvar [B][var_ch][P]: d loss / d input (NULL to skip).
 *   neg_flag: device int, set to 1 if any var < 0 (the reference raises ValueError, losses.py:199-200).
 *   scratch: 16 bytes of device memory. */
int ub200_mgnll_forward(const float* pred, long long pred_sb, const float* target, long long targ_sb, const float* var,
                        long long var_sb, int var_ch, int B, int P, float eps, float* loss, float* dpred, float* dvar,
                        int* neg_flag, void* scratch, void* stream);

/* GaussianNLLLoss.forward -> gaussian_nll_loss (model/src/losses.py:46-128,222-284; `--loss GNLL`, covmode 'uni'), reduction='mean':
 *   loss = mean over all B*13*P elements of 1/2 (log v + (pred - target)^2 / v) (+ 1/2 log(2 pi) if full), v = max(var, eps) with
 *   identity gradient.  pred/target/var as in ub200_mgnll_forward with 13 variance planes; dpred/dvar [B][13][P] (NULL to skip);
 *   var_out [B][13][P] = the clamped variance, the reference's second return value (NULL to skip); scratch: 16 bytes. */
int ub200_gnll_forward(const float* pred, long long pred_sb, const float* target, long long targ_sb, const float* var,
                       long long var_sb, int B, int P, float eps, int full, float* loss, float* dpred, float* dvar, float* var_out,
                       int* neg_flag, void* scratch, void* stream);

/* reduction='none' of the two losses (losses.py:213-218, :122-128).  grad_loss == NULL: forward, writes `loss` (MGNLL: [P][B],
 * the nested vmap's output order [H][W][B]; GNLL: [B][13][P]) and neg_flag (and GNLL's clamped variance var_out, optional).
 * grad_loss != NULL (same shape as loss): backward, writes dpred [B][13][P] and dvar [B][var_ch | 13][P].
 * reduction='sum' needs no entry point: it is B*P (GNLL: B*13*P) times the 'mean' value. */
int ub200_mgnll_none(const float* pred, long long pred_sb, const float* target, long long targ_sb, const float* var, long long var_sb,
                     int var_ch, int B, int P, float eps, const float* grad_loss, float* loss, float* dpred, float* dvar, int* neg_flag,
                     void* stream);
int ub200_gnll_none(const float* pred, long long pred_sb, const float* target, long long targ_sb, const float* var, long long var_sb,
                    int B, int P, float eps, int full, const float* grad_loss, float* loss, float* var_out, float* dpred, float* dvar,
                    int* neg_flag, void* stream);

/* out[i] = in[i] * grad_loss[0]  (chain rule with the upstream gradient of the scalar loss). */
int ub200_scale_by_scalar(const float* in, const float* grad_loss, float* out, size_t n, void* stream);

/* Second return value of the loss: diag_embed(max(var, eps)) -> [B][1][13][13][H][W] (losses.py:145,211). */
int ub200_covariance(const float* var, long long var_sb, int var_ch, int B, int P, float eps, float* cov, void* stream);

/* One Adam step over flat fp32 buffers of n elements (torch.optim.Adam defaults: no amsgrad, L2 weight decay folded into the
 * gradient; replaces optimizer_G.step() + optimizer_G.zero_grad() of base_model.py:120-122 with ONE kernel):
 *   g' = grads*grad_scale (+ weight_decay*p);  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;
 *   p -= lr/(1-b1^step) * m / (sqrt(v)/sqrt(1-b2^step) + eps);   grads = 0 if zero_grad.
 * step >= 1 is the number of this update; ExponentialLR (base_model.py:51) is the caller changing lr. */
int ub200_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, size_t n, int step, float lr, float beta1,
                    float beta2, float eps, float weight_decay, float grad_scale, int zero_grad, void* stream);

/* Network input from the per-time-point tensors of a collated batch (prepare_data_multi, model/train_reconstruct.py:161-179:
 * torch.stack x2 + torch.cat) in one pass: x [B][T][c_s1 + c_s2][P] from src_table, a DEVICE array of 2T device pointers --
 * entries 0..T-1 the S1 tensors [B][c_s1][P] (ignored when c_s1 == 0), entries T..2T-1 the S2 tensors [B][c_s2][P].  P % 4 == 0. */
int ub200_assemble_input(const void* const* src_table, float* x, int B, int T, int c_s1, int c_s2, int P, void* stream);

/* img_metrics (model/src/learning/metrics.py:20-57) + SSIM (util/pytorch_ssim/__init__.py:17-37) for a batch of [B][13][H][W] images
 * in two kernels.  acc: [B][16] doubles (zeroed by the call): 0 sum e^2, 1 sum |e|, 2 sum of per-pixel spectral angles (deg),
 * 3..6 NaN-excluding sums of e^2, |e|, e and their count, 7..8 NaN-excluding sum of var and its count, 9 sum of the SSIM map.
 * var (optional) [B][13][H][W]; pixelwise (optional) [B][4][H*W]: band-nanmean of error, |error|, error^2, var per pixel. */
int ub200_img_metrics(const float* target, const float* pred, const float* var, int B, int H, int W, double* acc, float* pixelwise,
                      void* stream);

/* out_conv + head alone (uncrtaints.py:381,432-446; tests and the calibration sweep of BASELINE config #5):
 *   dec [B][P][128] pixel-major decoder output, w [out_dim][128], bias [out_dim] -> out [B][out_dim][P] with
 *   mean = scale_by*sigmoid(.) (or identity) on the first 13 planes and softplus(.)+var_eps on the rest.
 * Backward: ddec [B][P][128] is overwritten, dw / db are ACCUMULATED into.  P must be a multiple of 128. */
int ub200_head_forward(const float* dec, const float* w, const float* bias, float* out, int B, int out_dim, int P, float scale_by,
                       int mean_sigmoid, float var_eps, void* stream);
int ub200_head_backward(const float* grad_out, const float* out, const float* dec, const float* w, float* ddec, float* dw, float* db,
                        int B, int out_dim, int P, float scale_by, int mean_sigmoid, float var_eps, void* stream);

/* One MBConv block on pixel-major tensors (tests): x, out, dout, dx are [N][H*W][128].
 * block_params / block_grads: UB200_BLOCK_STRIDE pointers.  groups: 4 = GroupNorm(4), 0 = BatchNorm2d. */
size_t ub200_mbconv_workspace_bytes(int N, int H, int W);
int ub200_mbconv_forward(const float* x, const void* const* block_params, int N, int H, int W, int groups, int training,
                         float eps, float momentum, int gemm_backend, float* out, void* workspace, size_t workspace_bytes,
                         void* stream);
int ub200_mbconv_backward(const float* x, const void* const* block_params, const float* dout, void* const* block_grads,
                          int N, int H, int W, int groups, int training, int gemm_backend, float* dx, void* workspace,
                          size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNCRTAINTS_B200_H */
