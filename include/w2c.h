/*
 * w2c.h — C ABI of the B200-native When2com forward hot path (libw2c.so).
 *
 * The reference (GT-RIPL/MultiAgentPerception) has no FFI layer: its hot path is torch.nn modules called from
 * ptsemseg/models/agent.py.  Every entry point below replaces one group of reference module calls; the
 * reference file:line each one stands in for is cited on the declaration.  The Python host side
 * (multiagentperception_b200/) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - plain pointers and sizes only; all data pointers are DEVICE pointers unless the name ends in _host;
 *  - the caller owns every buffer (no allocation inside, no hidden global state);
 *  - every call enqueues work on the given CUDA stream and returns without synchronising (graph-capturable);
 *  - return value: 0 = W2C_OK, negative = error; w2c_last_error() returns a pointer to a thread-local message
 *    buffer owned by the library (valid until the next failing call on the same thread) - the reference has no FFI
 *    error convention to mirror; SURVEY 8(b) sketched a (char*, size_t) copy-out form, the pointer form was kept
 *    because ctypes reads it without a scratch buffer;
 *  - re-entrant across devices and streams: the only library-side state is per-device, immutable after first use
 *    (kernel attribute opt-ins and the SM count, keyed by the device current at the call - see csrc/common.cuh);
 *  - activations are NHWC.  "act" selects the storage:  W2C_ACT_BF16 = one bf16 plane per pixel;
 *    W2C_ACT_BF16X2 = two bf16 planes per pixel [hi(C) | lo(C)] with value = hi + lo (the "bf16x3" parity
 *    precision: every product is evaluated as hi*hi + hi*lo + lo*hi on the tensor cores, fp32 accumulate).
 */
#ifndef W2C_H_
#define W2C_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* w2c_stream_t; /* a cudaStream_t / CUstream */

enum {
  W2C_OK = 0,
  W2C_ERR_INVALID = -1,     /* bad argument (null pointer, shape not supported by the layout rules) */
  W2C_ERR_UNSUPPORTED = -2, /* valid request the library has no kernel for */
  W2C_ERR_CUDA = -3,        /* CUDA runtime error while launching */
  W2C_ERR_DRIVER = -4       /* driver entry point (tensor-map encode) unavailable */
};

/* W2C_ACT_FP16: one IEEE half plane per pixel (weights packed as half too): same speed and bytes as W2C_ACT_BF16 with
 * three more mantissa bits - logits within ~2e-3 of the fp32 reference instead of ~2e-2 - for activations that stay
 * below 65504 (BatchNorm-ed feature maps do). */
/* W2C_ACT_FP16X2: two IEEE-half planes per pixel [hi(C) | lo(C)], value = hi + lo (22 significant bits; lo may be
 * subnormal for |value| < 2^-3, absolute floor 2^-25).  Same three-pass product as BF16X2 by default, but a layer
 * may run ONE pass (w2c_conv_args.passes = 1: hi*hi only, i.e. plain fp16 arithmetic on the rounded operands, at a
 * third of the MMA work) while still writing both planes - the storage the "mixed" precision plan is built on. */
enum { W2C_ACT_BF16 = 0, W2C_ACT_BF16X2 = 1, W2C_ACT_FP16 = 2, W2C_ACT_FP16X2 = 3 };
enum { W2C_OUT_NHWC = 0, W2C_OUT_NCHW_F32 = 1 };
/* W2C_IMPL_TCGEN05 (the product setting) lets the library pick between its two tensor-core kernels: the persistent
 * warp-specialised kernel wherever a layer has at least one tile per SM, the one-tile-per-CTA kernel for the
 * sub-wave layers; _TC_PERSIST / _TC_TAPS force one (tests); _SIMT is the CUDA-core cross-check (tests only).
 * Value 3 is retired (the halo-tile experiment lives in experiments/csrc, outside the product build). */
enum { W2C_IMPL_TCGEN05 = 0, W2C_IMPL_SIMT = 1, W2C_IMPL_TC_TAPS = 2, W2C_IMPL_TC_PERSIST = 4 };
enum {
  W2C_CONV3X3_S1 = 0,   /* Conv2d k3 s1 p1                       */
  W2C_CONV3X3_S2 = 1,   /* Conv2d k3 s2 p1   (H, W even)         */
  W2C_DECONV3X3_S2 = 2, /* ConvTranspose2d k3 s2 p1 output_padding 1 */
  W2C_CONV1X1_S1 = 3,   /* Conv2d k1 s1 p0                       */
  W2C_CONV1X1_S2 = 4    /* Conv2d k1 s2 p0   (resnet downsample) */
};

/* Library / build identification. */
int w2c_version(void);
const char* w2c_last_error(void);
/* Number of kernel launches this process has enqueued through the library (all entry points). */
uint64_t w2c_launch_count(void);

/*
 * Fused  conv (or transposed conv) -> per-channel affine (folded eval-mode BatchNorm + conv bias) -> [+residual]
 * -> [ReLU].   Replaces conv2DBatchNormRelu.forward (ptsemseg/models/utils.py:87-120) and
 * deconv2DBatchNormRelu.forward (utils.py:148-168) as used by n_segnet_encoder/decoder (backbone.py:41-55,
 * 126-140), img_encoder.squeezer (agent.py:54-60), policy_net4 (agent.py:126-142); with residual it is the
 * BasicBlock tail of the resnet18 trunk used by resnet_encoder (backbone.py:72-96), and with scale = 1,
 * shift = bias it is the plain Conv2d(+ReLU) of simple_decoder.pred (backbone.py:150-154).
 *
 *   y[n, oh, ow, co] = act( scale[co] * sum_{tap,ci} x[n, ih, iw, ci] * w[co, tap, ci] + shift[co] (+ res) )
 *
 * x        NHWC, n x h_in x w_in pixels, each pixel x_cstride channels per plane; the conv reads channels
 *          [x_coffset, x_coffset + cin).  cin must be a multiple of 64.
 * w        packed by w2c_pack_conv_weight: bf16 [planes][cout_pad][ntaps*cin], k = tap*cin + ci.
 * scale, shift   fp32 [cout].
 * residual NHWC like y (same act, pixel stride y_cstride, offset y_coffset) or NULL.
 * y        W2C_OUT_NHWC: NHWC in the same act storage, pixel stride y_cstride, written at channel y_coffset;
 *          W2C_OUT_NCHW_F32: fp32 [n][cout][h_out][w_out] (the logits layout the reference returns).
 */
typedef struct w2c_conv_args {
  const void* x;
  const void* w;
  const float* scale;
  const float* shift;
  const void* residual;
  void* y;
  int32_t n, h_in, w_in;
  int32_t cin, cout;
  int32_t x_cstride, x_coffset;
  int32_t y_cstride, y_coffset;
  int32_t kind;     /* W2C_CONV3X3_S1 ... */
  int32_t relu;     /* 0 / 1 */
  int32_t act;      /* W2C_ACT_* (storage of x, residual, and of y when out_fmt = NHWC) */
  int32_t out_fmt;  /* W2C_OUT_* */
  int32_t impl;     /* W2C_IMPL_TCGEN05 (product) ... W2C_IMPL_SIMT (on-GPU cross-check, tests only) */
  int32_t block_n;  /* 0 = auto; else 16/32/64/128/256 */
  /* W2C_OUT_NCHW_F32 only, cout <= 32: also write labels[n][h_out][w_out] = argmax_co y (first maximal index; the
   * `outputs.data.max(1)[1]` of Trainer_MIMOcom.evaluate, trainer.py:804) from the same accumulators.  With labels
   * set, y may be NULL (label map only: the eval loop never reads the logits).  NULL = no label map. */
  uint8_t* labels;
  /* MMA passes over the operand planes of a two-plane act: 0 = the format's default (3: hi*hi + hi*lo + lo*hi),
   * 1 = hi*hi only.  Ignored (1) for one-plane formats.  The output is written in `act` either way. */
  int32_t passes;
  /* Train-mode BatchNorm statistics from the conv epilogue (NULL = off): fp64 [2 * cout], += the per-channel sum and
   * sum of squares of the output AS STORED (the raw conv output z of w2c_bn_train_fwd), so the separate statistics pass
   * over z is not needed (w2c_bn_train_from_sums_fwd).  Only where w2c_conv_fuses_bn_sums() says so. */
  double* bn_sums;
} w2c_conv_args;

int w2c_conv_bnrelu_fwd(const w2c_conv_args* args, w2c_stream_t stream);
/* 1 when w2c_conv_bnrelu_fwd would accumulate args->bn_sums for this launch (persistent kernel, NHWC output through the
 * TMA-store epilogue: cout % 64 == 0, a one-plane act), 0 when it would refuse it; < 0 on invalid arguments. */
int w2c_conv_fuses_bn_sums(const w2c_conv_args* args);

/*
 * Fused head of n_segnet_encoder: conv1 (3 -> 64, k3 s1) + BN + ReLU followed by conv2 (64 -> 64, k3 s2) + BN + ReLU
 * in ONE kernel, including the divide_inputs / cat regrouping of the views.  Replaces the first two layers of
 * n_segnet_encoder.forward (ptsemseg/models/backbone.py:19-20,42-43) under img_encoder.forward (agent.py:56-60) on
 * the input regrouped per agent.py:1088-1108, i.e. w2c_stem_conv3x3_fwd + w2c_conv_bnrelu_fwd(W2C_CONV3X3_S2) without
 * the 64-channel full-resolution map ever reaching HBM (csrc/enc_head.cu).  Results are bit-identical to that
 * pair of calls in the one-plane formats.
 *
 *   x        x_u8 = 0: the fp32 views (B, c_total, H, W); agents' channels [c_first, c_first + 3*n_agents)
 *            x_u8 = 1: the loader's raw frames, uint8 RGB (B, c_total = agents_total, H, W, 3), agents
 *                      [c_first, c_first + n_agents), mapped through lut (fp32 [3][256], see w2c_stem_conv3x3_u8_fwd)
 *   w1       conv1.weight fp32 [64][27]; scale1 / shift1: its folded BatchNorm (w2c_fold_bn)
 *   w2       conv2 weight packed by w2c_pack_conv_weight(cout 64, cin 64, 9 taps, act); scale2 / shift2 likewise
 *   y        NHWC (B*n_agents, H/2, W/2, y_cstride) in `act`, written at channel y_coffset; images agent-major
 *   act      W2C_ACT_BF16 / W2C_ACT_FP16, or a two-plane format: the conv1 map inside the kernel is ONE plane of the
 *            format's element type and both convs run one MMA pass (the "mixed" precision's one-pass layers); the
 *            output is written in both planes.  H and W must be even.
 */
typedef struct w2c_enc_head_args {
  const void* x;
  const float* lut;
  const float* w1;
  const float* scale1;
  const float* shift1;
  const void* w2;
  const float* scale2;
  const float* shift2;
  void* y;
  int32_t x_u8;
  int32_t b, n_agents, c_total, c_first;
  int32_t h, w;
  int32_t act;
  int32_t y_cstride, y_coffset;
} w2c_enc_head_args;

int w2c_enc_head_fwd(const w2c_enc_head_args* args, w2c_stream_t stream);

/*
 * Train-mode BatchNorm2d on a conv output, in place.  Replaces nn.BatchNorm2d in training mode inside
 * conv2DBatchNormRelu / deconv2DBatchNormRelu (ptsemseg/models/utils.py:110-114,152-164) as Trainer_*.train() runs
 * them after model.train() (trainer.py:659-669), plus the ReLU (and the BasicBlock residual add) that follows.
 *   z        the raw conv output (w2c_conv_bnrelu_fwd with scale = 1, shift = conv bias, relu = 0), NHWC in `act`,
 *            n_px pixels (N*H*W), channels [coffset, coffset + c) of cstride; overwritten with
 *            act(gamma * (z - mean) / sqrt(var + eps) + beta (+ residual)), mean / var = batch statistics (var biased)
 *   running_mean, running_var, num_batches_tracked: the module's own buffers, updated in place exactly like
 *            nn.BatchNorm2d (momentum, UNBIASED variance, counter + 1); NULL = track_running_stats off
 *   sums_ws  fp64 [2*c], ZEROED by the caller once (the call leaves it zeroed again); scale_ws / shift_ws fp32 [c].
 * c must be a multiple of 8 with c / 8 dividing 256.  Three launches: statistics, finalize, apply.
 *   y_out    NULL: normalise z in place.  Otherwise the result goes to y_out (NHWC in `act`, channels [y_coffset,
 *            y_coffset + c) of y_cstride; a residual is then laid out like y_out) and z keeps the raw conv output -
 *            what the backward pass (w2c_bn_train_bwd) needs.
 *   stats_out  NULL, or fp32 [2*c] receiving mean | 1/sqrt(var + eps) of this batch for the backward pass.
 */
int w2c_bn_train_fwd(void* z, const void* residual, int64_t n_px, int32_t c, int32_t cstride, int32_t coffset,
                     int32_t act, int32_t relu, const float* gamma, const float* beta, float eps, float momentum,
                     float* running_mean, float* running_var, int64_t* num_batches_tracked, double* sums_ws,
                     float* scale_ws, float* shift_ws, void* y_out, int32_t y_cstride, int32_t y_coffset,
                     float* stats_out, w2c_stream_t stream);
/* The same without the statistics pass: sums_ws already holds sum(z) | sum(z^2) of this batch, accumulated by the conv
 * that wrote z (w2c_conv_args.bn_sums).  Two launches: finalize, apply. */
int w2c_bn_train_from_sums_fwd(void* z, const void* residual, int64_t n_px, int32_t c, int32_t cstride, int32_t coffset,
                               int32_t act, int32_t relu, const float* gamma, const float* beta, float eps,
                               float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                               double* sums_ws, float* scale_ws, float* shift_ws, void* y_out, int32_t y_cstride,
                               int32_t y_coffset, float* stats_out, w2c_stream_t stream);
/* The same on an fp32 NCHW map [n][c][hw] (the logits layer: deconv12 is conv + BatchNorm + ReLU, backbone.py:124). */
int w2c_bn_train_nchw_fwd(float* z, int32_t n, int32_t c, int64_t hw, int32_t relu, const float* gamma,
                          const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                          int64_t* num_batches_tracked, double* sums_ws, float* scale_ws, float* shift_ws,
                          float* y_out, float* stats_out, w2c_stream_t stream);
/* First layers without the ReLU (the raw conv output train-mode BatchNorm starts from); arguments as
 * w2c_stem_conv3x3_fwd / w2c_stem_conv7x7s2_fwd. */
int w2c_stem_conv3x3_raw_fwd(const float* x, const float* w, const float* scale, const float* shift, void* y, int32_t b,
                             int32_t n_agents, int32_t c_total, int32_t c_first, int32_t h, int32_t w_px, int32_t cout,
                             int32_t act, int32_t n_split, w2c_stream_t stream);
int w2c_stem_conv7x7s2_raw_fwd(const float* x, const float* w, const float* scale, const float* shift, void* y,
                               int32_t b, int32_t n_agents, int32_t c_total, int32_t c_first, int32_t h, int32_t w_px,
                               int32_t cout, int32_t act, int32_t n_split, w2c_stream_t stream);

/* cout rounded up to the row padding the packed weight layout uses. */
int32_t w2c_cout_pad(int32_t cout);
/* Bytes of the packed weight buffer for a conv of the given geometry. */
size_t w2c_packed_weight_bytes(int32_t cout, int32_t cin, int32_t ntaps, int32_t act);

/*
 * Pack a PyTorch-layout fp32 conv weight into the K-major bf16 layout the conv kernels read.
 *   transposed = 0: w is Conv2d.weight          [cout][cin_real][kh][kw]
 *   transposed = 1: w is ConvTranspose2d.weight [cin_real][cout][kh][kw]
 * cin_real <= cin (extra input channels are zero-filled); ntaps = kh*kw (9 or 1).
 * act = W2C_ACT_BF16X2 additionally writes the low-order plane (w - bf16(w)).
 */
int w2c_pack_conv_weight(const float* w, int32_t cout, int32_t cin_real, int32_t cin, int32_t ntaps,
                         int32_t transposed, int32_t act, void* packed, w2c_stream_t stream);

/*
 * Fold eval-mode BatchNorm2d (+ the conv bias) into the per-channel affine the conv epilogue applies
 * (nn.BatchNorm2d inside cbr_unit / dcbr_unit, utils.py:110-114,152-164):
 *   scale = gamma / sqrt(var + eps);  shift = beta + (bias - mean) * scale.
 * Any of gamma/beta/mean/var may be NULL together (no BN): scale = 1, shift = bias.  bias may be NULL.
 */
int w2c_fold_bn(const float* conv_bias, const float* gamma, const float* beta, const float* mean,
                const float* var, float eps, int32_t cout, float* scale, float* shift, w2c_stream_t stream);

/*
 * First encoder layer: Conv2d(3 -> cout, k3 s1 p1) + BN + ReLU reading the caller's fp32 NCHW batch directly.
 * Replaces divide_inputs + cat (agent.py:1088-1108) and n_segnet_encoder.conv1 (backbone.py:19,42).
 *   x     fp32 [b][c_total][h][w]  (views concatenated on the channel axis, trainer.py:651); agent a reads
 *         channels [c_first + 3a, c_first + 3a + 3)
 *   w     fp32 [cout][27]  (k = ci*9 + kh*3 + kw, i.e. Conv2d.weight flattened), scale/shift fp32 [cout]
 *   y     NHWC [(n_agents*b)][h][w][cout], image index = agent*b + batch ("agent-major", agent.py:1103-1108)
 * cout must be a multiple of 32 and <= 128.  n_split = 2 (cout = 128: two encoders' first layers fused, the image is
 * read once): y receives TWO dense NHWC maps of 64 channels, one after the other, instead of one of 128.
 */
int w2c_stem_conv3x3_fwd(const float* x, const float* w, const float* scale, const float* shift, void* y,
                         int32_t b, int32_t n_agents, int32_t c_total, int32_t c_first, int32_t h, int32_t w_px,
                         int32_t cout, int32_t act, int32_t n_split, w2c_stream_t stream);

/*
 * The same first layer reading the loader's RAW frames: uint8 RGB HWC [b][agents_total][h][w][3] (what
 * airsim_loader.__getitem__ holds after cv2.cvtColor, airsim_loader.py:496), with the loader transform
 * (airsim_loader.py:515-527: RGB->BGR, - mean, / 255, HWC->CHW) and the trainer's channel concat (trainer.py:651)
 * fused into the im2col gather.  lut fp32 [3][256]: lut[ci][v] = float32((float64(v) - mean[ci]) / 255) for BGR channel
 * ci, built by the caller in float64 exactly as the loader computes it, so the result equals w2c_stem_conv3x3_fwd on
 * the transformed fp32 tensor bit for bit.  Agents [agent_first, agent_first + n_agents) are convolved.
 */
int w2c_stem_conv3x3_u8_fwd(const uint8_t* frames, const float* lut, const float* w, const float* scale,
                            const float* shift, void* y, int32_t b, int32_t n_agents, int32_t agents_total,
                            int32_t agent_first, int32_t h, int32_t w_px, int32_t cout, int32_t act, int32_t n_split,
                            w2c_stream_t stream);

/* ---- evaluation-loop glue ----------------------------------------------------------- */
/* labels[n][p] = argmax_c logits[n][c][p] (first maximal index), fp32 NCHW logits -> uint8; trainer.py:804. */
int w2c_argmax_labels_fwd(const float* logits, uint8_t* labels, int32_t n, int32_t c, int64_t hw,
                          w2c_stream_t stream);
/* runningScore._fast_hist accumulated on the device (metrics.py:99-108): hist[n_class*gt + pred] += 1 for every
 * pixel with 0 <= gt < n_class.  hist int64 [n_class][n_class], caller-zeroed; gt uint8 or int64 [count]. */
enum { W2C_GT_U8 = 0, W2C_GT_I64 = 1 };
int w2c_confusion_update(const uint8_t* pred, const void* gt, int32_t gt_dtype, int64_t count, int32_t n_class,
                         int64_t* hist, w2c_stream_t stream);

/* runningScore.update_div (metrics.py:70-97): the same histogram split per IMAGE into the "normal" and the "noisy"
 * matrix (confusion_matrix_pos / _neg).  pred / gt hold n_img images of px_per_img pixels; img_flag uint8 [n_img]:
 * 1 -> hist_pos, 0 -> hist_neg (the caller derives it from commun_label exactly as update_div does: 'mimo'
 * commun_label[:,0,:] == 0 transposed to agent-major, 'when2com' commun_label == -1). */
int w2c_confusion_update_div(const uint8_t* pred, const void* gt, int32_t gt_dtype, const uint8_t* img_flag,
                             int32_t n_img, int64_t px_per_img, int32_t n_class, int64_t* hist_pos, int64_t* hist_neg,
                             w2c_stream_t stream);

/* runningScore.update_selection (metrics.py:23-68) accumulated on the device: counters int64 [3] =
 * {total_agent, correct_when2com, correct_who2com}, caller-zeroed.
 *   mode 0 'mimo':     action int64 [b][n] (forward()'s action_argmax), commun_label int64 [b][2][n]
 *   mode 1 'when2com': action int64 [b] (arg-max link), commun_label int64 [b] in -1 .. n-2
 *   mode 2 'when2com': action fp32 [b][n] (the thresholded weights 'activated' returns), commun_label int64 [b] */
int w2c_selection_update(const void* action, const int64_t* commun_label, int32_t b_sz, int32_t n, int32_t mode,
                         int64_t* counters, w2c_stream_t stream);

/*
 * Key / query heads: flatten -> Linear -> ReLU -> Linear -> ReLU -> Linear.  Replaces km_generator.forward and
 * linear.forward (agent.py:145-178).  feat is the NHWC policy feature map [m][s][s][256]; w0 must already be
 * permuted to NHWC flatten order (the reference flattens NCHW, agent.py:158).
 *   out fp32 [m][out_dim].  ws: fp32 scratch of m*(256+128) floats.
 */
int w2c_kq_mlp_fwd(const void* feat, int32_t act, int32_t m, int32_t n_feat, const float* w0, const float* b0,
                   const float* w1, const float* b1, const float* w2, const float* b2, int32_t out_dim,
                   float* out, float* ws, w2c_stream_t stream);

/*
 * Up to two heads over the SAME feature map in one pair of launches (key_net and query_net both read the policy
 * map, agent.py:1132-1147).  ws: fp32 scratch of n_heads*m*256 floats.  Results are bit-identical to
 * w2c_kq_mlp_fwd called per head.
 */
typedef struct w2c_mlp_head {
  const float* w0; /* [256][n_feat], NHWC flatten order */
  const float* b0;
  const float* w1; /* [128][256] */
  const float* b1;
  const float* w2; /* [out_dim][128] */
  const float* b2;
  float* out;      /* [m][out_dim] */
  int32_t out_dim;
} w2c_mlp_head;
int w2c_kq_mlp_heads_fwd(const void* feat, int32_t act, int32_t m, int32_t n_feat, const w2c_mlp_head* heads,
                         int32_t n_heads, float* ws, w2c_stream_t stream);

/*
 * Communication graph + fusion.  Replaces MIMOGeneralDotProductAttention.forward (agent.py:252-286),
 * GeneralDotProductAttention / ScaledDotProductAttention (agent.py:194-213,345-368), the +0.001*I bias
 * (agent.py:1164-1167), activated_select / argmax_select (agent.py:1036-1078) and agents2batch (1080-1086).
 *
 *   qt[b,j,:] = Wq * query[b,j,:] + bq      (skipped when wq == NULL: qt = query, needs q_dim == k_dim)
 *   S[b,i,j]  = <key[b,i,:], qt[b,j,:]> / temperature
 *   P[b,:,j]  = softmax_i S   (or sparsemax_i when sparse != 0; MIMOcomWho: mask_self removes i == j)
 *   prob_out  = P + diag_bias * I
 *   coef      = P                                   (mode SOFTMAX: fuse with the un-biased P)
 *             = prob_out * [prob_out > thresh]      (mode ACTIVATED)
 *             = onehot_i(argmax_i prob_out)         (mode ARGMAX)
 *   fused[j*b_sz + b] = sum_i coef[b,i,j] * val[i*b_sz + b]          (agent-major images, NHWC)
 *   action[b,j] = argmax_i coef'   (coef' = prob_out for SOFTMAX, coef otherwise), int64
 *   connect[0] += #{(b,i,j): i != j, coef != 0}     (int32 counter; caller zeroes it)
 *
 * keys fp32 [n_k*b_sz][k_dim] and queries fp32 [n_q*b_sz][q_dim] are agent-major like the images.
 * val   NHWC [(n_k*b_sz)][hw][c] in `act` storage;  fused the same with n_q images.
 * prob_out fp32 [b_sz][n_k][n_q];  coef_out fp32 [b_sz][n_k][n_q] (may be NULL);  action int64 [b_sz][n_q].
 * n_k, n_q <= 8.
 */
enum { W2C_FUSE_SOFTMAX = 0, W2C_FUSE_ACTIVATED = 1, W2C_FUSE_ARGMAX = 2 };
typedef struct w2c_attn_args {
  const float* keys;
  const float* queries;
  const float* wq;
  const float* bq;
  const void* val;
  void* fused;
  float* prob_out;
  float* coef_out;
  int64_t* action;
  int32_t* connect;
  int32_t b_sz, n_k, n_q;
  int32_t k_dim, q_dim;
  int32_t hw, c; /* pixels per image, channels per plane */
  int32_t fused_cstride, fused_coffset; /* pixel stride / channel offset of `fused` (concat buffers) */
  int32_t act;
  int32_t mode;      /* W2C_FUSE_* */
  int32_t sparse;    /* 0 softmax, 1 sparsemax */
  int32_t mask_self; /* 1: drop i == j before the softmax (MIMOcomWho, agent.py:306-343) */
  float temperature; /* 1 for the "general" attention, sqrt(128) for ScaledDotProductAttention */
  float diag_bias;   /* 0.001 for MIMOcom, 0 otherwise */
  float thresh;      /* 0.2 */
  /* Agent sharding (one process per GPU): the score matrix is always computed for all n_k x n_q pairs, but only
   * queries [q_first, q_first + q_count) are fused and written, as images 0..q_count*b_sz of `fused`
   * (q_count = 0: all n_q).  keys / queries / val may point into an all-gathered buffer in which every rank
   * contributed agents_per_rank agents: agent i is row-block (i % agents_per_rank) of rank segment
   * (i / agents_per_rank), segments *_rank_stride ELEMENTS apart (agents_per_rank = 0: dense agent-major).  * Alignment: val and fused must be 16-byte aligned (bulk copies / 16-byte stores), wq 16-byte aligned when
 * q_dim is a multiple of 4 (vector loads of the projection rows); checked, W2C_ERR_INVALID otherwise.
 */
  int32_t q_first, q_count;
  int32_t agents_per_rank;
  int64_t keys_rank_stride, queries_rank_stride, val_rank_stride;
} w2c_attn_args;

int w2c_attn_fuse_fwd(const w2c_attn_args* args, w2c_stream_t stream);

/* ---- resnet18 trunk / simple_decoder specifics (backbone.py:58-96,143-164) ------------------------- */

/* Conv2d(3 -> 64, k7 s2 p3, no bias) + BN + ReLU on the fp32 NCHW batch; same input convention as the stem.  Runs on
 * the tensor cores (software im2col, K = 147 + the shift column padded to 160).  cout = 64, or 128 for two encoders'
 * first layers fused (w [128][147]); n_split = 2 then writes two dense 64-channel maps like w2c_stem_conv3x3_fwd. */
int w2c_stem_conv7x7s2_fwd(const float* x, const float* w, const float* scale, const float* shift, void* y,
                           int32_t b, int32_t n_agents, int32_t c_total, int32_t c_first, int32_t h, int32_t w_px,
                           int32_t cout, int32_t act, int32_t n_split, w2c_stream_t stream);
/* The same on the loader's raw uint8 RGB HWC frames (see w2c_stem_conv3x3_u8_fwd). */
int w2c_stem_conv7x7s2_u8_fwd(const uint8_t* frames, const float* lut, const float* w, const float* scale,
                              const float* shift, void* y, int32_t b, int32_t n_agents, int32_t agents_total,
                              int32_t agent_first, int32_t h, int32_t w_px, int32_t cout, int32_t act,
                              int32_t n_split, w2c_stream_t stream);
/* MaxPool2d(k3 s2 p1) on NHWC. */
int w2c_maxpool3x3s2_fwd(const void* x, void* y, int32_t n, int32_t h, int32_t w_px, int32_t c, int32_t act,
                         w2c_stream_t stream);
/* F.interpolate(mode='bilinear', align_corners=False) by an integer factor, fp32 NCHW in -> fp32 NCHW out. */
int w2c_bilinear_up_fwd(const float* x, float* y, int32_t n, int32_t c, int32_t h, int32_t w_px, int32_t factor,
                        w2c_stream_t stream);

/* Indexed copy of image groups between NHWC maps: dst image (g*b + i), channels [dst_coffset, dst_coffset + c)  <-
 * src image (sel[g]*b + i), channels [src_coffset, src_coffset + c), for g < n_groups, i < b.  sel is a DEVICE int32
 * array (NULL = identity): the random-selection baselines (All_agents / MIMO_All_agents with shuffle_features =
 * 'selection', agent.py:447-452,934-947) redraw it per forward without rebuilding the captured program; it is also
 * the concat of `torch.cat(feature maps, 1)` (agent.py:460,970).  Channel counts / strides / offsets % 8 == 0. */
int w2c_gather_images_fwd(const void* src, void* dst, const int32_t* sel, int32_t n_groups, int32_t b, int32_t h,
                          int32_t w_px, int32_t c, int32_t src_cstride, int32_t src_coffset, int32_t dst_cstride,
                          int32_t dst_coffset, int32_t act, w2c_stream_t stream);

/* The same up-sampling fused with the arg-max over the c channels: labels uint8 [n][h*factor][w*factor] only (the
 * label-map output of a simple_decoder model; equals w2c_bilinear_up_fwd + w2c_argmax_labels_fwd bit for bit). */
int w2c_bilinear_argmax_fwd(const float* x, uint8_t* labels, int32_t n, int32_t c, int32_t h, int32_t w_px,
                            int32_t factor, w2c_stream_t stream);

/* ---- backward pass (second half): loss.backward() of Trainer_*.train(), ptsemseg/trainer.py:668-670 ---
 * Gradient maps are NHWC like the activations, in their own storage `act_g` (bf16 / bf16 hi|lo planes: gradients need
 * the fp32 exponent range); `act_f` is the storage of the forward maps they are combined with.  Parameter gradients
 * are fp32 and ACCUMULATED (+=) like autograd's .grad. */

/*
 * Weight gradient of a conv / transposed conv on the tensor cores (csrc/wgrad.cu).  Replaces the weight-gradient half of
 * autograd's conv backward for the Conv2d / ConvTranspose2d of conv2DBatchNormRelu / deconv2DBatchNormRelu
 * (ptsemseg/models/utils.py:87-120,148-168).
 *   x    the layer's forward INPUT map, n x h_in x w_in pixels, channels [x_coffset, x_coffset + cin) of x_cstride
 *   dy   gradient w.r.t. the conv's raw OUTPUT (h_out x w_out as the kind implies), channels [dy_coffset, + cout)
 *   dw   fp32, +=, 16-byte aligned:  Conv2d kinds [cout][ntaps][cin];  W2C_DECONV3X3_S2 [cin][ntaps][cout]   (tap = kh*3 + kw; both are
 *        the parameter's [d0][d1][kh][kw] layout with the last three axes permuted: .view(d0,3,3,d1).permute(0,3,1,2))
 * cin and cout as the MAPS hold them: multiples of 64 (an 11-class logits gradient lives in a 64-channel padded map),
 * and 64 or a multiple of 128 on the small-grid operand (dy for convs, x for the transposed conv).
 * act_x and act_dy must agree in plane count and element type (a mixed f16 x bf16 MMA faults on B200); passes as in
 * w2c_conv_args.
 */
typedef struct w2c_wgrad_args {
  const void* x;
  const void* dy;
  float* dw;
  int32_t n, h_in, w_in;
  int32_t cin, cout;
  int32_t x_cstride, x_coffset;
  int32_t dy_cstride, dy_coffset;
  int32_t kind;
  int32_t act_x, act_dy;
  int32_t passes;
} w2c_wgrad_args;
int w2c_conv_wgrad(const w2c_wgrad_args* args, w2c_stream_t stream);

/* w2c_pack_conv_weight with an optional tap flip (tap -> ntaps-1-tap): the operand of the DATA-gradient convs, which run
 * on w2c_conv_bnrelu_fwd (scale = 1, shift = 0) over dL/dy:
 *   Conv2d k3 s1 [co][ci][k]      -> W2C_CONV3X3_S1,   pack(w, cout=ci, cin=co, transposed=1, flip=1)
 *   Conv2d k3 s2                  -> W2C_DECONV3X3_S2, pack(w, cout=ci, cin=co, transposed=1, flip=0)
 *   ConvTranspose2d k3 s2 [ci][co][k] -> W2C_CONV3X3_S2, pack(w, cout=ci, cin=co, transposed=0, flip=0)
 *   Conv2d k1 (s1, or s2 followed by w2c_upsample_zero2) -> W2C_CONV1X1_S1, pack(w, cout=ci, cin=co, transposed=1) */
int w2c_pack_conv_weight_ex(const float* w, int32_t cout, int32_t cin_real, int32_t cin, int32_t ntaps,
                            int32_t transposed, int32_t flip, int32_t act, void* packed, w2c_stream_t stream);

/*
 * Many w2c_pack_conv_weight_ex / w2c_fold_bn calls in ONE launch each: a training step re-derives every layer's packed
 * operand and folded bias from the live parameters (an optimizer step changes them in place, trainer.py:668-670), 43-47
 * layers x three tiny launches per step otherwise.  items: HOST array of n entries (copied into the kernel parameters,
 * 48 per launch); every entry means exactly what the arguments of the single call mean; results are bit-identical.
 */
typedef struct w2c_pack_item {
  const float* w;
  void* packed;
  int32_t cout, cin_real, cin, ntaps;
  int32_t transposed, flip;
} w2c_pack_item;
int w2c_pack_conv_weights_batch(const w2c_pack_item* items, int32_t n, int32_t act, w2c_stream_t stream);

typedef struct w2c_fold_item {
  const float* conv_bias;
  const float* gamma;
  const float* beta;
  const float* mean;
  const float* var;
  float* scale;
  float* shift;
  float eps;
  int32_t cout;
} w2c_fold_item;
int w2c_fold_bn_batch(const w2c_fold_item* items, int32_t n, w2c_stream_t stream);

/*
 * Backward of train-mode BatchNorm2d (+ residual) (+ ReLU) between the gradient of the unit's output and the gradient
 * of the raw conv output (nn.BatchNorm2d + nn.ReLU of cbr_unit / dcbr_unit, utils.py:110-114,152-164; BasicBlock tail).
 *   du = dy * [y > 0] (relu);  dres = du;  dbeta += sum du;  dgamma += sum du * xhat;
 *   dz = gamma * invstd * (du - mean(du) - xhat * mean(du * xhat)),  xhat = (z - mean) * invstd
 * stats = the fp32 [2c] mean | invstd the train forward saved; stats == NULL: no BatchNorm (dz = du, dbeta = conv-bias
 * gradient).  y is only read when relu != 0, z only with stats.  sums_ws fp64 [2c] zeroed by the caller once (left zeroed),
 * coef_ws fp32 [3c].  dres may be NULL.  Channel strides 0 = c.
 */
typedef struct w2c_bn_bwd_args {
  const void* dy;
  const void* y;
  const void* z;
  void* dz;
  void* dres;
  int64_t n_px;
  int32_t c;
  int32_t dy_cstride, dy_coffset;
  int32_t y_cstride, y_coffset;
  int32_t z_cstride, z_coffset;
  int32_t dz_cstride, dz_coffset;
  int32_t dres_cstride, dres_coffset;
  int32_t act_f, act_g, relu;
  const float* gamma;
  const float* stats;
  float* dgamma;
  float* dbeta;
  double* sums_ws;
  float* coef_ws;
  /* Optional: the scale_ws / shift_ws the train forward left behind.  The ReLU mask is then taken from
   * z * scale + shift > 0 (exactly what the forward evaluated) and y is not read at all: two of the seven map passes
   * less.  Units without a residual only (a residual changes the sign test). */
  const float* fwd_scale;
  const float* fwd_shift;
} w2c_bn_bwd_args;
int w2c_bn_train_bwd(const w2c_bn_bwd_args* args, w2c_stream_t stream);
/* The same for the fp32 NCHW logits layer: dy / y / z fp32 [n][c][hw]; dz is written as an NHWC gradient map of c_pad
 * channels (c .. c_pad-1 zero) at [dz_coffset, dz_coffset + c_pad) of dz_cstride - the input of the data- and
 * weight-gradient convs of that layer. */
int w2c_bn_train_nchw_bwd(const float* dy, const float* y, const float* z, void* dz, int32_t n, int32_t c, int64_t hw,
                          int32_t c_pad, int32_t dz_cstride, int32_t dz_coffset, int32_t act_g, int32_t relu,
                          const float* gamma, const float* stats, float* dgamma, float* dbeta, double* sums_ws,
                          float* coef_ws, w2c_stream_t stream);

/*
 * Backward of w2c_attn_fuse_fwd in its differentiable mode (W2C_FUSE_SOFTMAX, what forward(training=True) runs,
 * agent.py:1170-1179): autograd through MIMOGeneralDotProductAttention.forward (agent.py:252-286) and the
 * single-request attentions (agent.py:194-213,345-368).  prob = the UN-biased probabilities the forward fused with
 * (its coef_out).  Outputs: dval (NHWC act_g, dense, n_k*b_sz images; dval_accumulate != 0 adds to what is there),
 * dkeys fp32 [n_k*b_sz][k_dim], dqueries fp32 [n_q*b_sz][q_dim] (may be NULL), dwq / dbq fp32 += (NULL without a
 * projection).  dp_ws: fp32 [b_sz][n_k][n_q] zeroed by the caller once (left zeroed).  Dense agent-major inputs only.
 */
typedef struct w2c_attn_bwd_args {
  const float* keys;
  const float* queries;
  const float* wq;
  const float* bq;
  const void* val;
  const void* dfused;
  const float* prob;
  void* dval;
  float* dkeys;
  float* dqueries;
  float* dwq;
  float* dbq;
  float* dp_ws;
  int32_t b_sz, n_k, n_q;
  int32_t k_dim, q_dim;
  int32_t hw, c;
  int32_t dfused_cstride, dfused_coffset;
  int32_t act_f, act_g;
  int32_t sparse;
  int32_t dval_accumulate;
  float temperature;
} w2c_attn_bwd_args;
int w2c_attn_fuse_bwd(const w2c_attn_bwd_args* args, w2c_stream_t stream);

/*
 * Backward of w2c_kq_mlp_heads_fwd (km_generator / linear, agent.py:145-178).  heads = the forward's weights (out
 * unused), ws_fwd = the forward's scratch (the first hidden layer per head).  Per head: dout fp32 [m][out_dim] in,
 * parameter gradients += (dw0 in the NHWC flatten order of w0).  dfeat: NHWC act_g gradient of the policy map, the sum
 * over the heads.  ws: fp32 scratch of n_heads*m*512 + 256 floats.
 */
typedef struct w2c_mlp_head_grad {
  const float* dout;
  float* dw0;
  float* db0;
  float* dw1;
  float* db1;
  float* dw2;
  float* db2;
} w2c_mlp_head_grad;
int w2c_kq_mlp_heads_bwd(const void* feat, int32_t act_f, int32_t m, int32_t n_feat, const w2c_mlp_head* heads,
                         const w2c_mlp_head_grad* grads, int32_t n_heads, const float* ws_fwd, void* dfeat,
                         int32_t act_g, float* ws, w2c_stream_t stream);

/* Weight gradient of the 3-channel first layers from the fp32 NCHW views: ksize 3 = Conv2d(3, cout, 3, 1, 1)
 * (n_segnet_encoder.conv1), 7 = Conv2d(3, cout, 7, 2, 3) (resnet18 conv1).  dz: NHWC act_g gradient of the raw conv
 * output, channels [dz_coffset, + cout) of dz_cstride; dw fp32 [cout][3][k][k] +=; cout <= 64; other arguments as
 * w2c_stem_conv3x3_fwd. */
int w2c_stem_conv_wgrad(const float* x, const void* dz, float* dw, int32_t ksize, int32_t b, int32_t n_agents,
                        int32_t c_total, int32_t c_first, int32_t h, int32_t w_px, int32_t cout, int32_t dz_cstride,
                        int32_t dz_coffset, int32_t act_g, w2c_stream_t stream);

/* MaxPool2d(3, 2, 1) backward (the gradient goes to the first maximum of each window, like torch). */
int w2c_maxpool3x3s2_bwd(const void* x, const void* dy, void* dx, int32_t n, int32_t h, int32_t w_px, int32_t c,
                         int32_t act_f, int32_t act_g, w2c_stream_t stream);
/* Adjoint of w2c_bilinear_up_fwd: dy fp32 [n][c][h*factor][w*factor] -> dx fp32 [n][c][h][w]. */
int w2c_bilinear_up_bwd(const float* dy, float* dx, int32_t n, int32_t c, int32_t h, int32_t w_px, int32_t factor,
                        w2c_stream_t stream);
/* dst[n][2i][2j] = src[n][i][j], zero elsewhere (+ add, same layout as dst, may be NULL or dst): completes the data
 * gradient of a stride-2 1x1 conv.  h, w_px: extent of dst.  Dense NHWC maps in `act`. */
int w2c_upsample_zero2(const void* src, const void* add, void* dst, int32_t n, int32_t h, int32_t w_px, int32_t c,
                       int32_t act, w2c_stream_t stream);
/* dst slice = a slice (+ b slice) over n_px pixels of c channels: gradient accumulation / concat-slice extraction. */
int w2c_grad_add(const void* a, int32_t a_cstride, int32_t a_coffset, const void* b, int32_t b_cstride,
                 int32_t b_coffset, void* dst, int32_t d_cstride, int32_t d_coffset, int64_t n_px, int32_t c,
                 int32_t act, w2c_stream_t stream);

/* The trainers' loss with its gradient in one pass: cross_entropy2d (ptsemseg/loss/loss.py:5-18; F.cross_entropy over
 * the pixels of fp32 NCHW logits [n][c][hw], int64 target [n][hw], ignore_index 250, mean over the counted pixels).
 * dlogits receives softmax - onehot per counted pixel (zero for ignored ones), UNSCALED; totals fp64 [2] += {sum of the
 * per-pixel losses, number of counted pixels} (caller-zeroed): loss = totals[0] / totals[1], dL/dlogits = dlogits /
 * totals[1].  c <= 32. */
int w2c_cross_entropy2d(const float* logits, const int64_t* target, int32_t n, int32_t c, int64_t hw,
                        int64_t ignore_index, float* dlogits, double* totals, w2c_stream_t stream);

/* ---- layout helpers --------------------------------------------------------------------------------- */
/* NHWC activation (act storage) -> fp32 NCHW, and back.  Used at module boundaries and by the tests. */
int w2c_nhwc_to_nchw_f32(const void* x, float* y, int32_t n, int32_t h, int32_t w_px, int32_t c, int32_t cstride,
                         int32_t coffset, int32_t act, w2c_stream_t stream);
int w2c_nchw_f32_to_nhwc(const float* x, void* y, int32_t n, int32_t h, int32_t w_px, int32_t c, int32_t cstride,
                         int32_t coffset, int32_t act, w2c_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* W2C_H_ */
