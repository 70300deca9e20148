/*
 * ppyolo_b200.h -- C ABI of the B200-native PP-YOLO hot path (libppyolo_b200.so).
 *
 * Every entry point takes raw device pointers, explicit sizes/strides and a CUDA stream, allocates
 * nothing, keeps no global state besides a launch counter, and returns a ppy_status (0 = ok, <0 = error;
 * no exceptions cross this boundary).  Kernels are enqueued on `stream` and are NOT synchronised.
 *
 * The reference (miemie2013/Pytorch-PPYOLO) has no live C/FFI boundary on this path -- its only native
 * extension, external/DCNv2 (`_ext`, src/vision.cpp:3-8, src/dcn_v2.h:9-144), is dead code
 * (model/custom_layers.py:14-19, :102-105).  The entry points below are therefore what a binding of the
 * reference's *Python* operators would call; each cites the reference operator it replaces.
 *
 * Activation layout: NHWC ("channels last"), element (n,h,w,c) at  base + ((n*H + h)*W + w)*ld + c,
 * where ld >= C is the pixel stride in elements (lets a kernel read/write a channel slice of a wider
 * concat buffer).  dtype is PPY_F32 or PPY_BF16.
 */
#ifndef PPYOLO_B200_H
#define PPYOLO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ppy_stream_t; /* cudaStream_t */

enum ppy_status {
  PPY_OK = 0,
  PPY_ERR_INVALID = -1,     /* bad argument (null pointer, unsupported size, misalignment) */
  PPY_ERR_WORKSPACE = -2,   /* workspace too small */
  PPY_ERR_CUDA = -3,        /* a CUDA runtime/driver call failed; see ppy_last_cuda_error() */
  PPY_ERR_UNSUPPORTED = -4  /* configuration not built (e.g. tcgen05 path on a non-sm_100 device) */
};
/* PPY_F16X2: an fp32-grade value carried as TWO fp16 numbers (hi = fp16(v), lo = fp16(v - hi), v ~ hi + lo to 2^-22) in
 * two planes of the same geometry: element i of the hi plane at base + i, of the lo plane at base + plane + i (plane = an
 * element offset passed next to the pointer).  The operand format of the fp32-grade tensor-core path (ppy_conv_f16x2). */
enum ppy_dtype { PPY_F32 = 0, PPY_BF16 = 1, PPY_F16X2 = 2 };
enum ppy_act { PPY_ACT_NONE = 0, PPY_ACT_RELU = 1, PPY_ACT_LEAKY = 2, PPY_ACT_MISH = 3 };

int ppy_abi_version(void);
const char* ppy_status_string(int status);
int ppy_last_cuda_error(void);                 /* cudaError_t of the last failing call on this thread */
long long ppy_kernel_launch_count(void);       /* kernels launched by this library since load */

/* ------------------------------------------------------------------------------------------------
 * Layout / glue (replace torch permute/cat/pool kernels on the path)
 * ---------------------------------------------------------------------------------------------- */
/* NCHW fp32 (the tensor Decode.predict uploads, model/decode_np.py:142-147) -> NHWC, channels
 * [C, y_ld) zero-filled. */
int ppy_nchw_to_nhwc(const float* x, void* y, int n, int c, int h, int w, int y_ld, int y_dtype, ppy_stream_t s);
/* NHWC -> NCHW fp32 (module-level interop and tests). */
int ppy_nhwc_to_nchw(const void* x, int x_ld, int x_dtype, float* y, int n, int c, int h, int w, ppy_stream_t s);
/* MaxPool2d(3,2,1), model/resnet_vd.py:103. Output (h+1)/2 x (w+1)/2. */
int ppy_maxpool3x3s2(const void* x, int x_ld, void* y, int y_ld, int n, int h, int w, int c, int dtype, ppy_stream_t s);
/* AvgPool2d(2,2,0) of the "vd" shortcut, model/resnet_vd.py:30. */
int ppy_avgpool2x2(const void* x, int x_ld, void* y, int y_ld, int n, int h, int w, int c, int dtype, ppy_stream_t s);
/* SPP, model/custom_layers.py:275-290: y[..., 0:c]=x, [c:2c]=maxpool5, [2c:3c]=maxpool9, [3c:4c]=maxpool13. */
int ppy_spp(const void* x, int x_ld, void* y, int y_ld, int n, int h, int w, int c, int dtype, ppy_stream_t s);
/* Backward of ppy_spp (training head): dx = dy[..., 0:c] + the three max-pools' gradients routed to each window's arg-max (torch
 * max_pool2d semantics: the first maximum in row-major window order).  x: the forward input, dy: [n,h,w,>=4c]; c % 32 == 0. */
int ppy_spp_backward(const void* x, int x_ld, const void* dy, int dy_ld, void* dx, int dx_ld, int n, int h, int w, int c, int dtype,
                     ppy_stream_t s);
/* nearest x2 upsample (model/head.py:362) written into a channel slice: y is [n,2h,2w,*]. */
int ppy_upsample2x(const void* x, int x_ld, void* y, int y_ld, int n, int h, int w, int c, int dtype, ppy_stream_t s);
/* strided channel-slice copy (concat), rows = n*h*w pixels. */
int ppy_copy_channels(const void* x, int x_ld, void* y, int y_ld, long long rows, int c, int dtype, ppy_stream_t s);
/* CoordConv channels (model/custom_layers.py:256-272): y[h,w,0]=x coord in [-1,1], y[h,w,1]=y coord,
 * channels [2, y_ld) zero. One image, NHWC. */
int ppy_coord_channels(void* y, int y_ld, int h, int w, int dtype, ppy_stream_t s);
/* elementwise activation in place on a dense buffer (Mish etc.; module-level only). */
int ppy_activation(void* x, long long count, int act, int dtype, ppy_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * conv + folded norm + activation (+ residual), Conv2dUnit.forward model/custom_layers.py:243-253,
 * block residual add+ReLU model/resnet_vd.py:54-56/:84-86, DCNv2.forward model/custom_layers.py:551-677
 * ---------------------------------------------------------------------------------------------- */
/* OIHW fp32 conv weight -> packed K-major [cout_pad][k_pad] rows, K index = (kh*KW + kw)*cin_pad + c,
 * zero padded; channels [c_begin, c_begin+c_count) of the source are taken (CoordConv weight split). */
int ppy_pack_conv_weight(const float* w_oihw, int cout, int cin_total, int kh, int kw, int c_begin, int c_count,
                         void* packed, int cout_pad, int cin_pad, int k_pad, int dtype, ppy_stream_t s);
/* Packed weight of the INPUT-GRADIENT conv of a k x k conv (training head; torch autograd's conv backward, reference train.py:437):
 * the dgrad conv reads dY (o_pad >= cout channels) and writes c_main channels with the 180-degree-rotated, transposed weight --
 * packed[ci][tap * o_pad + co] = w[co][ci][kh*kw - 1 - tap], rows ci in [c_main, rows_pad) and columns co >= cout zero.
 * Same layout as ppy_pack_conv_weight(cout = c_main, cin = o_pad); PPY_BF16 only. */
int ppy_pack_conv_weight_dgrad(const float* w_oihw, int cout, int cin_total, int kh, int kw, int c_main, void* packed, int rows_pad, int o_pad,
                               int k_pad, int dtype, ppy_stream_t s);

typedef struct ppy_conv_params {
  const void* x;            /* input NHWC, dtype = in_dtype */
  int x_ld;
  int n, h, w, cin;         /* cin = channels read per tap (multiple of 8) */
  const void* weight;       /* packed by ppy_pack_conv_weight, same dtype as x */
  int cout, kh, kw, stride, pad;
  int k_pad;                /* row length of packed weight (multiple of 64) */
  int cout_pad;             /* rows of packed weight */
  const float* scale;       /* [cout] folded norm scale (or 1) */
  const float* shift;       /* [cout] folded norm shift / bias */
  const float* bias_map;    /* optional [ho*wo][cout] fp32 added to the accumulator (CoordConv fold) */
  const void* residual;     /* optional NHWC [n,ho,wo,cout], dtype = out_dtype, added before act */
  int res_ld;
  int act;                  /* ppy_act */
  void* y;                  /* output NHWC */
  int y_ld;
  int out_dtype;            /* PPY_F32 or PPY_BF16 */
  int upsample2x;           /* 1: write every output pixel to its 2x2 block of an [n,2ho,2wo,*] buffer */
  const float* offset_mask; /* non-null => DCNv2: [n,ho,wo,om_ld] fp32, ch 2t=dy 2t+1=dx, 18+t=mask logit */
  int om_ld;
  /* --- bf16 path only: partial-sum launches (K splits; the weight-gradient GEMM of the training step) ---------------
   * accumulate=1: results are ADDED (fp32 atomics) into y, which the caller zeroed; act none, no residual/bias_map,
   * shift applied once.  split_k: K splits per output tile (0 = enough to fill the SMs).
   * wgrad_taps / wgrad_pitch / wgrad_tap_stride describe the weight-gradient GEMM of a kxk stride-1 conv as ONE such launch:
   * a 1x1 "conv" whose rows are the conv's output channels (x = dY transposed, [cout][pixels]), whose packed "weight" is
   * the transposed zero-bordered input ([cin][pixels]) and whose K runs over the pixels; wgrad_taps=9 repeats the GEMM per
   * tap with the B operand read at flat pixel offset (ky-1)*wgrad_pitch + (kx-1) and the result written to columns
   * [tap*wgrad_tap_stride, ..).  (Kernel instantiation only, NOT validated: a round-2c attempt to drive it from bordered transposes faulted on the device; the
   * training step uses wgrad_taps = 0 with the K-major im2col operand of ppy_im2col_kmajor.) */
  int accumulate;
  int split_k;
  int wgrad_taps;
  int wgrad_pitch;
  int wgrad_tap_stride;
  /* CoordConv fold for 1x1 stride-1 convs (model/custom_layers.py:256-272 feeding a 1x1 Conv2dUnit): the two coordinate
   * channels contribute wx[co]*xc(ox) + wy[co]*yc(oy) with xc = ox/(wo-1)*2-1, yc = oy/(ho-1)*2-1 -- a rank-2 term the
   * epilogue adds to the accumulator (before scale/shift) from coord_w = [wx[cout] | wy[cout]] fp32, instead of reading a
   * per-pixel bias_map.  NULL = off; mutually exclusive with bias_map; requires kh == kw == 1. */
  const float* coord_w;
  /* --- PPY_F16X2 path only (ppy_conv_f16x2): element offsets from the hi plane to the lo plane of x / y / residual (the packed
   * weight's lo plane follows its hi plane at cout_pad*k_pad); overflow: optional device int, set to 1 when an output value
   * leaves the range an fp16 pair can carry (|v| > 65504) -- the caller must then discard the results. */
  long long x_plane;
  long long y_plane;
  long long res_plane;
  int* overflow;
  /* --- PPY_F16X2 path, plain 1x1 stride-1 convs only: a SECOND K source.  K blocks [cin/64, cin/64 + x2_kb) of the GEMM are
   * 64-channel blocks of the pair tensor x2 ([n,ho,wo,*], the output's pixel grid); the packed weight rows are
   * k_pad = cin + 64*x2_kb long.  Two uses (model/resnet_vd.py:44-56, :75-86 -- the bottleneck's `conv3(y) + shortcut`):
   *   x2_tiled = 0  K-concatenation: x2 channels [0, 64*x2_kb) -- conv3 and the projection shortcut conv4 as ONE GEMM
   *                 [W3*s3 | W4*s4] . [y ; x] (both norms' scales folded into the weight rows by the caller);
   *   x2_tiled = 1  the identity shortcut added BY THE TENSOR CORE: for the N tile starting at column n0 the blocks are x2
   *                 channels [n0, n0 + 64*x2_kb) and the weight block holds chan_scale[co] (a power of two: exact) on the
   *                 diagonal, so the residual arrives through the deep TMA operand pipeline instead of the epilogue's
   *                 latency-exposed loads; requires 64*x2_kb == the kernel's N tile (256 when cout % 256 == 0 and
   *                 k_pad <= 512, else 128) and the lo plane of those weight blocks to be zero (it is not even loaded).
   * x2_kb = 0: off.  Mutually exclusive with `residual`. */
  const void* x2;
  int x2_ld;
  long long x2_plane;
  int x2_kb;
  int x2_tiled;
  /* x2_row_mod > 0 (x2_tiled = 0; also 3x3 stride-1 pad-1 convs): x2 is BATCH-INVARIANT -- one image's ho*wo = x2_row_mod rows,
   * repeated so that the buffer holds x2_rows >= x2_row_mod + 127 rows (a 128-row tile starting anywhere inside an image reads
   * contiguous rows).  This is how CoordConv (model/custom_layers.py:256-272) enters the pair path: the two coordinate channels
   * (for a 3x3 conv: their zero-padded values at the 9 taps, 18 columns) are one extra 64-wide K block against the coordinate
   * columns of the weight, instead of a per-pixel fp32 bias map read by the epilogue. */
  int x2_row_mod;
  int x2_rows;
} ppy_conv_params;

/* Stem conv1_1 fused with the NCHW->NHWC change: NCHW fp32 images -> conv 3x3/s2/p1 (3 -> 32, model/resnet_vd.py:100)
 * + folded BN + act -> NHWC.  weight (OIHW [32,3,3,3]), scale, shift are HOST pointers (they travel in the kernel
 * parameter / constant bank).  y_dtype PPY_BF16 with y_ld == 32, w % 4 == 0 and a 16-byte aligned x runs on tcgen05
 * (image and weights rounded to bf16, fp32 accumulate -- the precision of every other bf16 conv); PPY_F32 output and the
 * remaining shapes run the fp32 SIMT kernel. */
int ppy_stem_conv3x3s2(const float* x_nchw, int n, int h, int w, const float* weight_oihw_host, const float* scale_host,
                       const float* shift_host, int cout, int act, void* y, int y_ld, int y_dtype, ppy_stream_t s);

/* The same layer with a PPY_F16X2 output (fp32 SIMT math; results split into the hi plane at y and the lo plane at
 * y + y_plane): first kernel of the fp32-grade tensor-core engine. */
int ppy_stem_conv3x3s2_f16x2(const float* x_nchw, int n, int h, int w, const float* weight_oihw_host, const float* scale_host,
                             const float* shift_host, int cout, int act, void* y, int y_ld, long long y_plane, ppy_stream_t s);

/* The same layer fed the RESIZED uint8 RGB batch [n, h, w, 3] (HWC: what cv2.resize hands Decode.process_image,
 * model/decode_np.py:125-134) instead of the normalised float CHW tensor: NormalizeImage (tools/transform.py:891-921) and Permute
 * (:1020-1055) happen inside the kernel through `lut` (DEVICE pointer, [3][256] floats: lut[c][u] = the reference's numpy
 * expression ((u / 255) - mean[c]) / std[c] evaluated on the host, so the result is bit-identical) -- a quarter of the upload
 * bytes.  y_dtype PPY_F32 / PPY_BF16 / PPY_F16X2 (then y_plane > 0). */
int ppy_stem_conv3x3s2_u8(const uint8_t* x_nhwc_u8, int n, int h, int w, const float* lut, const float* weight_oihw_host,
                          const float* scale_host, const float* shift_host, int cout, int act, void* y, int y_ld, int y_dtype,
                          long long y_plane, ppy_stream_t s);

/* Glue of the fp32-grade tensor-core path on PPY_F16X2 tensors (hi plane at the pointer, lo plane `plane` elements further):
 * MaxPool2d(3,2,1) model/resnet_vd.py:103, AvgPool2d(2,2) :30, SPP model/custom_layers.py:275-290 (y[..., 0:c]=x,
 * [c:2c]=maxpool5, [2c:3c]=maxpool9, [3c:4c]=maxpool13).  Max-type results equal the fp32 kernels' on the joined tensors. */
int ppy_maxpool3x3s2_f16x2(const void* x, int x_ld, long long x_plane, void* y, int y_ld, long long y_plane, int n, int h, int w,
                           int c, ppy_stream_t s);
int ppy_avgpool2x2_f16x2(const void* x, int x_ld, long long x_plane, void* y, int y_ld, long long y_plane, int n, int h, int w, int c,
                         ppy_stream_t s);
int ppy_spp_f16x2(const void* x, int x_ld, long long x_plane, void* y, int y_ld, long long y_plane, int n, int h, int w, int c,
                  ppy_stream_t s);
/* fp32 rows [rows][x_ld] -> pair planes, and back (module-level interop, tests). */
int ppy_split_f16x2(const float* x, int x_ld, void* y, int y_ld, long long y_plane, long long rows, int c, ppy_stream_t s);
int ppy_join_f16x2(const void* x, int x_ld, long long x_plane, float* y, int y_ld, long long rows, int c, ppy_stream_t s);

/* Stage 1 of the two-kernel DCNv2 (model/custom_layers.py:551-674): bilinear sample x sigmoid(mask) for every
 * (output pixel, tap) -> out[m][tap*c + ch], the K-major A matrix a 1x1 ppy_conv_* over [n,ho,wo,k*k*c] consumes with
 * the SAME packed 3x3 weight.  offset_mask as in ppy_conv_params. */
int ppy_dcn_gather(const void* x, int x_ld, int n, int h, int w, int c, const float* offset_mask, int om_ld, int k,
                   int stride, int pad, void* out, int dtype, ppy_stream_t s);

/* fp32 SIMT implicit GEMM (the 1e-4 parity path).  x/weight/residual fp32. */
int ppy_conv_f32(const ppy_conv_params* p, ppy_stream_t s);
/* bf16 tcgen05 implicit GEMM, fp32 accumulate in TMEM, TMA-fed weights (the throughput path). */
int ppy_conv_bf16(const ppy_conv_params* p, ppy_stream_t s);
/* 1 when the tcgen05 path can run on the current device (compute capability 10.x). */
int ppy_conv_bf16_supported(void);
/* fp32-grade tcgen05 implicit GEMM: x, weight (ppy_pack_conv_weight with PPY_F16X2), residual and y (out_dtype PPY_F16X2;
 * PPY_F32 also allowed for y) are fp16 hi/lo pairs; every K block issues hi*hi + hi*lo + lo*hi into the fp32 TMEM
 * accumulator (kind::f16, fp16 operands): products carry 22 bits, the dropped lo*lo term is 2^-22 relative -- the accuracy of an
 * fp32 FMA chain at a third of the bf16 tensor rate.  Same operator contract as ppy_conv_bf16 (Conv2dUnit.forward,
 * model/custom_layers.py:243-253; DCNv2.forward :551-677); no partial-sum (accumulate) launches. */
int ppy_conv_f16x2(const ppy_conv_params* p, ppy_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * training side (train.py:427-442; frozen-backbone BNs run on BATCH statistics, custom_layers.py:122)
 * ---------------------------------------------------------------------------------------------- */
/* Per-channel batch statistics of an NHWC tensor -> folded scale = gamma/sqrt(var+eps), shift = beta - mean*scale;
 * running_mean/var (optional) updated in place with torch's momentum / unbiased-variance rule (one launch: the last block to
 * finish finalizes). workspace: 2*c + 1 doubles (sums, sums of squares, block counter); zeroed by the call. */
int ppy_bn_batch_stats(const void* x, int x_ld, long long rows, int c, int dtype, const float* gamma, const float* beta,
                       float eps, float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                       double* workspace, ppy_stream_t s);
/* ppy_bn_batch_stats + ppy_scale_shift_act of one layer in ONE cooperative launch (statistics, grid barrier, normalise +
 * activation + residual; the second read of x mostly hits L2): the train-mode BatchNorm of a frozen-backbone layer as one graph node
 * instead of four.  scale / shift receive the folded parameters as before.  workspace: PPY_BN_WORKSPACE_DOUBLES(c) doubles (up to 16
 * replicas of the 2c per-channel sums -- CTAs of one channel column add to different addresses -- and a counter) that are ZERO on
 * entry; the call leaves them zero (no memset).  save_mean / save_invstd (optional, [c]): the batch mean and 1/sqrt(var + eps) a BatchNorm
 * backward needs.  c <= 2048; PPY_ERR_UNSUPPORTED when the device cannot launch cooperatively. */
#define PPY_BN_WORKSPACE_DOUBLES(c) (32 * (size_t)(c) + 1)
int ppy_bn_train_fused(const void* x, int x_ld, void* y, int y_ld, long long rows, int c, int dtype, const float* gamma, const float* beta,
                       float eps, float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                       const void* residual, int res_ld, int act, double* workspace, float* save_mean, float* save_invstd, ppy_stream_t s);
/* Backward of train-mode BatchNorm2d + relu / leaky(0.1) of a head Conv2dUnit (model/custom_layers.py:243-253 under autograd) in ONE
 * cooperative launch: g = dy * act'(y); dbeta = sum g; dgamma = sum g * xhat; dx = gamma * invstd * (g - dbeta/N - xhat * dgamma/N),
 * xhat = (x - save_mean) * save_invstd (the forward's batch statistics, ppy_bn_train_fused).  y (the forward's output) is needed for
 * relu / leaky only.  dgamma / dbeta: fp32 [c], optional.  workspace: PPY_BN_WORKSPACE_DOUBLES(c) doubles, ZERO on entry, left zero. */
int ppy_bn_act_backward(const void* dy, int dy_ld, const void* x, int x_ld, const void* y, int y_ld, void* dx, int dx_ld, long long rows, int c,
                        int dtype, const float* gamma, const float* save_mean, const float* save_invstd, int act, float* dgamma, float* dbeta,
                        double* workspace, ppy_stream_t s);
/* y = act(x*scale[c] + shift[c] (+ residual)) on NHWC rows. */
int ppy_scale_shift_act(const void* x, int x_ld, void* y, int y_ld, long long rows, int c, int dtype, const float* scale,
                        const float* shift, const void* residual, int res_ld, int act, ppy_stream_t s);
/* DropBlock, model/custom_layers.py:293-342 (train only; one per detection block, block_size 3, keep_prob 0.9), without a host
 * round trip: ppy_dropblock_mask draws Bernoulli(gamma) seeds with Philox4x32-10 -- one call per four consecutive LOGICAL
 * (n,c,h,w) elements, key/counter = the (seed, offset) pair in DEVICE memory at rng_state[0..1] (the kernel advances the offset,
 * so a captured CUDA graph draws fresh numbers at every replay) --, grows them with the 3x3 / stride 1 / padding 1 max-pool
 * (:333-334) into mask = 1 - pooled and accumulates sum(mask) in *count (device).  seeds and mask are uint8 tensors sharing the
 * element strides (sn, sc, sh, sw) of the activation (NCHW-contiguous or channels_last).  ppy_dropblock_apply then computes
 * y = x * mask * numel / sum(mask) in the reference's operation order (:341) over the flat storage -- forward (x -> y) and
 * backward (dy -> dx) alike.  block_size != 3 is PPY_ERR_UNSUPPORTED (the reference's own pooling only keeps the shape for 3).
 * ppy_dropblock_mask_from_seeds runs the second stage on caller-provided seeds (tests). */
int ppy_dropblock_mask(uint8_t* seeds, uint8_t* mask, int n, int c, int h, int w, long long sn, long long sc, long long sh, long long sw,
                       int block_size, float gamma, unsigned long long* rng_state, unsigned int* count, ppy_stream_t s);
int ppy_dropblock_mask_from_seeds(const uint8_t* seeds, uint8_t* mask, int n, int c, int h, int w, long long sn, long long sc, long long sh,
                                  long long sw, unsigned long long* rng_state, unsigned int* count, ppy_stream_t s);
int ppy_dropblock_apply(const void* x, void* y, const uint8_t* mask, const unsigned int* count, long long numel, int dtype, ppy_stream_t s);
/* torch.optim.SGD(momentum, weight_decay) update of one fp32 tensor, gradient pre-scaled by grad_scale (1/world). */
int ppy_sgd_momentum(float* param, const float* grad, float* momentum_buf, long long n, float lr, float momentum,
                     float weight_decay, float grad_scale, int first_step, ppy_stream_t s);
/* The whole optimizer step in ONE launch: for every trainable tensor t the ppy_sgd_momentum update read from the (all-reduced)
 * flat gradient bucket -- params[t] (device pointer table), grad_flat / momentum_flat ranges [offsets[t], offsets[t+1]), learning
 * rate lr * lr_mult[t], weight decay weight_decay[t] (the reference's per-layer groups, model/custom_layers.py:167-241) -- followed,
 * when shadow_flat != NULL, by ppy_ema_update of the new value into shadow_flat + shadow_offsets[t] (model/EMA.py:31-45; the
 * reference copies every parameter to the host for this each step).  Bit-identical to the two separate kernels. */
int ppy_sgd_ema_multi(float* const* params, const float* grad_flat, float* momentum_flat, float* shadow_flat, const long long* offsets,
                      const long long* shadow_offsets, const float* lr_mult, const float* weight_decay, int num_tensors, float lr,
                      float momentum, float grad_scale, int first_step, float ema_decay, float ema_one_minus_decay, ppy_stream_t s);
/* The exchange step of data-parallel training (train.py:437-442 with replicas) as ONE kernel per rank over NVLink /
 * NVSwitch peer memory: all-reduce(sum) of the flat fp32 gradient bucket + the ppy_sgd_ema_multi update.  The bucket of every rank
 * lives in symmetric memory: grad_local = this rank's mapping, grad_peers = DEVICE array [world] of every rank's bucket as mapped
 * into this process, grad_multicast = the NVLS multicast address of all of them (NULL when the box has none: the kernel then sums /
 * stores through grad_peers).  signal_pads = DEVICE array [world] of zero-initialised uint32 pads in peer memory (words
 * [pad_slot, pad_slot + world) are used); seq must grow by one per call on every rank (barrier sequence); done = TWO zeroed device
 * uint32 (CTA counter; error word, set to 1 when a barrier wait gave up after 3 s because a peer never arrived).  Every rank must make the same call (same seq) on its own device: the kernels rendezvous on the pads -- rank r reduces
 * slice r with multimem.ld_reduce (the switch adds) and multimem.st (the switch replicates), then all ranks run the optimizer on
 * the reduced bucket.  total_padded: bucket length in floats, multiple of 4 (tail zero); remaining arguments as ppy_sgd_ema_multi.
 * Cooperative launch (one CTA per SM); PPY_ERR_UNSUPPORTED if the device cannot. */
int ppy_allreduce_sgd_ema(float* grad_local, float* grad_multicast, float* const* grad_peers, unsigned int* const* signal_pads,
                          int pad_slot, int rank, int world, unsigned int seq, unsigned int* done, long long total_padded,
                          float* const* params, float* momentum_flat, float* shadow_flat, const long long* offsets,
                          const long long* shadow_offsets, const float* lr_mult, const float* weight_decay, int num_tensors, float lr,
                          float momentum, float grad_scale, int first_step, float ema_decay, float ema_one_minus_decay, ppy_stream_t s);
/* K-major operand of the weight-gradient GEMM of a k x k stride-1 conv (training step, conv_autograd.py): NHWC bf16 x
 * [n,h,w,c] -> out [c*k*k rows][m_pad] bf16 with out[(ch*k*k + ky*k + kx)][m] = x[pixel m shifted by (ky-pad, kx-pad)][ch], zero
 * outside the image and for m in [n*h*w, m_pad); rows follow the OIHW weight order, k = 1 is the plain transpose (also used
 * for dY).  m_pad % 64 == 0. */
int ppy_im2col_kmajor(const void* x, int x_ld, int n, int h, int w, int c, int k, int pad, void* out, long long m_pad, ppy_stream_t s);
/* The same operand for a STRIDED conv (the 3x3 / stride 2 convs of an unfrozen ResNet-vd stage, model/resnet_vd.py:19-22; the
 * DCNv2 offset conv of stage5_0): columns are the n*ho*wo OUTPUT pixels, out[(ch*k*k + ky*k + kx)][m] =
 * x[oy*stride + ky - pad][ox*stride + kx - pad][ch]. */
int ppy_im2col_kmajor_strided(const void* x, int x_ld, int n, int h, int w, int c, int k, int stride, int pad, void* out, long long m_pad,
                              ppy_stream_t s);
/* DCNv2 backward, sampling side (model/custom_layers.py:551-677 under torch autograd; the native spec the reference vendors but never
 * loads is external/DCNv2/src/cuda/dcn_v2_im2col_cuda.cu:197-327): given dcol = dY . Wt ([n*ho*wo][k*k*c] fp32, column tap*c + ch --
 * the layout ppy_dcn_gather writes) it ACCUMULATES the input gradient into dx (fp32 NHWC [n,h,w,dx_ld], caller-zeroed or holding the
 * offset conv's input gradient; vector atomics) and writes the gradient of the offset/mask conv's output into d_offset_mask
 * ([n*ho*wo][om_ld] fp32: 2t = d dy, 2t+1 = d dx, 2*k*k + t = d mask-logit; columns >= 3*k*k untouched).  x: NHWC, dtype PPY_F32 or
 * PPY_BF16.  The weight gradient is the conv kernel's partial-sum GEMM over ppy_dcn_gather's matrix (conv_autograd.py). */
int ppy_dcn_backward_sample(const void* x, int x_ld, int n, int h, int w, int c, const float* offset_mask, int om_ld, int k, int stride,
                            int pad, const float* dcol, float* dx, int dx_ld, float* d_offset_mask, int dtype, ppy_stream_t s);
/* ExponentialMovingAverage.update, model/EMA.py:31-45, on the device for all trainable tensors in one launch:
 * shadow[offsets[t] + i] = decay * shadow[..] + one_minus_decay * params[t][i]  (numpy float32 operation order; the reference
 * round-trips every parameter through host memory each step).  params: device array of num_tensors device pointers;
 * offsets: device array of num_tensors + 1 element offsets into shadow_flat. */
int ppy_ema_update(float* shadow_flat, const float* const* params, const long long* offsets, int num_tensors, float decay,
                   float one_minus_decay, ppy_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * head post-processing
 * ---------------------------------------------------------------------------------------------- */
/* get_iou_aware_score, model/head.py:138-141, on an NHWC fp32 head output: [.., A*(6+C)] -> [.., A*(5+C)]. */
int ppy_iou_aware_score(const float* x, int x_ld, float* y, int y_ld, long long pixels, int an_num, int num_classes,
                        double factor, ppy_stream_t s);
/* Fused get_iou_aware_score + yolo_box (model/head.py:21-80) for one scale.
 * head: NHWC fp32 [n,size,size,ld]; anchors: 2*an_num host floats; im_size: device [n,2] = (h,w).
 * boxes [n,total_boxes,4], scores [n,total_boxes,C]; this scale fills rows [box_offset, box_offset+size*size*an). */
int ppy_yolo_decode(const float* head, int ld, int n, int size, int an_num, int num_classes, const float* anchors,
                    int stride, double scale_x_y, const float* im_size, int clip_bbox, int iou_aware, double factor,
                    float* boxes, float* scores, int box_offset, int total_boxes, ppy_stream_t s);
/* Same, and every score > score_threshold is also counted in the per-image score histogram at the head of a Matrix-NMS
 * workspace (ppy_matrix_nms_workspace_bytes; zero it once per batch with ppy_nms_candidates_reset(ws, n, 1) before the
 * first scale), so that ppy_matrix_nms_batched_hist can skip its own histogram pass over the scores. */
int ppy_yolo_decode_hist(const float* head, int ld, int n, int size, int an_num, int num_classes, const float* anchors,
                         int stride, double scale_x_y, const float* im_size, int clip_bbox, int iou_aware, double factor,
                         float* boxes, float* scores, int box_offset, int total_boxes, float score_threshold,
                         void* nms_workspace, ppy_stream_t s);
/* jaccard, model/matrix_nms.py:33-47. */
int ppy_pairwise_iou(const float* a, int na, const float* b, int nb, float* out, ppy_stream_t s);

/* Batched Matrix-NMS, model/matrix_nms.py:102-151 applied to every image (replaces the loop at
 * model/head.py:462-464).  out: [n,keep_top_k,6] rows [label,score,x0,y0,x1,y1]; counts: [n] int32 rows
 * valid per image (0 => the reference's [[-1]*6] sentinel; <0 => candidate overflow). */
int ppy_matrix_nms_workspace_bytes(int n, int num_boxes, int num_classes, size_t* bytes);
int ppy_matrix_nms_batched(const float* boxes, const float* scores, int n, int num_boxes, int num_classes,
                           float score_threshold, float post_threshold, int nms_top_k, int keep_top_k,
                           int use_gaussian, float gaussian_sigma, float* out, int* counts, void* workspace,
                           size_t workspace_bytes, ppy_stream_t s);
/* ppy_matrix_nms_batched with the score histogram already in the workspace (ppy_yolo_decode_hist). */
int ppy_matrix_nms_batched_hist(const float* boxes, const float* scores, int n, int num_boxes, int num_classes,
                                float score_threshold, float post_threshold, int nms_top_k, int keep_top_k,
                                int use_gaussian, float gaussian_sigma, float* out, int* counts, void* workspace,
                                size_t workspace_bytes, ppy_stream_t s);

/* Sparse post-processing for the whole-network path (no dense score tensor): ppy_yolo_decode_candidates writes the boxes
 * of one scale and appends every (box, class) score > score_threshold to a per-image candidate list + score histogram
 * in `workspace` (ppy_nms_candidate_workspace_bytes; zero the counters once per batch with ppy_nms_candidates_reset
 * before the first scale); ppy_matrix_nms_candidates then runs Matrix-NMS for the whole batch from those lists.
 * Results are bit-identical to ppy_yolo_decode + ppy_matrix_nms_batched.  `cap` = list capacity per image; an image
 * with more candidates is flagged counts[i] = -2. */
int ppy_nms_candidate_workspace_bytes(int n, int cap, size_t* bytes);
int ppy_nms_candidates_reset(void* workspace, int n, int cap, ppy_stream_t s);
int ppy_yolo_decode_candidates(const float* head, int ld, int n, int size, int an_num, int num_classes,
                               const float* anchors, int stride, double scale_x_y, const float* im_size, int clip_bbox,
                               int iou_aware, double factor, float* boxes, int box_offset, int total_boxes,
                               float score_threshold, void* workspace, int cap, ppy_stream_t s);
int ppy_matrix_nms_candidates(const float* boxes, int n, int num_boxes, int num_classes, float score_threshold,
                              float post_threshold, int nms_top_k, int keep_top_k, int use_gaussian, float gaussian_sigma,
                              float* out, int* counts, void* workspace, int cap, ppy_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * Training: fused fine-grained YOLOv3 loss for one output scale (model/losses.py:121-356 _get_fine_grained_loss + _calc_obj_loss,
 * model/iou_losses.py IouLoss :15-191, IouAwareLoss :194-246) and the target assignment of the data pipeline
 * (tools/transform.py:1318-1421 Gt2YoloTargetSingle).
 *   out     [n, a*(5|6+C), size, size] fp32 NCHW raw head output (iou_aware: the first `a` channels are the IoU logits)
 *   target  [n, a, 6+C, size, size] fp32 (tx, ty, tw, th, tscale, tobj, one-hot class)
 *   gt_box  [n, g, 4] fp32 normalised (cx, cy, w, h), 16-byte aligned (zero rows = padding)
 *   anchors_host: 2*a floats (w, h in pixels) of this scale's anchors, HOST pointer
 * forward: ADDS this scale's six losses (xy, wh, obj, cls, iou, iou_aware; each the batch mean of per-image sums) to
 * losses[6] (device; zero it before the first scale) and writes the no-object mask [n, a, size, size] the backward needs.
 * workspace: ppy_yolo_loss_workspace_bytes() bytes, 16-byte aligned, its first 4 bytes zero before the first use.
 * backward: grad_out[n, a*(5|6+C), size, size] = d(sum_k grad_losses[k] * loss_k)/d(out); grad_losses is a DEVICE pointer to 6
 * floats (no host synchronisation; CUDA-graph capturable).  ciou_term is not supported (no config enables it). */
int ppy_yolo_loss_workspace_bytes(int n, int a, int size);
int ppy_yolo_loss_forward(const float* out, const float* target, const float* gt_box, int n, int a, int num_classes, int size, int g,
                          const float* anchors_host, int stride, double scale_x_y, float ignore_thresh, int iou_aware, int has_iou_loss,
                          float iou_loss_weight, int loss_square, float iou_aware_weight, int match_score, float* noobj_mask,
                          void* workspace, float* losses, ppy_stream_t s);
int ppy_yolo_loss_backward(const float* out, const float* target, int n, int a, int num_classes, int size, const float* anchors_host,
                           int stride, double scale_x_y, int iou_aware, int has_iou_loss, float iou_loss_weight, int loss_square,
                           float iou_aware_weight, const float* noobj_mask, const float* grad_losses, float* grad_out, ppy_stream_t s);
/* target [n, mask_len, 6+C, img_h/downsample, img_w/downsample] for one scale from padded ground truth: gt_bbox [n, g, 4]
 * normalised (cx, cy, w, h), gt_class [n, g] int32, gt_score [n, g]; anchors_host = all 2*num_anchors anchor sizes (ints, pixels),
 * mask_host = this scale's anchor indices (HOST pointers).  Sequential over a sample's boxes like the reference: later boxes
 * overwrite earlier ones in the same cell, class bits accumulate. */
int ppy_gt2yolo_target(const float* gt_bbox, const int* gt_class, const float* gt_score, int n, int g, const int* anchors_host,
                       int num_anchors, const int* mask_host, int mask_len, int num_classes, int img_h, int img_w, int downsample,
                       float iou_thresh, float* target, ppy_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * Pre-processing: ResizeImage of the eval / test pipeline (tools/transform.py:923-1018 with interp = cv2.INTER_CUBIC,
 * model/decode_np.py:125-134) + decodeImage's BGR -> RGB, for a batch of differently sized uint8 images in one launch.
 *   src   packed source images (each tightly packed HWC uint8, 3 channels), DEVICE memory
 *   meta  [n][3] int64, DEVICE: {byte offset of image i in src, height, width}
 *   dst   [n, dsize, dsize, 3] uint8 -- the input format of ppy_stem_conv3x3s2_u8
 * Arithmetic: OpenCV's own 8-bit bicubic resize restated operation by operation (fixed-point horizontal pass, float vertical
 * pass, round-to-nearest-even): bit-identical to cv2.resize wherever OpenCV runs its own code (builds without Intel IPP, or
 * cv2.ipp.setUseIPP(False)); IPP-enabled wheels differ from that by +-1 on ~3 % of the pixels. */
int ppy_resize_cubic_u8_batch(const uint8_t* src, const long long* meta, int n, uint8_t* dst, int dsize, int swap_rb, ppy_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* PPYOLO_B200_H */
