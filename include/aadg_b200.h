/*
 * aadg_b200 — C ABI of the B200-native AADG hot path (libaadg_b200.so, sm_100a).
 *
 * The reference (CRazorback/AADG) has no FFI: its hot path sits behind Python call shapes and runs
 * on Pillow (CPU), geomloss/KeOps and cuDNN.  Each entry point below names the reference
 * interface it replaces (file:line under /root/reference).  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; device pointers unless a parameter says HOST;
 *   - the caller owns every buffer (inputs, outputs, workspace); nothing is allocated, freed or
 *     synchronised inside; all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return 0 on success, a negative AADG_E* code otherwise; `aadg_last_error()` returns the
 *     calling thread's last message; no exception crosses the boundary;
 *   - re-entrant; no global mutable state.
 */
#ifndef AADG_B200_H
#define AADG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AADG_OK 0
#define AADG_EINVAL (-1)   /* bad argument */
#define AADG_ENOSPC (-2)   /* workspace too small */
#define AADG_ECUDA (-3)    /* CUDA launch / runtime failure */

#define AADG_ABI_VERSION 1

int aadg_version(void);
const char* aadg_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * uint8 augmentation bank — replaces data/basic.py:70-167,231-260 (the ten live Pillow ops and the
 * geometric ops), data/policy.py:15-61 (Policy / DGMultiPolicy application) and, for the epilogue,
 * data/transform.py:97-236 (DGRandomScaleCrop, Normalize_dg, ToTensor) + :323-340 (collate order).
 * ---------------------------------------------------------------------------------------------- */

#define AADG_MAX_OPS 4

enum aadg_op {           /* index into the reference's augment_list(), data/basic.py:231-243 ... */
  AADG_OP_AUTOCONTRAST = 0, AADG_OP_INVERT = 1, AADG_OP_EQUALIZE = 2, AADG_OP_SOLARIZE = 3,
  AADG_OP_POSTERIZE = 4, AADG_OP_CONTRAST = 5, AADG_OP_COLOR = 6, AADG_OP_BRIGHTNESS = 7,
  AADG_OP_SHARPNESS = 8, AADG_OP_CUTOUT = 9,
  /* ... then the ops the reference defines but never samples, data/basic.py:12-67,82 */
  AADG_OP_SHEAR_X = 10, AADG_OP_SHEAR_Y = 11, AADG_OP_TRANSLATE_X = 12, AADG_OP_TRANSLATE_Y = 13,
  AADG_OP_ROTATE = 14, AADG_OP_FLIP = 15,
  AADG_OP_COUNT = 16
};

enum aadg_dataset { AADG_DATASET_OPTIC = 0, AADG_DATASET_VESSEL = 1 };

/* One output image = one row of the decision table (every random draw of the reference resolved to
 * integers / C floats on the host; see aadg_b200/data/decisions.py).  160 bytes, little endian. */
typedef struct aadg_aug_row {
  int32_t src;                      /* source image index                                         */
  int32_t n_ops;                    /* ops of the chosen sub-policy, applied in order              */
  int32_t op[AADG_MAX_OPS];         /* enum aadg_op                                               */
  float fparam[AADG_MAX_OPS];       /* Contrast/Color/Brightness/Sharpness: blend factor          */
  int32_t iparam[AADG_MAX_OPS][6];  /* Solarize: [0]=ceil(threshold); Posterize: [0]=AND mask;    */
                                    /* Cutout: inclusive x0,y0,x1,y1 (x1<x0: no-op);              */
                                    /* Shear/Translate/Rotate: Pillow 16.16 a0,a1,a2,a3,a4,a5     */
  int32_t do_scale, scale_w, scale_h; /* DGRandomScaleCrop.scale (data/transform.py:104-112)      */
  int32_t pad, crop_x, crop_y;        /* RandomCrop (data/transform.py:35-55)                      */
} aadg_aug_row_t;

/* Bytes of workspace that is always enough for the u8 entry points below (worst case: every row
 * materialises AADG_MAX_OPS-1 intermediate images).  0 for non-positive sizes. */
size_t aadg_u8_workspace_bytes(int n_rows, int n_src, int height, int width);

/* Post-policy images: out_u8[r] = chain(rows[r])(src_images[rows[r].src])   — what
 * DGMultiPolicy.__call__ returns as sample['aug_images'] (data/policy.py:51-61).
 *   src_images uint8 [n_src,H,W,3]; src_masks uint8 [n_src,H,W] or NULL; rows HOST [n_rows];
 *   out_u8 uint8 [n_rows,H,W,3]; out_masks uint8 [n_rows,H,W] or NULL (Cutout/geometric ops edit
 *   the mask like data/basic.py does; the train transform later discards it).                    */
int aadg_u8_apply_policy(const uint8_t* src_images, const uint8_t* src_masks,
                         const aadg_aug_row_t* rows, int n_rows, int n_src, int height, int width,
                         uint8_t* out_u8, uint8_t* out_masks, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Policy + Normalize_dg + ToTensor in one pass (no scale/crop): data/policy.py:51-61 followed by
 * data/transform.py:149-186,217-236 and the collate order of :323-340 (row order is the caller's).
 *   out_images float32 [n_rows,3,H,W] = x/127.5-1; out_labels float32 [n_rows,C,H,W] from the
 *   ORIGINAL masks, C=2 (optic multilabel) or 1 (vessel); either output may be NULL.             */
int aadg_u8_policy_normalize(const uint8_t* src_images, const uint8_t* src_masks,
                             const aadg_aug_row_t* rows, int n_rows, int n_src, int height,
                             int width, int dataset, float* out_images, float* out_labels,
                             void* workspace, size_t workspace_bytes, void* stream);

/* DGRandomScaleCrop + Normalize_dg + ToTensor (data/transform.py:97-236), bit-exact with Pillow's BILINEAR /
 * NEAREST resize, ImageOps.expand zero padding and crop, driven by the row's do_scale, scale_w, scale_h, pad,
 * crop_x, crop_y.  images uint8 [*,H,W,3]: image r is images[r] if image_by_row else images[rows[r].src];
 * masks uint8 [n_src,H,W]: the ORIGINAL masks, indexed rows[r].src (data/transform.py:127-131), may be NULL;
 * out_images float32 [n_rows,3,crop_h,crop_w]; out_labels float32 [n_rows,C,crop_h,crop_w]; either may be NULL. */
size_t aadg_u8_scale_crop_workspace_bytes(int n_rows, int max_scale_w, int max_scale_h);
int aadg_u8_scale_crop_normalize(const uint8_t* images, int image_by_row, const uint8_t* masks,
                                 const aadg_aug_row_t* rows, int n_rows, int n_src, int height, int width,
                                 int crop_w, int crop_h, int dataset, float* out_images, float* out_labels,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Float tensor augmentation bank — replaces the 19 operations of data/operations.py:142-399 /
 * data/functional.py:110-271 (+ data/kernels.py:9-13), forward values:
 *     out = clamp(mask_b * op(x, mag_b) + (1 - mask_b) * x, 0, 1)          (operations.py:73-100)
 * x, out float32 [batch,3,h,w] in [0,1] (out != x); mag, mask float32 [batch] on the device (NULL: mag 0,
 * mask 1); perm int32 [batch] (SamplePairing only).  op = index in data/operations.py __all__:
 * ShearX 0, ShearY 1, TranslateX 2, TranslateY 3, HorizontalFlip 4, VerticalFlip 5, Rotate 6, Invert 7,
 * Solarize 8, Posterize 9, Gray 10, Contrast 11, AutoContrast 12, Saturate 13, Brightness 14, Hue 15,
 * SamplePairing 16, Equalize 17, Sharpness 18.
 * ---------------------------------------------------------------------------------------------- */
size_t aadg_f32_workspace_bytes(int batch);
int aadg_f32_op(int op, const float* x, int batch, int h, int w, const float* mag, const float* mask,
                const int32_t* perm, float* out, void* workspace, size_t workspace_bytes, void* stream);
/* Backward of aadg_f32_op (the point of the reference's bank: operations.py:73-108 draws a RelaxedBernoulli mask and
 * functional.py:21-46 uses straight-through estimators so that probabilities and magnitudes can be learned): given go =
 * d(loss)/d(out) writes gx = d/dx [batch,3,h,w] and gmag / gmask float32 [batch] (either may be NULL) = d/d(mag_b),
 * d/d(mask_b).  torch autograd's semantics through the reference code: clamp gates inclusive, Solarize / Posterize send
 * their gradient to the magnitude only (summed), AutoContrast / Equalize straight to the image. */
int aadg_f32_op_backward(int op, const float* x, const float* go, int batch, int h, int w, const float* mag,
                         const float* mask, const int32_t* perm, float* gx, float* gmag, float* gmask, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sinkhorn diversity reward — replaces geomloss.SamplesLoss("sinkhorn", cost=<cosine KeOps formula>,
 * backend="online") constructed at search_dg.py:116 / search_dg_2d.py:116 and called at
 * search_dg.py:158-160 / search_dg_2d.py:159-161 (p=2, blur=.05, scaling=.5, debiased, uniform
 * weights, forward value only), and the reward assembly of search_dg.py:150-162.
 * ---------------------------------------------------------------------------------------------- */

/* Largest cloud (points) the one-launch shared-memory kernels accept. */
int aadg_sinkhorn_small_max_points(void);

/* n_problems independent divergences in ONE launch, no host sync.
 *   points float32 [*, dim] (device); problems int32 [n_problems,4] (device) = x_off, x_n, y_off, y_n
 *   (row ranges of `points`, each 1..aadg_sinkhorn_small_max_points()); out float32 [n_problems]. */
int aadg_sinkhorn_small_batched(const float* points, const int32_t* problems, int n_problems, int dim,
                                float* out, void* stream);

size_t aadg_sinkhorn_rewards_workspace_bytes(int n_policies, int n_domains);

/* search_dg.py:150-162 in one launch: for every policy j the rows j::n_policies of `features`
 * float32 [n_rows,dim] are split by argmax(domain_code[row,:]) (float32 [n_rows,n_domains]) into
 * domain clouds; pair_values float32 [n_policies, n_pairs] receives the pairwise divergences in the
 * reference's call order ((1,2),(2,3),(1,3) for three domains) and rewards[j] += (d12+d13)+d23.
 * A cloud that is empty or larger than the small-kernel limit yields NaN for that pair. */
int aadg_sinkhorn_diversity_rewards(const float* features, const float* domain_code, int n_rows, int dim,
                                    int n_domains, int n_policies, float* rewards, float* pair_values,
                                    void* workspace, size_t workspace_bytes, void* stream);

size_t aadg_sinkhorn_large_workspace_bytes(int n, int m, int dim);

/* One divergence between big clouds x float32 [n,dim], y float32 [m,dim]; out float32 [1] (device).
 * The four cost matrices are materialised in the workspace (about 4*4*n*m bytes).  diameter > 0
 * skips the data-diameter reduction and its host sync (geomloss' `diameter=` argument); otherwise
 * the stream is synchronised once, like the reference's `.item()`.  n_iterations_out (HOST, may be
 * NULL) receives the number of epsilon values (soft-min sweeps executed = that + 2). */
int aadg_sinkhorn_large(const float* x, int n, const float* y, int m, int dim, float diameter, float* out,
                        int* n_iterations_out, void* workspace, size_t workspace_bytes, void* stream);
/* Only the set-up of aadg_sinkhorn_large (norms, diameter, cost matrices) — for benchmarks that report the
 * epsilon iterations separately from the one-off cost build. */
int aadg_sinkhorn_large_setup(const float* x, int n, const float* y, int m, int dim, float diameter,
                              void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Segmentation-net convolutions — replace the cuDNN convolutions behind `model(input)` and
 * `seg_loss.backward()` (search_dg.py:132,171; models/__init__.py:17-23: smp.DeepLabV3Plus / Unet).
 * Activations bf16 NHWC with a channel stride (`ld*`, elements) so a tensor may be a channel slice
 * of a wider (concat) buffer; fp32 accumulation on tcgen05 tensor cores; channel counts, strides and
 * offsets are multiples of 8; stride 1 or 2; any dilation; filters up to 7x7.
 * ---------------------------------------------------------------------------------------------- */

/* y[n,oy,ox,y_c_off+co] (+)= sum x[n, oy*stride-pad+r*dil, ox*stride-pad+s*dil, ci] * w[r*S+s][co][ci]
 *   x bf16 [n,h,w,ldx]; w bf16 [R*S][cout][cin]; y bf16 [n,ho,wo,ldy]. */
int aadg_conv_fprop_bf16(const void* x, int n, int h, int w, int cin, int ldx, const void* wgt, int cout, int r,
                         int s, int stride, int pad, int dil, void* y, int ho, int wo, int ldy, int y_c_off,
                         int accumulate, void* stream);
/* the same forward convolution plus the batch-norm statistics of its output in the same pass: stat_sum / stat_sq
 * (fp32 [cout], ACCUMULATED -- zero them first) receive the per-channel sum and sum of squares of the bf16 values
 * written to y (replaces aadg_bn_stats over y: BatchNorm2d batch statistics, smp Conv2dReLU / torchvision blocks) */
int aadg_conv_fprop_stats_bf16(const void* x, int n, int h, int w, int cin, int ldx, const void* wgt, int cout, int r,
                               int s, int stride, int pad, int dil, void* y, int ho, int wo, int ldy, int y_c_off,
                               float* stat_sum, float* stat_sq, void* stream);

/* Data gradient of the same convolution: dx (+)= conv_transpose(dy, w).  wgt_t bf16 [R*S][cin][cout]
 * (channel axes swapped); all geometry arguments are the forward convolution's. */
int aadg_conv_dgrad_bf16(const void* dy, int n, int ho, int wo, int cout, int lddy, const void* wgt_t, int cin,
                         int r, int s, int stride, int pad, int dil, void* dx, int h, int w, int lddx,
                         int dx_c_off, int accumulate, void* stream);

/* Weight gradient: dw fp32 [R*S][cout][cin] += sum over pixels of dy (x) x  (zero dw for a fresh one). */
int aadg_conv_wgrad_bf16(const void* x, int n, int h, int w, int cin, int ldx, const void* dy, int ho, int wo,
                         int cout, int lddy, int r, int s, int stride, int pad, int dil, float* dw, void* stream);

/* "Window" convolutions (the ResNet stem, smp encoder conv1 = Conv2d(3, 64, 7, stride 2, padding 3) behind
 * models/__init__.py:17-23): a VALID, stride-1 convolution whose input pixel pitch `ldx` may be SMALLER than `cin` --
 * every input "pixel" is a window of `cin` consecutive elements overlapping its neighbours.  The caller picks the
 * output extent (ho <= h - r + 1, wo <= w - s + 1) and guarantees the reached windows lie inside the allocation (the
 * last ones end cin - ldx elements past the tensor's nominal end).  Over the space-to-depth image of aadg_stem_s2d
 * the 7x7 / stride-2 stem is r = 4, s = 1, cin = 64, ldx = 16.  stat_sum / stat_sq (both or neither): fused
 * batch-norm statistics as in aadg_conv_fprop_stats_bf16.  wgt bf16 [r*s][cout][cin]. */
int aadg_conv_fprop_windows_bf16(const void* x, int n, int h, int w, int cin, int ldx, const void* wgt, int cout, int r,
                                 int s, void* y, int ho, int wo, int ldy, float* stat_sum, float* stat_sq, void* stream);
/* its weight gradient: dw fp32 [r*s][cout][cin] += sum over output pixels of dy (x) window */
int aadg_conv_wgrad_windows_bf16(const void* x, int n, int h, int w, int cin, int ldx, const void* dy, int ho, int wo,
                                 int cout, int lddy, int r, int s, float* dw, void* stream);

/* ------------------------------------------------------------------------------------------------
 * HBM-bound layers of the segmentation net (bf16 NHWC, channel strides `ld*` in elements, channel
 * counts multiples of 8, <= 2048) — the cuDNN/ATen batch-norm, ReLU, pooling, up-sampling and
 * optimiser kernels behind smp.DeepLabV3Plus / torch.optim.Adam (models/__init__.py:17-23,
 * search_dg.py:132,170-172, scheduler.py:10-11).
 * ---------------------------------------------------------------------------------------------- */

/* sum[c] += sum_p x[p][c]; sumsq[c] += sum_p x[p][c]^2 (fp32; zero them first) */
int aadg_bn_stats(const void* x, long long pixels, int c, int ld, float* sum, float* sumsq, void* stream);
/* mean, 1/sqrt(var+eps), scale = gamma*invstd, shift = beta - mean*scale, running stats (may be NULL);
 * reset_sums != 0 zeroes sum / sumsq after reading them (ready for the next accumulation) */
int aadg_bn_finalize(float* sum, float* sumsq, const float* gamma, const float* beta, int c, float count,
                     float eps, float momentum, float* mean, float* invstd, float* scale, float* shift,
                     float* run_mean, float* run_var, int reset_sums, void* stream);
/* y = act(x*scale + shift (+ res)) ; flags bit0 = ReLU, bit5 (32) = ReLU6 (with bit0), bit1 = Dropout(0.5) keyed by
 * (seed, element); bit6 (64, here and in the backward calls): `seed` is the DEVICE ADDRESS of the 64-bit seed, read by
 * the kernel (a captured CUDA graph replays the launch while the host advances the seed in device memory);
 * relu_bits (may be NULL): uint8 [pixels][c/8], bit i of byte g = (pre-activation of channel 8g+i > 0) */
int aadg_bn_apply(const void* x, int ldx, const float* scale, const float* shift, const void* res, int ldr, void* y,
                  int ldy, long long pixels, int c, int flags, unsigned long long seed, void* relu_bits,
                  void* stream);
/* backward of aadg_bn_apply(training statistics): dgamma, dbeta (overwritten), dx, optional dres (+=).
 * flags bit2 (4): no residual was added, recompute the ReLU mask from x and the forward pass's `shift`
 * vector (y and ldy unused); flags bit3 (8): `y` points to the relu_bits written by aadg_bn_apply;
 * flags bit4 (16): dgamma / dbeta are NOT cleared first -- they must be zero on entry (a gradient buffer the caller
 * zeroed once per step), which saves two memset launches per layer */
int aadg_bn_backward(const void* dy, int lddy, const void* x, int ldx, const void* y, int ldy, const float* mean,
                     const float* invstd, const float* gamma, const float* shift, long long pixels, int c, int flags,
                     unsigned long long seed, float* dgamma, float* dbeta, void* dx, int lddx, void* dres, int lddr,
                     int dres_accumulate, void* stream);
/* the same backward with the incoming gradient given as dy + dy2 (two gradient branches meeting at a residual
 * block's input are added on load instead of by a read-modify-write pass) */
int aadg_bn_backward2(const void* dy, int lddy, const void* dy2, int lddy2, const void* x, int ldx, const void* y, int ldy,
                      const float* mean, const float* invstd, const float* gamma, const float* shift, long long pixels,
                      int c, int flags, unsigned long long seed, float* dgamma, float* dbeta, void* dx, int lddx,
                      void* dres, int lddr, int dres_accumulate, void* stream);
/* The two halves of aadg_bn_backward / aadg_bn_backward2 for batch statistics shared across ranks (the SyncBN option
 * of SURVEY.md §8e(3); torch.nn.SyncBatchNorm semantics): `reduce` ACCUMULATES this rank's sum(g*xhat) / sum(g) into
 * dgamma_sum / dbeta_sum (fp32 [c], zero them first); the caller all-reduces them; `apply` takes the global sums and
 * inv_count = 1 / (global pixel count) and writes dx (and dres).  dy2 may be NULL. */
int aadg_bn_backward_reduce(const void* dy, int lddy, const void* dy2, int lddy2, const void* x, int ldx, const void* y,
                            int ldy, const float* mean, const float* invstd, const float* gamma, const float* shift,
                            long long pixels, int c, int flags, unsigned long long seed, float* dgamma_sum,
                            float* dbeta_sum, void* stream);
int aadg_bn_backward_apply(const void* dy, int lddy, const void* dy2, int lddy2, const void* x, int ldx, const void* y,
                           int ldy, const float* mean, const float* invstd, const float* gamma, const float* shift,
                           long long pixels, int c, int flags, unsigned long long seed, const float* dgamma_sum,
                           const float* dbeta_sum, float inv_count, void* dx, int lddx, void* dres, int lddr,
                           int dres_accumulate, void* stream);
int aadg_add_bf16(void* a, int lda, const void* b, int ldb, long long pixels, int c, void* stream);
/* MaxPool2d(3, stride 2, padding 1): argmax uint8 [n,ho,wo,c] */
int aadg_maxpool3x3s2_fwd(const void* x, int n, int h, int w, int c, void* y, void* argmax, void* stream);
int aadg_maxpool3x3s2_bwd(const void* dy, const void* argmax, int n, int h, int w, int c, void* dx, void* stream);
/* UpsamplingBilinear2d (align_corners=True) and its transpose */
int aadg_upsample_bilinear_fwd(const void* x, int n, int h, int w, int c, int ldx, void* y, int ho, int wo, int ldy,
                               void* stream);
int aadg_upsample_bilinear_bwd(const void* dy, int n, int ho, int wo, int c, int lddy, void* dx, int h, int w, int lddx,
                               void* stream);
/* nearest x2 up-sampling into a channel slice (smp Unet DecoderBlock) and its transpose; plain slice copy */
int aadg_upsample_nearest2x_fwd(const void* x, int n, int h, int w, int c, int ldx, void* y, int ldy, void* stream);
int aadg_upsample_nearest2x_bwd(const void* dy, int n, int h, int w, int c, int lddy, void* dx, int lddx, void* stream);
int aadg_copy_bf16(const void* x, int ldx, void* y, int ldy, long long pixels, int c, void* stream);
/* out fp32 [n,c] = scale * sum over the hw pixels of x bf16 [n,hw,ld] (AdaptiveAvgPool2d(1) with scale = 1/hw) */
int aadg_global_sum(const void* x, int n, int hw, int c, int ld, float* out, float scale, void* stream);
/* y[n, p, 0:c] = v[n, 0:c]  (bilinear resize of a 1x1 map) */
int aadg_broadcast_pixels(const void* v, int n, int c, void* y, int hw, int ldy, void* stream);
/* y bf16 [n, hw, ldy] (c channels) += scale * v fp32 [n, c] broadcast over the pixels (backward of the pooled branch) */
int aadg_broadcast_add_pixels(const float* v, int n, int c, void* y, int hw, int ldy, float scale, void* stream);
int aadg_f32_to_bf16(const float* x, void* y, long long count, float scale, void* stream);
/* depthwise 3x3, stride 1, padding = dilation; w fp32 [9][c]; direction 0 forward, 1 data gradient */
/* (direction bit 1 (value 2): y += result) */
int aadg_dwconv3x3(const void* x, int n, int h, int w, int c, int ldx, const float* wgt, int dil, int direction, void* y,
                   int ldy, void* stream);
int aadg_dwconv3x3_wgrad(const void* x, int n, int h, int w, int c, int ldx, const void* dy, int lddy, int dil, float* dw,
                         void* stream);
/* strided depthwise 3x3 (padding = dilation; ho = (h-1)/stride + 1): MobileNetV2's down-sampling blocks.  direction 1:
 * `x` is dy [n,ho,wo,ldx] and `y` is dx [n,h,w,ldy].  stride 1 forwards to the kernels above. */
int aadg_dwconv3x3_strided(const void* x, int n, int h, int w, int c, int ldx, const float* wgt, int dil, int stride,
                           int direction, void* y, int ho, int wo, int ldy, void* stream);
int aadg_dwconv3x3_strided_wgrad(const void* x, int n, int h, int w, int c, int ldx, const void* dy, int ho, int wo, int lddy,
                                 int dil, int stride, float* dw, void* stream);
/* img fp32 [n,3,h,w] -> col bf16 [n*ho*wo][kp], k = (r*S+s)*3 + c, zero padded to kp */
int aadg_im2col_stem(const float* img, int n, int h, int w, int r, int s, int stride, int pad, int kp, void* col,
                     void* stream);
/* the same patches in a row-pitched layout k = r*row_pitch + s*3 + c (row_pitch % 8 == 0, >= 3*s; kp % 8 == 0,
 * >= r*row_pitch; even stride): every filter row starts on a 16-byte boundary, the pass is a run copy */
int aadg_im2col_stem_rows(const float* img, int n, int h, int w, int r, int s, int stride, int pad, int row_pitch,
                          int kp, void* col, void* stream);
/* stem input in space-to-depth form (replaces the im2col patch buffer for the 7x7 / stride-2 stem): img fp32
 * [n,3,h,w] (h, w even; the Normalize_dg / ToTensor output of data/transform.py:149-236) -> out bf16
 * [n, h/2+3, w/2+3, 16]: s2d pixel (Y+2, X+2) holds input pixels (2Y+py, 2X+px) at channel (py*2+px)*3 + c, channels
 * 12..15 and the border (two rows / columns before, one after) are zero.  `out` needs 64 spare elements at its end. */
int aadg_stem_s2d(const float* img, int n, int h, int w, void* out, void* stream);
/* torch.optim.Adam step `step` (1-based) over flat fp32 buffers */
int aadg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long count, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int step, void* stream);
/* the same step with its scalars in device memory: hyper = float32 {lr, beta1, beta2, eps, weight_decay, grad_scale}
 * (grad_scale multiplies the gradient on load: 1/world after a summing all-reduce, models/__init__.py:39 DDP averaging),
 * step = int64 1-based step count.  Nothing is passed by value, so a captured CUDA graph replays it unchanged. */
int aadg_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long count,
                       const float* hyper, const long long* step, void* stream);
/* fp32 master weights [taps][cout][cin] -> bf16 copy (+ transposed [taps][cin][cout] copy); descs (device):
 * per weight { int64 off_master, off_bf16, off_bf16_t (-1: none); int32 taps, cout, cin, 0 } */
int aadg_weight_prep(const float* master, void* w_bf16, void* w_bf16_t, const void* descs, int n_descs, void* stream);

/* Segmentation head + loss (search_dg.py:140-142,164-165; losses.py:21-23; smp SegmentationHead):
 * logits z fp32 [pixels][classes] at decoder resolution; classes <= 2 */
int aadg_seg_head_fwd(const void* a, long long pixels, int c, int lda, const float* w, const float* bias, int classes,
                      float* z, void* stream);
/* upsample(align_corners=True) -> sigmoid -> BCELoss sum (double, +=) and TP/FP/FN counts int32
 * [n][classes][3] (+=) at threshold thr; logits_out fp32 [n,classes,H,W] or NULL; target fp32 [n,classes,H,W] */
int aadg_seg_loss_fwd(const float* z, int n, int h, int w, int classes, const float* target, int H, int W, float thr,
                      double* loss_sum, int* counts, float* logits_out, void* stream);
int aadg_seg_loss_bwd(const float* z, int n, int h, int w, int classes, const float* target, int H, int W,
                      float grad_scale, float* dz, void* stream);
/* transpose of the head's UpsamplingBilinear2d for an ARBITRARY gradient: dz fp32 [n,h,w,classes] from dlogits fp32
 * [n,classes,H,W] -- what `seg_loss.backward()` (search_dg.py:170) sends into the network when the caller computes its
 * own loss on `model(x)`'s logits (the autograd surface of aadg_b200.nn) */
int aadg_upsample_logits_bwd(const float* dlogits, int n, int h, int w, int classes, int H, int W, float* dz, void* stream);
int aadg_seg_head_bwd(const float* dz, const void* a, long long pixels, int c, int lda, const float* w, int classes,
                      void* da, int ldda, float* dw, float* db, void* stream);

/* smp Unet SegmentationHead: Conv2d(c, classes, 3, padding=1) at full resolution, c <= 64, classes <= 2.
 * w fp32 [classes][9][c]; z fp32 [n,h,w,classes]; backward: da bf16, dw (+=), db (+=) */
int aadg_seg_head3x3_fwd(const void* a, int n, int h, int w, int c, int lda, const float* wgt, const float* bias,
                         int classes, float* z, void* stream);
int aadg_seg_head3x3_bwd(const float* dz, const void* a, int n, int h, int w, int c, int lda, const float* wgt,
                         int classes, void* da, int ldda, float* dw, float* db, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Validation metric — replaces medpy.metric.binary.hd95 as validate() calls it per image and class on
 * the CPU (search_dg.py:246-260, train_dg.py): surfaces = mask xor 4-neighbour erosion, exact Euclidean
 * distances to the other surface, numpy.percentile(.., 95) of the union of both directed sets.
 * ---------------------------------------------------------------------------------------------- */
size_t aadg_hd95_workspace_bytes(int n_pairs, int h, int w);
/* result, reference: uint8 [n_pairs][h][w] (non-zero = foreground); out float64 [n_pairs] in pixels;
 * status int32 [n_pairs]: 0 ok, 1 = `result` empty, 2 = `reference` empty (medpy raises; out = NaN).
 * `percentile` in [0,100] (95 for hd95, 100 = the Hausdorff distance). */
int aadg_hd95(const unsigned char* result, const unsigned char* reference, int n_pairs, int h, int w, double percentile,
              double* out, int* status, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Policy controller — replaces the per-decision LSTMCell / Linear / softmax / multinomial launches of
 * models/controller.py:73-145 (sample, evaluate) and the autograd backward of the PPO update
 * (losses.py:127-157) with one launch each.  `params` / `grads`: HOST arrays of nine device pointers in module
 * order: embedding.weight [n_ops+n_mags, e], lstm.weight_ih [4h, e], lstm.weight_hh [4h, h], lstm.bias_ih [4h],
 * lstm.bias_hh [4h], outop.weight [n_ops, h], outop.bias, outmag.weight [n_mags, h], outmag.bias.
 * ---------------------------------------------------------------------------------------------- */
/* mode 0 = sample (policies int64 [batch, q*l*2] written; decision t of row m draws from Philox4x32-10 keyed by
 * `seed` with counter (m, t, call)), mode 1 = evaluate (policies read).  step_log_prob, step_entropy float32
 * [batch, q*l*2]; step_probs float32 [batch, q*l*2, max(n_ops, n_mags)] or NULL; saved float32 [batch, q*l*2, 6h]
 * (activations for the backward) or NULL. */
int aadg_controller_walk(const float* const* params, int n_ops, int n_mags, int q, int l, int e, int h, float c, float t,
                         int batch, int mode, unsigned long long seed, unsigned long long call, long long* policies,
                         float* step_log_prob, float* step_entropy, float* step_probs, float* saved, void* stream);
/* grads[i] += d( sum_m grad_log_prob[m] * sum_t log pi(a_t | m) ) / d params[i] (fp32, accumulated) */
int aadg_controller_backward(const float* const* params, int n_ops, int n_mags, int q, int l, int e, int h, float c,
                             float t, int batch, const long long* policies, const float* saved,
                             const float* grad_log_prob, float* const* grads, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AADG_B200_H */
