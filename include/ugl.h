/*
 * ugl.h — C-ABI of the B200-native photometric view-synthesis loss library (libugl_b200.so).
 *
 * The reference (jianfenglihg/Unsupervised_depth_OpticalFlow_egomotion) is pure Python/PyTorch and
 * has no FFI boundary of its own; the functions below are the entry points a binding for its loss
 * path would call.  Each one cites the reference callable (file:line under core/networks/) whose
 * device work it replaces.  See INTEGRATION.md for the ctypes binding and the drop-in Python layer.
 *
 * Conventions
 *  - all tensors are fp32, NCHW, contiguous, DEVICE pointers owned by the caller (the library never
 *    allocates, frees or retains pointers); workspaces are caller-provided;
 *  - every call is an asynchronous launch on `stream` (a cudaStream_t passed as void*), performs no
 *    host synchronisation and no D2H copy, and is CUDA-graph capturable;
 *  - return value: 0 on success, a negative UGL_E* for bad arguments, a positive cudaError_t if a
 *    launch failed; ugl_last_error() returns a thread-local description;
 *  - results are deterministic (bit-reproducible run to run): no floating-point atomics.
 */
#ifndef UGL_H_
#define UGL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UGL_VERSION 100          /* 0.1.0 */
#define UGL_MAX_LEVELS 6
#define UGL_FLOW_NSTATS 12       /* per-(sample, level) sums kept between forward and backward */

#define UGL_OK 0
#define UGL_EINVAL (-1)          /* null pointer / bad shape */
#define UGL_EALIGN (-2)          /* pointer not 4-byte aligned */
#define UGL_EWORKSPACE (-3)      /* workspace too small */
#define UGL_EUNSUPPORTED (-4)    /* shape outside the supported range */

int ugl_version(void);
const char* ugl_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Fused flow-mode loss — replaces the loss body of Model_flow.forward (model_flow.py:232-254):
 * warp_flow_pyramid (:66-70 -> structures/net_utils.py:16-54), compute_diff_weight (:105-138),
 * compute_loss_with_mask (:94-103), compute_loss_ssim (:141-152 -> pytorch_ssim/ssim.py:4-19),
 * compute_loss_flow_smooth (:173-181) and compute_loss_flow_consis (:184-199), for all pyramid
 * levels of a batch in one launch (+ a finalize launch).
 *
 *  loss  (4,B): rows = loss_flow_pixel, loss_flow_ssim, loss_flow_smooth, loss_flow_consis
 *  stats (B,scales,UGL_FLOW_NSTATS): written by forward, read by backward
 *  backward: grad_loss (4,B) -> grad_flow_fwd/bwd[l] (B,2,h_l,w_l) for l < scales
 * ------------------------------------------------------------------------------------------- */
typedef struct UglFlowLossArgs {
  int32_t batch;
  int32_t levels;                              /* entries filled in the arrays below            */
  int32_t scales;                              /* levels that carry a loss (num_scales) <= levels */
  int32_t height[UGL_MAX_LEVELS];
  int32_t width[UGL_MAX_LEVELS];
  const float* img_l[UGL_MAX_LEVELS];          /* (B,3,h,w) left-frame pyramid                   */
  const float* img[UGL_MAX_LEVELS];            /* (B,3,h,w) centre-frame pyramid                 */
  const float* img_r[UGL_MAX_LEVELS];          /* (B,3,h,w) right-frame pyramid                  */
  const float* flow_fwd[UGL_MAX_LEVELS];       /* (B,2,h,w) centre -> right                      */
  const float* flow_bwd[UGL_MAX_LEVELS];       /* (B,2,h,w) centre -> left                       */
  float* loss;                                 /* (4,B)                                          */
  float* stats;                                /* (B,scales,UGL_FLOW_NSTATS)                     */
  const float* grad_loss;                      /* (4,B)            [backward]                    */
  float* grad_flow_fwd[UGL_MAX_LEVELS];        /* (B,2,h,w)        [backward]                    */
  float* grad_flow_bwd[UGL_MAX_LEVELS];        /* (B,2,h,w)        [backward]                    */
  void* workspace;
  uint64_t workspace_bytes;
  void* stream;                                /* cudaStream_t                                   */
  float* basis[UGL_MAX_LEVELS];                /* (B,14,h,w) gradient basis maps  [single-pass mode] */
} UglFlowLossArgs;

uint64_t ugl_flow_loss_workspace_bytes(const UglFlowLossArgs* args);
int ugl_flow_loss_forward(const UglFlowLossArgs* args);
int ugl_flow_loss_backward(const UglFlowLossArgs* args);
/* number of kernel launches one forward / backward call issues (for launch accounting) */
int ugl_flow_loss_launches(int backward);
/* Single-pass mode: forward_grad computes the losses AND the un-normalised gradient of every term w.r.t. both
 * flows into basis[l] (B,UGL_FLOW_BASIS_PLANES,h,w) in one stencil kernel (+ finalize); combine is the element-wise
 * backward  grad = sum_k scale_k(sample, level) * basis_k.  Same results as forward + backward (recompute) with the
 * photometry / SSIM work executed once instead of twice; costs 14 floats per pixel of saved state. */
#define UGL_FLOW_BASIS_PLANES 14
int ugl_flow_loss_forward_grad(const UglFlowLossArgs* args);
/* The single-pass forward exists in several internal forms with the same interface and the same results (tests cross-check
 * them); ugl_flow_loss_forward_grad / ugl_geom_flow_forward_grad use UGL_SINGLE_PASS_SPLIT.
 *   FUSED      one tile kernel: photometry on the tile + 2-pixel halo, then the stencils, all in shared memory (round-1 kernel)
 *   SPLIT      photometry kernel (one thread per pixel, no halo) -> photometry planes in the workspace -> stencil kernel that
 *              stages them with TMA (cp.async.bulk.tensor, zero fill outside the image) when every level's width is a
 *              multiple of 4 and the bases are 16-byte aligned, with plain loads otherwise
 *   SPLIT_PLAIN / SPLIT_TMA   force the staging form (SPLIT_TMA fails with UGL_EUNSUPPORTED where TMA is not possible) */
#define UGL_SINGLE_PASS_FUSED 0
#define UGL_SINGLE_PASS_SPLIT 1
#define UGL_SINGLE_PASS_SPLIT_PLAIN 2
#define UGL_SINGLE_PASS_SPLIT_TMA 3
int ugl_flow_loss_forward_grad_ex(const UglFlowLossArgs* args, int32_t variant);
/* Fused forward + backward for a training step, where the upstream gradient is known before the forward runs (train.py:211-215:
 * d total / d loss_k[b] = w_k / B): loss (4,B) AND grad_flow_fwd/bwd[l] in four launches (photometry kernel, weight sums, stencil
 * kernel, finalize) chained by programmatic dependent launch on args->stream (the stencil kernel starts under the weight sums and
 * waits for them before its last phase; stream order towards other work is unchanged).  The per-sample normalisers only depend on the photometry kernel's sums, so the stencil kernel scales and
 * adds the four gradient terms itself: no basis planes (args->basis is ignored), no combine launch.  Same results as
 * ugl_flow_loss_forward_grad + ugl_flow_loss_combine. */
int ugl_flow_loss_step(const UglFlowLossArgs* args);
/* The same step launching only the selected kernels (bit mask), in order -- for per-kernel timing with events around each
 * launch (bench.py's roofline); the kernels of a later part read what the earlier parts of a previous call left in the workspace. */
#define UGL_STEP_PHOTO 1
#define UGL_STEP_NORM 2
#define UGL_STEP_STENCIL 4
#define UGL_STEP_FINALIZE 8
#define UGL_STEP_ALL 15
int ugl_flow_loss_step_parts(const UglFlowLossArgs* args, int32_t parts);
int ugl_flow_loss_combine(const UglFlowLossArgs* args);

/* ---------------------------------------------------------------------------------------------
 * Flow branch of the geom-mode loss — replaces, per level of Model_geometry.forward's loop (model_geometry.py:845-919):
 *   warp_flow x2 (:847-848), compute_occ_weight (:850 -> :105-132, hard masks [1 - softmax > 0.48]), calculate_rigid_flow
 *   + compute_dynamic_mask (:866-870 -> :698-707), the masked L1 terms split by the dynamic mask with weights 1 / 2
 *   (:905-908), the masked SSIM terms (:910-911), compute_loss_flow_smooth (:913-914) and compute_loss_flow_consis with
 *   mask 1 - occ_fwd (:917-918).  Same single-pass structure as ugl_flow_loss_forward_grad / _combine (gradients flow to
 *   the two flows only: every mask is detached in the reference).
 *   flow.loss (4,B): flow_pixel, flow_ssim, flow_smooth, flow_consis summed over levels;  flow.stats (B,scales,UGL_GEOM_NSTATS)
 *   mask_bytes[l] (B,h,w) uint8 OUTPUT (read again by _combine and by ugl_depth_photo_*): bit0 valid_bwd, bit1 valid_fwd,
 *   bit2 occ_bwd, bit3 occ_fwd, bit4 dyn_bwd, bit5 dyn_fwd  (bwd = centre->left flow, fwd = centre->right flow).
 *   disp[l] (B,1,h,w), Kinv[l] (B,3,3), P_bwd/P_fwd[l] (B,3,4) = K_l [R|t] of the centre->left / centre->right pose.
 * ------------------------------------------------------------------------------------------- */
#define UGL_GEOM_NSTATS 16
#define UGL_MASK_VALID_BWD 1
#define UGL_MASK_VALID_FWD 2
#define UGL_MASK_OCC_BWD 4
#define UGL_MASK_OCC_FWD 8
#define UGL_MASK_DYN_BWD 16
#define UGL_MASK_DYN_FWD 32
typedef struct UglGeomFlowArgs {
  UglFlowLossArgs flow;                        /* images, flows, loss, stats, grads, workspace, stream, basis */
  const float* disp[UGL_MAX_LEVELS];
  const float* Kinv[UGL_MAX_LEVELS];
  const float* P_bwd[UGL_MAX_LEVELS];
  const float* P_fwd[UGL_MAX_LEVELS];
  uint8_t* mask_bytes[UGL_MAX_LEVELS];
  float alpha, beta;                           /* flow_consist_alpha / flow_consist_beta */
} UglGeomFlowArgs;
int ugl_geom_flow_forward_grad(const UglGeomFlowArgs* args);
int ugl_geom_flow_forward_grad_ex(const UglGeomFlowArgs* args, int32_t variant);   /* variant: UGL_SINGLE_PASS_* */
int ugl_geom_flow_combine(const UglGeomFlowArgs* args);
/* Fused forward + backward of the geom-mode flow branch for a training step (the counterpart of ugl_flow_loss_step; train.py:211-215:
 * d total / d loss_k[b] = w_k / B known before the forward runs): flow.loss (4,B), mask_bytes AND flow.grad_flow_fwd/bwd[l] in four
 * chained launches; flow.basis is ignored, no combine launch.  Same results as ugl_geom_flow_forward_grad + ugl_geom_flow_combine. */
int ugl_geom_flow_step(const UglGeomFlowArgs* args);
/* The same step in two calls: parts = UGL_STEP_PHOTO (photometry kernel: mask_bytes are final when it completes), then
 * parts = UGL_STEP_ALL & ~UGL_STEP_PHOTO (weight sums, stencil kernel, finalize).  Lets the caller start the consumers of the mask bytes
 * (ugl_depth_photo_*, ugl_geom_rigid_*) on other streams while the stencil kernel runs.  Same results as ugl_geom_flow_step. */
int ugl_geom_flow_step_parts(const UglGeomFlowArgs* args, int32_t parts);

/* ---------------------------------------------------------------------------------------------
 * Image pyramid — replaces generate_img_pyramid: model_flow.py:58-64 (mode 0: adaptive average
 * pooling = 2^s x 2^s box mean) and model_geometry.py:65-72 / model_depth.py:44-50 (mode 1:
 * bilinear, align_corners=False = mean of the central 2x2 of every 2^s block), and the 'area'
 * resize of reconstruction (model_geometry.py:91) which equals mode 0.
 * out[l] is (B,C,H>>l,W>>l) for l = 1..levels-1 (level 0 is the input itself; out[0] is ignored).
 * H and W must be divisible by 2^(levels-1).
 * ------------------------------------------------------------------------------------------- */
int ugl_image_pyramid(const float* img, int32_t batch, int32_t channels, int32_t height, int32_t width,
                      int32_t levels, int32_t mode, float* const* out, void* stream);

/* The same pyramids for up to three images, both modes and all levels (2..4) in ONE launch: box[i][l] / bil[i][l] are the
 * level-l outputs (B,C,H>>l,W>>l) of image i in mode 0 / mode 1, NULL = not wanted.  H*W planes must be 16-byte aligned
 * (W divisible by 2^(levels-1) and a 16-byte aligned base).  Same per-output arithmetic as ugl_image_pyramid. */
#define UGL_PYRAMID_MAX_IMAGES 3
typedef struct UglPyramidArgs {
  int32_t batch, channels, height, width, levels, images;
  const float* img[UGL_PYRAMID_MAX_IMAGES];
  float* box[UGL_PYRAMID_MAX_IMAGES][UGL_MAX_LEVELS];
  float* bil[UGL_PYRAMID_MAX_IMAGES][UGL_MAX_LEVELS];
  void* stream;
} UglPyramidArgs;
int ugl_image_pyramid_multi(const UglPyramidArgs* args);

/* ---------------------------------------------------------------------------------------------
 * warp_flow — structures/net_utils.py:16-54.  out = grid_sample(x, (j+u, i+v)) [* keep mask].
 * backward: grad_flow (B,2,H,W) always; grad_x (B,C,H,W) if non-null (deterministic: fixed-point
 * accumulation).  mask (B,1,H,W) optional output of the {0,1} keep map.
 * ------------------------------------------------------------------------------------------- */
int ugl_warp_flow_forward(const float* x, const float* flow, int32_t batch, int32_t channels, int32_t height,
                          int32_t width, int32_t use_mask, float* out, float* mask, void* stream);
int ugl_warp_flow_backward(const float* x, const float* flow, const float* grad_out, int32_t batch, int32_t channels,
                           int32_t height, int32_t width, int32_t use_mask, float* grad_flow, float* grad_x,
                           void* workspace, uint64_t workspace_bytes, void* stream);
/* scatter form of grad_x (same bits either way; the global form is kept as the cross-check): TILE_LOCAL accumulates a CTA's taps in
 * a shared-memory window (tile + 8 pixels) and touches each global cell once; GLOBAL issues one 64-bit global atomic per tap corner. */
#define UGL_SCATTER_TILE_LOCAL 0
#define UGL_SCATTER_GLOBAL 1
int ugl_warp_flow_backward_ex(const float* x, const float* flow, const float* grad_out, int32_t batch, int32_t channels,
                              int32_t height, int32_t width, int32_t use_mask, float* grad_flow, float* grad_x,
                              void* workspace, uint64_t workspace_bytes, int32_t scatter, void* stream);
uint64_t ugl_warp_flow_backward_workspace_bytes(int32_t batch, int32_t channels, int32_t height, int32_t width,
                                                int32_t need_grad_x);

/* PWC-Net cost volume — replaces PWC_tf.corr_naive (structures/pwc_tf.py:97-106; SURVEY 8(f) rank 2, the caller of warp_flow at
 * pwc_tf.py:94-95, 121-160): out (B,(2d+1)^2,H,W), out[b, i*(2d+1)+j, y, x] = mean_c f1[b,c,y,x] * f2pad[b,c,y+i,x+j] with f2 zero-padded
 * by d (1 <= d <= 4; the reference uses 4).  backward: grad_out -> grad_f1 / grad_f2 (B,C,H,W), either may be NULL; gather form,
 * deterministic. */
int ugl_cost_volume_forward(const float* f1, const float* f2, int32_t batch, int32_t channels, int32_t height, int32_t width, int32_t d,
                            float* out, void* stream);
int ugl_cost_volume_backward(const float* f1, const float* f2, const float* grad_out, int32_t batch, int32_t channels, int32_t height,
                             int32_t width, int32_t d, float* grad_f1, float* grad_f2, void* stream);

/* EXTENSION — forward splat (`transformerFwd`): Model_flow.get_occlusion_mask_from_flow (model_flow.py:33-39) calls it but
 * the reference never defines it (dead code; upstream TrianFlow semantics, parity unpinned).  out[b,c,y',x'] accumulates
 * x[b,c,i,j] * bilinear weight over the four integer neighbours of (j+u, i+v); out-of-range corners are dropped;
 * clamp01 != 0 applies clamp(., 0, 1) (:37-38).  Deterministic.  Not differentiable. */
uint64_t ugl_forward_splat_workspace_bytes(int32_t batch, int32_t channels, int32_t height, int32_t width);
int ugl_forward_splat(const float* x, const float* flow, int32_t batch, int32_t channels, int32_t height, int32_t width,
                      int32_t clamp01, float* out, void* workspace, uint64_t workspace_bytes, void* stream);
int ugl_forward_splat_ex(const float* x, const float* flow, int32_t batch, int32_t channels, int32_t height, int32_t width,
                      int32_t clamp01, float* out, void* workspace, uint64_t workspace_bytes, int32_t scatter, void* stream);

/* Dataset glue: uint8 frames -> fp32 frames in [0,1], the arithmetic of core/dataset/kitti_prepared.py:89
 * (`img / 255.0` in float64, then `.float()`): for every byte value the result equals the correctly rounded fp32 quotient this
 * computes.  Lets a trainer ship the frames over PCIe as bytes (a quarter of the traffic).  `frames` (1..4) device pointers of
 * `elements` bytes / floats each, 16-byte aligned. */
int ugl_frames_u8_to_float(const void* const* src, void* const* dst, int32_t frames, uint64_t elements, void* stream);

/* Device self-test of the packed fp32 pair arithmetic (FADD2 / FMUL2 / FFMA2) the single-pass kernels run their SSIM stencil on:
 * `windows_per_thread` random 3x3 windows per thread (blocks x 128 threads, two windows per pair) are evaluated with the packed
 * pair functions and with the scalar functions of the per-method / recompute kernels; mismatch[2][14] (device, uint64) counts,
 * per pair lane, the windows whose sx, sy, sxx, syy, sxy, mu_x, n1, n2, d1, d2, S, cA, cB, cC differ in ANY bit.  All zeros is
 * the contract (ptxas contracts packed .rn products into sums unless prevented, see ugl_common.cuh: acc2_rn). */
int ugl_selftest_packed_pairs(uint64_t* mismatch, int32_t blocks, int32_t windows_per_thread, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Generic per-sample reductions: workspace for every *_forward / *_backward below that takes one
 * (>= ugl_reduce_workspace_bytes(B,H,W) bytes, 8-byte aligned).
 * ------------------------------------------------------------------------------------------- */
uint64_t ugl_reduce_workspace_bytes(int32_t batch, int32_t height, int32_t width);

/* P(d, m) = mean_{c,h,w}(d*m) / (mean_{h,w}(m) + 1e-12), one pyramid level -> out (B,), den (B,) kept
 * for backward.  mode 0: d = |a - b| — compute_photometric_loss (model_geometry.py:143-153,
 * model_depth.py:92-103), compute_loss_pixel (model_flow.py:72-81).  mode 1: d = a —
 * compute_loss_with_mask (model_flow.py:94-103), compute_depth_flow_consis_loss
 * (model_geometry.py:716-732), compute_consis_loss (model_geometry.py:182-193), and with
 * mask == NULL the plain mean of compute_epipolar_loss (model_geometry.py:413-418).
 * mask is (B,1,H,W) and is repeated over the C channels of d. */
int ugl_masked_mean_forward(const float* a, const float* b, const float* mask, int32_t batch, int32_t channels, int32_t height,
                            int32_t width, int32_t mode, float* out, float* den, void* workspace, uint64_t workspace_bytes,
                            void* stream);
int ugl_masked_mean_backward(const float* a, const float* b, const float* mask, const float* den, const float* grad_out,
                             int32_t batch, int32_t channels, int32_t height, int32_t width, int32_t mode, float* grad_a,
                             float* grad_b, void* stream);

/* compute_occ_weight (model_geometry.py:105-132, soft = 0: [1 - softmax > 0.48]) and
 * compute_diff_weight (model_flow.py:105-138, soft = 1: 2 exp(-(w-0.5)^2/0.03) * valid); all outputs
 * (B,1,H,W); diff_* = mean_c |img - from_*| may be NULL.  The diffs are differentiable w.r.t. the
 * warped images through ugl_channel_mean_abs_diff_backward. */
int ugl_occlusion_weights(const float* from_l, const float* img, const float* from_r, int32_t batch, int32_t height, int32_t width,
                          int32_t soft, float* w_bwd, float* w_fwd, float* valid_bwd, float* valid_fwd, float* diff_bwd,
                          float* diff_fwd, void* stream);
int ugl_channel_mean_abs_diff_backward(const float* img, const float* warped, const float* grad_diff, int32_t batch,
                                       int32_t channels, int32_t height, int32_t width, float* grad_warped, void* stream);

/* compute_texture_mask (model_geometry.py:134-140, model_depth.py:84-90) */
int ugl_texture_mask(const float* img, const float* rec, const float* src, int32_t batch, int32_t height, int32_t width, float* mask,
                     void* stream);

/* compute_dynamic_mask body (model_geometry.py:698-711) given the rigid flow: flow_diff = |rf - f| (B,2,H,W),
 * dyn = [n(fd)^2 < alpha (n(f)^2 + n(rf)^2) + beta], score = 1/(1e-4 + n(fd)) (may be NULL).
 * ugl_abs_diff_backward is the backward of d = |a - b| (element-wise). */
int ugl_dynamic_mask_forward(const float* flow, const float* rigid_flow, int32_t batch, int32_t height, int32_t width, float alpha,
                             float beta, float* flow_diff, float* dyn_mask, float* score, void* stream);
int ugl_abs_diff_backward(const float* a, const float* b, const float* grad_out, int64_t n, float* grad_a, float* grad_b, void* stream);

/* fusion_mask / fusion_mask_2item / fusion_mask_4item (model_geometry.py:735-765, model_depth.py:262-269):
 * product of up to 4 maps, optionally inverted (1 - m).  get_rigid_mask (model_geometry.py:420-425). */
int ugl_mask_product(const float* const* masks, const int32_t* invert, int32_t n_masks, int64_t n, float* out, void* stream);
int ugl_rigid_mask(const float* dist, int64_t n, float rigid_thres, float inlier_thres, float* rigid, float* inlier, float* score,
                   void* stream);

/* Step glue of the mode assemblies (T0; model_geometry.py:883-951 + train.py:211-214) without library kernels:
 *   ugl_accumulate_multi      dst[i] += src[3 i] (+ src[3 i + 1] + src[3 i + 2], null = absent) for up to 24 distinct destinations in one
 *                             launch -- the gradient accumulation autograd performs with one `add` launch per contribution when a
 *                             tensor (disparity, flow, K[R|t]) feeds several loss terms; `src` holds 3 n pointers;
 *   ugl_weighted_total_*      total = sum_k weights[k] * mean_b loss[k][b] over a (terms, batch) matrix, and grad_loss[k][b] =
 *                             grad_out * weights[k] / batch;
 *   ugl_assemble_rows         dst[i][b] = sum_{r < nsum[i]} src[i][r * batch + b]: the (B,) outputs of the per-term kernels -> rows of
 *                             one (n, batch) loss matrix (nsum > 1: the three compute_smooth_loss calls, summed in call order). */
int ugl_accumulate_multi(float* const* dst, const float* const* src, const int64_t* numel, int32_t n, void* stream);
int ugl_assemble_rows(const float* const* src, const int32_t* nsum, int32_t n, int32_t batch, float* dst, void* stream);
int ugl_weighted_total_forward(const float* loss, const float* weights, int32_t terms, int32_t batch, float* out, void* stream);
int ugl_weighted_total_backward(const float* grad_out, const float* weights, int32_t terms, int32_t batch, float* grad_loss, void* stream);

/* cal_grad2_error(flow/20, img) for one level (model_geometry.py:254-279, model_flow.py:156-181) -> (B,) */
int ugl_flow_smooth_forward(const float* flow, const float* img, int32_t batch, int32_t height, int32_t width, float* out,
                            void* workspace, uint64_t workspace_bytes, void* stream);
int ugl_flow_smooth_backward(const float* flow, const float* img, const float* grad_out, int32_t batch, int32_t height, int32_t width,
                             float* grad_flow, void* stream);

/* compute_loss_flow_consis for one level (model_geometry.py:195-210, model_flow.py:184-199): mask = 1 - occ,
 * gradient reaches the forward flow only. */
int ugl_flow_consis_forward(const float* fwd, const float* bwd, const float* occ, int32_t batch, int32_t height, int32_t width,
                            float* out, float* den, void* workspace, uint64_t workspace_bytes, void* stream);
int ugl_flow_consis_backward(const float* fwd, const float* bwd, const float* occ, const float* den, const float* grad_out,
                             int32_t batch, int32_t height, int32_t width, float* grad_fwd, void* stream);

/* depth difference map of compute_consis_loss (model_geometry.py:186-188, model_depth.py:158-160):
 * clamp(|comp - proj| / |comp + proj|, 0, 1), element-wise. */
int ugl_depth_diff_forward(const float* comp, const float* proj, int64_t n, float* out, void* stream);
int ugl_depth_diff_backward(const float* comp, const float* proj, const float* grad_out, int64_t n, float* grad_comp, float* grad_proj,
                            void* stream);

/* compute_smooth_loss (model_geometry.py:225-252, model_depth.py:220-247): all `levels` disparity maps
 * (B,1,h_l,w_l) are bilinearly up-sampled to (H,W) inside the kernel; out (B,) = sum over levels. */
int ugl_disp_smooth_forward(const float* img, const float* const* disps, const int32_t* heights, const int32_t* widths, int32_t levels,
                            int32_t batch, int32_t height, int32_t width, float* out, void* workspace, uint64_t workspace_bytes,
                            void* stream);
uint64_t ugl_disp_smooth_backward_workspace_bytes(int32_t batch, int32_t height, int32_t width);
int ugl_disp_smooth_backward(const float* img, const float* const* disps, const int32_t* heights, const int32_t* widths, int32_t levels,
                             const float* grad_out, int32_t batch, int32_t height, int32_t width, float* const* grad_disps,
                             void* workspace, uint64_t workspace_bytes, void* stream);

/* Single-pass form of the same term for up to UGL_DISP_SMOOTH_MAX_LISTS (image, disparity pyramid) lists at once — the
 * three compute_smooth_loss calls of model_geometry.py:938-940 / model_depth.py:281-283 in one launch:
 *   forward_grad: out (lists,B) and, where G[list][l] is non-NULL, G (B,1,H,W) = d out / d up_l(Y,X) at full resolution
 *                 (un-normalised: without the upstream gradient), the image edge weights formed once per tile for all levels;
 *   combine:      grad_disp[list][l] (B,1,h_l,w_l) = grad_out[list,b] * (bilinear up-sampling)^T G[list][l]. */
#define UGL_DISP_SMOOTH_MAX_LISTS 3
typedef struct UglDispSmoothArgs {
  int32_t batch, lists, levels, height, width;
  int32_t lheight[UGL_MAX_LEVELS], lwidth[UGL_MAX_LEVELS];
  const float* img[UGL_DISP_SMOOTH_MAX_LISTS];                   /* (B,3,H,W)                 */
  const float* disp[UGL_DISP_SMOOTH_MAX_LISTS][UGL_MAX_LEVELS];  /* (B,1,h_l,w_l)             */
  float* out;                                                    /* (lists,B)                 */
  float* G[UGL_DISP_SMOOTH_MAX_LISTS][UGL_MAX_LEVELS];           /* (B,1,H,W) or NULL         */
  const float* grad_out;                                         /* (lists,B)      [combine]  */
  float* grad_disp[UGL_DISP_SMOOTH_MAX_LISTS][UGL_MAX_LEVELS];   /* (B,1,h_l,w_l)  [combine]  */
  void* workspace;
  uint64_t workspace_bytes;
  void* stream;
  int32_t grad_out_shared;                                       /* [combine] 1: grad_out is ONE (B,) row used by every list (the three
                                                                    compute_smooth_loss calls of model_geometry.py:938-940 share one weight) */
} UglDispSmoothArgs;
uint64_t ugl_disp_smooth_fused_workspace_bytes(const UglDispSmoothArgs* args);
int ugl_disp_smooth_forward_grad(const UglDispSmoothArgs* args);
int ugl_disp_smooth_combine(const UglDispSmoothArgs* args);

/* ---------------------------------------------------------------------------------------------
 * inverse_warp2 (structures/inverse_warp.py:263-303 with pixel2cam :30-45 and cam2pixel2 :227-260).
 * Kinv (B,3,3) = intrinsics.inverse() and P (B,3,4) = intrinsics @ pose_vec2mat(pose) are built by the
 * caller; backward returns grad_depth (B,1,H,W), grad_P (B,3,4) and, if requested, the deterministic
 * scatter gradients grad_img (B,C,H,W) / grad_ref_depth (B,1,H,W).  go_* may be NULL (= zero).
 * calculate_rigid_flow (:311-342) shares the projection.
 * ------------------------------------------------------------------------------------------- */
int ugl_reproject_forward(const float* img, const float* depth, const float* ref_depth, const float* Kinv, const float* P,
                          int32_t batch, int32_t channels, int32_t height, int32_t width, float* out, float* valid,
                          float* proj_depth, float* comp_depth, void* stream);
uint64_t ugl_reproject_backward_workspace_bytes(int32_t batch, int32_t channels, int32_t height, int32_t width,
                                                int32_t need_grad_img, int32_t need_grad_ref);
int ugl_reproject_backward(const float* img, const float* depth, const float* ref_depth, const float* Kinv, const float* P,
                           const float* go_img, const float* go_proj, const float* go_comp, int32_t batch, int32_t channels,
                           int32_t height, int32_t width, float* grad_depth, float* grad_P, float* grad_img, float* grad_ref_depth,
                           void* workspace, uint64_t workspace_bytes, void* stream);
int ugl_rigid_flow_forward(const float* depth, const float* Kinv, const float* P, int32_t batch, int32_t height, int32_t width,
                           float* out, void* stream);
int ugl_rigid_flow_backward(const float* depth, const float* Kinv, const float* P, const float* grad_out, int32_t batch, int32_t height,
                            int32_t width, float* grad_depth, float* grad_P, void* workspace, uint64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused reprojection-photometric term of the depth / geom modes: for every level and both directions
 * reconstruction (model_geometry.py:80-103 -> inverse_warp2, structures/inverse_warp.py:263-303) +
 * compute_texture_mask (:134-140) + mask fusion (model_depth.py:262-269 valid * texture, or the flow-branch
 * mask * texture of model_geometry.py:854-855 when ext_mask is given) + compute_photometric_loss (:143-153),
 * summed over directions and levels -> loss (B,) = loss_depth_pixel.  dir 0 = left source frame / pose[:,0],
 * dir 1 = right source frame / pose[:,1].  backward: grad_loss (B,) -> grad_disp[l] (B,1,h,w), grad_P[dir][l] (B,3,4).
 * ------------------------------------------------------------------------------------------- */
typedef struct UglDepthPhotoArgs {
  int32_t batch;
  int32_t scales;
  int32_t height[UGL_MAX_LEVELS];
  int32_t width[UGL_MAX_LEVELS];
  const float* img[UGL_MAX_LEVELS];            /* (B,3,h,w) centre frame, bilinear pyramid                 */
  const float* src_area[2][UGL_MAX_LEVELS];    /* (B,3,h,w) source frames, area pyramid (sampled)          */
  const float* src_bil[2][UGL_MAX_LEVELS];     /* (B,3,h,w) source frames, bilinear pyramid (texture mask) */
  const float* disp[UGL_MAX_LEVELS];           /* (B,1,h,w) centre disparity                                */
  const float* Kinv[UGL_MAX_LEVELS];           /* (B,3,3) inverse of the level's intrinsics                 */
  const float* P[2][UGL_MAX_LEVELS];           /* (B,3,4) K_s [R|t]                                         */
  const float* ext_mask[2][UGL_MAX_LEVELS];    /* (B,1,h,w) or NULL (= use the reprojection valid mask)     */
  float* valid_out[2][UGL_MAX_LEVELS];         /* optional (B,1,h,w) reprojection valid mask                */
  float* tex_out[2][UGL_MAX_LEVELS];           /* optional (B,1,h,w) texture mask                           */
  float* loss;                                 /* (B,)                                                      */
  float* den;                                  /* (B,scales,2) written by forward, read by backward         */
  const float* grad_loss;                      /* (B,)            [backward]                                */
  float* grad_disp[UGL_MAX_LEVELS];            /* (B,1,h,w)       [backward]                                */
  float* grad_P[2][UGL_MAX_LEVELS];            /* (B,3,4)         [backward]                                */
  void* workspace;
  uint64_t workspace_bytes;
  void* stream;
  /* geom mode, instead of ext_mask: packed masks written by ugl_geom_flow_forward_grad; the base mask of direction d is
   * [ (ext_bytes & ext_need[d]) == ext_need[d] ]  (model_geometry.py:857-858: valid * occ * dyn = bits 1|4|16 / 2|8|32) */
  const uint8_t* ext_bytes[UGL_MAX_LEVELS];    /* (B,h,w) or NULL */
  int32_t ext_need[2];
} UglDepthPhotoArgs;

uint64_t ugl_depth_photo_workspace_bytes(const UglDepthPhotoArgs* args);
int ugl_depth_photo_forward(const UglDepthPhotoArgs* args);
int ugl_depth_photo_backward(const UglDepthPhotoArgs* args);

/* Single-pass variant of the above (what the autograd wrapper uses when a gradient is needed): forward_grad also writes the
 * UN-normalised d loss / d disparity through each direction, basis[l] (B,2,h,w), and the un-normalised d loss / d P sums,
 * psum (B,scales,2,12), while the taps are in registers; combine (element-wise) applies grad_loss (B,) and the normalisers:
 * grad_disp[l] (B,1,h,w), grad_P[dir][l] (B,3,4).  One gather kernel per step instead of two.  Workspace: >= ..._grad_workspace_bytes. */
typedef struct UglDepthPhotoGradArgs {
  UglDepthPhotoArgs photo;
  float* basis[UGL_MAX_LEVELS];
  float* psum;
} UglDepthPhotoGradArgs;

uint64_t ugl_depth_photo_grad_workspace_bytes(const UglDepthPhotoGradArgs* args);
int ugl_depth_photo_forward_grad(const UglDepthPhotoGradArgs* args);
int ugl_depth_photo_combine(const UglDepthPhotoGradArgs* args);

/* Depth-consistency term of the depth mode for both source frames and all levels: compute_consis_loss (unmasked,
 * model_depth.py:154-163; enabled at model_depth_texture.py:308-309) on the projected / computed depths of inverse_warp2
 * (structures/inverse_warp.py:263-303): loss (B,) = sum_{frames, levels} mean(clamp(|comp - proj| / |comp + proj|, 0, 1)).
 * ref_disp[d][l] (B,1,h,w) = the source frame's disparity pyramid (d = 0 left / pose[:,0], 1 = right).  backward:
 * grad_loss (B,) -> grad_disp[l], grad_ref[d][l] (deterministic fixed-point scatter), grad_P[d][l]. */
typedef struct UglDepthConsisArgs {
  int32_t batch, scales;
  int32_t height[UGL_MAX_LEVELS], width[UGL_MAX_LEVELS];
  const float* disp[UGL_MAX_LEVELS];
  const float* ref_disp[2][UGL_MAX_LEVELS];
  const float* Kinv[UGL_MAX_LEVELS];
  const float* P[2][UGL_MAX_LEVELS];
  float* loss;
  const float* grad_loss;
  float* grad_disp[UGL_MAX_LEVELS];
  float* grad_ref[2][UGL_MAX_LEVELS];
  float* grad_P[2][UGL_MAX_LEVELS];
  void* workspace;                             /* >= ugl_depth_consis_workspace_bytes, 8-byte aligned */
  uint64_t workspace_bytes;
  void* stream;
} UglDepthConsisArgs;
uint64_t ugl_depth_consis_workspace_bytes(const UglDepthConsisArgs* args);
int ugl_depth_consis_forward(const UglDepthConsisArgs* args);
int ugl_depth_consis_backward(const UglDepthConsisArgs* args);

/* Depth mode WITH the SSIM term (model_depth_texture.py:296-301; BASELINE configs[2]): loss_depth_pixel (L1 under valid * texture,
 * compute_photometric_depth_loss) AND loss_depth_ssim (compute_ssim_loss under the reprojection valid mask, model_depth.py /
 * model_geometry.py:212-223 + pytorch_ssim/ssim.py:4-19) of both source frames and all levels in the single-pass tile kernel
 * (reprojection warps in place of flow warps).  `photo` carries the inputs exactly as for ugl_depth_photo_* (ext_* unused; its
 * `loss` / `den` / `grad_loss` fields are ignored).  forward_grad: loss4 (4,B) rows = depth_pixel, depth_ssim, 0, 0; stats
 * (B,scales,UGL_GEOM_NSTATS); basis[l] (B,UGL_DEPTH_BASIS_PLANES,h,w) = un-normalised d loss / d (projected u, v) per term and
 * direction.  combine: grad_loss4 (2,B) = upstream gradients of depth_pixel / depth_ssim -> photo.grad_disp[l], photo.grad_P[d][l]. */
#define UGL_DEPTH_BASIS_PLANES 8
typedef struct UglDepthSsimArgs {
  UglDepthPhotoArgs photo;
  float* loss4;                                /* (4,B)                                         */
  float* stats;                                /* (B,scales,UGL_GEOM_NSTATS)                    */
  float* basis[UGL_MAX_LEVELS];                /* (B,8,h,w)                                     */
  const float* grad_loss4;                     /* (2,B)                 [combine]               */
} UglDepthSsimArgs;
uint64_t ugl_depth_ssim_workspace_bytes(const UglDepthSsimArgs* args);
int ugl_depth_ssim_forward_grad(const UglDepthSsimArgs* args);
int ugl_depth_ssim_combine(const UglDepthSsimArgs* args);

/* Per-step matrix set-up of the depth / geom modes, one launch for every level and both poses:
 *   K_s = K with rows 0-1 divided by downscales[s] (model_geometry.py:92-93); Kinv_out[s] (B,3,3) = K_s^-1 (inverse_warp.py:284);
 *   [R|t] = pose_vec2mat(pose[:,k]) (euler, inverse_warp.py:110-145, 172-187); P_out[k*S+s] (B,3,4) = K_s [R|t] (:289);
 *   F_out[k] (B,3,3) = K_inv^T [t]x R K_inv (compute_essential_matrix :354-364 + model_geometry.py:355-370), optional.
 * pose (B,n,6) = [tx,ty,tz,rx,ry,rz], n <= 2.  backward: grad_P[k*S+s] / grad_F[k] (entries may be NULL = zero)
 * -> grad_pose (B,n,6).  downscales is a HOST array of S floats. */
int ugl_pose_setup_forward(const float* pose, const float* K, const float* K_inv, const float* downscales, int32_t batch, int32_t poses,
                           int32_t levels, float* const* Kinv_out, float* const* P_out, float* const* F_out, void* stream);
int ugl_pose_setup_backward(const float* pose, const float* K, const float* K_inv, const float* downscales, int32_t batch, int32_t poses,
                            int32_t levels, const float* const* grad_P, const float* const* grad_F, float* grad_pose, void* stream);

/* Level-0 rigid-consistency terms of Model_geometry.forward in one forward / one backward kernel (+ finalize each):
 *   loss_dfc (B,) = loss_depth_flow_consis: calculate_rigid_flow (inverse_warp.py:311-342) for both poses, |rigid - flow| (2 ch,
 *                   compute_dynamic_mask's flow_diff, model_geometry.py:698-700) under valid*occ*dyn (mask_bytes & need[d] == need[d]),
 *                   compute_depth_flow_consis_loss with scales=1 (:716-732, called at :925-926);
 *   loss_epi (B,) = loss_epipolar: compute_epipolar_map (:355-403) for both poses, plain per-sample mean (:413-418).
 * den (B,2) is written by forward and read by backward.  backward: grad_dfc / grad_epi (B,) (either may be NULL = zero) ->
 * grad_flow_bwd/fwd (B,2,H,W), grad_disp (B,1,H,W), grad_P_bwd/fwd (B,3,4), grad_F_bwd/fwd (B,3,3). */
typedef struct UglGeomRigidArgs {
  int32_t batch, height, width;
  int32_t need[2];                             /* mask bits per direction (bwd, fwd)             */
  const float *flow_bwd, *flow_fwd;            /* (B,2,H,W) level-0 flows                        */
  const float* disp;                           /* (B,1,H,W) centre disparity, level 0            */
  const uint8_t* mask_bytes;                   /* (B,H,W) from ugl_geom_flow_forward_grad        */
  const float *Kinv, *P_bwd, *P_fwd;           /* (B,3,3), (B,3,4) x2 of level 0                 */
  const float *F_bwd, *F_fwd;                  /* (B,3,3) fundamental matrices                   */
  float *loss_dfc, *loss_epi, *den;            /* (B,), (B,), (B,2)                              */
  const float *grad_dfc, *grad_epi;            /* (B,)                      [backward]           */
  float *grad_flow_bwd, *grad_flow_fwd, *grad_disp, *grad_P_bwd, *grad_P_fwd, *grad_F_bwd, *grad_F_fwd;
  void* workspace;                             /* >= ugl_geom_rigid_workspace_bytes              */
  uint64_t workspace_bytes;
  void* stream;
} UglGeomRigidArgs;
uint64_t ugl_geom_rigid_workspace_bytes(int32_t batch, int32_t height, int32_t width);
int ugl_geom_rigid_forward(const UglGeomRigidArgs* args);
int ugl_geom_rigid_backward(const UglGeomRigidArgs* args);

/* compute_epipolar_map (model_geometry.py:355-403) given F (B,3,3) = K^-T [t]x R K^-1: dist (B,1,H,W);
 * backward: grad_flow (B,2,H,W, may be NULL) and grad_F (B,3,3). */
int ugl_epipolar_forward(const float* flow, const float* F, int32_t batch, int32_t height, int32_t width, float* out, void* stream);
int ugl_epipolar_backward(const float* flow, const float* F, const float* grad_out, int32_t batch, int32_t height, int32_t width,
                          float* grad_flow, float* grad_F, void* workspace, uint64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SSIM (pytorch_ssim/ssim.py:4-19): the map and its backward; and the masked SSIM loss of
 * compute_ssim_loss / compute_loss_ssim (model_geometry.py:212-223, model_flow.py:141-152, one level).
 * ------------------------------------------------------------------------------------------- */
int ugl_ssim_forward(const float* x, const float* y, int32_t batch, int32_t channels, int32_t height, int32_t width, float* out,
                     void* stream);
int ugl_ssim_backward(const float* x, const float* y, const float* grad_out, int32_t batch, int32_t channels, int32_t height,
                      int32_t width, float* grad_x, float* grad_y, void* stream);
uint64_t ugl_ssim_loss_workspace_bytes(int32_t batch, int32_t channels, int32_t height, int32_t width);
int ugl_ssim_loss_forward(const float* img, const float* warped, const float* mask, int32_t batch, int32_t channels, int32_t height,
                          int32_t width, float* out, float* den, void* workspace, uint64_t workspace_bytes, void* stream);
int ugl_ssim_loss_backward(const float* img, const float* warped, const float* mask, const float* den, const float* grad_out,
                           int32_t batch, int32_t channels, int32_t height, int32_t width, float* grad_img, float* grad_warped,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UGL_H_ */
