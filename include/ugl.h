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
} UglFlowLossArgs;

uint64_t ugl_flow_loss_workspace_bytes(const UglFlowLossArgs* args);
int ugl_flow_loss_forward(const UglFlowLossArgs* args);
int ugl_flow_loss_backward(const UglFlowLossArgs* args);
/* number of kernel launches one forward / backward call issues (for launch accounting) */
int ugl_flow_loss_launches(int backward);

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
uint64_t ugl_warp_flow_backward_workspace_bytes(int32_t batch, int32_t channels, int32_t height, int32_t width,
                                                int32_t need_grad_x);

#ifdef __cplusplus
}
#endif
#endif /* UGL_H_ */
