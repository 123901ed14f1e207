// Per-pixel logic of the primitive ops (pyramid, warp_flow) as host/device functions, shared by
// ugl_primitives.cu and the host emulator.
#pragma once

#include "ugl_common.cuh"

namespace ugl {

// ---- image pyramid -------------------------------------------------------------------------------
// mode 0: box mean of the 2^l x 2^l block, accumulated row-major then divided by the count
//         (adaptive_avg_pool2d / interpolate 'area'; model_flow.py:62, model_geometry.py:91).
// mode 1: bilinear align_corners=False at an exact power-of-two ratio = the central 2x2 of the block
//         with weights 1/4 (model_geometry.py:70).
UGL_HD float pyramid_pixel(const float* __restrict__ plane, int W, int l, int mode, int oi, int oj) {
  const int f = 1 << l;
  if (mode == 0) {
    float s = 0.f;
    for (int dy = 0; dy < f; ++dy)
      for (int dx = 0; dx < f; ++dx) s = add_rn(s, plane[(long)(oi * f + dy) * W + (oj * f + dx)]);
    return div_rn(s, (float)(f * f));
  }
  const int y = oi * f + f / 2 - 1, x = oj * f + f / 2 - 1;
  const float* p = plane + (long)y * W + x;
  // Horizontal lerp, then vertical: the order of ATen's CUDA kernel and of its vectorised multi-threaded CPU
  // kernel at production sizes.  (ATen's CPU result is itself not bit-stable: small or single-threaded
  // calls take a path that accumulates the four taps sequentially and differs in the last ulp.)
  const float top = add_rn(mul_rn(0.5f, p[0]), mul_rn(0.5f, p[1]));
  const float bot = add_rn(mul_rn(0.5f, p[W]), mul_rn(0.5f, p[W + 1]));
  return add_rn(mul_rn(0.5f, top), mul_rn(0.5f, bot));
}

// ---- warp_flow (structures/net_utils.py:16-54) ------------------------------------------------------
// forward for one pixel, all channels; returns the keep value (1 when use_mask == 0)
UGL_HD float warp_pixel_forward(const float* __restrict__ x, const float* __restrict__ flow, int C, const WarpGeom& geom,
                                int b, int i, int j, int use_mask, float* __restrict__ out) {
  const int H = geom.H, W = geom.W;
  const long plane = (long)H * W, pix = (long)i * W + j;
  const float u = flow[((long)b * 2) * plane + pix], v = flow[((long)b * 2 + 1) * plane + pix];
  const Tap t = flow_tap(j, i, u, v, geom);
  const float keep = use_mask ? tap_keep(t) : 1.0f;
  for (int c = 0; c < C; ++c) {
    const Corners k = tap_fetch(x + ((long)b * C + c) * plane, W, t);
    const float val = corners_value(k, t);
    out[((long)b * C + c) * plane + pix] = use_mask ? val * keep : val;
  }
  return keep;
}

// backward w.r.t. the flow for one pixel
UGL_HD void warp_pixel_backward_flow(const float* __restrict__ x, const float* __restrict__ flow,
                                     const float* __restrict__ gout, int C, const WarpGeom& geom, int b, int i, int j,
                                     int use_mask, float* __restrict__ gflow) {
  const int H = geom.H, W = geom.W;
  const long plane = (long)H * W, pix = (long)i * W + j;
  const float u = flow[((long)b * 2) * plane + pix], v = flow[((long)b * 2 + 1) * plane + pix];
  const Tap t = flow_tap(j, i, u, v, geom);
  const float keep = use_mask ? tap_keep(t) : 1.0f;
  float gx = 0.f, gy = 0.f;
  for (int c = 0; c < C; ++c) {
    const Corners k = tap_fetch(x + ((long)b * C + c) * plane, W, t);
    const float g = gout[((long)b * C + c) * plane + pix] * keep;
    gx += g * corners_ddx(k, t);
    gy += g * corners_ddy(k, t);
  }
  gflow[((long)b * 2) * plane + pix] = gx * geom.sx;
  gflow[((long)b * 2 + 1) * plane + pix] = gy * geom.sy;
}

}  // namespace ugl
