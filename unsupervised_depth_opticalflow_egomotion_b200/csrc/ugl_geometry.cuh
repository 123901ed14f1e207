// Per-pixel projection math shared by the geometry kernels: pixel2cam / cam2pixel2 of
// structures/inverse_warp.py:30-45, 227-260 and their backward.
#pragma once

#include "ugl_common.cuh"

namespace ugl {

constexpr float kDepthMin = 1e-3f;   // inverse_warp.py:247,301

struct Projected {
  float dir[3];   // K^-1 [j,i,1]
  float cam[3];   // dir * depth
  float X, Y, q2; // K [R|t] cam (q2 before the clamp)
  float Z;        // max(q2, 1e-3)
  float u, v;     // X/Z, Y/Z
};

// pixel2cam (inverse_warp.py:30-45) + cam2pixel (:47-78 / :227-260): matmul rows as fma chains
UGL_HD Projected project_pixel(const float* __restrict__ Kinv, const float* __restrict__ P, float D, int j, int i) {
  Projected r;
  const float fj = (float)j, fi = (float)i;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    r.dir[k] = fma_rn(Kinv[k * 3 + 2], 1.0f, fma_rn(Kinv[k * 3 + 1], fi, mul_rn(Kinv[k * 3 + 0], fj)));
    r.cam[k] = mul_rn(r.dir[k], D);
  }
  float q[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    q[k] = add_rn(fma_rn(P[k * 4 + 2], r.cam[2], fma_rn(P[k * 4 + 1], r.cam[1], mul_rn(P[k * 4 + 0], r.cam[0]))), P[k * 4 + 3]);
  r.X = q[0]; r.Y = q[1]; r.q2 = q[2];
  r.Z = q[2] < kDepthMin ? kDepthMin : q[2];   // clamp(min=1e-3); NaN propagates like torch
  if (!(q[2] >= kDepthMin) && !(q[2] < kDepthMin)) r.Z = q[2];
  r.u = div_rn(r.X, r.Z);
  r.v = div_rn(r.Y, r.Z);
  return r;
}

// normalised sampling coordinates of cam2pixel2 (:250-257): out-of-range values are replaced by 2
struct NormCoord { float gx, gy; bool ox, oy; };
UGL_HD NormCoord normalise(const Projected& r, const WarpGeom& g) {
  NormCoord n;
  n.gx = sub_rn(div_c(mul_rn(2.0f, r.u), g.dw, g.rdw), 1.0f);
  n.gy = sub_rn(div_c(mul_rn(2.0f, r.v), g.dh, g.rdh), 1.0f);
  n.ox = (n.gx > 1.0f) || (n.gx < -1.0f);
  n.oy = (n.gy > 1.0f) || (n.gy < -1.0f);
  if (n.ox) n.gx = 2.0f;
  if (n.oy) n.gy = 2.0f;
  return n;
}


// chain d loss / d(u, v, Z) back to depth and the 12 entries of P (accumulated in acc[0..11])
UGL_HD float project_backward(const Projected& r, const float* __restrict__ P, float g_u, float g_v, float g_Z,
                                                  float* acc) {
  const float iz = 1.0f / r.Z;
  const float gX = g_u * iz, gY = g_v * iz;
  float gq2 = -(g_u * r.X + g_v * r.Y) * iz * iz + g_Z;
  if (!(r.q2 >= kDepthMin)) gq2 = 0.f;             // clamp(min) passes the gradient where q2 >= 1e-3
  const float gq[3] = {gX, gY, gq2};
  float gD = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    gD += gq[k] * (P[k * 4 + 0] * r.dir[0] + P[k * 4 + 1] * r.dir[1] + P[k * 4 + 2] * r.dir[2]);
    acc[k * 4 + 0] += gq[k] * r.cam[0];
    acc[k * 4 + 1] += gq[k] * r.cam[1];
    acc[k * 4 + 2] += gq[k] * r.cam[2];
    acc[k * 4 + 3] += gq[k];
  }
  return gD;
}


// ---- epipolar distance of (j+u, i+v) to the line F [j,i,1]^T (model_geometry.py:355-403) ----
struct Epi { float l[3], p2[3], n, r, d, dist; };
UGL_HD Epi epipolar_pixel(const float* __restrict__ F, float u, float v, int j, int i) {
  Epi e;
  const float fj = (float)j, fi = (float)i;
#pragma unroll
  for (int k = 0; k < 3; ++k) e.l[k] = fma_rn(F[k * 3 + 2], 1.0f, fma_rn(F[k * 3 + 1], fi, mul_rn(F[k * 3 + 0], fj)));
  e.p2[0] = add_rn(fj, u); e.p2[1] = add_rn(fi, v); e.p2[2] = 1.0f;
  e.n = add_rn(add_rn(mul_rn(e.p2[0], e.l[0]), mul_rn(e.p2[1], e.l[1])), mul_rn(e.p2[2], e.l[2]));
  e.r = sqrt_rn(add_rn(mul_rn(e.l[0], e.l[0]), mul_rn(e.l[1], e.l[1])));
  e.d = add_rn(e.r, 1e-6f);
  e.dist = div_rn(fabsf(e.n), e.d);
  return e;
}


}  // namespace ugl
