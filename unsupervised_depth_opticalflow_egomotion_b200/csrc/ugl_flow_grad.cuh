// Single-pass variant of the fused flow-mode loss: ONE stencil kernel computes the loss sums AND the
// un-normalised gradient of every term w.r.t. both flows ("gradient basis maps"); the backward pass is
// then an element-wise combine  grad = sum_k scale_k(sample, level) * basis_k  once the per-sample
// normalisers (mean of the occlusion weights) and the upstream gradients are known.
//
// Why: the kernels are bound by fp32 instruction issue and latency, not by HBM (profiles/), and the recompute scheme of
// ugl_flow_loss.cuh executes the photometry and the SSIM moments twice (forward, then backward).  This variant executes them
// once, trading 14 floats/pixel of extra HBM traffic (far from binding) for ~1/3 fewer instructions.
//
// The two warp directions of a pixel (.x = forward flow / right frame / frame 0, .y = backward flow / left frame / frame 1) travel
// through every phase as ONE packed fp32 pair: sm_100a's FADD2 / FMUL2 / FFMA2 give two individually IEEE-rounded results per
// instruction, so the fp32 instruction count halves while every rounding of the scalar formulation is kept (ugl_common.cuh:
// add2 / mul2 / fma2 and the contraction-proof acc2_rn / sub2_rn).  Phase order per tile: phase 1 (photometry, halo 2) ->
// per channel [phase 2 (SSIM windows -> coefficient pairs) -> phase 3 accumulate (box sums x Jacobian)] -> phase 3 store +
// phase 4a (signed second differences) -> phase 4b (smoothness gather) -> block reduction.
//
// Basis planes per level, layout (B, 14, h, w):
//   0,1  Gp_f  = w_f * sum_c sign(Wf_c - I_c) * keep * dWf_c/d(u,v)            x g_pix  / (3 hw den_f)
//   2,3  Gs_f  = w_f * sum_c (sA + 2 y sB + x sC) * keep * dWf_c/d(u,v)        x g_ssim / (27 hw den_f)
//   4,5  Gm_f  = sum_t coef_t (wx_t sign(dxx_t) / nx + wy_t sign(dyy_t) / ny)  x g_smooth / 40
//   6,7  Gc    = (1 - w_f) * d(|f^_f + f^_b|_1)/d(u_f, v_f)                    x g_consis / (2 hw den_c)
//   8,9  Gp_b, 10,11 Gs_b, 12,13 Gm_b  (same for the backward flow; the consistency term has no bwd gradient)
#pragma once

#include "ugl_flow_loss.cuh"
#include "ugl_geometry.cuh"

namespace ugl {

constexpr int kBasisPlanes = 14;

// ---- unconditional clamped 2x2 gather ----------------------------------------------------------------------------
// Bilinear weights and the in-bounds tests are separable (w_nw = ox*oy, mask = [x0 in range]*[y0 in range]).  The kernel
// always loads the in-image 2x2 block at columns (xa, xa+1), rows (ya, ya+1) with xa = clamp(x0, 0, W-2), ya likewise, so
// the three neighbours are fixed offsets (+1, +W, +W+1) of ONE address per channel, and folds "which of my two columns is
// x0 / x0+1, if any" into per-column weights cw[2] (rows: rw[2]).  The products cw*rw are exactly the ATen weights
// (ox*oy etc.) or exact zeros, accumulated in ATen's nw, ne, sw, se order, so the sampled value is bit-identical to
// tap_fetch/corners_value; out-of-range corners contribute an exact +0.
struct TapC {
  int off;                    // ya * W + xa
  float w[4];                 // nw, ne, sw, se weights of the loaded block
  float dx[4], dy[4];         // d w / d ix, d w / d iy
  float keep;
};

template <bool kGrad>
UGL_HD TapC tap_clamped_norm(float gx, float gy, const WarpGeom& g) {
  float ix = unnormalize(gx, g.W), iy = unnormalize(gy, g.H);
  ix = fminf(fmaxf(ix, -2.0f), (float)g.W + 1.0f);      // beyond that every corner is out of range anyway
  iy = fminf(fmaxf(iy, -2.0f), (float)g.H + 1.0f);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float tx = ix - fx, ty = iy - fy;
  const float ox = (fx + 1.0f) - ix, oy = (fy + 1.0f) - iy;
  const int xa = imin(imax(x0, 0), g.W - 2), ya = imin(imax(y0, 0), g.H - 2);
  // column xa holds x0 (weight ox) or x0+1 (weight tx) or neither; column xa+1 likewise
  const float l0 = (x0 == xa) ? 1.f : 0.f, r0 = (x0 + 1 == xa) ? 1.f : 0.f;
  const float l1 = (x0 == xa + 1) ? 1.f : 0.f, r1 = (x0 == xa) ? 1.f : 0.f;
  const float t0 = (y0 == ya) ? 1.f : 0.f, b0 = (y0 + 1 == ya) ? 1.f : 0.f;
  const float t1 = (y0 == ya + 1) ? 1.f : 0.f, b1 = (y0 == ya) ? 1.f : 0.f;
  // per-corner weights as ATen forms them: (x-part) * (y-part); exactly one of the two terms of a part is non-zero
  const float cx0 = ox * l0 + tx * r0, cx1 = ox * l1 + tx * r1;
  const float ry0 = oy * t0 + ty * b0, ry1 = oy * t1 + ty * b1;
  TapC t;
  t.off = ya * g.W + xa;
  t.w[0] = cx0 * ry0; t.w[1] = cx1 * ry0; t.w[2] = cx0 * ry1; t.w[3] = cx1 * ry1;
  // coverage = sum of the in-bounds weights in ATen's corner order (nw, ne, sw, se of the ORIGINAL footprint): the
  // original corners map to block positions in the same relative order, zeros are exact
  const float wnw = (ox * oy) * ((x0 >= 0 && x0 < g.W && y0 >= 0 && y0 < g.H) ? 1.f : 0.f);
  const float wne = (tx * oy) * ((x0 + 1 >= 0 && x0 + 1 < g.W && y0 >= 0 && y0 < g.H) ? 1.f : 0.f);
  const float wsw = (ox * ty) * ((x0 >= 0 && x0 < g.W && y0 + 1 >= 0 && y0 + 1 < g.H) ? 1.f : 0.f);
  const float wse = (tx * ty) * ((x0 + 1 >= 0 && x0 + 1 < g.W && y0 + 1 >= 0 && y0 + 1 < g.H) ? 1.f : 0.f);
  t.keep = add_rn(add_rn(add_rn(wnw, wne), wsw), wse) >= 0.9999f ? 1.0f : 0.0f;
  if (kGrad) {
    const float dcx0 = r0 - l0, dcx1 = r1 - l1;          // d cx / d ix  (d ox = -1, d tx = +1)
    const float dry0 = b0 - t0, dry1 = b1 - t1;
    t.dx[0] = dcx0 * ry0; t.dx[1] = dcx1 * ry0; t.dx[2] = dcx0 * ry1; t.dx[3] = dcx1 * ry1;
    t.dy[0] = cx0 * dry0; t.dy[1] = cx1 * dry0; t.dy[2] = cx0 * dry1; t.dy[3] = cx1 * dry1;
  }
  return t;
}


// ---- the same gather set-up for BOTH warp directions of a pixel as packed fp32 pairs (.x = forward flow / right frame,
// .y = backward flow / left frame): identical roundings (every packed op is the .rn form of the scalar one; where ptxas may
// contract a product into a sum the product is by an exact 0 / 1 factor), half the fp32 instructions.  The clamps, floor,
// float->int conversions and integer compares stay per lane.
struct TapC2 {
  int off[2];                    // ya * W + xa per direction
  float2 w[4];                   // nw, ne, sw, se weights of the loaded block
  float2 dx[4], dy[4];           // d w / d ix, d w / d iy
  float2 keep;
};

UGL_HD float flag(bool b) { return b ? 1.0f : 0.0f; }

template <bool kGrad>
UGL_HD TapC2 tap_clamped_norm2(float2 gx, float2 gy, const WarpGeom& g) {
  const float2 one = splat2(1.0f);
  float2 ix = fma2(add2(gx, one), splat2(0.5f * (float)g.W), splat2(-0.5f));     // unnormalize(), both directions
  float2 iy = fma2(add2(gy, one), splat2(0.5f * (float)g.H), splat2(-0.5f));
  const float xhi = (float)g.W + 1.0f, yhi = (float)g.H + 1.0f;                  // beyond that every corner is out of range anyway
  ix = make_float2(fminf(fmaxf(ix.x, -2.0f), xhi), fminf(fmaxf(ix.y, -2.0f), xhi));
  iy = make_float2(fminf(fmaxf(iy.x, -2.0f), yhi), fminf(fmaxf(iy.y, -2.0f), yhi));
  const float2 fx = make_float2(floorf(ix.x), floorf(ix.y)), fy = make_float2(floorf(iy.x), floorf(iy.y));
  const int x0[2] = {(int)fx.x, (int)fx.y}, y0[2] = {(int)fy.x, (int)fy.y};
  const float2 tx = sub2(ix, fx), ty = sub2(iy, fy);
  const float2 ox = sub2(add2(fx, one), ix), oy = sub2(add2(fy, one), iy);
  TapC2 t;
  float l0[2], r0[2], l1[2], r1[2], t0[2], b0[2], t1[2], b1[2], mnw[2], mne[2], msw[2], mse[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const int xa = imin(imax(x0[d], 0), g.W - 2), ya = imin(imax(y0[d], 0), g.H - 2);
    t.off[d] = ya * g.W + xa;
    // column xa holds x0 (weight ox) or x0+1 (weight tx) or neither; column xa+1 likewise
    l0[d] = flag(x0[d] == xa); r0[d] = flag(x0[d] + 1 == xa); l1[d] = flag(x0[d] == xa + 1); r1[d] = l0[d];
    t0[d] = flag(y0[d] == ya); b0[d] = flag(y0[d] + 1 == ya); t1[d] = flag(y0[d] == ya + 1); b1[d] = t0[d];
    const bool xl = x0[d] >= 0 && x0[d] < g.W, xr = x0[d] + 1 >= 0 && x0[d] + 1 < g.W;
    const bool yt = y0[d] >= 0 && y0[d] < g.H, yb = y0[d] + 1 >= 0 && y0[d] + 1 < g.H;
    mnw[d] = flag(xl && yt); mne[d] = flag(xr && yt); msw[d] = flag(xl && yb); mse[d] = flag(xr && yb);
  }
  const float2 L0 = make_float2(l0[0], l0[1]), R0 = make_float2(r0[0], r0[1]), L1 = make_float2(l1[0], l1[1]), R1 = make_float2(r1[0], r1[1]);
  const float2 T0 = make_float2(t0[0], t0[1]), B0 = make_float2(b0[0], b0[1]), T1 = make_float2(t1[0], t1[1]), B1 = make_float2(b1[0], b1[1]);
  // per-corner weights as ATen forms them: (x-part) * (y-part); exactly one of the two terms of a part is non-zero
  const float2 cx0 = fma2(ox, L0, mul2(tx, R0)), cx1 = fma2(ox, L1, mul2(tx, R1));
  const float2 ry0 = fma2(oy, T0, mul2(ty, B0)), ry1 = fma2(oy, T1, mul2(ty, B1));
  t.w[0] = mul2(cx0, ry0); t.w[1] = mul2(cx1, ry0); t.w[2] = mul2(cx0, ry1); t.w[3] = mul2(cx1, ry1);
  // coverage = sum of the in-bounds weights in ATen's corner order (nw, ne, sw, se of the ORIGINAL footprint)
  const float2 wnw = mul2(mul2(ox, oy), make_float2(mnw[0], mnw[1])), wne = mul2(mul2(tx, oy), make_float2(mne[0], mne[1]));
  const float2 wsw = mul2(mul2(ox, ty), make_float2(msw[0], msw[1])), wse = mul2(mul2(tx, ty), make_float2(mse[0], mse[1]));
  const float2 cov = add2(add2(add2(wnw, wne), wsw), wse);
  t.keep = make_float2(flag(cov.x >= 0.9999f), flag(cov.y >= 0.9999f));
  if (kGrad) {
    const float2 dcx0 = sub2(R0, L0), dcx1 = sub2(R1, L1);          // d cx / d ix  (d ox = -1, d tx = +1)
    const float2 dry0 = sub2(B0, T0), dry1 = sub2(B1, T1);
    t.dx[0] = mul2(dcx0, ry0); t.dx[1] = mul2(dcx1, ry0); t.dx[2] = mul2(dcx0, ry1); t.dx[3] = mul2(dcx1, ry1);
    t.dy[0] = mul2(cx0, dry0); t.dy[1] = mul2(cx1, dry0); t.dy[2] = mul2(cx0, dry1); t.dy[3] = mul2(cx1, dry1);
  }
  return t;
}

// both flow warps of pixel (j, i): u = (u_fwd, u_bwd), v likewise
template <bool kGrad>
UGL_HD TapC2 flow_tap_clamped2(int j, int i, float2 u, float2 v, const WarpGeom& g) {
  const float2 two = splat2(2.0f), one = splat2(1.0f);
  const float2 gx = sub2(div_c2(mul2(two, add2(splat2((float)j), u)), g.dw, g.rdw), one);
  const float2 gy = sub2(div_c2(mul2(two, add2(splat2((float)i), v)), g.dh, g.rdh), one);
  return tap_clamped_norm2<kGrad>(gx, gy, g);
}

struct DirectLoads { float uf, vf, ub, vb, I[3]; bool inside; };

UGL_HD float ld_once(const float* p) {
#if defined(__CUDA_ARCH__)
  return __ldcg(p);
#else
  return *p;
#endif
}

UGL_HD float2 ld_once2(const float* p) {
#if defined(__CUDA_ARCH__)
  return __ldcg(reinterpret_cast<const float2*>(p));
#else
  return *reinterpret_cast<const float2*>(p);
#endif
}
UGL_HD float4 ld_once4(const float* p) {
#if defined(__CUDA_ARCH__)
  return __ldcg(reinterpret_cast<const float4*>(p));
#else
  return *reinterpret_cast<const float4*>(p);
#endif
}

// the coalesced (non-gather) loads of one halo pixel: issued one pixel ahead of use to overlap their latency
template <int PW, int R>
UGL_HD DirectLoads load_direct(const FlowLevelDesc& L, const TileCoord& tc, int idx, int& i, int& j) {
  DirectLoads d;
  const int ly = idx / PW, lx = idx - ly * PW;
  i = tc.y0 - R + ly; j = tc.x0 - R + lx;
  d.inside = (i >= 0 && i < L.h && j >= 0 && j < L.w);
  d.uf = d.vf = d.ub = d.vb = 0.f; d.I[0] = d.I[1] = d.I[2] = 0.f;
  if (d.inside) {
    const int plane = L.h * L.w, pix = i * L.w + j;
    const float* ff = L.flow_f + (long)tc.b * 2 * plane;
    const float* fb = L.flow_b + (long)tc.b * 2 * plane;
    const float* ic = L.img + (long)tc.b * 3 * plane;
    // read once per CTA: bypass L1 so its 56 KB stay with the gather rows of the two resident CTAs
    d.uf = ld_once(ff + pix); d.vf = ld_once(ff + plane + pix);
    d.ub = ld_once(fb + pix); d.vb = ld_once(fb + plane + pix);
    d.I[0] = ld_once(ic + pix); d.I[1] = ld_once(ic + plane + pix); d.I[2] = ld_once(ic + 2 * plane + pix);
  }
  return d;
}

// kGeom = false: Model_flow weights (soft, model_flow.py:105-138); true: Model_geometry (hard occlusion * valid, model_geometry.py:105-132, 857-858)
template <bool kGrad, bool kGeom>
UGL_HD void flow_photo_pixel_c(const FlowLevelDesc& L, int b, int i, int j, const DirectLoads& d, Photo& P, float* dW) {
  const int plane = L.h * L.w, W = L.w;
  const float* ir = L.img_r + (long)b * 3 * plane;
  const float* il = L.img_l + (long)b * 3 * plane;
  const TapC2 t = flow_tap_clamped2<kGrad>(j, i, make_float2(d.uf, d.ub), make_float2(d.vf, d.vb), L.geom);
  const float* pr = ir + t.off[0];
  const float* pl = il + t.off[1];
  const float2 ksx = mul2(t.keep, splat2(L.geom.sx)), ksy = mul2(t.keep, splat2(L.geom.sy));
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    P.I[c] = d.I[c];
    // the four taps of both directions as pairs; the fused multiply-add chain is the one nvcc contracts the scalar form
    // (and ATen's CUDA sampler) into: v = nw * w_nw; v = fma(ne, w_ne, v); ...
    const float2 q0 = make_float2(pr[0], pl[0]), q1 = make_float2(pr[1], pl[1]), q2 = make_float2(pr[W], pl[W]), q3 = make_float2(pr[W + 1], pl[W + 1]);
    pr += plane; pl += plane;
    const float2 v = mul2(fma2(q3, t.w[3], fma2(q2, t.w[2], fma2(q1, t.w[1], mul2(q0, t.w[0])))), t.keep);
    P.Wf[c] = v.x;
    P.Wb[c] = v.y;
    if (kGrad) {
      const float2 gu = mul2(ksx, fma2(q3, t.dx[3], fma2(q2, t.dx[2], fma2(q1, t.dx[1], mul2(q0, t.dx[0])))));
      const float2 gv = mul2(ksy, fma2(q3, t.dy[3], fma2(q2, t.dy[2], fma2(q1, t.dy[1], mul2(q0, t.dy[0])))));
      dW[2 * c + 0] = gu.x; dW[6 + 2 * c + 0] = gu.y;
      dW[2 * c + 1] = gv.x; dW[6 + 2 * c + 1] = gv.y;
    }
  }
  const float valid_f = (P.Wf[0] == 0.f && P.Wf[1] == 0.f && P.Wf[2] == 0.f) ? 0.f : 1.f;
  const float valid_b = (P.Wb[0] == 0.f && P.Wb[1] == 0.f && P.Wb[2] == 0.f) ? 0.f : 1.f;
  P.d_f = mean3_abs_diff(P.I, P.Wf);
  P.d_b = mean3_abs_diff(P.I, P.Wb);
  float wl, wr;
  one_minus_softmax2(P.d_b, P.d_f, wl, wr);
  if (kGeom) {
    P.occ_b = wl > 0.48f ? 1.f : 0.f;
    P.occ_f = wr > 0.48f ? 1.f : 0.f;
    P.valid_b = valid_b; P.valid_f = valid_f;
    P.w_b = valid_b * P.occ_b;
    P.w_f = valid_f * P.occ_f;
  } else {
    P.w_b = soft_occ_weight(wl) * valid_b;
    P.w_f = soft_occ_weight(wr) * valid_f;
  }
}

// shared-memory planes of the single-pass kernel (halo-2 tile).  x = I*w and y = W*w are stored pre-multiplied, and the two
// warp directions of a pixel sit side by side as one float2 (.x = direction 0: forward flow / frame 0, .y = direction 1), the
// operand form of the packed fp32 instructions (add2 / mul2 / fma2) the stencil phases run on.
//   scalar planes: I (3);  pair planes (2 floats per pixel): X[c], Y[c] (c = 0..2), W over the directions; the pre-scaled
//   flows as (u, v) pairs per direction (the smoothness phases run on them)
enum GradPlane { GP_I0 = 0, GP_F2 = 3 /* pair plane (u, v) / 20 of the forward flow */, GP_B2 = 5 /* of the backward flow */, GP_X2 = 7 /* pair planes X[c] at GP_X2 + 2c */, GP_Y2 = 13, GP_W2 = 19, GP_COUNT = 21 };

template <int PN>
UGL_HD void store_grad_planes(float* sm, int idx, const Photo& P, float uf, float vf, float ub, float vb) {
  const float2 w2 = make_float2(P.w_f, P.w_b);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    sm[(GP_I0 + c) * PN + idx] = P.I[c];
    *reinterpret_cast<float2*>(sm + (GP_X2 + 2 * c) * PN + 2 * idx) = mul2(splat2(P.I[c]), w2);
    *reinterpret_cast<float2*>(sm + (GP_Y2 + 2 * c) * PN + 2 * idx) = mul2(make_float2(P.Wf[c], P.Wb[c]), w2);
  }
  *reinterpret_cast<float2*>(sm + GP_W2 * PN + 2 * idx) = w2;
  constexpr float r20 = 1.0f / 20.0f;
  *reinterpret_cast<float2*>(sm + GP_F2 * PN + 2 * idx) = div_c2(make_float2(uf, vf), 20.0f, r20);
  *reinterpret_cast<float2*>(sm + GP_B2 * PN + 2 * idx) = div_c2(make_float2(ub, vb), 20.0f, r20);
}

struct FlowGradParams {
  FlowLossParams base;
  float one = 1.0f;           // an opaque 1.0 for acc2_rn (ugl_common.cuh): keeps ptxas from contracting packed products into their sums
  float* basis[kMaxLevels];   // (B, 14, h, w) per level
  float* scratch[kMaxLevels]; // split kernels (ugl_flow_split.cuh): (B, kPhotoPairs = 15, h, w, 2) photometry pair planes per level
  // split kernels: the photometry kernel has its own tile grid (PhotoTiling) and its own partial-sum rows
  struct PhotoTiling {
    int tiles_x[kMaxLevels], tile_begin[kMaxLevels];   // tiles per row; first tile id of the level (ids ordered level, sample, ty, tx)
    int per_img[kMaxLevels];                           // tiles per sample of the level
    int per_sample;                                    // tiles of one sample over all levels (grid.x)
    int nacc;                                          // floats per partial row
    float* partials;                                   // [tiles][nacc]; nullptr: the stencil-tile rows hold every column (fused kernel)
  } photo;
  float* step_scales = nullptr;   // step mode: [B][scales][8] = FlowCombineScales of every (sample, level), written by flow_photo_norm_kernel
  int step = 0;               // split kernels, fused forward + backward (ugl_flow_loss_step): 1 = the photometry kernel hands its L1 /
                              // consistency gradient bases to the stencil kernel (3 more pair planes), which writes d loss / d flow itself
  // geom mode only (Model_geometry): in-kernel rigid flow -> dynamic mask, packed masks out
  const float* disp[kMaxLevels];          // (B,1,h,w) centre disparity
  const float* Kinv[kMaxLevels];          // (B,3,3)
  const float* P[2][kMaxLevels];          // (B,3,4): 0 = centre->left (bwd flow), 1 = centre->right (fwd flow)
  unsigned char* mask_bytes[kMaxLevels];  // (B,h,w): bit0 valid_b, bit1 valid_f, bit2 occ_b, bit3 occ_f, bit4 dyn_b, bit5 dyn_f
  float alpha, beta;                      // flow_consist_alpha / beta
  // depth mode only (Model_depth, SSIM variant): reprojection warps instead of flow warps.  Direction d = source frame d
  // (0 = left / pose[:,0], 1 = right / pose[:,1]) with P[d]; uses disp, Kinv, P above and:
  const float* src_area[2][kMaxLevels];   // (B,3,h,w) source frames, area pyramid (sampled)
  const float* src_bil[2][kMaxLevels];    // (B,3,h,w) source frames, bilinear pyramid (texture mask)
  float* valid_out[2][kMaxLevels];        // optional (B,1,h,w)
  float* tex_out[2][kMaxLevels];          // optional (B,1,h,w)
};

// kernel modes of the single-pass tile kernel
constexpr int kModeFlow = 0, kModeGeom = 1, kModeDepth = 2;

// geom mode splits the L1 term by the dynamic mask (weights 1 and 2, model_geometry.py:905-908): four more accumulators
enum GeomAcc { GA_PIXD_F = FA_COUNT, GA_WD_F, GA_PIXD_B, GA_WD_B, GA_COUNT };
constexpr unsigned kMaskValidB = 1u, kMaskValidF = 2u, kMaskOccB = 4u, kMaskOccF = 8u, kMaskDynB = 16u, kMaskDynF = 32u;

// ================================================================================================
// the single-pass stencil kernel's tile logic
// ================================================================================================
template <int TW, int TH, int NT, int kMode = kModeFlow>
struct FlowGradTile {
  static constexpr bool kGeom = (kMode == kModeGeom), kDepth = (kMode == kModeDepth);
  static constexpr int kAcc = kMode == kModeFlow ? (int)FA_COUNT : (int)GA_COUNT;
  // gradient-basis layout: flow / geom 14 planes (direction stride 8); depth 8 planes: [Gp_u, Gp_v, Gs_u, Gs_v] per direction
  static constexpr int kPlanes = kDepth ? 8 : 14, kDirStride = kDepth ? 4 : 8;
  static_assert(TW % 2 == 0, "1x2 micro-tiles need an even tile width");
  static constexpr int R = 2;
  static constexpr int PW = TW + 2 * R, PH = TH + 2 * R, PN = PW * PH;   // photometry planes (halo 2)
  static constexpr int CW = TW + 2, CH = TH + 2, CN = CW * CH;           // coefficient / edge planes (halo 1)
  static constexpr int TN = TW * TH;
  static constexpr int kOffCoef = GP_COUNT * PN;                          // 3 pair planes A, B, C of the CURRENT channel (both directions)
  static constexpr int kOffEdge = kOffCoef + 6 * CN;                      // wx, wy
  static constexpr int kOffS4 = GP_X2 * PN;                               // phase 4: 8 planes of signed weights over the (dead) X / Y pair planes
  static_assert(8 * CN <= 12 * PN, "phase-4 planes must fit into the X / Y pair planes");
  static constexpr int kOffDW = kOffEdge + 2 * CN;                        // 6 pair planes keep*dW_c/d(u|v) [2c+uv] + 2 pair planes L1 sign sums (u, v)
  static constexpr int kSmemFloats = kOffDW + 16 * TN;
  static_assert(PW % 2 == 0 && CW % 2 == 0 && PN % 4 == 0 && CN % 2 == 0 && TN % 2 == 0,
                "pair planes are read as float4 (two pixels x two directions): every plane must start 16-byte aligned");

  // phase 1: photometry on the halo-2 tile; interior pixels also: L1/weight/consistency sums, warp Jacobians,
  // L1 sign sums (shared memory) and the consistency basis (global)
  // mats (geom mode): K^-1 (9) then P_bwd (12), P_fwd (12) of this sample and level
  static UGL_HD void phase1(const FlowGradParams& gp, const TileCoord& tc, int tid, int nt, float* sm, float* acc, const float* mats = nullptr) {
    const FlowLossParams& p = gp.base;
    const FlowLevelDesc& L = p.lv[tc.level];
    const int plane = L.h * L.w;
    float* basis = gp.basis[tc.level] + (long)tc.b * kPlanes * plane;
    int i = 0, j = 0, ni = 0, nj = 0;
    DirectLoads cur = load_direct<PW, R>(L, tc, tid < PN ? tid : 0, i, j);
    for (int idx = tid; idx < PN; idx += nt) {
      // software pipeline: the coalesced loads of this thread's NEXT pixel are in flight while this one is processed
      const int nidx = idx + nt;
      DirectLoads nxt = cur;
      if (nidx < PN) nxt = load_direct<PW, R>(L, tc, nidx, ni, nj);
      const int ly = idx / PW, lx = idx - ly * PW;
      Photo P;
      const float uf = cur.uf, vf = cur.vf, ub = cur.ub, vb = cur.vb;
      if (cur.inside) {
        const int pix = i * L.w + j;
        const bool interior = (ly >= R && ly < R + TH && lx >= R && lx < R + TW);
        // one uniform code path for interior and halo pixels (a warp straddling both would otherwise execute the
        // gradient and the non-gradient variant back to back, and the second copy doubles the I-cache footprint)
        float dW[12];
        flow_photo_pixel_c<true, kGeom>(L, tc.b, i, j, cur, P, dW);
        if (interior) {
          const int t = (ly - R) * TW + (lx - R);
          float* o = sm + kOffDW + 2 * t;
#pragma unroll
          for (int k = 0; k < 6; ++k) *reinterpret_cast<float2*>(o + k * 2 * TN) = make_float2(dW[k], dW[6 + k]);
          float sgu[2], sgv[2];
#pragma unroll
          for (int dir = 0; dir < 2; ++dir) {
            float su = 0.f, sv = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float sg = sgnf((dir == 0 ? P.Wf[c] : P.Wb[c]) - P.I[c]);
              su += sg * dW[6 * dir + 2 * c];
              sv += sg * dW[6 * dir + 2 * c + 1];
            }
            sgu[dir] = su; sgv[dir] = sv;
          }
          *reinterpret_cast<float2*>(o + 12 * TN) = make_float2(sgu[0], sgu[1]);
          *reinterpret_cast<float2*>(o + 14 * TN) = make_float2(sgv[0], sgv[1]);
          float om;   // mask of the direction-consistency term: 1 - w_f (flow mode) / 1 - occ_f (geom mode)
          if (kGeom) {
            // rigid flow of the centre disparity under both poses -> dynamic masks; L1 split into rigid / dynamic parts
            const float D = gp.disp[tc.level][(long)tc.b * plane + pix];
            const Projected qb = project_pixel(mats, mats + 9, D, j, i), qf = project_pixel(mats, mats + 21, D, j, i);
            const float dyn_b = dynamic_mask_value(ub, vb, sub_rn(qb.u, (float)j), sub_rn(qb.v, (float)i), gp.alpha, gp.beta);
            const float dyn_f = dynamic_mask_value(uf, vf, sub_rn(qf.u, (float)j), sub_rn(qf.v, (float)i), gp.alpha, gp.beta);
            acc[FA_PIX_F] += P.d_f * (P.w_f * dyn_f);           acc[FA_W_F] += P.w_f * dyn_f;
            acc[GA_PIXD_F] += P.d_f * (P.w_f * (1.f - dyn_f));  acc[GA_WD_F] += P.w_f * (1.f - dyn_f);
            acc[FA_PIX_B] += P.d_b * (P.w_b * dyn_b);           acc[FA_W_B] += P.w_b * dyn_b;
            acc[GA_PIXD_B] += P.d_b * (P.w_b * (1.f - dyn_b));  acc[GA_WD_B] += P.w_b * (1.f - dyn_b);
            const unsigned bits = (P.valid_b != 0.f ? kMaskValidB : 0u) | (P.valid_f != 0.f ? kMaskValidF : 0u) |
                                  (P.occ_b != 0.f ? kMaskOccB : 0u) | (P.occ_f != 0.f ? kMaskOccF : 0u) |
                                  (dyn_b != 0.f ? kMaskDynB : 0u) | (dyn_f != 0.f ? kMaskDynF : 0u);
            gp.mask_bytes[tc.level][(long)tc.b * plane + pix] = (unsigned char)bits;
            om = 1.0f - P.occ_f;
          } else {
            acc[FA_PIX_F] += P.d_f * P.w_f;
            acc[FA_W_F] += P.w_f;
            acc[FA_PIX_B] += P.d_b * P.w_b;
            acc[FA_W_B] += P.w_b;
            om = 1.0f - P.w_f;
          }
          // direction consistency: value and un-normalised gradient w.r.t. the forward flow
          const float rf = sqrt_rn(uf * uf + vf * vf), rb = sqrt_rn(ub * ub + vb * vb);
          const float inf_ = fast_div(1.0f, rf + 1e-12f), inb_ = fast_div(1.0f, rb + 1e-12f);
          const float cu = uf * inf_ + ub * inb_, cv = vf * inf_ + vb * inb_;
          acc[FA_CONS] += (fabsf(cu) + fabsf(cv)) * om;
          acc[FA_CONS_W] += om;
          const float su = sgnf(cu) * om, sv = sgnf(cv) * om;
          const float gn = -(su * uf + sv * vf) * inf_ * inf_;
          const float ir = rf > 0.f ? fast_div(1.0f, rf) : 0.f;
          basis[6 * plane + pix] = su * inf_ + gn * uf * ir;
          basis[7 * plane + pix] = sv * inf_ + gn * vf * ir;
        }
      } else {
        zero_photo(P);
      }
      store_grad_planes<PN>(sm, idx, P, uf, vf, ub, vb);
      cur = nxt; i = ni; j = nj;
    }
  }

  // phase 1, depth mode: the warps are depth + pose reprojections (inverse_warp2, structures/inverse_warp.py:263-303) of the
  // area-resized source frames; SSIM weight = valid (model_depth_texture.py:300-301 -> compute_ssim_loss), L1 weight = valid *
  // texture mask (:296-297 -> compute_photometric_depth_loss).  Gradients are taken w.r.t. the projected pixel coordinates (u, v);
  // the combine kernel chains them to the disparity and to P.  mats: K^-1 (9), P[0] (12), P[1] (12).
  static UGL_HD void phase1_depth(const FlowGradParams& gp, const TileCoord& tc, int tid, int nt, float* sm, float* acc, const float* mats) {
    const FlowLevelDesc& L = gp.base.lv[tc.level];
    const int plane = L.h * L.w, W = L.w;
    const float* ic = L.img + (long)tc.b * 3 * plane;
    const float* dp = gp.disp[tc.level] + (long)tc.b * plane;
    for (int idx = tid; idx < PN; idx += nt) {
      const int ly = idx / PW, lx = idx - ly * PW;
      const int i = tc.y0 - R + ly, j = tc.x0 - R + lx;
      Photo P;
      if (i >= 0 && i < L.h && j >= 0 && j < W) {
        const int pix = i * W + j;
        const bool interior = (ly >= R && ly < R + TH && lx >= R && lx < R + TW);
        const float D = dp[pix];
        P.I[0] = ic[pix]; P.I[1] = ic[plane + pix]; P.I[2] = ic[2 * plane + pix];
        float dW[12], valid[2], Wd[2][3];
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const Projected pr = project_pixel(mats, mats + 9 + 12 * d, D, j, i);
          const NormCoord nc = normalise(pr, L.geom);
          valid[d] = (fabsf(nc.gx) <= 1.0f && fabsf(nc.gy) <= 1.0f) ? 1.f : 0.f;
          const TapC t = tap_clamped_norm<true>(nc.gx, nc.gy, L.geom);
          const float* ps = gp.src_area[d][tc.level] + (long)tc.b * 3 * plane + t.off;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float f0 = ps[0], f1 = ps[1], f2 = ps[W], f3 = ps[W + 1];
            ps += plane;
            float v = f0 * t.w[0]; v += f1 * t.w[1]; v += f2 * t.w[2]; v += f3 * t.w[3];
            Wd[d][c] = v;
            dW[6 * d + 2 * c + 0] = L.geom.sx * (f0 * t.dx[0] + f1 * t.dx[1] + f2 * t.dx[2] + f3 * t.dx[3]);
            dW[6 * d + 2 * c + 1] = L.geom.sy * (f0 * t.dy[0] + f1 * t.dy[1] + f2 * t.dy[2] + f3 * t.dy[3]);
          }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) { P.Wf[c] = Wd[0][c]; P.Wb[c] = Wd[1][c]; }
        P.w_f = valid[0]; P.w_b = valid[1];
        P.d_f = P.d_b = 0.f;
        if (interior) {
          const int t = (ly - R) * TW + (lx - R);
          float* o = sm + kOffDW + 2 * t;
#pragma unroll
          for (int k = 0; k < 6; ++k) *reinterpret_cast<float2*>(o + k * 2 * TN) = make_float2(dW[k], dW[6 + k]);
#pragma unroll
          for (int d = 0; d < 2; ++d) {
            const float* sb = gp.src_bil[d][tc.level] + (long)tc.b * 3 * plane + pix;
            float a = 0.f, s = 0.f, su = 0.f, sv = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              a = add_rn(a, fabsf(sub_rn(P.I[c], Wd[d][c])));
              s = add_rn(s, fabsf(sub_rn(P.I[c], sb[c * plane])));
              const float sg = sgnf(Wd[d][c] - P.I[c]);
              su += sg * dW[6 * d + 2 * c];
              sv += sg * dW[6 * d + 2 * c + 1];
            }
            const float r3 = 1.0f / 3.0f;
            const float tex = div_c(a, 3.0f, r3) < div_c(s, 3.0f, r3) ? 1.f : 0.f;     // compute_texture_mask
            const float m = mul_rn(valid[d], tex);
            o[12 * TN + d] = su * tex;                // phase 3 multiplies by the SSIM weight (valid): valid * tex in total
            o[14 * TN + d] = sv * tex;
            acc[d == 0 ? FA_PIX_F : FA_PIX_B] += a * m;
            acc[d == 0 ? FA_W_F : FA_W_B] += valid[d];
            acc[d == 0 ? GA_WD_F : GA_WD_B] += m;
            if (gp.valid_out[d][tc.level]) gp.valid_out[d][tc.level][(long)tc.b * plane + pix] = valid[d];
            if (gp.tex_out[d][tc.level]) gp.tex_out[d][tc.level][(long)tc.b * plane + pix] = tex;
          }
        }
      } else {
        zero_photo(P);
      }
      store_grad_planes<PN>(sm, idx, P, 0.f, 0.f, 0.f, 0.f);
    }
  }

  // phase 2, unit c (both directions at once).  c = 0..2: SSIM of channel c for every 1x2 strip of the halo-1 region.  A unit
  // loads its 3x4 taps once as (direction 0, direction 1) pairs (128-bit shared loads: two pixels x two directions), forms the
  // products once and accumulates the two 3x3 windows in the reference's row-major order (bit-identical SSIM), all in packed
  // fp32 instructions: one FADD2 / FMUL2 / FFMA2 does the work of both directions.  It writes the SSIM backward coefficients
  // of this channel as pairs; window centres that are interior pixels also add their SSIM loss value.  c = 3: the smoothness
  // edge weights of the strip (flow / geom modes).  The kernel runs channel by channel (phase 2 -> phase 3 accumulate) so only
  // ONE channel's coefficient planes live in shared memory: 88 KB per CTA keeps two CTAs per SM inside the 196 KB carve-out and
  // leaves 60 KB of L1 for the warps' gathers (with all three channels resident the carve-out grows to 228 KB and the gathers
  // of phase 1 miss: measured 0.424 vs 0.347 ms).
  static UGL_HD void phase2(const FlowGradParams& gp, const TileCoord& tc, int c, int tid, int nt, float* sm, float* acc) {
    const FlowLevelDesc& L = gp.base.lv[tc.level];
    constexpr int SW = CW / 2;                       // strips per row
    constexpr int NS = SW * CH;                      // strips per tile (238 for the 32x12 tile: one per thread)
    float2 ssim_sum = make_float2(0.f, 0.f);
    const float2 one = splat2(gp.one);
    for (int s = tid; s < NS; s += nt) {
      const int ly = s / SW, lx = (s - ly * SW) * 2;          // halo-1 coordinates of the left centre
      const int i = tc.y0 - 1 + ly, j0 = tc.x0 - 1 + lx;
      const int c0 = (ly + 1) * PW + (lx + 1);                // photometry-plane index of the left centre (odd: c0 - 1 is 16-byte aligned in a pair plane)
      const bool row_in = (i >= 0 && i < L.h);
      const bool in0 = row_in && j0 >= 0 && j0 < L.w, in1 = row_in && j0 + 1 >= 0 && j0 + 1 < L.w;
      if (c < 3) {
        const float* xpl = sm + (GP_X2 + 2 * c) * PN + 2 * (c0 - 1);
        const float* ypl = sm + (GP_Y2 + 2 * c) * PN + 2 * (c0 - 1);
        Moments2 m[2];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int o = 2 * (r - 1) * PW;
          const float4 xa = *reinterpret_cast<const float4*>(xpl + o), xb = *reinterpret_cast<const float4*>(xpl + o + 4);
          const float4 ya = *reinterpret_cast<const float4*>(ypl + o), yb = *reinterpret_cast<const float4*>(ypl + o + 4);
          const float2 x[4] = {lo2(xa), hi2(xa), lo2(xb), hi2(xb)}, y[4] = {lo2(ya), hi2(ya), lo2(yb), hi2(yb)};
          float2 xx[4], yy[4], xy[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) { xx[k] = mul2(x[k], x[k]); yy[k] = mul2(y[k], y[k]); xy[k] = mul2(x[k], y[k]); }
#pragma unroll
          for (int w = 0; w < 2; ++w)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              if (r == 0 && k == 0) {     // 0 + v == v: the first tap initialises the sums
                m[w].sx = x[w]; m[w].sy = y[w]; m[w].sxx = xx[w]; m[w].syy = yy[w]; m[w].sxy = xy[w];
              } else {
                m[w].sx = add2(m[w].sx, x[w + k]); m[w].sy = add2(m[w].sy, y[w + k]);
                m[w].sxx = acc2_rn(m[w].sxx, xx[w + k], one); m[w].syy = acc2_rn(m[w].syy, yy[w + k], one); m[w].sxy = acc2_rn(m[w].sxy, xy[w + k], one);
              }
            }
        }
        float2 cA[2], cB[2], cC[2];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          // both windows are evaluated unconditionally (a window centred outside the image sees zero planes: finite values) and
          // masked afterwards: no divergent region around 60 % of the phase's arithmetic
          const bool in = (w == 0 ? in0 : in1);
          const SsimTerms2 t = ssim_terms2(m[w], one);
          const float2 v = ssim_half_one_minus2(t.S, one);
          // g = 0 zeroes every coefficient of a masked window
          const float2 g = make_float2((in && v.x >= 0.f && v.x <= 1.f) ? -0.5f : 0.f, (in && v.y >= 0.f && v.y <= 1.f) ? -0.5f : 0.f);
          ssim_partials2(t, g, cA[w], cB[w], cC[w]);
          const bool interior = in && (ly >= 1 && ly <= TH && lx + w >= 1 && lx + w <= TW);
          ssim_sum.x += interior ? (v.x < 0.f ? 0.f : (v.x > 1.f ? 1.f : v.x)) : 0.f;
          ssim_sum.y += interior ? (v.y < 0.f ? 0.f : (v.y > 1.f ? 1.f : v.y)) : 0.f;
        }
        float* oc = sm + kOffCoef + 2 * (ly * CW + lx);
        *reinterpret_cast<float4*>(oc) = make_float4(cA[0].x, cA[0].y, cA[1].x, cA[1].y);
        *reinterpret_cast<float4*>(oc + 2 * CN) = make_float4(cB[0].x, cB[0].y, cB[1].x, cB[1].y);
        *reinterpret_cast<float4*>(oc + 4 * CN) = make_float4(cC[0].x, cC[0].y, cC[1].x, cC[1].y);
      } else {
        float wx[2] = {0.f, 0.f}, wy[2] = {0.f, 0.f};
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          if (o == 0 ? in0 : in1) {
            const int cc = c0 + o, j = j0 + o;
            const float Ic[3] = {sm[GP_I0 * PN + cc], sm[(GP_I0 + 1) * PN + cc], sm[(GP_I0 + 2) * PN + cc]};
            if (j >= 1 && j <= L.w - 2) {
              const float Iq[3] = {sm[GP_I0 * PN + cc + 1], sm[(GP_I0 + 1) * PN + cc + 1], sm[(GP_I0 + 2) * PN + cc + 1]};
              wx[o] = edge_weight10(Ic, Iq);
            }
            if (i >= 1 && i <= L.h - 2) {
              const float Iq[3] = {sm[GP_I0 * PN + cc + PW], sm[(GP_I0 + 1) * PN + cc + PW], sm[(GP_I0 + 2) * PN + cc + PW]};
              wy[o] = edge_weight10(Ic, Iq);
            }
          }
        }
        *reinterpret_cast<float2*>(sm + kOffEdge + ly * CW + lx) = make_float2(wx[0], wx[1]);
        *reinterpret_cast<float2*>(sm + kOffEdge + CN + ly * CW + lx) = make_float2(wy[0], wy[1]);
      }
    }
    acc[FA_SSIM_F] += ssim_sum.x;
    acc[FA_SSIM_B] += ssim_sum.y;
  }

  // phase 3 works on 1x2 strips of the interior; a thread owns strips tid, tid + nt, ... (kP3 of them) and keeps their SSIM
  // gradient sums g[n] = {gsu(left), gsu(right), gsv(left), gsv(right)} (direction pairs) in registers across the channel passes
  static constexpr int kP3 = ((TW / 2) * TH + NT - 1) / NT;

  // phase 3, channel c: 3x3 box sums of this channel's coefficient pairs (vertical sums shared by the strip's two outputs),
  // chained through the warp Jacobian of channel c into the running sums
  static UGL_HD void phase3_accumulate(const FlowGradParams& gp, const TileCoord& tc, int c, int tid, int nt, const float* sm, float2 (*g)[4]) {
    const FlowLevelDesc& L = gp.base.lv[tc.level];
    constexpr int SW = TW / 2;
    int n = 0;
    for (int s = tid; s < SW * TH; s += nt, ++n) {
      const int ty = s / SW, tx = (s - ty * SW) * 2;
      if (tc.y0 + ty >= L.h || tc.x0 + tx >= L.w) continue;
      const int c0 = (ty + R) * PW + (tx + R);       // photometry planes, left pixel (even)
      const int q0 = (ty + 1) * CW + (tx + 1);       // coefficient planes, left pixel (odd)
      const int t0 = ty * TW + tx;
      float2 sum[3][2];
#pragma unroll
      for (int k = 0; k < 3; ++k) {                 // A, B, C planes
        const float* cf = sm + kOffCoef + k * 2 * CN + 2 * (q0 - 1);
        float2 col[4];
#pragma unroll
        for (int r = -1; r <= 1; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(cf + 2 * r * CW);
          const float4 b2 = *reinterpret_cast<const float4*>(cf + 2 * r * CW + 4);
          if (r == -1) { col[0] = lo2(a); col[1] = hi2(a); col[2] = lo2(b2); col[3] = hi2(b2); }
          else { col[0] = add2(col[0], lo2(a)); col[1] = add2(col[1], hi2(a)); col[2] = add2(col[2], lo2(b2)); col[3] = add2(col[3], hi2(b2)); }
        }
        const float2 mid = add2(col[1], col[2]);
        sum[k][0] = add2(col[0], mid);
        sum[k][1] = add2(mid, col[3]);
      }
      const float4 wq = *reinterpret_cast<const float4*>(sm + GP_W2 * PN + 2 * c0);
      const float4 Xv = *reinterpret_cast<const float4*>(sm + (GP_X2 + 2 * c) * PN + 2 * c0);
      const float4 Yv = *reinterpret_cast<const float4*>(sm + (GP_Y2 + 2 * c) * PN + 2 * c0);
      const float4 du = *reinterpret_cast<const float4*>(sm + kOffDW + (2 * c) * 2 * TN + 2 * t0);
      const float4 dv = *reinterpret_cast<const float4*>(sm + kOffDW + (2 * c + 1) * 2 * TN + 2 * t0);
      const float2 wv[2] = {lo2(wq), hi2(wq)}, X2[2] = {lo2(Xv), hi2(Xv)}, Y2[2] = {lo2(Yv), hi2(Yv)};
      const float2 du2[2] = {lo2(du), hi2(du)}, dv2[2] = {lo2(dv), hi2(dv)};
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        // A + 2 y B + x C cancels heavily where the warp matches the frame: fused multiply-adds keep the rounding noise down
        const float2 gW = mul2(fma2(X2[o], sum[2][o], fma2(add2(Y2[o], Y2[o]), sum[1][o], sum[0][o])), wv[o]);
        g[n][o] = fma2(gW, du2[o], g[n][o]);
        g[n][2 + o] = fma2(gW, dv2[o], g[n][2 + o]);
      }
    }
  }

  // phase 3, closing: the L1 basis (sign sums x weight) and the accumulated SSIM basis of both directions -> global memory
  static UGL_HD void phase3_store(const FlowGradParams& gp, const TileCoord& tc, int tid, int nt, const float* sm, const float2 (*g)[4]) {
    const FlowLevelDesc& L = gp.base.lv[tc.level];
    const int plane = L.h * L.w;
    float* basis = gp.basis[tc.level] + (long)tc.b * kPlanes * plane;
    constexpr int SW = TW / 2;
    const bool vec = (L.w & 1) == 0;                  // even width: the strip's two pixels are one aligned 64-bit store
    int n = 0;
    for (int s = tid; s < SW * TH; s += nt, ++n) {
      const int ty = s / SW, tx = (s - ty * SW) * 2;
      const int i = tc.y0 + ty, j = tc.x0 + tx;
      if (i >= L.h || j >= L.w) continue;
      const int c0 = (ty + R) * PW + (tx + R);
      const int t0 = ty * TW + tx;
      const float4 wq = *reinterpret_cast<const float4*>(sm + GP_W2 * PN + 2 * c0);
      const float4 pu = *reinterpret_cast<const float4*>(sm + kOffDW + 12 * TN + 2 * t0);
      const float4 pv = *reinterpret_cast<const float4*>(sm + kOffDW + 14 * TN + 2 * t0);
      const float2 gpu[2] = {mul2(lo2(pu), lo2(wq)), mul2(hi2(pu), hi2(wq))}, gpv[2] = {mul2(lo2(pv), lo2(wq)), mul2(hi2(pv), hi2(wq))};
      const float2 *gsu = g[n], *gsv = g[n] + 2;
      const int pix = i * L.w + j;
      float* b0 = basis + pix;                               // direction 0 planes 0..3, direction 1 planes kDirStride + 0..3
      float* b1 = basis + (long)kDirStride * plane + pix;
      if (vec) {                                             // j even, width even: j + 1 < w and pix is even
        *reinterpret_cast<float2*>(b0) = make_float2(gpu[0].x, gpu[1].x);
        *reinterpret_cast<float2*>(b0 + plane) = make_float2(gpv[0].x, gpv[1].x);
        *reinterpret_cast<float2*>(b0 + 2 * (long)plane) = make_float2(gsu[0].x, gsu[1].x);
        *reinterpret_cast<float2*>(b0 + 3 * (long)plane) = make_float2(gsv[0].x, gsv[1].x);
        *reinterpret_cast<float2*>(b1) = make_float2(gpu[0].y, gpu[1].y);
        *reinterpret_cast<float2*>(b1 + plane) = make_float2(gpv[0].y, gpv[1].y);
        *reinterpret_cast<float2*>(b1 + 2 * (long)plane) = make_float2(gsu[0].y, gsu[1].y);
        *reinterpret_cast<float2*>(b1 + 3 * (long)plane) = make_float2(gsv[0].y, gsv[1].y);
      } else {
        b0[0] = gpu[0].x; b0[plane] = gpv[0].x; b0[2 * (long)plane] = gsu[0].x; b0[3 * (long)plane] = gsv[0].x;
        b1[0] = gpu[0].y; b1[plane] = gpv[0].y; b1[2 * (long)plane] = gsu[0].y; b1[3 * (long)plane] = gsv[0].y;
        if (j + 1 < L.w) {
          b0[1] = gpu[1].x; b0[plane + 1] = gpv[1].x; b0[2 * (long)plane + 1] = gsu[1].x; b0[3 * (long)plane + 1] = gsv[1].x;
          b1[1] = gpu[1].y; b1[plane + 1] = gpv[1].y; b1[2 * (long)plane + 1] = gsu[1].y; b1[3 * (long)plane + 1] = gsv[1].y;
        }
      }
    }
  }

  // phase 4a: every centre of the halo-1 region computes its signed, edge-weighted second differences ONCE
  // (s = w * sign(d2 f) per flow component and axis, 8 planes written over the no longer needed X / Y pair planes);
  // interior centres also add the smoothness loss value.  phase 4b then gathers 3 taps per axis instead of re-deriving
  // the second differences of its neighbours.
  static UGL_HD void phase4a(const FlowGradParams& gp, const TileCoord& tc, int tid, int nt, float* sm, float* acc) {
    const FlowLevelDesc& L = gp.base.lv[tc.level];
    for (int idx = tid; idx < CN; idx += nt) {
      const int ly = idx / CW, lx = idx - ly * CW;
      const int c0 = (ly + 1) * PW + (lx + 1);
      const float wx = sm[kOffEdge + idx], wy = sm[kOffEdge + CN + idx];      // zero where the centre does not exist
      const bool interior = (ly >= 1 && ly <= TH && lx >= 1 && lx <= TW) && (tc.y0 + ly - 1 < L.h) && (tc.x0 + lx - 1 < L.w);
#pragma unroll
      for (int d = 0; d < 2; ++d) {                    // forward / backward flow, (u, v) as one packed pair
        const float* f = sm + (GP_F2 + 2 * d) * PN + 2 * c0;
        // the edge weight is 0 for centres whose neighbours fall outside the image, so the reads below stay inside the
        // halo-2 planes and contribute nothing there
        const float2 c = *reinterpret_cast<const float2*>(f);
        const float2 xm = *reinterpret_cast<const float2*>(f - 2), xp = *reinterpret_cast<const float2*>(f + 2);
        const float2 ym = *reinterpret_cast<const float2*>(f - 2 * PW), yp = *reinterpret_cast<const float2*>(f + 2 * PW);
        const float2 dxx = sub2(sub2(xp, c), sub2(c, xm)), dyy = sub2(sub2(yp, c), sub2(c, ym));       // second_diff's order
        *reinterpret_cast<float2*>(sm + kOffS4 + (2 * d) * 2 * CN + 2 * idx) = make_float2(wx * sgnf(dxx.x), wx * sgnf(dxx.y));
        *reinterpret_cast<float2*>(sm + kOffS4 + (2 * d + 1) * 2 * CN + 2 * idx) = make_float2(wy * sgnf(dyy.x), wy * sgnf(dyy.y));
        if (interior) {
          acc[d == 0 ? FA_SMX_F : FA_SMX_B] += wx * fabsf(dxx.x);
          acc[d == 0 ? FA_SMY_F : FA_SMY_B] += wy * fabsf(dyy.x);
          acc[d == 0 ? FA_SMX_F : FA_SMX_B] += wx * fabsf(dxx.y);
          acc[d == 0 ? FA_SMY_F : FA_SMY_B] += wy * fabsf(dyy.y);
        }
      }
    }
  }

  static UGL_HD void phase4b(const FlowGradParams& gp, const TileCoord& tc, int tid, int nt, const float* sm) {
    const FlowLevelDesc& L = gp.base.lv[tc.level];
    const int plane = L.h * L.w;
    float* basis = gp.basis[tc.level] + (long)tc.b * kBasisPlanes * plane;
    // scales of a gradient basis: 1e-4 contract
    const float2 inx = splat2(fast_rcp(2.0f * (float)L.h * (float)(L.w - 2))), iny = splat2(fast_rcp(2.0f * (float)(L.h - 2) * (float)L.w));
    for (int idx = tid; idx < TN; idx += nt) {
      const int ty = idx / TW, tx = idx - ty * TW;
      const int i = tc.y0 + ty, j = tc.x0 + tx;
      if (i >= L.h || j >= L.w) continue;
      const int q0 = (ty + 1) * CW + (tx + 1);
      float2 g[2];
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const float* sx = sm + kOffS4 + (2 * d) * 2 * CN + 2 * q0;
        const float* sy = sm + kOffS4 + (2 * d + 1) * 2 * CN + 2 * q0;
        const float2 xc = *reinterpret_cast<const float2*>(sx), yc = *reinterpret_cast<const float2*>(sy);
        const float2 gx = add2(fma2(xc, splat2(-2.0f), *reinterpret_cast<const float2*>(sx - 2)), *reinterpret_cast<const float2*>(sx + 2));
        const float2 gy = add2(fma2(yc, splat2(-2.0f), *reinterpret_cast<const float2*>(sy - 2 * CW)), *reinterpret_cast<const float2*>(sy + 2 * CW));
        g[d] = fma2(gy, iny, mul2(gx, inx));
      }
      const int pix = i * L.w + j;
      basis[4 * plane + pix] = g[0].x; basis[5 * plane + pix] = g[0].y;
      basis[12 * plane + pix] = g[1].x; basis[13 * plane + pix] = g[1].y;
    }
  }
};

// ---- geom mode closing formulas (model_geometry.py:905-919): out = flow_pixel, flow_ssim, flow_smooth, flow_consis of one level
UGL_HD void geom_level_losses(const float* S, int h, int w, float* out) {
  const float hw = (float)h * (float)w;
  const float den_rf = S[FA_W_F] / hw + 1e-12f, den_df = S[GA_WD_F] / hw + 1e-12f;
  const float den_rb = S[FA_W_B] / hw + 1e-12f, den_db = S[GA_WD_B] / hw + 1e-12f;
  const float den_sf = (S[FA_W_F] + S[GA_WD_F]) / hw + 1e-12f, den_sb = (S[FA_W_B] + S[GA_WD_B]) / hw + 1e-12f;   // mean(valid * occ)
  out[0] = (S[FA_PIX_B] / hw) / den_rb + (S[FA_PIX_F] / hw) / den_rf + 2.0f * ((S[GA_PIXD_B] / hw) / den_db) + 2.0f * ((S[GA_PIXD_F] / hw) / den_df);
  out[1] = (S[FA_SSIM_B] / (3.0f * hw)) / den_sb + (S[FA_SSIM_F] / (3.0f * hw)) / den_sf;
  const float nx = 2.0f * (float)h * (float)(w - 2), ny = 2.0f * (float)(h - 2) * (float)w;
  out[2] = (S[FA_SMX_F] / nx + S[FA_SMY_F] / ny) * 0.5f + (S[FA_SMX_B] / nx + S[FA_SMY_B] / ny) * 0.5f;
  out[3] = (S[FA_CONS] / (2.0f * hw)) / (S[FA_CONS_W] / hw + 1e-12f);
}

struct GeomCombineScales { float pix_r[2], pix_d[2], ssim[2], sm, cons; };   // [0] = fwd flow, [1] = bwd flow

UGL_HD GeomCombineScales geom_combine_scales(const float* S, int h, int w, const float* gloss, int B, int b) {
  GeomCombineScales k;
  const float hw = (float)h * (float)w;
  const float g_pix = gloss[0 * B + b], g_ssim = gloss[1 * B + b], g_sm = gloss[2 * B + b], g_cons = gloss[3 * B + b];
  k.pix_r[0] = g_pix / hw / (S[FA_W_F] / hw + 1e-12f) / 3.0f;
  k.pix_r[1] = g_pix / hw / (S[FA_W_B] / hw + 1e-12f) / 3.0f;
  k.pix_d[0] = 2.0f * g_pix / hw / (S[GA_WD_F] / hw + 1e-12f) / 3.0f;
  k.pix_d[1] = 2.0f * g_pix / hw / (S[GA_WD_B] / hw + 1e-12f) / 3.0f;
  k.ssim[0] = g_ssim / (3.0f * hw) / ((S[FA_W_F] + S[GA_WD_F]) / hw + 1e-12f) / 9.0f;
  k.ssim[1] = g_ssim / (3.0f * hw) / ((S[FA_W_B] + S[GA_WD_B]) / hw + 1e-12f) / 9.0f;
  k.sm = g_sm * 0.5f / 20.0f;
  k.cons = g_cons / (2.0f * hw) / (S[FA_CONS_W] / hw + 1e-12f);
  return k;
}

UGL_HD void geom_combine_pixel(const float* __restrict__ basis, const unsigned char* __restrict__ mask, int plane, int pix,
                               const GeomCombineScales& k, float* __restrict__ gf, float* __restrict__ gb) {
  const unsigned bits = mask[pix];
  const float kpf = (bits & kMaskDynF) ? k.pix_r[0] : k.pix_d[0], kpb = (bits & kMaskDynB) ? k.pix_r[1] : k.pix_d[1];
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    gf[ch * plane + pix] = kpf * basis[(0 + ch) * plane + pix] + k.ssim[0] * basis[(2 + ch) * plane + pix]
                           + k.sm * basis[(4 + ch) * plane + pix] + k.cons * basis[(6 + ch) * plane + pix];
    gb[ch * plane + pix] = kpb * basis[(8 + ch) * plane + pix] + k.ssim[1] * basis[(10 + ch) * plane + pix]
                           + k.sm * basis[(12 + ch) * plane + pix];
  }
}

// ---- depth mode closing formulas: out[0] = loss_depth_pixel, out[1] = loss_depth_ssim of one level (both directions) ----
UGL_HD void depth_level_losses(const float* S, int h, int w, float* out) {
  const float hw = (float)h * (float)w;
  out[0] = (S[FA_PIX_F] / (3.0f * hw)) / (S[GA_WD_F] / hw + 1e-12f) + (S[FA_PIX_B] / (3.0f * hw)) / (S[GA_WD_B] / hw + 1e-12f);
  out[1] = (S[FA_SSIM_F] / (3.0f * hw)) / (S[FA_W_F] / hw + 1e-12f) + (S[FA_SSIM_B] / (3.0f * hw)) / (S[FA_W_B] / hw + 1e-12f);
  out[2] = 0.f; out[3] = 0.f;
}
// scales of the L1 / SSIM basis of direction d given the upstream gradients gloss (2,B)
UGL_HD void depth_combine_scales(const float* S, int h, int w, const float* gloss, int B, int b, int d, float& k_pix, float& k_ssim) {
  const float hw = (float)h * (float)w;
  k_pix = gloss[0 * B + b] / hw / (S[d == 0 ? GA_WD_F : GA_WD_B] / hw + 1e-12f) / 3.0f;
  k_ssim = gloss[1 * B + b] / (3.0f * hw) / (S[d == 0 ? FA_W_F : FA_W_B] / hw + 1e-12f) / 9.0f;
}

// ---- backward = element-wise combine -----------------------------------------------------------------------------
// grad_f = kp_f Gp_f + ks_f Gs_f + ksm Gm_f + kc Gc ;  grad_b = kp_b Gp_b + ks_b Gs_b + ksm Gm_b
struct FlowCombineScales { float pix[2], ssim[2], sm, cons; };

UGL_HD FlowCombineScales flow_combine_scales(const float* S, int h, int w, const float* gloss, int B, int b) {
  FlowCombineScales k;
  const float hw = (float)h * (float)w;
  const float den_f = S[FA_W_F] / hw + 1e-12f, den_b = S[FA_W_B] / hw + 1e-12f;
  const float g_pix = gloss[0 * B + b], g_ssim = gloss[1 * B + b], g_sm = gloss[2 * B + b], g_cons = gloss[3 * B + b];
  k.pix[0] = g_pix / hw / den_f / 3.0f;
  k.pix[1] = g_pix / hw / den_b / 3.0f;
  k.ssim[0] = g_ssim / (3.0f * hw) / den_f / 9.0f;
  k.ssim[1] = g_ssim / (3.0f * hw) / den_b / 9.0f;
  k.sm = g_sm * 0.5f / 20.0f;
  k.cons = g_cons / (2.0f * hw) / (S[FA_CONS_W] / hw + 1e-12f);
  return k;
}

UGL_HD void flow_combine_pixel(const float* __restrict__ basis, int plane, int pix, const FlowCombineScales& k,
                               float* __restrict__ gf, float* __restrict__ gb) {
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    gf[ch * plane + pix] = k.pix[0] * basis[(0 + ch) * plane + pix] + k.ssim[0] * basis[(2 + ch) * plane + pix]
                           + k.sm * basis[(4 + ch) * plane + pix] + k.cons * basis[(6 + ch) * plane + pix];
    gb[ch * plane + pix] = k.pix[1] * basis[(8 + ch) * plane + pix] + k.ssim[1] * basis[(10 + ch) * plane + pix]
                           + k.sm * basis[(12 + ch) * plane + pix];
  }
}

}  // namespace ugl
