// Fused flow-mode photometric loss (reference model_flow.py:232-254):
// per pyramid level and sample, in ONE pass over the pixels:
//   backward-warp of the left/right frame by the bwd/fwd flow with validity mask   (W1, net_utils.py:16-54)
//   valid-pixel mask, soft bidirectional occlusion weights                          (M2/M3, model_flow.py:105-138)
//   weighted L1                                                                     (L1, model_flow.py:94-103)
//   3x3 SSIM with the weight multiplied into both inputs                            (L2, model_flow.py:141-152, ssim.py:4-19)
//   edge-aware second-order flow smoothness                                         (L4, model_flow.py:156-181)
//   forward/backward flow direction consistency                                     (L5, model_flow.py:184-199)
// The forward kernel produces per-tile partial sums; a finalize kernel turns them into the four
// (B,) loss vectors and keeps the per-(sample, level) sums for the backward pass, which recomputes
// the per-pixel quantities from the inputs (nothing per-pixel is saved) and writes d loss / d flow.
//
// Tile logic is written as phase functions (tid, shared buffer) so the same code runs inside the
// CUDA kernels (ugl_flow_loss.cu) and in the host emulator (tests/hostemu/).
//
// The kernels are instruction-issue bound, not HBM bound (profiles/): the arithmetic below is
// organised to keep exact-rounding where parity needs it (bilinear cell, SSIM moments) at the lowest
// instruction count (div_c instead of IEEE division, flows pre-scaled once, edge weights computed once).
#pragma once

#include "ugl_common.cuh"

namespace ugl {

// tile shapes (interior pixels per CTA) of the forward and backward kernels
#ifndef UGL_BTW
#define UGL_BTW 32
#endif
#ifndef UGL_BTH
#define UGL_BTH 13     /* swept on B200 (profiles/r1e_tile_sweep.txt, r1m): 32x13x256 is the fastest single-pass tile that keeps two CTAs inside the 196 KB carve-out */
#endif
#ifndef UGL_BMINB
#define UGL_BMINB 2
#endif
constexpr int kFTW = 32, kFTH = 16;
constexpr int kBTW = UGL_BTW, kBTH = UGL_BTH;   // backward / single-pass tiles (tunable at build time for experiments)

// per-(sample, level) accumulators
enum FlowAcc {
  FA_PIX_F = 0,  // sum d_r * w_f
  FA_W_F,        // sum w_f
  FA_PIX_B,      // sum d_l * w_b
  FA_W_B,        // sum w_b
  FA_SSIM_F,     // sum_{c,p} clamp((1-ssim)/2)
  FA_SSIM_B,
  FA_SMX_F,      // sum_{ch,p} wx * |dxx f/20|
  FA_SMY_F,
  FA_SMX_B,
  FA_SMY_B,
  FA_CONS,       // sum_{ch,p} |f^_fwd + f^_bwd| * (1 - w_f)
  FA_CONS_W,     // sum (1 - w_f)
  FA_COUNT
};

struct FlowLevelDesc {
  int h, w;
  int tiles_x, tiles_y;
  int tile_begin;                       // first tile id of this level; tiles ordered (level, b, ty, tx)
  WarpGeom geom;                        // (w-1), (h-1) normalisers and their reciprocals
  const float* img_l;                   // (B,3,h,w) left  frame pyramid level (warped by the bwd flow)
  const float* img;                     // (B,3,h,w) centre frame
  const float* img_r;                   // (B,3,h,w) right frame (warped by the fwd flow)
  const float* flow_f;                  // (B,2,h,w) centre -> right
  const float* flow_b;                  // (B,2,h,w) centre -> left
  float* gflow_f;                       // (B,2,h,w) backward outputs
  float* gflow_b;
};

struct FlowLossParams {
  int B;
  int scales;                           // levels that carry a loss (reference: self.num_scales)
  int total_tiles;
  FlowLevelDesc lv[kMaxLevels];
  float* partials;                      // [total_tiles][FA_COUNT]
  float* stats;                         // [B][scales][FA_COUNT] level sums, kept for backward
  float* loss;                          // [4][B]: pixel, ssim, smooth, consis
  const float* gloss;                   // [4][B] upstream gradient (backward only)
};

struct TileCoord { int level, b, x0, y0; };

template <int TW, int TH>
UGL_HD TileCoord decode_tile(const FlowLossParams& p, int tile) {
  TileCoord tc;
  int l = 0;
  while (l + 1 < p.scales && tile >= p.lv[l + 1].tile_begin) ++l;
  const FlowLevelDesc& L = p.lv[l];
  int r = tile - L.tile_begin;
  const int per_img = L.tiles_x * L.tiles_y;
  tc.level = l;
  tc.b = r / per_img;
  r -= tc.b * per_img;
  tc.y0 = (r / L.tiles_x) * TH;
  tc.x0 = (r % L.tiles_x) * TW;
  return tc;
}

// The single-pass kernel is launched on a (tiles per sample over all levels, B) grid: the sample is blockIdx.y, so the decode
// needs one integer division (row / column of the tile) instead of three.  tile_id keeps the (level, b, ty, tx) order of the
// partial-sum rows the finalize kernel reads.
template <int TW, int TH>
UGL_HD TileCoord decode_tile_2d(const FlowLossParams& p, int r, int b, int& tile_id) {
  TileCoord tc;
  int l = 0, per_img = p.lv[0].tiles_x * p.lv[0].tiles_y;
  while (l + 1 < p.scales && r >= per_img) {     // r counts this sample's tiles over all levels
    r -= per_img;
    ++l;
    per_img = p.lv[l].tiles_x * p.lv[l].tiles_y;
  }
  const FlowLevelDesc& L = p.lv[l];
  tile_id = L.tile_begin + b * per_img + r;
  const int ty = r / L.tiles_x;
  tc.level = l;
  tc.b = b;
  tc.y0 = ty * TH;
  tc.x0 = (r - ty * L.tiles_x) * TW;
  return tc;
}

// ---- per-pixel photometry shared by forward and backward ----------------------------------------
struct Photo {
  float I[3], Wf[3], Wb[3];   // centre image, warped-from-right (fwd flow), warped-from-left (bwd flow)
  float d_f, d_b;             // mean_c |I - Wf|, mean_c |I - Wb|  (reference: img_diff_r, img_diff_l)
  float w_f, w_b;             // flow mode: soft occlusion weights * valid; geom mode: valid * hard occlusion mask
  float occ_f, occ_b;         // geom mode: hard occlusion masks [1 - softmax > 0.48]
  float valid_f, valid_b;     // geom mode: warp valid masks
};

UGL_HD float mean3_abs_diff(const float* a, const float* b) {
  const float s = add_rn(add_rn(fabsf(sub_rn(a[0], b[0])), fabsf(sub_rn(a[1], b[1]))), fabsf(sub_rn(a[2], b[2])));
  return div_c(s, 3.0f, 1.0f / 3.0f);
}

// edge weight exp(-10 * mean_c |I(q) - I(p)|) of the smoothness term (model_flow.py:163-164)
UGL_HD float edge_weight10(const float* Ip, const float* Iq) {
  const float s = add_rn(add_rn(fabsf(sub_rn(Iq[0], Ip[0])), fabsf(sub_rn(Iq[1], Ip[1]))), fabsf(sub_rn(Iq[2], Ip[2])));
  return expf(-10.0f * div_c(s, 3.0f, 1.0f / 3.0f));
}

// kGrad: also return keep * d W_c / d(u,v) for both directions (12 floats: f: c0u,c0v,c1u,..., then b)
template <bool kGrad>
UGL_HD void flow_photo_pixel(const FlowLevelDesc& L, int b, int i, int j, float uf, float vf, float ub, float vb,
                             Photo& P, float* dW) {
  const long plane = (long)L.h * L.w;
  const long pix = (long)i * L.w + j;
  const float* ic = L.img + (long)b * 3 * plane;
  const float* ir = L.img_r + (long)b * 3 * plane;
  const float* il = L.img_l + (long)b * 3 * plane;
  const Tap tf = flow_tap(j, i, uf, vf, L.geom);
  const Tap tb = flow_tap(j, i, ub, vb, L.geom);
  const float keep_f = tap_keep(tf), keep_b = tap_keep(tb);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    P.I[c] = ic[c * plane + pix];
    const Corners cf = tap_fetch(ir + c * plane, L.w, tf);
    const Corners cb = tap_fetch(il + c * plane, L.w, tb);
    P.Wf[c] = corners_value(cf, tf) * keep_f;
    P.Wb[c] = corners_value(cb, tb) * keep_b;
    if (kGrad) {
      dW[2 * c + 0] = keep_f * corners_ddx(cf, tf) * L.geom.sx;
      dW[2 * c + 1] = keep_f * corners_ddy(cf, tf) * L.geom.sy;
      dW[6 + 2 * c + 0] = keep_b * corners_ddx(cb, tb) * L.geom.sx;
      dW[6 + 2 * c + 1] = keep_b * corners_ddy(cb, tb) * L.geom.sy;
    }
  }
  const float valid_f = (P.Wf[0] == 0.f && P.Wf[1] == 0.f && P.Wf[2] == 0.f) ? 0.f : 1.f;
  const float valid_b = (P.Wb[0] == 0.f && P.Wb[1] == 0.f && P.Wb[2] == 0.f) ? 0.f : 1.f;
  P.d_f = mean3_abs_diff(P.I, P.Wf);
  P.d_b = mean3_abs_diff(P.I, P.Wb);
  float wl, wr;
  one_minus_softmax2(P.d_b, P.d_f, wl, wr);     // channel 0 = from-left (bwd), channel 1 = from-right (fwd)
  P.w_b = soft_occ_weight(wl) * valid_b;
  P.w_f = soft_occ_weight(wr) * valid_f;
}

// direction consistency value for one pixel: (|f^u_f + f^u_b| + |f^v_f + f^v_b|)  (model_flow.py:184-199)
UGL_HD float consis_value(float uf, float vf, float ub, float vb) {
  const float inf_ = fast_div(1.0f, sqrt_rn(uf * uf + vf * vf) + 1e-12f);
  const float inb_ = fast_div(1.0f, sqrt_rn(ub * ub + vb * vb) + 1e-12f);
  return fabsf(uf * inf_ + ub * inb_) + fabsf(vf * inf_ + vb * inb_);
}

// shared-memory planes of the halo'd tile (structure of arrays).  Flows are stored pre-divided by 20
// (model_flow.py:177 `flow/20.0`), the only form the stencils need.
enum FlowPlane { PL_I0 = 0, PL_I1, PL_I2, PL_F0, PL_F1, PL_F2, PL_B0, PL_B1, PL_B2, PL_WF, PL_WB, PL_UF, PL_VF, PL_UB, PL_VB, PL_COUNT };
static_assert(PL_B0 == PL_F0 + 3 && PL_WB == PL_WF + 1, "the single-pass kernel indexes the direction planes arithmetically");

template <int PN>
UGL_HD void store_photo_planes(float* sm, int idx, const Photo& P, float uf, float vf, float ub, float vb) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    sm[(PL_I0 + c) * PN + idx] = P.I[c];
    sm[(PL_F0 + c) * PN + idx] = P.Wf[c];
    sm[(PL_B0 + c) * PN + idx] = P.Wb[c];
  }
  sm[PL_WF * PN + idx] = P.w_f;
  sm[PL_WB * PN + idx] = P.w_b;
  constexpr float r20 = 1.0f / 20.0f;
  sm[PL_UF * PN + idx] = div_c(uf, 20.0f, r20);
  sm[PL_VF * PN + idx] = div_c(vf, 20.0f, r20);
  sm[PL_UB * PN + idx] = div_c(ub, 20.0f, r20);
  sm[PL_VB * PN + idx] = div_c(vb, 20.0f, r20);
}

UGL_HD void zero_photo(Photo& P) {
#pragma unroll
  for (int c = 0; c < 3; ++c) P.I[c] = P.Wf[c] = P.Wb[c] = 0.f;
  P.w_f = P.w_b = P.d_f = P.d_b = 0.f;   // zero padding of the 3x3 pooling: x = y = 0 outside the image
}

// 3x3 window sums of x = I*w, y = W*w around plane index c0 (row pitch PW), one direction, one channel
template <int PW>
UGL_HD Moments window_moments(const float* ipl, const float* ypl, const float* wpl, int c0) {
  Moments m = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int q = c0 + dy * PW + dx;
      const float wq = wpl[q];
      moments_add(m, mul_rn(ipl[q], wq), mul_rn(ypl[q], wq));
    }
  return m;
}

// second difference of the pre-scaled flow around f[0] with stride s, in the reference's order
UGL_HD float second_diff(const float* f, int s) { return sub_rn(sub_rn(f[s], f[0]), sub_rn(f[0], f[-s])); }

// ================================================================================================
// forward
// ================================================================================================
template <int TW, int TH>
struct FlowFwdTile {
  static constexpr int R = 1;
  static constexpr int PW = TW + 2 * R, PH = TH + 2 * R, PN = PW * PH;
  static constexpr int kSmemFloats = PL_COUNT * PN;

  // phase 1: photometry on the halo'd tile -> shared planes; L1 / weight / consistency sums for interior pixels
  static UGL_HD void phase1(const FlowLossParams& p, const TileCoord& tc, int tid, int nt, float* sm, float* acc) {
    const FlowLevelDesc& L = p.lv[tc.level];
    const long plane = (long)L.h * L.w;
    for (int idx = tid; idx < PN; idx += nt) {
      const int ly = idx / PW, lx = idx - ly * PW;
      const int i = tc.y0 - R + ly, j = tc.x0 - R + lx;
      Photo P;
      float uf = 0.f, vf = 0.f, ub = 0.f, vb = 0.f;
      if (i >= 0 && i < L.h && j >= 0 && j < L.w) {
        const long pix = (long)i * L.w + j;
        const float* ff = L.flow_f + (long)tc.b * 2 * plane;
        const float* fb = L.flow_b + (long)tc.b * 2 * plane;
        uf = ff[pix]; vf = ff[plane + pix];
        ub = fb[pix]; vb = fb[plane + pix];
        flow_photo_pixel<false>(L, tc.b, i, j, uf, vf, ub, vb, P, nullptr);
        const bool interior = (ly >= R && ly < R + TH && lx >= R && lx < R + TW);
        if (interior) {
          acc[FA_PIX_F] += P.d_f * P.w_f;
          acc[FA_W_F] += P.w_f;
          acc[FA_PIX_B] += P.d_b * P.w_b;
          acc[FA_W_B] += P.w_b;
          const float om = 1.0f - P.w_f;
          acc[FA_CONS] += consis_value(uf, vf, ub, vb) * om;
          acc[FA_CONS_W] += om;
        }
      } else {
        zero_photo(P);
      }
      store_photo_planes<PN>(sm, idx, P, uf, vf, ub, vb);
    }
  }

  // phase 2: SSIM and smoothness for interior pixels
  static UGL_HD void phase2(const FlowLossParams& p, const TileCoord& tc, int tid, int nt, const float* sm, float* acc) {
    const FlowLevelDesc& L = p.lv[tc.level];
    for (int idx = tid; idx < TW * TH; idx += nt) {
      const int ty = idx / TW, tx = idx - ty * TW;
      const int i = tc.y0 + ty, j = tc.x0 + tx;
      if (i >= L.h || j >= L.w) continue;
      const int c0 = (ty + R) * PW + (tx + R);
      // ---- SSIM, both directions, three channels
#pragma unroll
      for (int dir = 0; dir < 2; ++dir) {
        const float* wpl = sm + (dir == 0 ? PL_WF : PL_WB) * PN;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const Moments m = window_moments<PW>(sm + (PL_I0 + c) * PN, sm + ((dir == 0 ? PL_F0 : PL_B0) + c) * PN, wpl, c0);
          s += ssim_loss_value(ssim_from_sums(m));
        }
        acc[dir == 0 ? FA_SSIM_F : FA_SSIM_B] += s;
      }
      // ---- edge-aware second-order smoothness (flow / 20), centre-pixel form:
      //      the weight of edge (c, c+1) pairs with the second difference centred at c
      const float Ic[3] = {sm[PL_I0 * PN + c0], sm[PL_I1 * PN + c0], sm[PL_I2 * PN + c0]};
      if (j >= 1 && j <= L.w - 2) {
        const float Iq[3] = {sm[PL_I0 * PN + c0 + 1], sm[PL_I1 * PN + c0 + 1], sm[PL_I2 * PN + c0 + 1]};
        const float wx = edge_weight10(Ic, Iq);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k < 2 ? FA_SMX_F : FA_SMX_B] += wx * fabsf(second_diff(sm + (PL_UF + k) * PN + c0, 1));
      }
      if (i >= 1 && i <= L.h - 2) {
        const float Iq[3] = {sm[PL_I0 * PN + c0 + PW], sm[PL_I1 * PN + c0 + PW], sm[PL_I2 * PN + c0 + PW]};
        const float wy = edge_weight10(Ic, Iq);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k < 2 ? FA_SMY_F : FA_SMY_B] += wy * fabsf(second_diff(sm + (PL_UF + k) * PN + c0, PW));
      }
    }
  }
};

// ---- finalize: level sums -> the four per-sample loss terms of one level -------------------------
// out[0..3] = pixel, ssim, smooth, consis  (model_flow.py:244-254)
UGL_HD void flow_level_losses(const float* S, int h, int w, float* out) {
  const float hw = (float)h * (float)w;
  const float den_f = S[FA_W_F] / hw + 1e-12f, den_b = S[FA_W_B] / hw + 1e-12f;
  out[0] = (S[FA_PIX_F] / hw) / den_f + (S[FA_PIX_B] / hw) / den_b;
  out[1] = (S[FA_SSIM_F] / (3.0f * hw)) / den_f + (S[FA_SSIM_B] / (3.0f * hw)) / den_b;
  const float nx = 2.0f * (float)h * (float)(w - 2), ny = 2.0f * (float)(h - 2) * (float)w;
  out[2] = (S[FA_SMX_F] / nx + S[FA_SMY_F] / ny) * 0.5f + (S[FA_SMX_B] / nx + S[FA_SMY_B] / ny) * 0.5f;
  out[3] = (S[FA_CONS] / (2.0f * hw)) / (S[FA_CONS_W] / hw + 1e-12f);
}

// per-(sample, level) scale factors of the backward pass
struct FlowBwdCoef {
  float pix[2];     // multiplies w * d(d)/dW   (includes 1/hw, the divider and the 1/3 of the channel mean)
  float ssim[2];    // multiplies d clamp/dS chain (includes 1/(3hw), the divider and 1/9)
  float smx, smy;   // includes go/2, 1/20 and the element counts
  float cons;
};

UGL_HD FlowBwdCoef flow_bwd_coef(const float* S, int h, int w, const float* gloss, int B, int b) {
  FlowBwdCoef k;
  const float hw = (float)h * (float)w;
  const float den_f = S[FA_W_F] / hw + 1e-12f, den_b = S[FA_W_B] / hw + 1e-12f;
  const float g_pix = gloss[0 * B + b], g_ssim = gloss[1 * B + b], g_sm = gloss[2 * B + b], g_cons = gloss[3 * B + b];
  k.pix[0] = g_pix / hw / den_f / 3.0f;
  k.pix[1] = g_pix / hw / den_b / 3.0f;
  k.ssim[0] = g_ssim / (3.0f * hw) / den_f / 9.0f;
  k.ssim[1] = g_ssim / (3.0f * hw) / den_b / 9.0f;
  k.smx = g_sm * 0.5f / (2.0f * (float)h * (float)(w - 2)) / 20.0f;
  k.smy = g_sm * 0.5f / (2.0f * (float)(h - 2) * (float)w) / 20.0f;
  k.cons = g_cons / (2.0f * hw) / (S[FA_CONS_W] / hw + 1e-12f);
  return k;
}

// ================================================================================================
// backward (recompute)
// ================================================================================================
// Shared memory: 15 photometry planes on the halo-2 tile; for ONE direction at a time 9 SSIM coefficient
// planes on the halo-1 tile; 2 edge-weight planes (halo 1); 16 interior planes (keep * dW/d(u,v) and the
// L1 sign sums).  The two directions reuse the coefficient planes.
template <int TW, int TH, int NT>
struct FlowBwdTile {
  static constexpr int R = 2;
  static constexpr int PW = TW + 2 * R, PH = TH + 2 * R, PN = PW * PH;   // photometry planes (halo 2)
  static constexpr int CW = TW + 2, CH = TH + 2, CN = CW * CH;           // coefficient / edge-weight planes (halo 1)
  static constexpr int TN = TW * TH;
  static constexpr int PPT = (TN + NT - 1) / NT;                          // interior pixels owned by one thread
  static constexpr int kOffCoef = PL_COUNT * PN;                          // 9 planes: [c][A,B,C] of the current direction
  static constexpr int kOffEdge = kOffCoef + 9 * CN;                      // 2 planes: wx (edge c,c+1), wy
  static constexpr int kOffDW = kOffEdge + 2 * CN;                        // 12 planes keep*dW/d(u,v) + 4 planes L1 sign sums
  static constexpr int kSmemFloats = kOffDW + 16 * TN;

  static UGL_HD void phase1(const FlowLossParams& p, const TileCoord& tc, int tid, int nt, float* sm) {
    const FlowLevelDesc& L = p.lv[tc.level];
    const long plane = (long)L.h * L.w;
    for (int idx = tid; idx < PN; idx += nt) {
      const int ly = idx / PW, lx = idx - ly * PW;
      const int i = tc.y0 - R + ly, j = tc.x0 - R + lx;
      Photo P;
      float uf = 0.f, vf = 0.f, ub = 0.f, vb = 0.f;
      if (i >= 0 && i < L.h && j >= 0 && j < L.w) {
        const long pix = (long)i * L.w + j;
        const float* ff = L.flow_f + (long)tc.b * 2 * plane;
        const float* fb = L.flow_b + (long)tc.b * 2 * plane;
        uf = ff[pix]; vf = ff[plane + pix];
        ub = fb[pix]; vb = fb[plane + pix];
        const bool interior = (ly >= R && ly < R + TH && lx >= R && lx < R + TW);
        if (interior) {
          float dW[12];
          flow_photo_pixel<true>(L, tc.b, i, j, uf, vf, ub, vb, P, dW);
          const int t = (ly - R) * TW + (lx - R);
          float* o = sm + kOffDW + t;
#pragma unroll
          for (int k = 0; k < 12; ++k) o[k * TN] = dW[k];
          // L1 term: sum_c sign(W_c - I_c) * keep * dW_c/d(u,v), per direction
#pragma unroll
          for (int dir = 0; dir < 2; ++dir) {
            float su = 0.f, sv = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float sg = sgnf((dir == 0 ? P.Wf[c] : P.Wb[c]) - P.I[c]);
              su += sg * dW[6 * dir + 2 * c];
              sv += sg * dW[6 * dir + 2 * c + 1];
            }
            o[(12 + 2 * dir) * TN] = su;
            o[(13 + 2 * dir) * TN] = sv;
          }
        } else {
          flow_photo_pixel<false>(L, tc.b, i, j, uf, vf, ub, vb, P, nullptr);
        }
      } else {
        zero_photo(P);
      }
      store_photo_planes<PN>(sm, idx, P, uf, vf, ub, vb);
    }
  }

  // phase 2 (per direction): SSIM backward coefficients for every window centre in the halo-1 region;
  // on the first direction also the smoothness edge weights of that pixel
  static UGL_HD void phase2(const FlowLossParams& p, const TileCoord& tc, int dir, int tid, int nt, float* sm) {
    const FlowLevelDesc& L = p.lv[tc.level];
    const float* wpl = sm + (dir == 0 ? PL_WF : PL_WB) * PN;
    for (int idx = tid; idx < CN; idx += nt) {
      const int ly = idx / CW, lx = idx - ly * CW;
      const int i = tc.y0 - 1 + ly, j = tc.x0 - 1 + lx;
      const bool inside = (i >= 0 && i < L.h && j >= 0 && j < L.w);
      const int c0 = (ly + 1) * PW + (lx + 1);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float cA = 0.f, cB = 0.f, cC = 0.f;
        if (inside) {
          const Moments m = window_moments<PW>(sm + (PL_I0 + c) * PN, sm + ((dir == 0 ? PL_F0 : PL_B0) + c) * PN, wpl, c0);
          ssim_backward_coeffs(m, cA, cB, cC);
        }
        float* o = sm + kOffCoef + (c * 3) * CN + idx;
        o[0] = cA; o[CN] = cB; o[2 * CN] = cC;
      }
      if (dir == 0) {
        float wx = 0.f, wy = 0.f;
        if (inside) {
          const float Ic[3] = {sm[PL_I0 * PN + c0], sm[PL_I1 * PN + c0], sm[PL_I2 * PN + c0]};
          if (j >= 1 && j <= L.w - 2) {
            const float Iq[3] = {sm[PL_I0 * PN + c0 + 1], sm[PL_I1 * PN + c0 + 1], sm[PL_I2 * PN + c0 + 1]};
            wx = edge_weight10(Ic, Iq);
          }
          if (i >= 1 && i <= L.h - 2) {
            const float Iq[3] = {sm[PL_I0 * PN + c0 + PW], sm[PL_I1 * PN + c0 + PW], sm[PL_I2 * PN + c0 + PW]};
            wy = edge_weight10(Ic, Iq);
          }
        }
        sm[kOffEdge + idx] = wx;          // zero where that centre does not exist
        sm[kOffEdge + CN + idx] = wy;
      }
    }
  }

  // phase 3 (per direction): photometric gradient of this direction's flow for the owned interior pixels
  static UGL_HD void phase3(const FlowLossParams& p, const TileCoord& tc, const FlowBwdCoef& k, int dir, int tid, int nt,
                            const float* sm, float (*g)[4]) {
    const FlowLevelDesc& L = p.lv[tc.level];
    const float kp = k.pix[dir], ks = k.ssim[dir];
    int n = 0;
    for (int idx = tid; idx < TN; idx += nt, ++n) {
      const int ty = idx / TW, tx = idx - ty * TW;
      if (tc.y0 + ty >= L.h || tc.x0 + tx >= L.w) continue;
      const int c0 = (ty + R) * PW + (tx + R);       // photometry planes
      const int q0 = (ty + 1) * CW + (tx + 1);       // coefficient planes
      const float wq = sm[(dir == 0 ? PL_WF : PL_WB) * PN + c0];
      const float* dw = sm + kOffDW + idx;
      float gu = kp * wq * dw[(12 + 2 * dir) * TN], gv = kp * wq * dw[(13 + 2 * dir) * TN];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float Ic = sm[(PL_I0 + c) * PN + c0];
        const float Wc = sm[((dir == 0 ? PL_F0 : PL_B0) + c) * PN + c0];
        const float* cf = sm + kOffCoef + (c * 3) * CN + q0;
        float sA = 0.f, sB = 0.f, sC = 0.f;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const int q = dy * CW + dx;
            sA += cf[q]; sB += cf[CN + q]; sC += cf[2 * CN + q];
          }
        const float gW = (sA + 2.0f * (Wc * wq) * sB + (Ic * wq) * sC) * (ks * wq);   // d L_ssim / d W_c
        gu += gW * dw[(6 * dir + 2 * c) * TN];
        gv += gW * dw[(6 * dir + 2 * c + 1) * TN];
      }
      g[n][2 * dir] = gu;
      g[n][2 * dir + 1] = gv;
    }
  }

  // phase 4: smoothness + consistency gradients, final store
  static UGL_HD void phase4(const FlowLossParams& p, const TileCoord& tc, const FlowBwdCoef& k, int tid, int nt,
                            const float* sm, float (*g)[4]) {
    const FlowLevelDesc& L = p.lv[tc.level];
    const long plane = (long)L.h * L.w;
    int n = 0;
    for (int idx = tid; idx < TN; idx += nt, ++n) {
      const int ty = idx / TW, tx = idx - ty * TW;
      const int i = tc.y0 + ty, j = tc.x0 + tx;
      if (i >= L.h || j >= L.w) continue;
      const int c0 = (ty + R) * PW + (tx + R);
      const int q0 = (ty + 1) * CW + (tx + 1);
      float go[4] = {g[n][0], g[n][1], g[n][2], g[n][3]};
      // second-order smoothness: centres c in {j-1, j, j+1} / {i-1, i, i+1}; the edge-weight planes are
      // zero where a centre does not exist
#pragma unroll
      for (int t = -1; t <= 1; ++t) {
        const float coef = (t == 0) ? -2.0f : 1.0f;
        const float wx = sm[kOffEdge + q0 + t] * coef * k.smx;
        const float wy = sm[kOffEdge + CN + q0 + t * CW] * coef * k.smy;
#pragma unroll
        for (int f4 = 0; f4 < 4; ++f4) {
          const float* f = sm + (PL_UF + f4) * PN + c0;
          go[f4] += wx * sgnf(second_diff(f + t, 1)) + wy * sgnf(second_diff(f + t * PW, PW));
        }
      }
      // direction consistency: gradient reaches the forward flow only (bwd branch detached)
      const long pix = (long)i * L.w + j;
      {
        const float* ff = L.flow_f + (long)tc.b * 2 * plane;
        const float* fb = L.flow_b + (long)tc.b * 2 * plane;
        const float uf = ff[pix], vf = ff[plane + pix], ub = fb[pix], vb = fb[plane + pix];
        const float rf = sqrt_rn(uf * uf + vf * vf), rb = sqrt_rn(ub * ub + vb * vb);
        const float inf_ = fast_div(1.0f, rf + 1e-12f), inb_ = fast_div(1.0f, rb + 1e-12f);
        const float om = (1.0f - sm[PL_WF * PN + c0]) * k.cons;
        const float su = sgnf(uf * inf_ + ub * inb_) * om, sv = sgnf(vf * inf_ + vb * inb_) * om;   // d L / d f^_u, f^_v
        // f^ = f / n, n = r + eps; d n / d f = f / r (0 where r == 0, torch.norm backward)
        const float gn = -(su * uf + sv * vf) * inf_ * inf_;                                          // d L / d n
        const float ir = rf > 0.f ? fast_div(1.0f, rf) : 0.f;
        go[0] += su * inf_ + gn * uf * ir;
        go[1] += sv * inf_ + gn * vf * ir;
      }
      float* gf = L.gflow_f + (long)tc.b * 2 * plane;
      float* gb = L.gflow_b + (long)tc.b * 2 * plane;
      gf[pix] = go[0]; gf[plane + pix] = go[1];
      gb[pix] = go[2]; gb[plane + pix] = go[3];
    }
  }
};

}  // namespace ugl
