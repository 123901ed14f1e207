// Split form of the single-pass flow-mode / geom-mode loss (ugl_flow_grad.cuh): the gather-heavy photometry and the
// shared-memory stencils run as TWO kernels over the same tile decomposition.
//
//   photometry kernel  (FlowPhotoPixel)   one thread per pixel, no shared memory tile, no halo: both backward warps of the
//       pixel (24 bilinear taps), validity / occlusion weights, warp Jacobians, the L1 and consistency terms with their
//       gradient basis, the tile's L1 / weight / consistency sums.  Writes per pixel 12 (direction 0, direction 1) pairs of
//       "photometry planes" for the stencil kernel:  x[c] = I[c] * w,  y[c] = W[c] * w (c = 0..2: the SSIM operands),
//       w * keep * dW[c]/d(u|v)  (6 pairs).
//   stencil kernel     (FlowStencilTile)  stages the photometry planes of a tile + 2-pixel halo in shared memory — with TMA
//       (cp.async.bulk.tensor, out-of-image elements zero-filled by the copy engine) where the level's width allows 16-byte
//       global strides, with plain coalesced loads otherwise, identical shared-memory contents either way — and runs the
//       SSIM window / coefficient / box-sum phases and the smoothness phases of ugl_flow_grad.cuh on them, channel by
//       channel with the next channel's planes in flight (two-slot ring, one mbarrier per copy group).
//
// Why (profiles/r1v_*): fused, the photometry phase recomputed 1.47 pixels per pixel (2-pixel halo of a 32x13 tile), ran at
// 16 warps per SM next to 94 KB of stencil planes, and was 49 % of the kernel's time (long-scoreboard stalls on its gathers).
// Split, every pixel's photometry runs once in a kernel with no shared-memory footprint (occupancy bound by registers
// only), and the stencil kernel's footprint drops to 56 KB per CTA.  The price is 24 floats per pixel written and read
// back (mostly through L2); the step was nowhere near the HBM roofline.
//
// Same arithmetic as the fused kernel: x = I * w and y = W * w are the same packed products (formed by the photometry kernel:
// the stencil kernel's shared-memory pipe is its binding resource, a conversion pass over the staged planes cost 15 % of its
// traffic), the SSIM moments / terms / coefficients are the same functions, the L1 and consistency bases are the same
// expressions.  The only re-association is w * (keep dW) being formed before instead of after the box sums (gradient only).
#pragma once

#include "ugl_flow_grad.cuh"

namespace ugl {

// pair planes per sample and level: x[0..2], y[0..2], wdW[2c+uv] (c = 0..2, uv = 0,1); step mode only: the L1 basis (Gp_u, Gp_v as
// (fwd, bwd) pairs) and the consistency basis (u, v)
constexpr int kPhotoPairs = 15;
constexpr int kPhotoFloats = 2 * kPhotoPairs;
enum PhotoPair { PP_X0 = 0, PP_Y0 = 3, PP_DW0 = 6, PP_GPU = 12, PP_GPV = 13, PP_GC = 14 };

// ---- photometry kernel: one pixel -----------------------------------------------------------------------------------
// acc slots of the photometry kernel (subset of FlowAcc / GeomAcc written by this kernel; the stencil kernel writes the rest
// of the same partial row)
template <bool kGeom>
struct FlowPhotoPixel {
  static constexpr int kAcc = kGeom ? 10 : 6;
  // slot k of acc[] -> column of the tile's partial row
  static UGL_HD int column(int k) {
    constexpr int flow_cols[6] = {FA_PIX_F, FA_W_F, FA_PIX_B, FA_W_B, FA_CONS, FA_CONS_W};
    return k < 6 ? flow_cols[k] : (int)GA_PIXD_F + (k - 6);
  }

  static UGL_HD bool is_photo_column(int col) {
    return col == FA_PIX_F || col == FA_W_F || col == FA_PIX_B || col == FA_W_B || col == FA_CONS || col == FA_CONS_W || (kGeom && col >= (int)GA_PIXD_F);
  }

  // pixel (i, j) of sample b, level lv; mats (geom mode): K^-1 (9), P_bwd (12), P_fwd (12)
  // the coalesced loads of pixel (i, j): issued one pass ahead of their use by the kernel's loop
  static UGL_HD DirectLoads load(const FlowGradParams& gp, int lv, int b, int i, int j) {
    const FlowLevelDesc& L = gp.base.lv[lv];
    const int plane = L.h * L.w, pix = i * L.w + j;
    DirectLoads d;
    const float* ff = L.flow_f + (long)b * 2 * plane;
    const float* fb = L.flow_b + (long)b * 2 * plane;
    const float* ic = L.img + (long)b * 3 * plane;
    d.inside = true;
    d.uf = ld_once(ff + pix); d.vf = ld_once(ff + plane + pix);
    d.ub = ld_once(fb + pix); d.vb = ld_once(fb + plane + pix);
    d.I[0] = ic[pix]; d.I[1] = ic[plane + pix]; d.I[2] = ic[2 * plane + pix];   // re-read by the stencil kernel: keep in L2
    return d;
  }

  static UGL_HD void run(const FlowGradParams& gp, int lv, int b, int i, int j, const DirectLoads& d, float* acc, const float* mats) {
    const FlowLevelDesc& L = gp.base.lv[lv];
    const int plane = L.h * L.w, pix = i * L.w + j;
    Photo P;
    float dW[12];
    flow_photo_pixel_c<true, kGeom>(L, b, i, j, d, P, dW);
    const float uf = d.uf, vf = d.vf, ub = d.ub, vb = d.vb;
    const float2 w2 = make_float2(P.w_f, P.w_b);
    float* scr = gp.scratch[lv] + (long)b * kPhotoFloats * plane + 2 * (long)pix;
    const long pp = 2 * (long)plane;                                           // floats per pair plane
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      *reinterpret_cast<float2*>(scr + (PP_X0 + c) * pp) = mul2(splat2(P.I[c]), w2);
      *reinterpret_cast<float2*>(scr + (PP_Y0 + c) * pp) = mul2(make_float2(P.Wf[c], P.Wb[c]), w2);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) *reinterpret_cast<float2*>(scr + (PP_DW0 + k) * pp) = mul2(make_float2(dW[k], dW[6 + k]), w2);
    // L1 basis: w * sum_c sign(W_c - I_c) * keep * dW_c/d(u,v)
    float gpu[2], gpv[2];
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
      float su = 0.f, sv = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float sg = sgnf((dir == 0 ? P.Wf[c] : P.Wb[c]) - P.I[c]);
        su += sg * dW[6 * dir + 2 * c];
        sv += sg * dW[6 * dir + 2 * c + 1];
      }
      const float w = dir == 0 ? P.w_f : P.w_b;
      gpu[dir] = su * w;
      gpv[dir] = sv * w;
    }
    float* basis = gp.step ? nullptr : gp.basis[lv] + (long)b * kBasisPlanes * plane;
    if (gp.step) {
      *reinterpret_cast<float2*>(scr + PP_GPU * pp) = make_float2(gpu[0], gpu[1]);
      *reinterpret_cast<float2*>(scr + PP_GPV * pp) = make_float2(gpv[0], gpv[1]);
    } else {
      basis[0 * (long)plane + pix] = gpu[0]; basis[1 * (long)plane + pix] = gpv[0];
      basis[8 * (long)plane + pix] = gpu[1]; basis[9 * (long)plane + pix] = gpv[1];
    }
    float om;   // mask of the direction-consistency term: 1 - w_f (flow mode) / 1 - occ_f (geom mode)
    if (kGeom) {
      const float D = gp.disp[lv][(long)b * plane + pix];
      const Projected qb = project_pixel(mats, mats + 9, D, j, i), qf = project_pixel(mats, mats + 21, D, j, i);
      const float dyn_b = dynamic_mask_value(ub, vb, sub_rn(qb.u, (float)j), sub_rn(qb.v, (float)i), gp.alpha, gp.beta);
      const float dyn_f = dynamic_mask_value(uf, vf, sub_rn(qf.u, (float)j), sub_rn(qf.v, (float)i), gp.alpha, gp.beta);
      acc[0] += P.d_f * (P.w_f * dyn_f);           acc[1] += P.w_f * dyn_f;
      acc[6] += P.d_f * (P.w_f * (1.f - dyn_f));   acc[7] += P.w_f * (1.f - dyn_f);
      acc[2] += P.d_b * (P.w_b * dyn_b);           acc[3] += P.w_b * dyn_b;
      acc[8] += P.d_b * (P.w_b * (1.f - dyn_b));   acc[9] += P.w_b * (1.f - dyn_b);
      const unsigned bits = (P.valid_b != 0.f ? kMaskValidB : 0u) | (P.valid_f != 0.f ? kMaskValidF : 0u) |
                            (P.occ_b != 0.f ? kMaskOccB : 0u) | (P.occ_f != 0.f ? kMaskOccF : 0u) |
                            (dyn_b != 0.f ? kMaskDynB : 0u) | (dyn_f != 0.f ? kMaskDynF : 0u);
      gp.mask_bytes[lv][(long)b * plane + pix] = (unsigned char)bits;
      om = 1.0f - P.occ_f;
    } else {
      acc[0] += P.d_f * P.w_f;
      acc[1] += P.w_f;
      acc[2] += P.d_b * P.w_b;
      acc[3] += P.w_b;
      om = 1.0f - P.w_f;
    }
    // direction consistency: value and un-normalised gradient w.r.t. the forward flow (model_flow.py:184-199)
    const float rf = sqrt_rn(uf * uf + vf * vf), rb = sqrt_rn(ub * ub + vb * vb);
    const float inf_ = fast_div(1.0f, rf + 1e-12f), inb_ = fast_div(1.0f, rb + 1e-12f);
    const float cu = uf * inf_ + ub * inb_, cv = vf * inf_ + vb * inb_;
    acc[4] += (fabsf(cu) + fabsf(cv)) * om;
    acc[5] += om;
    const float su = sgnf(cu) * om, sv = sgnf(cv) * om;
    const float gn = -(su * uf + sv * vf) * inf_ * inf_;
    const float ir = rf > 0.f ? fast_div(1.0f, rf) : 0.f;
    const float gcu = su * inf_ + gn * uf * ir, gcv = sv * inf_ + gn * vf * ir;
    if (gp.step) {
      *reinterpret_cast<float2*>(scr + PP_GC * pp) = make_float2(gcu, gcv);
    } else {
      basis[6 * (long)plane + pix] = gcu;
      basis[7 * (long)plane + pix] = gcv;
    }
  }
};

// ---- stencil kernel: tile logic ---------------------------------------------------------------------------------------
// Shared-memory plan (floats; every slot starts 128-byte aligned: a TMA destination).  PN / CN / TN = pixels of the tile with
// halo 2 / halo 1 / no halo.
//   kOffI      3 scalar planes I[c] (halo 2)                       [whole kernel; edge weights of the smoothness term]
//   kOffStage  2 slots x { x[c], y[c] pair planes (halo 2) ; 2 pair planes w dW[c]/d(u|v) (no halo) }
//   kOffCoef   3 pair planes A, B, C of the current channel (halo 1)
// Smoothness phases (after the last channel): the four raw flow planes arrive in slot 1, are divided by 20 and interleaved into
// two (u, v) pair planes in slot 0; then the edge weights wx, wy (2 scalar planes, halo 1) are formed from I over the dead raw flow
// planes; the signed second differences (4 pair planes, halo 1) follow them, over the rest of slot 1 and the coefficient planes.
// 56 KB in all: four CTAs per SM.
template <int TW, int TH, int NT, bool kGeom>
struct FlowStencilTile {
  static_assert(TW % 2 == 0, "1x2 micro-tiles need an even tile width");
  static constexpr int R = 2;
  static constexpr int PW = TW + 2 * R, PH = TH + 2 * R, PN = PW * PH;
  static constexpr int CW = TW + 2, CH = TH + 2, CN = CW * CH;
  static constexpr int TN = TW * TH;
  // Halo planes: a TMA box must start on a 16-byte boundary of the global row, and column x0 - 2 of a 4-byte element is not
  // one, so every halo plane (scalar and pair alike: one index for all of them) is staged from column x0 - 4 with a row
  // pitch of SPW = TW + 8 pixels; the halo-2 pixel (ly, lx) sits at hp(ly, lx) = ly * SPW + lx + kScalX
  static constexpr int SPW = TW + 8, SPN = SPW * PH, kScalX = 2;
  static UGL_HD int hp(int ly, int lx) { return ly * SPW + lx + kScalX; }
  static constexpr int pad32(int n) { return (n + 31) & ~31; }
  static constexpr int kPairP = pad32(2 * SPN), kScalP = pad32(SPN), kPairC = pad32(2 * CN), kScalC = pad32(CN), kPairT = pad32(2 * TN);
  static constexpr int kOffI = 0;
  static constexpr int kOffStage = kOffI + 3 * kScalP;
  static constexpr int kStage = 2 * kPairP + 2 * kPairT;          // x, y (halo 2), w dW/du, w dW/dv (tile)
  static constexpr int kSlotY = kPairP, kSlotDu = 2 * kPairP, kSlotDv = 2 * kPairP + kPairT;
  static constexpr int kOffCoef = kOffStage + 2 * kStage;
  static constexpr int kSmemFloats = kOffCoef + 3 * kPairC;
  static constexpr int kOffRawFlow = kOffStage + kStage;          // 4 scalar planes (uf, vf, ub, vb), halo 2, in slot 1
  static constexpr int kOffEdge = kOffRawFlow;                    // wx, wy (halo 1): over the raw flow planes once those are converted
  static constexpr int kOffF2 = kOffStage;                        // 2 pair planes (u, v) / 20 of the fwd / bwd flow, in slot 0
  static constexpr int kOffS4 = kOffEdge + 2 * kScalC;            // 4 pair planes of signed weights (halo 1)
  static_assert(4 * kScalP <= kStage && 2 * kPairP <= kStage && 2 * kScalC <= kStage, "flow planes / edge planes must fit a ring slot");
  static_assert(kOffS4 + 4 * kPairC <= kSmemFloats, "phase-4 planes must fit behind the edge weights (rest of slot 1 + coefficient planes)");
  static_assert(SPW % 2 == 0 && kScalX % 2 == 0 && CW % 2 == 0, "pair planes are read as float4 (two pixels x two directions)");
  static constexpr int kAcc = 6;                                  // SSIM_F, SSIM_B, SMX_F, SMY_F, SMX_B, SMY_B
  static UGL_HD int column(int k) { return (int)FA_SSIM_F + k; }
  static_assert(FA_SSIM_B == FA_SSIM_F + 1 && FA_SMX_F == FA_SSIM_F + 2 && FA_SMY_B == FA_SSIM_F + 5, "stencil sums are contiguous columns");

  static UGL_HD float* stage(float* sm, int c) { return sm + kOffStage + (c & 1) * kStage; }
  static UGL_HD const float* stage(const float* sm, int c) { return sm + kOffStage + (c & 1) * kStage; }

  // ---- plain-load staging (levels whose width does not allow TMA strides; also the host emulator) ----
  static UGL_HD void load_pair_halo(float* dst, const float* src /* (h, w, 2) */, const FlowLevelDesc& L, const TileCoord& tc, int tid, int nt) {
    for (int idx = tid; idx < SPN; idx += nt) {
      const int ly = idx / SPW, lx = idx - ly * SPW;
      const int i = tc.y0 - R + ly, j = tc.x0 - R - kScalX + lx;
      float2 v = make_float2(0.f, 0.f);
      if (i >= 0 && i < L.h && j >= 0 && j < L.w) v = *reinterpret_cast<const float2*>(src + 2 * ((long)i * L.w + j));
      *reinterpret_cast<float2*>(dst + 2 * idx) = v;
    }
  }
  static UGL_HD void load_scalar_halo(float* dst, const float* src /* (h, w) */, const FlowLevelDesc& L, const TileCoord& tc, int tid, int nt) {
    for (int idx = tid; idx < SPN; idx += nt) {
      const int ly = idx / SPW, lx = idx - ly * SPW;
      const int i = tc.y0 - R + ly, j = tc.x0 - R - kScalX + lx;
      dst[idx] = (i >= 0 && i < L.h && j >= 0 && j < L.w) ? src[(long)i * L.w + j] : 0.f;
    }
  }
  static UGL_HD void load_pair_interior(float* dst, const float* src, const FlowLevelDesc& L, const TileCoord& tc, int tid, int nt) {
    for (int idx = tid; idx < TN; idx += nt) {
      const int ty = idx / TW, tx = idx - ty * TW;
      const int i = tc.y0 + ty, j = tc.x0 + tx;
      float2 v = make_float2(0.f, 0.f);
      if (i < L.h && j < L.w) v = *reinterpret_cast<const float2*>(src + 2 * ((long)i * L.w + j));
      *reinterpret_cast<float2*>(dst + 2 * idx) = v;
    }
  }
  // copy groups: 0 = I[0..2] + channel 0 -> slot 0; 1 = channel 1 -> slot 1; 2 = channel 2 -> slot 0; 3 = raw flows -> slot 1
  static UGL_HD void load_group_plain(const FlowGradParams& gp, const TileCoord& tc, int group, int tid, int nt, float* sm) {
    const FlowLevelDesc& L = gp.base.lv[tc.level];
    const long plane = (long)L.h * L.w;
    const float* scr = gp.scratch[tc.level] + (long)tc.b * kPhotoFloats * plane;
    if (group == 0) {
      for (int c = 0; c < 3; ++c) load_scalar_halo(sm + kOffI + c * kScalP, L.img + ((long)tc.b * 3 + c) * plane, L, tc, tid, nt);
    }
    if (group < 3) {
      const int c = group;
      float* st = stage(sm, c);
      load_pair_halo(st, scr + (PP_X0 + c) * 2 * plane, L, tc, tid, nt);
      load_pair_halo(st + kSlotY, scr + (PP_Y0 + c) * 2 * plane, L, tc, tid, nt);
      load_pair_interior(st + kSlotDu, scr + (PP_DW0 + 2 * c) * 2 * plane, L, tc, tid, nt);
      load_pair_interior(st + kSlotDv, scr + (PP_DW0 + 2 * c + 1) * 2 * plane, L, tc, tid, nt);
    } else {
      float* raw = sm + kOffRawFlow;
      load_scalar_halo(raw, L.flow_f + (long)tc.b * 2 * plane, L, tc, tid, nt);
      load_scalar_halo(raw + kScalP, L.flow_f + ((long)tc.b * 2 + 1) * plane, L, tc, tid, nt);
      load_scalar_halo(raw + 2 * kScalP, L.flow_b + (long)tc.b * 2 * plane, L, tc, tid, nt);
      load_scalar_halo(raw + 3 * kScalP, L.flow_b + ((long)tc.b * 2 + 1) * plane, L, tc, tid, nt);
    }
  }

  // raw flow planes (slot 1) -> (u, v) / 20 pair planes of both flows (slot 0): `flow / 20.0` of model_flow.py:177
  static UGL_HD void convert_flows(int tid, int nt, float* sm) {
    constexpr float r20 = 1.0f / 20.0f;
    const float* raw = sm + kOffRawFlow;
    for (int idx = tid; idx < SPN; idx += nt) {
      *reinterpret_cast<float2*>(sm + kOffF2 + 2 * idx) = div_c2(make_float2(raw[idx], raw[kScalP + idx]), 20.0f, r20);
      *reinterpret_cast<float2*>(sm + kOffF2 + kPairP + 2 * idx) = div_c2(make_float2(raw[2 * kScalP + idx], raw[3 * kScalP + idx]), 20.0f, r20);
    }
  }

  // Per-thread strip geometry of phases 2 and 3: the same for the three channels, so it is decoded once per tile (the channel loop is
  // rolled around CTA barriers: the compiler re-derived the divisions, plane offsets and bounds tests in every channel) and kept packed
  // in one register per strip.
  //   phase 2: bits 0..11 float offset / 2 of the strip's window in a halo-2 pair plane, 12..22 float offset / 2 of its coefficients,
  //            23 / 24 pixel 0 / 1 inside the image, 25 / 26 pixel 0 / 1 inside the tile as well (counts towards the SSIM sum)
  //   phase 3: bits 0..10 halo-2 index of the strip's left pixel, 11..20 halo-1 index, 21..30 tile index, 31 strip inside the image
  static constexpr int kNS2 = (CW / 2) * CH, kNS3 = (TW / 2) * TH;
  static constexpr int kP2 = (kNS2 + NT - 1) / NT, kP3 = (kNS3 + NT - 1) / NT;
  static_assert(SPN < 4096 && CN < 2048 && SPN < 2048 && CN < 1024 && TN < 1024, "packed strip geometry");
  static UGL_HD void strip_geometry(const FlowGradParams& gp, const TileCoord& tc, int tid, int nt, unsigned* p2, unsigned* p3) {
    const FlowLevelDesc& L = gp.base.lv[tc.level];
    constexpr int SW = CW / 2;
    // Strip order of phase 2: the first SW - 1 strips of every row, row by row, then the last strip of every row.  With SW - 1 a
    // multiple of 8 (TW = 32: 16) a quarter-warp's eight 16-byte reads stay inside one row = 128 contiguous bytes; the row-major
    // order (17 strips per row of pitch 320 bytes) split 45 % of the quarter-warps over two rows whose banks overlap (5.6
    // wavefronts per LDS.128 instead of 4).
    constexpr bool kSplitRows = ((SW - 1) % 8) == 0;
    constexpr int SM = kSplitRows ? SW - 1 : SW;
    int n = 0;
    for (int s = tid; s < kNS2; s += nt, ++n) {
      int ly, lx;
      if (s < SM * CH) { ly = s / SM; lx = (s - ly * SM) * 2; }
      else { ly = s - SM * CH; lx = 2 * SM; }
      const int i = tc.y0 - 1 + ly, j0 = tc.x0 - 1 + lx;
      const int c0 = hp(ly + 1, lx + 1);                    // staged-plane index of the left centre
      const bool row_in = (i >= 0 && i < L.h);
      const bool in0 = row_in && j0 >= 0 && j0 < L.w, in1 = row_in && j0 + 1 >= 0 && j0 + 1 < L.w;
      const bool t0 = in0 && ly >= 1 && ly <= TH && lx >= 1 && lx <= TW, t1 = in1 && ly >= 1 && ly <= TH && lx + 1 >= 1 && lx + 1 <= TW;
      p2[n] = (unsigned)(c0 - 1) | ((unsigned)(ly * CW + lx) << 12) | (in0 ? 1u << 23 : 0u) | (in1 ? 1u << 24 : 0u) | (t0 ? 1u << 25 : 0u) | (t1 ? 1u << 26 : 0u);
    }
    n = 0;
    constexpr int SW3 = TW / 2;
    for (int s = tid; s < kNS3; s += nt, ++n) {
      const int ty = s / SW3, tx = (s - ty * SW3) * 2;
      const bool in = tc.y0 + ty < L.h && tc.x0 + tx < L.w;
      p3[n] = (unsigned)hp(ty + R, tx + R) | ((unsigned)((ty + 1) * CW + (tx + 1)) << 11) | ((unsigned)(ty * TW + tx) << 21) | (in ? 1u << 31 : 0u);
    }
  }

  // SSIM of channel c (c < 3) for every 1x2 strip of the halo-1 region -> coefficient pairs + loss sums; c == 3: the smoothness
  // edge weights.  FlowGradTile::phase2 on the split kernel's planes.
  static UGL_HD void phase2(const FlowGradParams& gp, const TileCoord& tc, int c, int tid, int nt, float* sm, float* acc, const unsigned* p2) {
    if (c >= 3) {
      edge_weights(gp, tc, tid, nt, sm);
      return;
    }
    float2 ssim_sum = make_float2(0.f, 0.f);
    const float2 one = splat2(gp.one);
    int n = 0;
    for (int s = tid; s < kNS2; s += nt, ++n) {
      const unsigned geo = p2[n];
      const bool in0 = (geo >> 23) & 1u, in1 = (geo >> 24) & 1u;
      {
        const float* xpl = stage(sm, c) + 2 * (int)(geo & 4095u);
        const float* ypl = xpl + kSlotY;
        Moments2 m[2];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int o = 2 * (r - 1) * SPW;
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
              if (r == 0 && k == 0) {
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
          const bool in = (w == 0 ? in0 : in1);
          const SsimTerms2 t = ssim_terms2(m[w], one);
          const float2 v = ssim_half_one_minus2(t.S, one);
          const float2 g = make_float2((in && v.x >= 0.f && v.x <= 1.f) ? -0.5f : 0.f, (in && v.y >= 0.f && v.y <= 1.f) ? -0.5f : 0.f);
          ssim_partials2(t, g, cA[w], cB[w], cC[w]);
          const bool interior = (geo >> (25 + w)) & 1u;
          ssim_sum.x += interior ? (v.x < 0.f ? 0.f : (v.x > 1.f ? 1.f : v.x)) : 0.f;
          ssim_sum.y += interior ? (v.y < 0.f ? 0.f : (v.y > 1.f ? 1.f : v.y)) : 0.f;
        }
        float* oc = sm + kOffCoef + 2 * (int)((geo >> 12) & 2047u);
        *reinterpret_cast<float4*>(oc) = make_float4(cA[0].x, cA[0].y, cA[1].x, cA[1].y);
        *reinterpret_cast<float4*>(oc + kPairC) = make_float4(cB[0].x, cB[0].y, cB[1].x, cB[1].y);
        *reinterpret_cast<float4*>(oc + 2 * kPairC) = make_float4(cC[0].x, cC[0].y, cC[1].x, cC[1].y);
      }
    }
    acc[0] += ssim_sum.x;
    acc[1] += ssim_sum.y;
  }

  // smoothness edge weights wx, wy of every pixel of the halo-1 region, one pixel per thread (consecutive lanes read consecutive
  // words of the three image planes: the 1x2 strips of the SSIM passes read them with stride 2 = twice the wavefronts)
  static UGL_HD void edge_weights(const FlowGradParams& gp, const TileCoord& tc, int tid, int nt, float* sm) {
    const FlowLevelDesc& L = gp.base.lv[tc.level];
    const float* I0 = sm + kOffI, *I1 = I0 + kScalP, *I2 = I1 + kScalP;
    for (int s = tid; s < CN; s += nt) {
      const int ly = s / CW, lx = s - ly * CW;
      const int i = tc.y0 - 1 + ly, j = tc.x0 - 1 + lx;
      float wx = 0.f, wy = 0.f;
      if (i >= 0 && i < L.h && j >= 0 && j < L.w) {
        const int cc = hp(ly + 1, lx + 1);
        const float Ic[3] = {I0[cc], I1[cc], I2[cc]};
        if (j >= 1 && j <= L.w - 2) {
          const float Iq[3] = {I0[cc + 1], I1[cc + 1], I2[cc + 1]};
          wx = edge_weight10(Ic, Iq);
        }
        if (i >= 1 && i <= L.h - 2) {
          const float Iq[3] = {I0[cc + SPW], I1[cc + SPW], I2[cc + SPW]};
          wy = edge_weight10(Ic, Iq);
        }
      }
      sm[kOffEdge + s] = wx;
      sm[kOffEdge + kScalC + s] = wy;
    }
  }

  // 3x3 box sums of channel c's coefficient pairs, chained through w * keep * dW[c]/d(u, v) into the running sums
  static UGL_HD void phase3_accumulate(const FlowGradParams& gp, const TileCoord& tc, int c, int tid, int nt, const float* sm, float2 (*g)[4],
                                       const unsigned* p3) {
    const float* st = stage(sm, c);
    int n = 0;
    for (int s = tid; s < kNS3; s += nt, ++n) {
      const unsigned geo = p3[n];
      if (!(geo >> 31)) continue;
      const int c0 = (int)(geo & 2047u);
      const int q0 = (int)((geo >> 11) & 1023u);
      const int t0 = (int)((geo >> 21) & 1023u);
      float2 sum[3][2];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float* cf = sm + kOffCoef + k * kPairC + 2 * (q0 - 1);
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
      const float4 Xv = *reinterpret_cast<const float4*>(st + 2 * c0);
      const float4 Yv = *reinterpret_cast<const float4*>(st + kSlotY + 2 * c0);
      const float4 du = *reinterpret_cast<const float4*>(st + kSlotDu + 2 * t0);
      const float4 dv = *reinterpret_cast<const float4*>(st + kSlotDv + 2 * t0);
      const float2 X2[2] = {lo2(Xv), hi2(Xv)}, Y2[2] = {lo2(Yv), hi2(Yv)};
      const float2 du2[2] = {lo2(du), hi2(du)}, dv2[2] = {lo2(dv), hi2(dv)};
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        const float2 gW = fma2(X2[o], sum[2][o], fma2(add2(Y2[o], Y2[o]), sum[1][o], sum[0][o]));
        g[n][o] = fma2(gW, du2[o], g[n][o]);
        g[n][2 + o] = fma2(gW, dv2[o], g[n][2 + o]);
      }
    }
  }

  // the accumulated SSIM basis of both directions -> global memory (planes 2, 3 and 10, 11)
  static UGL_HD void phase3_store(const FlowGradParams& gp, const TileCoord& tc, int tid, int nt, const float2 (*g)[4]) {
    const FlowLevelDesc& L = gp.base.lv[tc.level];
    const long plane = (long)L.h * L.w;
    float* basis = gp.basis[tc.level] + (long)tc.b * kBasisPlanes * plane;
    constexpr int SW = TW / 2;
    const bool vec = (L.w & 1) == 0;
    int n = 0;
    for (int s = tid; s < SW * TH; s += nt, ++n) {
      const int ty = s / SW, tx = (s - ty * SW) * 2;
      const int i = tc.y0 + ty, j = tc.x0 + tx;
      if (i >= L.h || j >= L.w) continue;
      const float2 *gsu = g[n], *gsv = g[n] + 2;
      const long pix = (long)i * L.w + j;
      float* b0 = basis + 2 * plane + pix;
      float* b1 = basis + 10 * plane + pix;
      if (vec) {
        *reinterpret_cast<float2*>(b0) = make_float2(gsu[0].x, gsu[1].x);
        *reinterpret_cast<float2*>(b0 + plane) = make_float2(gsv[0].x, gsv[1].x);
        *reinterpret_cast<float2*>(b1) = make_float2(gsu[0].y, gsu[1].y);
        *reinterpret_cast<float2*>(b1 + plane) = make_float2(gsv[0].y, gsv[1].y);
      } else {
        b0[0] = gsu[0].x; b0[plane] = gsv[0].x; b1[0] = gsu[0].y; b1[plane] = gsv[0].y;
        if (j + 1 < L.w) { b0[1] = gsu[1].x; b0[plane + 1] = gsv[1].x; b1[1] = gsu[1].y; b1[plane + 1] = gsv[1].y; }
      }
    }
  }

  // signed, edge-weighted second differences once per halo-1 centre (FlowGradTile::phase4a)
  static UGL_HD void phase4a(const FlowGradParams& gp, const TileCoord& tc, int tid, int nt, float* sm, float* acc) {
    const FlowLevelDesc& L = gp.base.lv[tc.level];
    for (int idx = tid; idx < CN; idx += nt) {
      const int ly = idx / CW, lx = idx - ly * CW;
      const int c0 = hp(ly + 1, lx + 1);
      const float wx = sm[kOffEdge + idx], wy = sm[kOffEdge + kScalC + idx];
      const bool interior = (ly >= 1 && ly <= TH && lx >= 1 && lx <= TW) && (tc.y0 + ly - 1 < L.h) && (tc.x0 + lx - 1 < L.w);
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const float* f = sm + kOffF2 + d * kPairP + 2 * c0;
        const float2 c = *reinterpret_cast<const float2*>(f);
        const float2 xm = *reinterpret_cast<const float2*>(f - 2), xp = *reinterpret_cast<const float2*>(f + 2);
        const float2 ym = *reinterpret_cast<const float2*>(f - 2 * SPW), yp = *reinterpret_cast<const float2*>(f + 2 * SPW);
        const float2 dxx = sub2(sub2(xp, c), sub2(c, xm)), dyy = sub2(sub2(yp, c), sub2(c, ym));
        *reinterpret_cast<float2*>(sm + kOffS4 + (2 * d) * kPairC + 2 * idx) = make_float2(wx * sgnf(dxx.x), wx * sgnf(dxx.y));
        *reinterpret_cast<float2*>(sm + kOffS4 + (2 * d + 1) * kPairC + 2 * idx) = make_float2(wy * sgnf(dyy.x), wy * sgnf(dyy.y));
        if (interior) {
          acc[2 + 2 * d] += wx * fabsf(dxx.x);
          acc[3 + 2 * d] += wy * fabsf(dyy.x);
          acc[2 + 2 * d] += wx * fabsf(dxx.y);
          acc[3 + 2 * d] += wy * fabsf(dyy.y);
        }
      }
    }
  }

  // ---- fused forward + backward (step mode): the per-sample normalisers are known before this kernel starts (the photometry
  // kernel's weight sums), so the strips' owners combine the four gradient terms themselves and write d loss / d flow: no basis
  // planes, no combine launch.  pre[n][0..2]: the strip's L1 basis (u), L1 basis (v) as (fwd, bwd) pairs of its two pixels and
  // the consistency basis (u, v) of its two pixels, loaded from the photometry planes before phase 4a.
  static UGL_HD void prefetch_step(const FlowGradParams& gp, const TileCoord& tc, int tid, int nt, float4 (*pre)[3]) {
    const FlowLevelDesc& L = gp.base.lv[tc.level];
    const long plane = (long)L.h * L.w;
    const float* scr = gp.scratch[tc.level] + (long)tc.b * kPhotoFloats * plane;
    constexpr int SW = TW / 2;
    int n = 0;
    for (int s = tid; s < SW * TH; s += nt, ++n) {
      const int ty = s / SW, tx = (s - ty * SW) * 2;
      const int i = tc.y0 + ty, j = tc.x0 + tx;
      if (i >= L.h || j >= L.w) continue;
      const long pix = (long)i * L.w + j;
      const bool two = j + 1 < L.w;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float* q = scr + (PP_GPU + k) * 2 * plane + 2 * pix;
        if (two && (L.w & 1) == 0) {
          pre[n][k] = ld_once4(q);                                   // pix even, plane base 16-byte aligned
        } else {
          const float2 a = ld_once2(q);
          const float2 c = two ? ld_once2(q + 2) : make_float2(0.f, 0.f);
          pre[n][k] = make_float4(a.x, a.y, c.x, c.y);
        }
      }
    }
  }

  // ks[8]: scale factors of this (sample, level): L1 fwd / bwd (geom mode: rigid pixels), L1 fwd / bwd of the dynamic pixels (geom mode; flow
  // mode: the same two again), SSIM fwd / bwd, smoothness, consistency -- FlowCombineScales / GeomCombineScales as written by the weight-sum kernel
  static UGL_HD void phase4b_step(const FlowGradParams& gp, const TileCoord& tc, const float* ks, int tid, int nt, const float* sm,
                                  const float2 (*g)[4], const float4 (*pre)[3]) {
    const FlowLevelDesc& L = gp.base.lv[tc.level];
    const long plane = (long)L.h * L.w;
    float* gf = L.gflow_f + (long)tc.b * 2 * plane;
    float* gb = L.gflow_b + (long)tc.b * 2 * plane;
    const float2 inx = splat2(fast_rcp(2.0f * (float)L.h * (float)(L.w - 2))), iny = splat2(fast_rcp(2.0f * (float)(L.h - 2) * (float)L.w));
    constexpr int SW = TW / 2;
    const bool vec = (L.w & 1) == 0;
    int n = 0;
    for (int s = tid; s < SW * TH; s += nt, ++n) {
      const int ty = s / SW, tx = (s - ty * SW) * 2;
      const int i = tc.y0 + ty, j = tc.x0 + tx;
      if (i >= L.h || j >= L.w) continue;
      float o_fu[2], o_fv[2], o_bu[2], o_bv[2];
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        const int q0 = (ty + 1) * CW + (tx + o + 1);
        float2 gm[2];
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const float* sx = sm + kOffS4 + (2 * d) * kPairC + 2 * q0;
          const float* sy = sm + kOffS4 + (2 * d + 1) * kPairC + 2 * q0;
          const float2 xc = *reinterpret_cast<const float2*>(sx), yc = *reinterpret_cast<const float2*>(sy);
          const float2 gx = add2(fma2(xc, splat2(-2.0f), *reinterpret_cast<const float2*>(sx - 2)), *reinterpret_cast<const float2*>(sx + 2));
          const float2 gy = add2(fma2(yc, splat2(-2.0f), *reinterpret_cast<const float2*>(sy - 2 * CW)), *reinterpret_cast<const float2*>(sy + 2 * CW));
          gm[d] = fma2(gy, iny, mul2(gx, inx));
        }
        const float2 gpu = o == 0 ? lo2(pre[n][0]) : hi2(pre[n][0]), gpv = o == 0 ? lo2(pre[n][1]) : hi2(pre[n][1]);
        const float2 gc = o == 0 ? lo2(pre[n][2]) : hi2(pre[n][2]);
        const float2 gsu = g[n][o], gsv = g[n][2 + o];
        float kpf = ks[0], kpb = ks[1];
        if (kGeom && (o == 0 || j + 1 < L.w)) {              // L1 weight 1 on rigid pixels, 2 on dynamic ones (model_geometry.py:857-876)
          const unsigned bits = gp.mask_bytes[tc.level][(long)tc.b * plane + (long)i * L.w + j + o];
          kpf = (bits & kMaskDynF) ? ks[0] : ks[2];
          kpb = (bits & kMaskDynB) ? ks[1] : ks[3];
        }
        o_fu[o] = kpf * gpu.x + ks[4] * gsu.x + ks[6] * gm[0].x + ks[7] * gc.x;
        o_fv[o] = kpf * gpv.x + ks[4] * gsv.x + ks[6] * gm[0].y + ks[7] * gc.y;
        o_bu[o] = kpb * gpu.y + ks[5] * gsu.y + ks[6] * gm[1].x;
        o_bv[o] = kpb * gpv.y + ks[5] * gsv.y + ks[6] * gm[1].y;
      }
      const long pix = (long)i * L.w + j;
      if (vec) {
        *reinterpret_cast<float2*>(gf + pix) = make_float2(o_fu[0], o_fu[1]);
        *reinterpret_cast<float2*>(gf + plane + pix) = make_float2(o_fv[0], o_fv[1]);
        *reinterpret_cast<float2*>(gb + pix) = make_float2(o_bu[0], o_bu[1]);
        *reinterpret_cast<float2*>(gb + plane + pix) = make_float2(o_bv[0], o_bv[1]);
      } else {
        gf[pix] = o_fu[0]; gf[plane + pix] = o_fv[0]; gb[pix] = o_bu[0]; gb[plane + pix] = o_bv[0];
        if (j + 1 < L.w) { gf[pix + 1] = o_fu[1]; gf[plane + pix + 1] = o_fv[1]; gb[pix + 1] = o_bu[1]; gb[plane + pix + 1] = o_bv[1]; }
      }
    }
  }

  static UGL_HD void phase4b(const FlowGradParams& gp, const TileCoord& tc, int tid, int nt, const float* sm) {
    const FlowLevelDesc& L = gp.base.lv[tc.level];
    const long plane = (long)L.h * L.w;
    float* basis = gp.basis[tc.level] + (long)tc.b * kBasisPlanes * plane;
    const float2 inx = splat2(fast_rcp(2.0f * (float)L.h * (float)(L.w - 2))), iny = splat2(fast_rcp(2.0f * (float)(L.h - 2) * (float)L.w));
    for (int idx = tid; idx < TN; idx += nt) {
      const int ty = idx / TW, tx = idx - ty * TW;
      const int i = tc.y0 + ty, j = tc.x0 + tx;
      if (i >= L.h || j >= L.w) continue;
      const int q0 = (ty + 1) * CW + (tx + 1);
      float2 g[2];
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const float* sx = sm + kOffS4 + (2 * d) * kPairC + 2 * q0;
        const float* sy = sm + kOffS4 + (2 * d + 1) * kPairC + 2 * q0;
        const float2 xc = *reinterpret_cast<const float2*>(sx), yc = *reinterpret_cast<const float2*>(sy);
        const float2 gx = add2(fma2(xc, splat2(-2.0f), *reinterpret_cast<const float2*>(sx - 2)), *reinterpret_cast<const float2*>(sx + 2));
        const float2 gy = add2(fma2(yc, splat2(-2.0f), *reinterpret_cast<const float2*>(sy - 2 * CW)), *reinterpret_cast<const float2*>(sy + 2 * CW));
        g[d] = fma2(gy, iny, mul2(gx, inx));
      }
      const long pix = (long)i * L.w + j;
      basis[4 * plane + pix] = g[0].x; basis[5 * plane + pix] = g[0].y;
      basis[12 * plane + pix] = g[1].x; basis[13 * plane + pix] = g[1].y;
    }
  }
};

}  // namespace ugl
