// TEST INFRASTRUCTURE — host emulator of the CUDA kernels' per-tile logic.
//
// The kernels' phase functions are `__host__ __device__`; this file drives them tile by tile on
// the CPU (threads emulated sequentially, shared memory = a heap buffer) so that indexing and
// arithmetic can be checked against the oracle on the GPU-less build box.  It is compiled by
// tests/test_hostemu.py with g++ (-ffp-contract=off), is never part of libugl_b200.so and is never
// imported by the product package: the product has no CPU path.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/ugl.h"
#include "../../unsupervised_depth_opticalflow_egomotion_b200/csrc/ugl_flow_loss.cuh"
#include "../../unsupervised_depth_opticalflow_egomotion_b200/csrc/ugl_flow_grad.cuh"
#include "../../unsupervised_depth_opticalflow_egomotion_b200/csrc/ugl_flow_split.cuh"
#include "../../unsupervised_depth_opticalflow_egomotion_b200/csrc/ugl_primitives.cuh"

using namespace ugl;

template <int TW, int TH>
static void fill_params(const UglFlowLossArgs* a, bool backward, FlowLossParams& p) {
  p.B = a->batch;
  p.scales = a->scales;
  int tiles = 0;
  for (int l = 0; l < a->scales; ++l) {
    FlowLevelDesc& L = p.lv[l];
    L.h = a->height[l]; L.w = a->width[l];
    L.geom = make_warp_geom(L.w, L.h);
    L.img_l = a->img_l[l]; L.img = a->img[l]; L.img_r = a->img_r[l];
    L.flow_f = a->flow_fwd[l]; L.flow_b = a->flow_bwd[l];
    L.gflow_f = backward ? a->grad_flow_fwd[l] : nullptr;
    L.gflow_b = backward ? a->grad_flow_bwd[l] : nullptr;
    L.tiles_x = (L.w + TW - 1) / TW;
    L.tiles_y = (L.h + TH - 1) / TH;
    L.tile_begin = tiles;
    tiles += L.tiles_x * L.tiles_y * a->batch;
  }
  p.total_tiles = tiles;
  p.stats = a->stats;
  p.loss = a->loss;
  p.gloss = a->grad_loss;
  p.partials = nullptr;
}

extern "C" int emu_flow_loss_forward(const UglFlowLossArgs* a) {
  FlowLossParams p;
  fill_params<kFTW, kFTH>(a, false, p);
  using Tile = FlowFwdTile<kFTW, kFTH>;
  std::vector<float> partials((size_t)p.total_tiles * FA_COUNT, 0.f), sm(Tile::kSmemFloats);
  for (int tile = 0; tile < p.total_tiles; ++tile) {
    const TileCoord tc = decode_tile<kFTW, kFTH>(p, tile);
    float acc[FA_COUNT] = {0};
    Tile::phase1(p, tc, 0, 1, sm.data(), acc);
    Tile::phase2(p, tc, 0, 1, sm.data(), acc);
    for (int k = 0; k < FA_COUNT; ++k) partials[(size_t)tile * FA_COUNT + k] = acc[k];
  }
  for (int b = 0; b < p.B; ++b) {
    float tot[4] = {0, 0, 0, 0};
    for (int l = 0; l < p.scales; ++l) {
      const FlowLevelDesc& L = p.lv[l];
      const int per_img = L.tiles_x * L.tiles_y;
      double s[FA_COUNT] = {0};
      for (int t = 0; t < per_img; ++t)
        for (int k = 0; k < FA_COUNT; ++k) s[k] += partials[((size_t)L.tile_begin + (size_t)b * per_img + t) * FA_COUNT + k];
      float S[FA_COUNT], out[4];
      for (int k = 0; k < FA_COUNT; ++k) { S[k] = (float)s[k]; p.stats[((size_t)b * p.scales + l) * FA_COUNT + k] = S[k]; }
      flow_level_losses(S, L.h, L.w, out);
      for (int k = 0; k < 4; ++k) tot[k] += out[k];
    }
    for (int k = 0; k < 4; ++k) p.loss[k * p.B + b] = tot[k];
  }
  return 0;
}

extern "C" int emu_flow_loss_backward(const UglFlowLossArgs* a) {
  FlowLossParams p;
  fill_params<kBTW, kBTH>(a, true, p);
  using Tile = FlowBwdTile<kBTW, kBTH, 1>;   // one emulated thread owns every interior pixel
  std::vector<float> sm(Tile::kSmemFloats);
  std::vector<float> gbuf((size_t)Tile::PPT * 4);
  float (*g)[4] = reinterpret_cast<float (*)[4]>(gbuf.data());
  for (int tile = 0; tile < p.total_tiles; ++tile) {
    const TileCoord tc = decode_tile<kBTW, kBTH>(p, tile);
    const FlowLevelDesc& L = p.lv[tc.level];
    const FlowBwdCoef k = flow_bwd_coef(p.stats + ((size_t)tc.b * p.scales + tc.level) * FA_COUNT, L.h, L.w, p.gloss, p.B, tc.b);
    Tile::phase1(p, tc, 0, 1, sm.data());
    for (int dir = 0; dir < 2; ++dir) {
      Tile::phase2(p, tc, dir, 0, 1, sm.data());
      Tile::phase3(p, tc, k, dir, 0, 1, sm.data(), g);
    }
    Tile::phase4(p, tc, k, 0, 1, sm.data(), g);
  }
  return 0;
}

static void emu_finalize(const FlowLossParams& p, const std::vector<float>& partials) {
  for (int b = 0; b < p.B; ++b) {
    float tot[4] = {0, 0, 0, 0};
    for (int l = 0; l < p.scales; ++l) {
      const FlowLevelDesc& L = p.lv[l];
      const int per_img = L.tiles_x * L.tiles_y;
      double s[FA_COUNT] = {0};
      for (int t = 0; t < per_img; ++t)
        for (int k = 0; k < FA_COUNT; ++k) s[k] += partials[((size_t)L.tile_begin + (size_t)b * per_img + t) * FA_COUNT + k];
      float S[FA_COUNT], out[4];
      for (int k = 0; k < FA_COUNT; ++k) { S[k] = (float)s[k]; p.stats[((size_t)b * p.scales + l) * FA_COUNT + k] = S[k]; }
      flow_level_losses(S, L.h, L.w, out);
      for (int k = 0; k < 4; ++k) tot[k] += out[k];
    }
    for (int k = 0; k < 4; ++k) p.loss[k * p.B + b] = tot[k];
  }
}

// single-pass mode: losses + gradient basis, then the element-wise combine
extern "C" int emu_flow_loss_forward_grad(const UglFlowLossArgs* a) {
  FlowGradParams gp;
  fill_params<kBTW, kBTH>(a, false, gp.base);
  for (int l = 0; l < a->scales; ++l) gp.basis[l] = a->basis[l];
  using Tile = FlowGradTile<kBTW, kBTH, 1>;
  std::vector<float> partials((size_t)gp.base.total_tiles * FA_COUNT, 0.f), sm(Tile::kSmemFloats);
  for (int tile = 0; tile < gp.base.total_tiles; ++tile) {
    const TileCoord tc = decode_tile<kBTW, kBTH>(gp.base, tile);
    float acc[FA_COUNT] = {0};
    Tile::phase1(gp, tc, 0, 1, sm.data(), acc);
    std::vector<float2> g3v((size_t)Tile::kP3 * 4, make_float2(0.f, 0.f));
    float2 (*g3)[4] = reinterpret_cast<float2 (*)[4]>(g3v.data());
    for (int c = 0; c < 3; ++c) {
      Tile::phase2(gp, tc, c, 0, 1, sm.data(), acc);
      if (c == 0 && !Tile::kDepth) Tile::phase2(gp, tc, 3, 0, 1, sm.data(), acc);
      Tile::phase3_accumulate(gp, tc, c, 0, 1, sm.data(), g3);
    }
    Tile::phase3_store(gp, tc, 0, 1, sm.data(), g3);
    Tile::phase4a(gp, tc, 0, 1, sm.data(), acc);
    Tile::phase4b(gp, tc, 0, 1, sm.data());
    for (int k = 0; k < FA_COUNT; ++k) partials[(size_t)tile * FA_COUNT + k] = acc[k];
  }
  emu_finalize(gp.base, partials);
  return 0;
}

// split form (ugl_flow_split.cuh): photometry pixel by pixel -> photometry planes -> stencil tiles staged by the plain loader (the TMA copy
// engine fills the same shared-memory layout on the device).  step = 0: basis planes out (then emu_flow_loss_combine); step = 1: the fused
// training step, gradients written by the stencil tiles.  Photometry sums go through their own partial rows (here: one row per pixel row).
template <bool kGeom>
static int emu_flow_split(FlowGradParams& gp, int step) {
  constexpr int TW = kBTW, TH = kBTH;
  using Px = FlowPhotoPixel<kGeom>;
  using Tile = FlowStencilTile<TW, TH, 1, kGeom>;
  constexpr int ROW = kGeom ? (int)GA_COUNT : (int)FA_COUNT;
  FlowLossParams& p = gp.base;
  gp.step = step;
  std::vector<std::vector<float>> scratch(p.scales);
  for (int l = 0; l < p.scales; ++l) {
    scratch[l].assign((size_t)p.B * kPhotoFloats * p.lv[l].h * p.lv[l].w, 0.f);
    gp.scratch[l] = scratch[l].data();
  }
  // photometry: per (level, sample) sums in fp64 over per-row partials (any fixed order is a valid kernel schedule)
  std::vector<double> psum((size_t)p.B * p.scales * ROW, 0.0);
  for (int l = 0; l < p.scales; ++l)
    for (int b = 0; b < p.B; ++b) {
      float mats[33] = {0};
      if (kGeom) {
        for (int k = 0; k < 9; ++k) mats[k] = gp.Kinv[l][b * 9 + k];
        for (int k = 0; k < 12; ++k) { mats[9 + k] = gp.P[0][l][b * 12 + k]; mats[21 + k] = gp.P[1][l][b * 12 + k]; }
      }
      for (int i = 0; i < p.lv[l].h; ++i) {
        float acc[Px::kAcc] = {0};
        for (int j = 0; j < p.lv[l].w; ++j) {
          const DirectLoads d = Px::load(gp, l, b, i, j);
          Px::run(gp, l, b, i, j, d, acc, mats);
        }
        for (int k = 0; k < Px::kAcc; ++k) psum[((size_t)b * p.scales + l) * ROW + Px::column(k)] += acc[k];
      }
    }
  std::vector<float> scales_buf((size_t)p.B * p.scales * 8, 0.f);
  if (step) {
    for (int b = 0; b < p.B; ++b)
      for (int l = 0; l < p.scales; ++l) {
        float S[ROW];
        for (int k = 0; k < ROW; ++k) S[k] = (float)psum[((size_t)b * p.scales + l) * ROW + k];
        float* o = scales_buf.data() + ((size_t)b * p.scales + l) * 8;
        if (kGeom) {     // flow_photo_norm_kernel<true>: the eight factors of geom_combine_scales
          const GeomCombineScales k = geom_combine_scales(S, p.lv[l].h, p.lv[l].w, p.gloss, p.B, b);
          o[0] = k.pix_r[0]; o[1] = k.pix_r[1]; o[2] = k.pix_d[0]; o[3] = k.pix_d[1]; o[4] = k.ssim[0]; o[5] = k.ssim[1]; o[6] = k.sm; o[7] = k.cons;
        } else {
          const FlowCombineScales k = flow_combine_scales(S, p.lv[l].h, p.lv[l].w, p.gloss, p.B, b);
          o[0] = k.pix[0]; o[1] = k.pix[1]; o[2] = k.pix[0]; o[3] = k.pix[1]; o[4] = k.ssim[0]; o[5] = k.ssim[1]; o[6] = k.sm; o[7] = k.cons;
        }
      }
  }
  // stencil tiles
  std::vector<float> sm(Tile::kSmemFloats);
  std::vector<double> ssum((size_t)p.B * p.scales * ROW, 0.0);
  for (int tile = 0; tile < p.total_tiles; ++tile) {
    const TileCoord tc = decode_tile<TW, TH>(p, tile);
    float acc[Tile::kAcc] = {0};
    std::vector<float2> g3v((size_t)Tile::kP3 * 4, make_float2(0.f, 0.f));
    float2 (*g3)[4] = reinterpret_cast<float2 (*)[4]>(g3v.data());
    Tile::load_group_plain(gp, tc, 0, 0, 1, sm.data());
    Tile::load_group_plain(gp, tc, 1, 0, 1, sm.data());
    std::vector<unsigned> geo2(Tile::kP2), geo3(Tile::kP3);
    Tile::strip_geometry(gp, tc, 0, 1, geo2.data(), geo3.data());
    for (int c = 0; c < 3; ++c) {
      Tile::phase2(gp, tc, c, 0, 1, sm.data(), acc, geo2.data());
      Tile::phase3_accumulate(gp, tc, c, 0, 1, sm.data(), g3, geo3.data());
      if (c < 2) Tile::load_group_plain(gp, tc, c + 2, 0, 1, sm.data());
    }
    std::vector<float4> prev((size_t)Tile::kP3 * 3);
    float4 (*pre)[3] = reinterpret_cast<float4 (*)[3]>(prev.data());
    if (step) Tile::prefetch_step(gp, tc, 0, 1, pre);
    else Tile::phase3_store(gp, tc, 0, 1, g3);
    Tile::convert_flows(0, 1, sm.data());
    Tile::phase2(gp, tc, 3, 0, 1, sm.data(), acc, geo2.data());
    Tile::phase4a(gp, tc, 0, 1, sm.data(), acc);
    if (step) {
      const float* o = scales_buf.data() + ((size_t)tc.b * p.scales + tc.level) * 8;
      Tile::phase4b_step(gp, tc, o, 0, 1, sm.data(), g3, pre);
    } else {
      Tile::phase4b(gp, tc, 0, 1, sm.data());
    }
    for (int k = 0; k < Tile::kAcc; ++k) ssum[((size_t)tc.b * p.scales + tc.level) * ROW + Tile::column(k)] += acc[k];
  }
  for (int b = 0; b < p.B; ++b) {
    float tot[4] = {0, 0, 0, 0};
    for (int l = 0; l < p.scales; ++l) {
      float S[ROW], out[4];
      for (int k = 0; k < ROW; ++k) {
        S[k] = (float)(Px::is_photo_column(k) ? psum[((size_t)b * p.scales + l) * ROW + k] : ssum[((size_t)b * p.scales + l) * ROW + k]);
        p.stats[((size_t)b * p.scales + l) * ROW + k] = S[k];
      }
      if (kGeom) geom_level_losses(S, p.lv[l].h, p.lv[l].w, out); else flow_level_losses(S, p.lv[l].h, p.lv[l].w, out);
      for (int k = 0; k < 4; ++k) tot[k] += out[k];
    }
    for (int k = 0; k < 4; ++k) p.loss[k * p.B + b] = tot[k];
  }
  return 0;
}

extern "C" int emu_flow_loss_split_forward_grad(const UglFlowLossArgs* a) {
  FlowGradParams gp;
  fill_params<kBTW, kBTH>(a, false, gp.base);
  for (int l = 0; l < a->scales; ++l) gp.basis[l] = a->basis[l];
  return emu_flow_split<false>(gp, 0);
}

extern "C" int emu_flow_loss_step(const UglFlowLossArgs* a) {
  FlowGradParams gp;
  fill_params<kBTW, kBTH>(a, true, gp.base);
  return emu_flow_split<false>(gp, 1);
}

extern "C" int emu_geom_flow_split_forward_grad(const UglGeomFlowArgs* g) {
  const UglFlowLossArgs* a = &g->flow;
  FlowGradParams gp;
  fill_params<kBTW, kBTH>(a, false, gp.base);
  for (int l = 0; l < a->scales; ++l) {
    gp.basis[l] = a->basis[l]; gp.disp[l] = g->disp[l]; gp.Kinv[l] = g->Kinv[l];
    gp.P[0][l] = g->P_bwd[l]; gp.P[1][l] = g->P_fwd[l]; gp.mask_bytes[l] = g->mask_bytes[l];
  }
  gp.alpha = g->alpha; gp.beta = g->beta;
  return emu_flow_split<true>(gp, 0);
}

// the fused geom training step (ugl_geom_flow_step): gradients written by the stencil tiles, L1 scale per pixel from the mask byte
extern "C" int emu_geom_flow_step(const UglGeomFlowArgs* g) {
  const UglFlowLossArgs* a = &g->flow;
  FlowGradParams gp;
  fill_params<kBTW, kBTH>(a, true, gp.base);
  for (int l = 0; l < a->scales; ++l) {
    gp.disp[l] = g->disp[l]; gp.Kinv[l] = g->Kinv[l];
    gp.P[0][l] = g->P_bwd[l]; gp.P[1][l] = g->P_fwd[l]; gp.mask_bytes[l] = g->mask_bytes[l];
  }
  gp.alpha = g->alpha; gp.beta = g->beta;
  return emu_flow_split<true>(gp, 1);
}

extern "C" int emu_flow_loss_combine(const UglFlowLossArgs* a) {
  FlowGradParams gp;
  fill_params<kBTW, kBTH>(a, true, gp.base);
  const FlowLossParams& p = gp.base;
  for (int l = 0; l < p.scales; ++l) {
    const FlowLevelDesc& L = p.lv[l];
    const int plane = L.h * L.w;
    for (int b = 0; b < p.B; ++b) {
      const FlowCombineScales k = flow_combine_scales(p.stats + ((size_t)b * p.scales + l) * FA_COUNT, L.h, L.w, p.gloss, p.B, b);
      for (int pix = 0; pix < plane; ++pix)
        flow_combine_pixel(a->basis[l] + (size_t)b * kBasisPlanes * plane, plane, pix, k, L.gflow_f + (size_t)b * 2 * plane,
                           L.gflow_b + (size_t)b * 2 * plane);
    }
  }
  return 0;
}

// geom mode (Model_geometry's flow branch): same tile logic with kGeom = true
extern "C" int emu_geom_flow_forward_grad(const UglGeomFlowArgs* g) {
  const UglFlowLossArgs* a = &g->flow;
  FlowGradParams gp;
  fill_params<kBTW, kBTH>(a, false, gp.base);
  for (int l = 0; l < a->scales; ++l) {
    gp.basis[l] = a->basis[l]; gp.disp[l] = g->disp[l]; gp.Kinv[l] = g->Kinv[l];
    gp.P[0][l] = g->P_bwd[l]; gp.P[1][l] = g->P_fwd[l]; gp.mask_bytes[l] = g->mask_bytes[l];
  }
  gp.alpha = g->alpha; gp.beta = g->beta;
  using Tile = FlowGradTile<kBTW, kBTH, 1, true>;
  const FlowLossParams& p = gp.base;
  std::vector<float> partials((size_t)p.total_tiles * GA_COUNT, 0.f), sm(Tile::kSmemFloats);
  for (int tile = 0; tile < p.total_tiles; ++tile) {
    const TileCoord tc = decode_tile<kBTW, kBTH>(p, tile);
    float acc[GA_COUNT] = {0}, mats[33];
    for (int k = 0; k < 9; ++k) mats[k] = gp.Kinv[tc.level][tc.b * 9 + k];
    for (int k = 0; k < 12; ++k) { mats[9 + k] = gp.P[0][tc.level][tc.b * 12 + k]; mats[21 + k] = gp.P[1][tc.level][tc.b * 12 + k]; }
    Tile::phase1(gp, tc, 0, 1, sm.data(), acc, mats);
    std::vector<float2> g3v((size_t)Tile::kP3 * 4, make_float2(0.f, 0.f));
    float2 (*g3)[4] = reinterpret_cast<float2 (*)[4]>(g3v.data());
    for (int c = 0; c < 3; ++c) {
      Tile::phase2(gp, tc, c, 0, 1, sm.data(), acc);
      if (c == 0 && !Tile::kDepth) Tile::phase2(gp, tc, 3, 0, 1, sm.data(), acc);
      Tile::phase3_accumulate(gp, tc, c, 0, 1, sm.data(), g3);
    }
    Tile::phase3_store(gp, tc, 0, 1, sm.data(), g3);
    Tile::phase4a(gp, tc, 0, 1, sm.data(), acc);
    Tile::phase4b(gp, tc, 0, 1, sm.data());
    for (int k = 0; k < GA_COUNT; ++k) partials[(size_t)tile * GA_COUNT + k] = acc[k];
  }
  for (int b = 0; b < p.B; ++b) {
    float tot[4] = {0, 0, 0, 0};
    for (int l = 0; l < p.scales; ++l) {
      const FlowLevelDesc& L = p.lv[l];
      const int per_img = L.tiles_x * L.tiles_y;
      double s[GA_COUNT] = {0};
      for (int t = 0; t < per_img; ++t)
        for (int k = 0; k < GA_COUNT; ++k) s[k] += partials[((size_t)L.tile_begin + (size_t)b * per_img + t) * GA_COUNT + k];
      float S[GA_COUNT], out[4];
      for (int k = 0; k < GA_COUNT; ++k) { S[k] = (float)s[k]; p.stats[((size_t)b * p.scales + l) * GA_COUNT + k] = S[k]; }
      geom_level_losses(S, L.h, L.w, out);
      for (int k = 0; k < 4; ++k) tot[k] += out[k];
    }
    for (int k = 0; k < 4; ++k) p.loss[k * p.B + b] = tot[k];
  }
  return 0;
}

extern "C" int emu_geom_flow_combine(const UglGeomFlowArgs* g) {
  const UglFlowLossArgs* a = &g->flow;
  FlowGradParams gp;
  fill_params<kBTW, kBTH>(a, true, gp.base);
  const FlowLossParams& p = gp.base;
  for (int l = 0; l < p.scales; ++l) {
    const FlowLevelDesc& L = p.lv[l];
    const int plane = L.h * L.w;
    for (int b = 0; b < p.B; ++b) {
      const GeomCombineScales k = geom_combine_scales(p.stats + ((size_t)b * p.scales + l) * GA_COUNT, L.h, L.w, p.gloss, p.B, b);
      for (int pix = 0; pix < plane; ++pix)
        geom_combine_pixel(a->basis[l] + (size_t)b * kBasisPlanes * plane, g->mask_bytes[l] + (size_t)b * plane, plane, pix, k,
                           L.gflow_f + (size_t)b * 2 * plane, L.gflow_b + (size_t)b * 2 * plane);
    }
  }
  return 0;
}

// depth mode (reprojection warps + L1 under valid*texture + SSIM under valid): tile logic with kModeDepth, then the combine
extern "C" int emu_depth_ssim_forward_grad(const UglDepthSsimArgs* g) {
  const UglDepthPhotoArgs* a = &g->photo;
  FlowGradParams gp;
  FlowLossParams& p = gp.base;
  p.B = a->batch; p.scales = a->scales;
  int tiles = 0;
  for (int l = 0; l < a->scales; ++l) {
    FlowLevelDesc& L = p.lv[l];
    L.h = a->height[l]; L.w = a->width[l];
    L.geom = make_warp_geom(L.w, L.h);
    L.img = a->img[l];
    L.tiles_x = (L.w + kBTW - 1) / kBTW; L.tiles_y = (L.h + kBTH - 1) / kBTH;
    L.tile_begin = tiles;
    tiles += L.tiles_x * L.tiles_y * a->batch;
    gp.basis[l] = g->basis[l]; gp.disp[l] = a->disp[l]; gp.Kinv[l] = a->Kinv[l];
    for (int d = 0; d < 2; ++d) {
      gp.P[d][l] = a->P[d][l]; gp.src_area[d][l] = a->src_area[d][l]; gp.src_bil[d][l] = a->src_bil[d][l];
      gp.valid_out[d][l] = a->valid_out[d][l]; gp.tex_out[d][l] = a->tex_out[d][l];
    }
  }
  p.total_tiles = tiles; p.stats = g->stats; p.loss = g->loss4;
  using Tile = FlowGradTile<kBTW, kBTH, 1, kModeDepth>;
  std::vector<float> partials((size_t)tiles * GA_COUNT, 0.f), sm(Tile::kSmemFloats);
  for (int tile = 0; tile < tiles; ++tile) {
    const TileCoord tc = decode_tile<kBTW, kBTH>(p, tile);
    float acc[GA_COUNT] = {0}, mats[33];
    for (int k = 0; k < 9; ++k) mats[k] = gp.Kinv[tc.level][tc.b * 9 + k];
    for (int k = 0; k < 12; ++k) { mats[9 + k] = gp.P[0][tc.level][tc.b * 12 + k]; mats[21 + k] = gp.P[1][tc.level][tc.b * 12 + k]; }
    Tile::phase1_depth(gp, tc, 0, 1, sm.data(), acc, mats);
    std::vector<float2> g3v((size_t)Tile::kP3 * 4, make_float2(0.f, 0.f));
    float2 (*g3)[4] = reinterpret_cast<float2 (*)[4]>(g3v.data());
    for (int c = 0; c < 3; ++c) {
      Tile::phase2(gp, tc, c, 0, 1, sm.data(), acc);
      if (c == 0 && !Tile::kDepth) Tile::phase2(gp, tc, 3, 0, 1, sm.data(), acc);
      Tile::phase3_accumulate(gp, tc, c, 0, 1, sm.data(), g3);
    }
    Tile::phase3_store(gp, tc, 0, 1, sm.data(), g3);
    for (int k = 0; k < GA_COUNT; ++k) partials[(size_t)tile * GA_COUNT + k] = acc[k];
  }
  for (int b = 0; b < p.B; ++b) {
    float tot[4] = {0, 0, 0, 0};
    for (int l = 0; l < p.scales; ++l) {
      const FlowLevelDesc& L = p.lv[l];
      const int per_img = L.tiles_x * L.tiles_y;
      double s[GA_COUNT] = {0};
      for (int t = 0; t < per_img; ++t)
        for (int k = 0; k < GA_COUNT; ++k) s[k] += partials[((size_t)L.tile_begin + (size_t)b * per_img + t) * GA_COUNT + k];
      float S[GA_COUNT], out[4];
      for (int k = 0; k < GA_COUNT; ++k) { S[k] = (float)s[k]; p.stats[((size_t)b * p.scales + l) * GA_COUNT + k] = S[k]; }
      depth_level_losses(S, L.h, L.w, out);
      for (int k = 0; k < 4; ++k) tot[k] += out[k];
    }
    for (int k = 0; k < 4; ++k) p.loss[k * p.B + b] = tot[k];
  }
  return 0;
}

extern "C" int emu_depth_ssim_combine(const UglDepthSsimArgs* g) {
  const UglDepthPhotoArgs* a = &g->photo;
  for (int l = 0; l < a->scales; ++l) {
    const int h = a->height[l], w = a->width[l];
    const size_t plane = (size_t)h * w;
    for (int b = 0; b < a->batch; ++b) {
      for (size_t px = 0; px < plane; ++px) a->grad_disp[l][b * plane + px] = 0.f;
      for (int d = 0; d < 2; ++d) {
        float k_pix, k_ssim;
        depth_combine_scales(g->stats + ((size_t)b * a->scales + l) * GA_COUNT, h, w, g->grad_loss4, a->batch, b, d, k_pix, k_ssim);
        const float* bs = g->basis[l] + ((size_t)b * 8 + 4 * d) * plane;
        const float* P = a->P[d][l] + b * 12;
        double accd[12] = {0};
        for (size_t px = 0; px < plane; ++px) {
          const int i = (int)(px / w), j = (int)(px % w);
          const float gu = k_pix * bs[px] + k_ssim * bs[2 * plane + px];
          const float gv = k_pix * bs[plane + px] + k_ssim * bs[3 * plane + px];
          const Projected r = project_pixel(a->Kinv[l] + b * 9, P, a->disp[l][b * plane + px], j, i);
          float acc[12] = {0};
          a->grad_disp[l][b * plane + px] += project_backward(r, P, gu, gv, 0.f, acc);
          for (int k = 0; k < 12; ++k) accd[k] += acc[k];
        }
        for (int k = 0; k < 12; ++k) a->grad_P[d][l][b * 12 + k] = (float)accd[k];
      }
    }
  }
  return 0;
}

extern "C" int emu_image_pyramid(const float* img, int B, int C, int H, int W, int levels, int mode, float* const* out) {
  for (int l = 1; l < levels; ++l) {
    const int oh = H >> l, ow = W >> l;
    for (long pl = 0; pl < (long)B * C; ++pl)
      for (int oi = 0; oi < oh; ++oi)
        for (int oj = 0; oj < ow; ++oj)
          out[l][(pl * oh + oi) * ow + oj] = pyramid_pixel(img + pl * (long)H * W, W, l, mode, oi, oj);
  }
  return 0;
}

extern "C" int emu_warp_flow_forward(const float* x, const float* flow, int B, int C, int H, int W, int use_mask,
                                     float* out, float* mask) {
  const WarpGeom geom = make_warp_geom(W, H);
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < H; ++i)
      for (int j = 0; j < W; ++j) {
        const float keep = warp_pixel_forward(x, flow, C, geom, b, i, j, use_mask, out);
        if (mask) mask[((long)b * H + i) * W + j] = keep;
      }
  return 0;
}

extern "C" int emu_warp_flow_backward(const float* x, const float* flow, const float* gout, int B, int C, int H, int W,
                                      int use_mask, float* gflow, float* gx) {
  const long plane = (long)H * W;
  const WarpGeom geom = make_warp_geom(W, H);
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < H; ++i)
      for (int j = 0; j < W; ++j) warp_pixel_backward_flow(x, flow, gout, C, geom, b, i, j, use_mask, gflow);
  if (gx) {
    const long nx = (long)B * C * plane;
    float m = 0.f;
    for (long k = 0; k < nx; ++k) { const float a = fabsf(gout[k]); if (a > m && a <= 3.0e38f) m = a; }
    const int e = fixed_point_exponent(m, plane);
    std::vector<long long> acc(nx, 0);
    for (int b = 0; b < B; ++b)
      for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
          const long pix = (long)i * W + j;
          const Tap t = flow_tap(j, i, flow[((long)b * 2) * plane + pix], flow[((long)b * 2 + 1) * plane + pix], geom);
          const float keep = use_mask ? tap_keep(t) : 1.0f;
          if (keep == 0.f || t.inb == 0u) continue;
          const float wgt[4] = {t.wnw, t.wne, t.wsw, t.wse};
          const long off[4] = {0, 1, W, (long)W + 1};
          for (int c = 0; c < C; ++c) {
            const float g = gout[((long)b * C + c) * plane + pix] * keep;
            for (int k = 0; k < 4; ++k)
              if (t.inb & (1u << k))
                acc[((long)b * C + c) * plane + (long)t.y0 * W + t.x0 + off[k]] += llrint(ldexp((double)(g * wgt[k]), e));
          }
        }
    for (long k = 0; k < nx; ++k) gx[k] = (float)ldexp((double)acc[k], -e);
  }
  return 0;
}
