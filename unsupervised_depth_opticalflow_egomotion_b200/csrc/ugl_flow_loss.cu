// CUDA kernels + C-ABI for the fused flow-mode loss (see ugl_flow_loss.cuh for the math and the
// reference citations).  Three launches per training step: forward, finalize, backward.
#include "ugl_flow_loss.cuh"
#include "ugl_flow_grad.cuh"
#include "ugl_host.cuh"
#include "ugl_flow_split.cuh"
#include "ugl_flow_split_host.cuh"

namespace ugl {

constexpr int kFNT = 256;   // forward  threads per CTA (one CTA per tile)
#ifndef UGL_BNT
#define UGL_BNT 256
#endif
constexpr int kBNT = UGL_BNT;   // backward / single-pass threads per CTA

template <int TW, int TH, int NT>
__global__ void __launch_bounds__(NT) flow_loss_fwd_kernel(const __grid_constant__ FlowLossParams p) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[(NT / 32) * FA_COUNT];
  using Tile = FlowFwdTile<TW, TH>;
  const int tile = blockIdx.x;
  const TileCoord tc = decode_tile<TW, TH>(p, tile);
  float acc[FA_COUNT];
#pragma unroll
  for (int k = 0; k < FA_COUNT; ++k) acc[k] = 0.f;
  Tile::phase1(p, tc, threadIdx.x, NT, sm, acc);
  __syncthreads();
  Tile::phase2(p, tc, threadIdx.x, NT, sm, acc);
  const float v = block_reduce_n<NT, FA_COUNT>(acc, red);
  if (threadIdx.x < FA_COUNT) p.partials[(long)tile * FA_COUNT + threadIdx.x] = v;
}

// Finalize: one CTA per (level, sample).  256 threads stride over that sample-level's tile partials (fp64, fixed
// order), a shuffle + shared-memory tree reduces them (deterministic), thread 0 applies the closing formulas and
// stores the level sums for backward.  The CTA that finishes a sample last (ticket counter, zeroed by a memset node
// before the launch) adds the levels up in level order -> loss (4,B).
constexpr int kFinThreads = 256;
template <int kMode>
__global__ void __launch_bounds__(kFinThreads)
flow_loss_finalize_kernel(const __grid_constant__ FlowLossParams p, float* __restrict__ lvl_loss /* [B][scales][4] */,
                          unsigned* __restrict__ tickets /* [B] */, const __grid_constant__ FlowGradParams::PhotoTiling pt) {
  constexpr int FA_COUNT = kMode == kModeFlow ? (int)ugl::FA_COUNT : (int)GA_COUNT;   // geom / depth modes carry four more sums
  __shared__ double red[kFinThreads / 32][FA_COUNT];
  __shared__ bool last;
  const int l = blockIdx.x, b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const FlowLevelDesc& L = p.lv[l];
  const int per_img = L.tiles_x * L.tiles_y;
  griddep_wait();   // launched as a programmatic dependent of the stencil kernel in the fused step (resident under its last wave)
  const float* base = p.partials + ((long)L.tile_begin + (long)b * per_img) * FA_COUNT;
  double s[FA_COUNT];
#pragma unroll
  for (int k = 0; k < FA_COUNT; ++k) s[k] = 0.0;
  // split kernels: the photometry kernel's columns (L1 / weight / consistency sums) live in its own partial rows
  using Px = FlowPhotoPixel<kMode == kModeGeom>;
  const bool split = (kMode != kModeDepth) && pt.partials != nullptr;
  for (int t = threadIdx.x; t < per_img; t += kFinThreads) {
#pragma unroll
    for (int k = 0; k < FA_COUNT; ++k)
      if (!(split && Px::is_photo_column(k))) s[k] += (double)base[(long)t * FA_COUNT + k];
  }
  if (split) {
    const float* pb = pt.partials + ((long)pt.tile_begin[l] + (long)b * pt.per_img[l]) * Px::kAcc;
    for (int t = threadIdx.x; t < pt.per_img[l]; t += kFinThreads) {
#pragma unroll
      for (int k = 0; k < Px::kAcc; ++k) s[Px::column(k)] += (double)pb[(long)t * Px::kAcc + k];
    }
  }
#pragma unroll
  for (int k = 0; k < FA_COUNT; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_down_sync(0xffffffffu, s[k], o);
    if (lane == 0) red[warp][k] = s[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float S[FA_COUNT], out[4];
    float* st = p.stats + ((long)b * p.scales + l) * FA_COUNT;
#pragma unroll
    for (int k = 0; k < FA_COUNT; ++k) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < kFinThreads / 32; ++w) v += red[w][k];
      S[k] = (float)v;
      st[k] = S[k];
    }
    if (kMode == kModeGeom) geom_level_losses(S, L.h, L.w, out);
    else if (kMode == kModeDepth) depth_level_losses(S, L.h, L.w, out);
    else flow_level_losses(S, L.h, L.w, out);
    float* ll = lvl_loss + ((long)b * p.scales + l) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) ll[k] = out[k];
    __threadfence();
    last = (atomicAdd(&tickets[b], 1u) == (unsigned)(p.scales - 1));
  }
  __syncthreads();
  if (last && threadIdx.x < 4) {
    __threadfence();
    float t = 0.f;
    for (int l2 = 0; l2 < p.scales; ++l2) t += __ldcg(lvl_loss + ((long)b * p.scales + l2) * 4 + threadIdx.x);
    p.loss[threadIdx.x * p.B + b] = t;
  }
}

// scratch behind the tile partials: [B][scales][4] level losses, then B ticket counters
template <int kMode>
static float* finalize_level_losses(const FlowLossParams& p) { return p.partials + (size_t)p.total_tiles * (kMode == kModeFlow ? (int)FA_COUNT : (int)GA_COUNT); }

template <int kMode = kModeFlow>
static int finalize_reset_tickets(const FlowLossParams& p, cudaStream_t st) {
  unsigned* tickets = reinterpret_cast<unsigned*>(finalize_level_losses<kMode>(p) + (size_t)p.B * p.scales * 4);
  const cudaError_t e = cudaMemsetAsync(tickets, 0, sizeof(unsigned) * p.B, st);
  if (e != cudaSuccess) return fail((int)e, "flow_loss finalize: memset: %s", cudaGetErrorString(e));
  return UGL_OK;
}

// chained = true: the tickets were reset before the first kernel of the step (finalize_reset_tickets) and the launch is a programmatic
// dependent of the kernel before it on the stream (a memset node in between would break the chain)
template <int kMode = kModeFlow>
static int launch_finalize(const FlowLossParams& p, cudaStream_t st, const FlowGradParams::PhotoTiling* photo = nullptr, bool chained = false) {
  FlowGradParams::PhotoTiling pt;
  if (photo) pt = *photo; else pt.partials = nullptr;
  float* lvl_loss = finalize_level_losses<kMode>(p);
  unsigned* tickets = reinterpret_cast<unsigned*>(lvl_loss + (size_t)p.B * p.scales * 4);
  int rc;
  if (!chained && (rc = finalize_reset_tickets<kMode>(p, st))) return rc;
  return launch_kernel("flow_loss_finalize_kernel", flow_loss_finalize_kernel<kMode>, dim3(p.scales, p.B), dim3(kFinThreads), 0, st, chained, p, lvl_loss, tickets, pt);
}

template <int TW, int TH, int NT>
__global__ void __launch_bounds__(NT) flow_loss_bwd_kernel(const __grid_constant__ FlowLossParams p) {
  extern __shared__ __align__(16) float sm[];
  using Tile = FlowBwdTile<TW, TH, NT>;
  const int tile = blockIdx.x;
  const TileCoord tc = decode_tile<TW, TH>(p, tile);
  const FlowLevelDesc& L = p.lv[tc.level];
  const FlowBwdCoef k = flow_bwd_coef(p.stats + ((long)tc.b * p.scales + tc.level) * FA_COUNT, L.h, L.w, p.gloss, p.B, tc.b);
  float g[Tile::PPT][4];
  Tile::phase1(p, tc, threadIdx.x, NT, sm);
  __syncthreads();
#pragma unroll
  for (int dir = 0; dir < 2; ++dir) {
    Tile::phase2(p, tc, dir, threadIdx.x, NT, sm);
    __syncthreads();
    Tile::phase3(p, tc, k, dir, threadIdx.x, NT, sm, g);
    if (dir == 0) __syncthreads();   // the coefficient planes are reused by the second direction
  }
  Tile::phase4(p, tc, k, threadIdx.x, NT, sm, g);
}

// single-pass: losses + gradient basis maps
template <int TW, int TH, int NT, int kMode>
__global__ void __launch_bounds__(NT, UGL_BMINB) flow_loss_fwdgrad_kernel(const __grid_constant__ FlowGradParams gp) {
  extern __shared__ __align__(16) float sm[];
  using Tile = FlowGradTile<TW, TH, NT, kMode>;
  constexpr int FA_COUNT = Tile::kAcc;
  constexpr bool kMats = (kMode != kModeFlow);
  __shared__ float red[(NT / 32) * FA_COUNT];
  __shared__ float mats[kMats ? 33 : 1];   // geom / depth modes: K^-1, P[0], P[1] of this tile's sample and level
  int tile;
  const TileCoord tc = decode_tile_2d<TW, TH>(gp.base, blockIdx.x, blockIdx.y, tile);
  if (kMats) {
    if (threadIdx.x < 9) mats[threadIdx.x] = gp.Kinv[tc.level][tc.b * 9 + threadIdx.x];
    else if (threadIdx.x < 21) mats[threadIdx.x] = gp.P[0][tc.level][tc.b * 12 + threadIdx.x - 9];
    else if (threadIdx.x < 33) mats[threadIdx.x] = gp.P[1][tc.level][tc.b * 12 + threadIdx.x - 21];
    __syncthreads();
  }
  float acc[FA_COUNT];
#pragma unroll
  for (int k = 0; k < FA_COUNT; ++k) acc[k] = 0.f;
  if (kMode == kModeDepth) Tile::phase1_depth(gp, tc, threadIdx.x, NT, sm, acc, mats);
  else Tile::phase1(gp, tc, threadIdx.x, NT, sm, acc, mats);
  __syncthreads();
  float2 g3[Tile::kP3][4];
#pragma unroll
  for (int n = 0; n < Tile::kP3; ++n)
#pragma unroll
    for (int k = 0; k < 4; ++k) g3[n][k] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {      // channel by channel (rolled: one copy of the stencil phases in the instruction cache)
    Tile::phase2(gp, tc, c, threadIdx.x, NT, sm, acc);
    if (c == 0 && kMode != kModeDepth) Tile::phase2(gp, tc, 3, threadIdx.x, NT, sm, acc);   // smoothness edge weights
    __syncthreads();
    Tile::phase3_accumulate(gp, tc, c, threadIdx.x, NT, sm, g3);
    __syncthreads();                 // the coefficient planes are reused by the next channel; phase 4a overwrites the X / Y planes
  }
  Tile::phase3_store(gp, tc, threadIdx.x, NT, sm, g3);
  if (kMode != kModeDepth) {         // flow smoothness (flow / geom modes)
    Tile::phase4a(gp, tc, threadIdx.x, NT, sm, acc);
    __syncthreads();
    Tile::phase4b(gp, tc, threadIdx.x, NT, sm);
  }
  const float v = block_reduce_n<NT, FA_COUNT>(acc, red);
  if (threadIdx.x < FA_COUNT) gp.base.partials[(long)tile * FA_COUNT + threadIdx.x] = v;
}

// element-wise backward of the single-pass mode: grid (chunks, B, scales)
template <bool kGeom>
__global__ void __launch_bounds__(256) flow_combine_kernel(const __grid_constant__ FlowGradParams gp) {
  const FlowLossParams& p = gp.base;
  const int b = blockIdx.y, l = blockIdx.z;
  const FlowLevelDesc& L = p.lv[l];
  const int plane = L.h * L.w;
  const float* basis = gp.basis[l] + (long)b * kBasisPlanes * plane;
  float* gf = L.gflow_f + (long)b * 2 * plane;
  float* gb = L.gflow_b + (long)b * 2 * plane;
  if (kGeom) {   // the L1 scale depends on the pixel's dynamic-mask bit
    const GeomCombineScales kg = geom_combine_scales(p.stats + ((long)b * p.scales + l) * GA_COUNT, L.h, L.w, p.gloss, p.B, b);
    const unsigned char* mask = gp.mask_bytes[l] + (long)b * plane;
    if ((plane & 3) == 0) {   // four pixels per thread: 128-bit basis loads, one 32-bit load of the four mask bytes
      const float4* b4 = reinterpret_cast<const float4*>(basis);
      const unsigned* m4 = reinterpret_cast<const unsigned*>(mask);
      float4* gf4 = reinterpret_cast<float4*>(gf);
      float4* gb4 = reinterpret_cast<float4*>(gb);
      const int q = plane >> 2;
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < q; i += gridDim.x * blockDim.x) {
        const unsigned bits = m4[i];
        float kf[4], kb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const unsigned bt = bits >> (8 * e);
          kf[e] = (bt & kMaskDynF) ? kg.pix_r[0] : kg.pix_d[0];
          kb[e] = (bt & kMaskDynB) ? kg.pix_r[1] : kg.pix_d[1];
        }
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const float4 a0 = __ldcs(b4 + (0 + ch) * q + i), a1 = __ldcs(b4 + (2 + ch) * q + i), a2 = __ldcs(b4 + (4 + ch) * q + i),
                       a3 = __ldcs(b4 + (6 + ch) * q + i);
          const float4 c0 = __ldcs(b4 + (8 + ch) * q + i), c1 = __ldcs(b4 + (10 + ch) * q + i), c2 = __ldcs(b4 + (12 + ch) * q + i);
          float4 f, g;
          f.x = kf[0] * a0.x + kg.ssim[0] * a1.x + kg.sm * a2.x + kg.cons * a3.x;
          f.y = kf[1] * a0.y + kg.ssim[0] * a1.y + kg.sm * a2.y + kg.cons * a3.y;
          f.z = kf[2] * a0.z + kg.ssim[0] * a1.z + kg.sm * a2.z + kg.cons * a3.z;
          f.w = kf[3] * a0.w + kg.ssim[0] * a1.w + kg.sm * a2.w + kg.cons * a3.w;
          g.x = kb[0] * c0.x + kg.ssim[1] * c1.x + kg.sm * c2.x;
          g.y = kb[1] * c0.y + kg.ssim[1] * c1.y + kg.sm * c2.y;
          g.z = kb[2] * c0.z + kg.ssim[1] * c1.z + kg.sm * c2.z;
          g.w = kb[3] * c0.w + kg.ssim[1] * c1.w + kg.sm * c2.w;
          gf4[ch * q + i] = f;
          gb4[ch * q + i] = g;
        }
      }
      return;
    }
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < plane; pix += gridDim.x * blockDim.x)
      geom_combine_pixel(basis, mask, plane, pix, kg, gf, gb);
    return;
  }
  const FlowCombineScales k = flow_combine_scales(p.stats + ((long)b * p.scales + l) * FA_COUNT, L.h, L.w, p.gloss, p.B, b);
  if ((plane & 3) == 0) {   // 128-bit path (every plane start is then 16-byte aligned: torch allocations are 512-byte aligned)
    const float4* b4 = reinterpret_cast<const float4*>(basis);
    float4* gf4 = reinterpret_cast<float4*>(gf);
    float4* gb4 = reinterpret_cast<float4*>(gb);
    const int q = plane >> 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < q; i += gridDim.x * blockDim.x) {
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const float4 a0 = __ldcs(b4 + (0 + ch) * q + i), a1 = __ldcs(b4 + (2 + ch) * q + i), a2 = __ldcs(b4 + (4 + ch) * q + i),
                     a3 = __ldcs(b4 + (6 + ch) * q + i);
        const float4 c0 = __ldcs(b4 + (8 + ch) * q + i), c1 = __ldcs(b4 + (10 + ch) * q + i), c2 = __ldcs(b4 + (12 + ch) * q + i);
        float4 f, g;
        f.x = k.pix[0] * a0.x + k.ssim[0] * a1.x + k.sm * a2.x + k.cons * a3.x;
        f.y = k.pix[0] * a0.y + k.ssim[0] * a1.y + k.sm * a2.y + k.cons * a3.y;
        f.z = k.pix[0] * a0.z + k.ssim[0] * a1.z + k.sm * a2.z + k.cons * a3.z;
        f.w = k.pix[0] * a0.w + k.ssim[0] * a1.w + k.sm * a2.w + k.cons * a3.w;
        g.x = k.pix[1] * c0.x + k.ssim[1] * c1.x + k.sm * c2.x;
        g.y = k.pix[1] * c0.y + k.ssim[1] * c1.y + k.sm * c2.y;
        g.z = k.pix[1] * c0.z + k.ssim[1] * c1.z + k.sm * c2.z;
        g.w = k.pix[1] * c0.w + k.ssim[1] * c1.w + k.sm * c2.w;
        gf4[ch * q + i] = f;
        gb4[ch * q + i] = g;
      }
    }
  } else {
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < plane; pix += gridDim.x * blockDim.x)
      flow_combine_pixel(basis, plane, pix, k, gf, gb);
  }
}

// ---- host side --------------------------------------------------------------------------------
template <int TW, int TH>
static int build_params(const UglFlowLossArgs* a, bool backward, FlowLossParams& p) {
  if (!a) return fail(UGL_EINVAL, "flow_loss: null args");
  if (a->batch <= 0 || a->levels <= 0 || a->levels > UGL_MAX_LEVELS || a->scales <= 0 || a->scales > a->levels)
    return fail(UGL_EINVAL, "flow_loss: bad batch/levels/scales (%d/%d/%d)", a->batch, a->levels, a->scales);
  p.B = a->batch;
  p.scales = a->scales;
  int tiles = 0;
  for (int l = 0; l < a->scales; ++l) {
    FlowLevelDesc& L = p.lv[l];
    L.h = a->height[l];
    L.w = a->width[l];
    if (L.h < 3 || L.w < 3) return fail(UGL_EUNSUPPORTED, "flow_loss: level %d is %dx%d; need >= 3x3", l, L.h, L.w);
    const void* ptrs[5] = {a->img_l[l], a->img[l], a->img_r[l], a->flow_fwd[l], a->flow_bwd[l]};
    for (int k = 0; k < 5; ++k) {
      if (!ptrs[k]) return fail(UGL_EINVAL, "flow_loss: null input pointer at level %d", l);
      if (!aligned4(ptrs[k])) return fail(UGL_EALIGN, "flow_loss: misaligned input pointer at level %d", l);
    }
    L.geom = make_warp_geom(L.w, L.h);
    L.img_l = a->img_l[l]; L.img = a->img[l]; L.img_r = a->img_r[l];
    L.flow_f = a->flow_fwd[l]; L.flow_b = a->flow_bwd[l];
    L.gflow_f = backward ? a->grad_flow_fwd[l] : nullptr;
    L.gflow_b = backward ? a->grad_flow_bwd[l] : nullptr;
    if (backward && (!L.gflow_f || !L.gflow_b)) return fail(UGL_EINVAL, "flow_loss: null grad_flow pointer at level %d", l);
    L.tiles_x = (L.w + TW - 1) / TW;
    L.tiles_y = (L.h + TH - 1) / TH;
    L.tile_begin = tiles;
    tiles += L.tiles_x * L.tiles_y * a->batch;
  }
  p.total_tiles = tiles;
  if (!a->stats) return fail(UGL_EINVAL, "flow_loss: null stats");
  p.stats = a->stats;
  p.loss = a->loss;
  p.gloss = a->grad_loss;
  p.partials = static_cast<float*>(a->workspace);
  return UGL_OK;
}

}  // namespace ugl

using namespace ugl;

// tile partials + [B][scales][4] level losses + B ticket counters (finalize scratch)
static uint64_t flow_partials_bytes(const UglFlowLossArgs* a) {
  uint64_t tf = 0, tb = 0;   // only the shapes matter here; cover the forward and the single-pass tile shapes
  for (int l = 0; l < a->scales && l < UGL_MAX_LEVELS; ++l) {
    tf += (uint64_t)((a->width[l] + kFTW - 1) / kFTW) * ((a->height[l] + kFTH - 1) / kFTH) * a->batch;
    tb += (uint64_t)((a->width[l] + kBTW - 1) / kBTW) * ((a->height[l] + kBTH - 1) / kBTH) * a->batch;
  }
  return (tf > tb ? tf : tb) * GA_COUNT * sizeof(float) + (uint64_t)a->batch * a->scales * 4 * sizeof(float) + (uint64_t)a->batch * sizeof(unsigned);
}

// workspace layout: [tile partials, level losses, tickets] [photometry-kernel partial rows, 256-byte aligned] [photometry planes]
static char* split_photo_partials(const UglFlowLossArgs* a) {
  const uintptr_t p = reinterpret_cast<uintptr_t>(a->workspace) + flow_partials_bytes(a);
  return reinterpret_cast<char*>((p + 255) & ~(uintptr_t)255);
}

// + the photometry planes the split single-pass kernels hand from the photometry kernel to the stencil kernel
extern "C" uint64_t ugl_flow_loss_workspace_bytes(const UglFlowLossArgs* a) {
  if (!a) return 0;
  return flow_partials_bytes(a) + 256 + flow_split_photo_partials_bytes(a->height, a->width, a->scales, a->batch) +
         flow_split_scratch_bytes(a->height, a->width, a->scales, a->batch);
}

extern "C" int ugl_flow_loss_launches(int backward) { return backward ? 1 : 2; }   // recompute mode; single-pass: 3 forward (photometry, stencil, finalize), 1 combine

extern "C" int ugl_flow_loss_forward(const UglFlowLossArgs* a) {
  FlowLossParams p;
  int rc = build_params<kFTW, kFTH>(a, false, p);
  if (rc) return rc;
  if (!a->loss) return fail(UGL_EINVAL, "flow_loss_forward: null loss");
  if (!a->workspace || a->workspace_bytes < ugl_flow_loss_workspace_bytes(a))
    return fail(UGL_EWORKSPACE, "flow_loss_forward: workspace too small (%llu bytes given)", (unsigned long long)a->workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  using Tile = FlowFwdTile<kFTW, kFTH>;
  constexpr size_t smem = Tile::kSmemFloats * sizeof(float);
  auto kern = flow_loss_fwd_kernel<kFTW, kFTH, kFNT>;
  static_assert(smem <= 227 * 1024, "forward tile does not fit in shared memory");
  if ((rc = opt_in_smem(kern, smem))) return rc;
  kern<<<p.total_tiles, kFNT, smem, st>>>(p);
  if ((rc = check_launch("flow_loss_fwd_kernel"))) return rc;
  return launch_finalize(p, st);
}

extern "C" int ugl_flow_loss_forward_grad_ex(const UglFlowLossArgs* a, int variant) {
  FlowGradParams gp;
  int rc = build_params<kBTW, kBTH>(a, false, gp.base);
  if (rc) return rc;
  if (!a->loss) return fail(UGL_EINVAL, "flow_loss_forward_grad: null loss");
  if (variant < UGL_SINGLE_PASS_FUSED || variant > UGL_SINGLE_PASS_SPLIT_TMA) return fail(UGL_EINVAL, "flow_loss_forward_grad: unknown variant %d", variant);
  for (int l = 0; l < a->scales; ++l) {
    if (!a->basis[l]) return fail(UGL_EINVAL, "flow_loss_forward_grad: null basis pointer at level %d", l);
    if (reinterpret_cast<uintptr_t>(a->basis[l]) & 7u) return fail(UGL_EALIGN, "flow_loss_forward_grad: basis not 8-byte aligned");
    gp.basis[l] = a->basis[l];
  }
  if (!a->workspace || a->workspace_bytes < ugl_flow_loss_workspace_bytes(a))
    return fail(UGL_EWORKSPACE, "flow_loss_forward_grad: workspace too small (%llu bytes given)", (unsigned long long)a->workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  if (variant != UGL_SINGLE_PASS_FUSED) {
    char* pp = split_photo_partials(a);
    flow_split_assign_scratch(gp, pp + flow_split_photo_partials_bytes(a->height, a->width, a->scales, a->batch));
    if ((rc = launch_flow_split<false>(gp, pp, st, variant == UGL_SINGLE_PASS_SPLIT ? 1 : (variant == UGL_SINGLE_PASS_SPLIT_TMA ? 2 : 0)))) return rc;
    return launch_finalize(gp.base, st, &gp.photo);
  }
  using Tile = FlowGradTile<kBTW, kBTH, kBNT>;
  constexpr size_t smem = Tile::kSmemFloats * sizeof(float);
  static_assert(smem <= 227 * 1024, "single-pass tile does not fit in shared memory");
  auto kern = flow_loss_fwdgrad_kernel<kBTW, kBTH, kBNT, kModeFlow>;
  if ((rc = opt_in_smem(kern, smem))) return rc;
  kern<<<dim3(gp.base.total_tiles / gp.base.B, gp.base.B), kBNT, smem, st>>>(gp);
  if ((rc = check_launch("flow_loss_fwdgrad_kernel"))) return rc;
  return launch_finalize(gp.base, st);
}

extern "C" int ugl_flow_loss_forward_grad(const UglFlowLossArgs* a) { return ugl_flow_loss_forward_grad_ex(a, UGL_SINGLE_PASS_SPLIT); }

// fused forward + backward for a known upstream gradient: photometry kernel -> weight sums -> stencil kernel (writes the flow gradients)
// -> finalize (losses).  No basis planes, no combine launch.
extern "C" int ugl_flow_loss_step(const UglFlowLossArgs* a) { return ugl_flow_loss_step_parts(a, UGL_STEP_ALL); }

extern "C" int ugl_flow_loss_step_parts(const UglFlowLossArgs* a, int parts) {
  FlowGradParams gp;
  int rc = build_params<kBTW, kBTH>(a, true, gp.base);
  if (rc) return rc;
  if (!a->loss || !a->grad_loss) return fail(UGL_EINVAL, "flow_loss_step: null loss / grad_loss");
  if (!a->workspace || a->workspace_bytes < ugl_flow_loss_workspace_bytes(a))
    return fail(UGL_EWORKSPACE, "flow_loss_step: workspace too small (%llu bytes given)", (unsigned long long)a->workspace_bytes);
  for (int l = 0; l < a->scales; ++l)
    if ((reinterpret_cast<uintptr_t>(a->grad_flow_fwd[l]) | reinterpret_cast<uintptr_t>(a->grad_flow_bwd[l])) & 7u)
      return fail(UGL_EALIGN, "flow_loss_step: grad_flow not 8-byte aligned at level %d", l);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  gp.step = 1;
  char* pp = split_photo_partials(a);
  flow_split_assign_scratch(gp, pp + flow_split_photo_partials_bytes(a->height, a->width, a->scales, a->batch));
  const bool chained = parts == UGL_STEP_ALL;
  if (chained && (rc = finalize_reset_tickets(gp.base, st))) return rc;
  if ((rc = launch_flow_split<false>(gp, pp, st, 1, parts & 7))) return rc;
  return (parts & UGL_STEP_FINALIZE) ? launch_finalize(gp.base, st, &gp.photo, chained) : UGL_OK;
}

// ---- geom mode (Model_geometry's flow branch) -----------------------------------------------------
// step = true: the fused training step (forward inputs AND gradient outputs, no basis planes)
static int geom_params(const UglGeomFlowArgs* g, bool backward, FlowGradParams& gp, bool step = false) {
  if (!g) return fail(UGL_EINVAL, "geom_flow: null args");
  const UglFlowLossArgs* a = &g->flow;
  int rc = build_params<kBTW, kBTH>(a, backward || step, gp.base);
  if (rc) return rc;
  for (int l = 0; l < a->scales; ++l) {
    if (!g->mask_bytes[l]) return fail(UGL_EINVAL, "geom_flow: null mask_bytes pointer at level %d", l);
    gp.mask_bytes[l] = g->mask_bytes[l];
    if (!step) {
      if (!a->basis[l]) return fail(UGL_EINVAL, "geom_flow: null basis pointer at level %d", l);
      if (reinterpret_cast<uintptr_t>(a->basis[l]) & 7u) return fail(UGL_EALIGN, "geom_flow: basis not 8-byte aligned");
      gp.basis[l] = a->basis[l];
    }
    if (!backward) {
      const void* ptrs[4] = {g->disp[l], g->Kinv[l], g->P_bwd[l], g->P_fwd[l]};
      for (int k = 0; k < 4; ++k) {
        if (!ptrs[k]) return fail(UGL_EINVAL, "geom_flow: null disp / Kinv / P pointer at level %d", l);
        if (!aligned4(ptrs[k])) return fail(UGL_EALIGN, "geom_flow: misaligned disp / Kinv / P pointer at level %d", l);
      }
      gp.disp[l] = g->disp[l]; gp.Kinv[l] = g->Kinv[l]; gp.P[0][l] = g->P_bwd[l]; gp.P[1][l] = g->P_fwd[l];
    }
  }
  gp.alpha = g->alpha; gp.beta = g->beta;
  return UGL_OK;
}

extern "C" int ugl_geom_flow_forward_grad_ex(const UglGeomFlowArgs* g, int variant) {
  FlowGradParams gp;
  int rc = geom_params(g, false, gp);
  if (rc) return rc;
  const UglFlowLossArgs* a = &g->flow;
  if (!a->loss) return fail(UGL_EINVAL, "geom_flow_forward_grad: null loss");
  if (variant < UGL_SINGLE_PASS_FUSED || variant > UGL_SINGLE_PASS_SPLIT_TMA) return fail(UGL_EINVAL, "geom_flow_forward_grad: unknown variant %d", variant);
  if (!a->workspace || a->workspace_bytes < ugl_flow_loss_workspace_bytes(a))
    return fail(UGL_EWORKSPACE, "geom_flow_forward_grad: workspace too small (%llu bytes given)", (unsigned long long)a->workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  if (variant != UGL_SINGLE_PASS_FUSED) {
    char* pp = split_photo_partials(a);
    flow_split_assign_scratch(gp, pp + flow_split_photo_partials_bytes(a->height, a->width, a->scales, a->batch));
    if ((rc = launch_flow_split<true>(gp, pp, st, variant == UGL_SINGLE_PASS_SPLIT ? 1 : (variant == UGL_SINGLE_PASS_SPLIT_TMA ? 2 : 0)))) return rc;
    return launch_finalize<kModeGeom>(gp.base, st, &gp.photo);
  }
  using Tile = FlowGradTile<kBTW, kBTH, kBNT, kModeGeom>;
  constexpr size_t smem = Tile::kSmemFloats * sizeof(float);
  auto kern = flow_loss_fwdgrad_kernel<kBTW, kBTH, kBNT, kModeGeom>;
  if ((rc = opt_in_smem(kern, smem))) return rc;
  kern<<<dim3(gp.base.total_tiles / gp.base.B, gp.base.B), kBNT, smem, st>>>(gp);
  if ((rc = check_launch("flow_loss_fwdgrad_kernel<geom>"))) return rc;
  return launch_finalize<kModeGeom>(gp.base, st);
}

extern "C" int ugl_geom_flow_forward_grad(const UglGeomFlowArgs* g) { return ugl_geom_flow_forward_grad_ex(g, UGL_SINGLE_PASS_SPLIT); }

// fused forward + backward of the geom-mode flow branch for a known upstream gradient (ugl_flow_loss_step's counterpart): losses,
// packed mask bytes AND grad_flow_fwd/bwd in four chained launches, no basis planes, no combine launch
extern "C" int ugl_geom_flow_step(const UglGeomFlowArgs* g) { return ugl_geom_flow_step_parts(g, UGL_STEP_ALL); }

// parts = UGL_STEP_PHOTO, then UGL_STEP_ALL & ~UGL_STEP_PHOTO: the same step in two calls, so that the caller can start what only needs the
// photometry kernel's mask bytes (the reprojection term, the level-0 rigid terms) on other streams while the stencil kernel runs
extern "C" int ugl_geom_flow_step_parts(const UglGeomFlowArgs* g, int parts) {
  FlowGradParams gp;
  int rc = geom_params(g, false, gp, true);
  if (rc) return rc;
  const UglFlowLossArgs* a = &g->flow;
  if (!a->loss || !a->grad_loss) return fail(UGL_EINVAL, "geom_flow_step: null loss / grad_loss");
  if (!a->workspace || a->workspace_bytes < ugl_flow_loss_workspace_bytes(a))
    return fail(UGL_EWORKSPACE, "geom_flow_step: workspace too small (%llu bytes given)", (unsigned long long)a->workspace_bytes);
  for (int l = 0; l < a->scales; ++l)
    if ((reinterpret_cast<uintptr_t>(a->grad_flow_fwd[l]) | reinterpret_cast<uintptr_t>(a->grad_flow_bwd[l])) & 7u)
      return fail(UGL_EALIGN, "geom_flow_step: grad_flow not 8-byte aligned at level %d", l);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  gp.step = 1;
  char* pp = split_photo_partials(a);
  flow_split_assign_scratch(gp, pp + flow_split_photo_partials_bytes(a->height, a->width, a->scales, a->batch));
  if (parts != UGL_STEP_ALL && parts != UGL_STEP_PHOTO && parts != (UGL_STEP_ALL & ~UGL_STEP_PHOTO))
    return fail(UGL_EINVAL, "geom_flow_step_parts: parts must be all, the photometry kernel, or everything after it");
  if ((parts & UGL_STEP_PHOTO) && (rc = finalize_reset_tickets<kModeGeom>(gp.base, st))) return rc;
  if ((rc = launch_flow_split<true>(gp, pp, st, 1, parts & 7))) return rc;
  // chained (programmatic dependent launch): the finalize follows the stencil kernel on the stream in both forms
  return (parts & UGL_STEP_FINALIZE) ? launch_finalize<kModeGeom>(gp.base, st, &gp.photo, true) : UGL_OK;
}

extern "C" int ugl_geom_flow_combine(const UglGeomFlowArgs* g) {
  FlowGradParams gp;
  int rc = geom_params(g, true, gp);
  if (rc) return rc;
  const UglFlowLossArgs* a = &g->flow;
  if (!a->grad_loss) return fail(UGL_EINVAL, "geom_flow_combine: null grad_loss");
  int max_plane = 0;
  for (int l = 0; l < a->scales; ++l) {
    const int pl = a->height[l] * a->width[l];
    max_plane = pl > max_plane ? pl : max_plane;
  }
  int chunks = (max_plane + 256 * 4 - 1) / (256 * 4);
  chunks = chunks < 1 ? 1 : (chunks > 64 ? 64 : chunks);
  flow_combine_kernel<true><<<dim3(chunks, gp.base.B, gp.base.scales), 256, 0, static_cast<cudaStream_t>(a->stream)>>>(gp);
  return check_launch("flow_combine_kernel<geom>");
}

extern "C" int ugl_flow_loss_combine(const UglFlowLossArgs* a) {
  FlowGradParams gp;
  int rc = build_params<kBTW, kBTH>(a, true, gp.base);
  if (rc) return rc;
  if (!a->grad_loss) return fail(UGL_EINVAL, "flow_loss_combine: null grad_loss");
  int max_plane = 0;
  for (int l = 0; l < a->scales; ++l) {
    if (!a->basis[l]) return fail(UGL_EINVAL, "flow_loss_combine: null basis pointer at level %d", l);
    gp.basis[l] = a->basis[l];
    const int pl = a->height[l] * a->width[l];
    max_plane = pl > max_plane ? pl : max_plane;
  }
  int chunks = (max_plane + 256 * 4 - 1) / (256 * 4);
  chunks = chunks < 1 ? 1 : (chunks > 64 ? 64 : chunks);
  flow_combine_kernel<false><<<dim3(chunks, gp.base.B, gp.base.scales), 256, 0, static_cast<cudaStream_t>(a->stream)>>>(gp);
  return check_launch("flow_combine_kernel");
}

extern "C" int ugl_flow_loss_backward(const UglFlowLossArgs* a) {
  FlowLossParams p;
  int rc = build_params<kBTW, kBTH>(a, true, p);
  if (rc) return rc;
  if (!a->grad_loss) return fail(UGL_EINVAL, "flow_loss_backward: null grad_loss");
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  using Tile = FlowBwdTile<kBTW, kBTH, kBNT>;
  constexpr size_t smem = Tile::kSmemFloats * sizeof(float);
  static_assert(smem <= 227 * 1024, "backward tile does not fit in shared memory");
  auto kern = flow_loss_bwd_kernel<kBTW, kBTH, kBNT>;
  if ((rc = opt_in_smem(kern, smem))) return rc;
  kern<<<p.total_tiles, kBNT, smem, st>>>(p);
  return check_launch("flow_loss_bwd_kernel");
}

// ---- depth mode (Model_depth with SSIM: model_depth_texture.py:296-301) -----------------------------
// forward: the single-pass tile kernel in depth mode + finalize.  The backward (chain of the (u,v) basis through the
// projection to disp and P) lives in ugl_depth_photo.cu: ugl_depth_ssim_combine.
extern "C" int ugl_depth_ssim_forward_grad(const UglDepthSsimArgs* g) {
  if (!g) return fail(UGL_EINVAL, "depth_ssim: null args");
  const UglDepthPhotoArgs* a = &g->photo;
  if (a->batch <= 0 || a->scales <= 0 || a->scales > UGL_MAX_LEVELS) return fail(UGL_EINVAL, "depth_ssim: bad batch/scales (%d/%d)", a->batch, a->scales);
  FlowGradParams gp;
  FlowLossParams& p = gp.base;
  p.B = a->batch; p.scales = a->scales;
  int tiles = 0;
  for (int l = 0; l < a->scales; ++l) {
    FlowLevelDesc& L = p.lv[l];
    L.h = a->height[l]; L.w = a->width[l];
    if (L.h < 3 || L.w < 3) return fail(UGL_EUNSUPPORTED, "depth_ssim: level %d is %dx%d; need >= 3x3", l, L.h, L.w);
    const void* ptrs[10] = {a->img[l], a->disp[l], a->Kinv[l], a->src_area[0][l], a->src_area[1][l], a->src_bil[0][l], a->src_bil[1][l],
                            a->P[0][l], a->P[1][l], g->basis[l]};
    for (int k = 0; k < 10; ++k) {
      if (!ptrs[k]) return fail(UGL_EINVAL, "depth_ssim: null pointer (%d) at level %d", k, l);
      if (!aligned4(ptrs[k])) return fail(UGL_EALIGN, "depth_ssim: misaligned pointer (%d) at level %d", k, l);
    }
    if (reinterpret_cast<uintptr_t>(g->basis[l]) & 7u) return fail(UGL_EALIGN, "depth_ssim: basis not 8-byte aligned");
    L.geom = make_warp_geom(L.w, L.h);
    L.img = a->img[l]; L.img_l = nullptr; L.img_r = nullptr; L.flow_f = nullptr; L.flow_b = nullptr; L.gflow_f = nullptr; L.gflow_b = nullptr;
    L.tiles_x = (L.w + kBTW - 1) / kBTW;
    L.tiles_y = (L.h + kBTH - 1) / kBTH;
    L.tile_begin = tiles;
    tiles += L.tiles_x * L.tiles_y * a->batch;
    gp.basis[l] = g->basis[l]; gp.disp[l] = a->disp[l]; gp.Kinv[l] = a->Kinv[l];
    for (int d = 0; d < 2; ++d) {
      gp.P[d][l] = a->P[d][l]; gp.src_area[d][l] = a->src_area[d][l]; gp.src_bil[d][l] = a->src_bil[d][l];
      gp.valid_out[d][l] = a->valid_out[d][l]; gp.tex_out[d][l] = a->tex_out[d][l];
    }
  }
  p.total_tiles = tiles;
  if (!g->stats || !g->loss4) return fail(UGL_EINVAL, "depth_ssim: null stats / loss");
  p.stats = g->stats; p.loss = g->loss4; p.gloss = nullptr;
  const uint64_t need = (uint64_t)tiles * GA_COUNT * sizeof(float) + (uint64_t)a->batch * a->scales * 4 * sizeof(float) + (uint64_t)a->batch * sizeof(unsigned);
  if (!a->workspace || a->workspace_bytes < need)
    return fail(UGL_EWORKSPACE, "depth_ssim_forward_grad: workspace too small (%llu < %llu)", (unsigned long long)a->workspace_bytes, (unsigned long long)need);
  p.partials = static_cast<float*>(a->workspace);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  using Tile = FlowGradTile<kBTW, kBTH, kBNT, kModeDepth>;
  constexpr size_t smem = Tile::kSmemFloats * sizeof(float);
  auto kern = flow_loss_fwdgrad_kernel<kBTW, kBTH, kBNT, kModeDepth>;
  int rc;
  if ((rc = opt_in_smem(kern, smem))) return rc;
  kern<<<dim3(p.total_tiles / p.B, p.B), kBNT, smem, st>>>(gp);
  if ((rc = check_launch("flow_loss_fwdgrad_kernel<depth>"))) return rc;
  return launch_finalize<kModeDepth>(p, st);
}

extern "C" uint64_t ugl_depth_ssim_workspace_bytes(const UglDepthSsimArgs* g) {
  if (!g) return 0;
  const UglDepthPhotoArgs* a = &g->photo;
  uint64_t tiles = 0;
  long max_plane = 0;
  for (int l = 0; l < a->scales && l < UGL_MAX_LEVELS; ++l) {
    tiles += (uint64_t)((a->width[l] + kBTW - 1) / kBTW) * ((a->height[l] + kBTH - 1) / kBTH) * a->batch;
    const long pl = (long)a->height[l] * a->width[l];
    max_plane = pl > max_plane ? pl : max_plane;
  }
  const uint64_t fwd = tiles * GA_COUNT * sizeof(float) + (uint64_t)a->batch * a->scales * 4 * sizeof(float) + (uint64_t)a->batch * sizeof(unsigned);
  const uint64_t bwd = ugl_depth_photo_workspace_bytes(a);
  return fwd > bwd ? fwd : bwd;
}
