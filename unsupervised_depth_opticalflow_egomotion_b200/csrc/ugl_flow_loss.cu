// CUDA kernels + C-ABI for the fused flow-mode loss (see ugl_flow_loss.cuh for the math and the
// reference citations).  Three launches per training step: forward, finalize, backward.
#include "ugl_flow_loss.cuh"
#include "ugl_flow_grad.cuh"
#include "ugl_host.cuh"

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

// one CTA per sample, one warp per level: ordered (deterministic) fp64 sum of the tile partials
__global__ void flow_loss_finalize_kernel(const __grid_constant__ FlowLossParams p) {
  __shared__ float lvl_loss[kMaxLevels][4];
  const int b = blockIdx.x, lane = threadIdx.x & 31, l = threadIdx.x >> 5;
  if (l < p.scales) {
    const FlowLevelDesc& L = p.lv[l];
    const int per_img = L.tiles_x * L.tiles_y;
    const float* base = p.partials + ((long)L.tile_begin + (long)b * per_img) * FA_COUNT;
    double s[FA_COUNT];
#pragma unroll
    for (int k = 0; k < FA_COUNT; ++k) s[k] = 0.0;
    for (int t = lane; t < per_img; t += 32) {
#pragma unroll
      for (int k = 0; k < FA_COUNT; ++k) s[k] += (double)base[(long)t * FA_COUNT + k];
    }
    float S[FA_COUNT];
#pragma unroll
    for (int k = 0; k < FA_COUNT; ++k) {
      double v = s[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      S[k] = (float)__shfl_sync(0xffffffffu, v, 0);
    }
    if (lane == 0) {
      float* st = p.stats + ((long)b * p.scales + l) * FA_COUNT;
#pragma unroll
      for (int k = 0; k < FA_COUNT; ++k) st[k] = S[k];
      float out[4];
      flow_level_losses(S, L.h, L.w, out);
#pragma unroll
      for (int k = 0; k < 4; ++k) lvl_loss[l][k] = out[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float t = 0.f;
    for (int l2 = 0; l2 < p.scales; ++l2) t += lvl_loss[l2][threadIdx.x];
    p.loss[threadIdx.x * p.B + b] = t;
  }
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
template <int TW, int TH, int NT>
__global__ void __launch_bounds__(NT, UGL_BMINB) flow_loss_fwdgrad_kernel(const __grid_constant__ FlowGradParams gp) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[(NT / 32) * FA_COUNT];
  using Tile = FlowGradTile<TW, TH, NT>;
  const int tile = blockIdx.x;
  const TileCoord tc = decode_tile<TW, TH>(gp.base, tile);
  float acc[FA_COUNT];
#pragma unroll
  for (int k = 0; k < FA_COUNT; ++k) acc[k] = 0.f;
  Tile::phase1(gp, tc, threadIdx.x, NT, sm, acc);
  __syncthreads();
#pragma unroll 1
  for (int dir = 0; dir < 2; ++dir) {          // rolled: one copy of the stencil phases in the instruction cache
    Tile::phase2(gp, tc, dir, threadIdx.x, NT, sm, acc);
    __syncthreads();
    Tile::phase3(gp, tc, dir, threadIdx.x, NT, sm);
    if (dir == 0) __syncthreads();   // the coefficient planes are reused by the second direction
  }
  Tile::phase4(gp, tc, threadIdx.x, NT, sm, acc);
  const float v = block_reduce_n<NT, FA_COUNT>(acc, red);
  if (threadIdx.x < FA_COUNT) gp.base.partials[(long)tile * FA_COUNT + threadIdx.x] = v;
}

// element-wise backward of the single-pass mode: grid (chunks, B, scales)
__global__ void __launch_bounds__(256) flow_combine_kernel(const __grid_constant__ FlowGradParams gp) {
  const FlowLossParams& p = gp.base;
  const int b = blockIdx.y, l = blockIdx.z;
  const FlowLevelDesc& L = p.lv[l];
  const int plane = L.h * L.w;
  const FlowCombineScales k = flow_combine_scales(p.stats + ((long)b * p.scales + l) * FA_COUNT, L.h, L.w, p.gloss, p.B, b);
  const float* basis = gp.basis[l] + (long)b * kBasisPlanes * plane;
  float* gf = L.gflow_f + (long)b * 2 * plane;
  float* gb = L.gflow_b + (long)b * 2 * plane;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < plane; pix += gridDim.x * blockDim.x)
    flow_combine_pixel(basis, plane, pix, k, gf, gb);
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

extern "C" uint64_t ugl_flow_loss_workspace_bytes(const UglFlowLossArgs* a) {
  if (!a) return 0;
  uint64_t tf = 0, tb = 0;   // only the shapes matter here; cover the forward and the single-pass tile shapes
  for (int l = 0; l < a->scales && l < UGL_MAX_LEVELS; ++l) {
    tf += (uint64_t)((a->width[l] + kFTW - 1) / kFTW) * ((a->height[l] + kFTH - 1) / kFTH) * a->batch;
    tb += (uint64_t)((a->width[l] + kBTW - 1) / kBTW) * ((a->height[l] + kBTH - 1) / kBTH) * a->batch;
  }
  return (tf > tb ? tf : tb) * FA_COUNT * sizeof(float);
}

extern "C" int ugl_flow_loss_launches(int backward) { return backward ? 1 : 2; }

extern "C" int ugl_flow_loss_forward(const UglFlowLossArgs* a) {
  FlowLossParams p;
  int rc = build_params<kFTW, kFTH>(a, false, p);
  if (rc) return rc;
  if (!a->loss) return fail(UGL_EINVAL, "flow_loss_forward: null loss");
  if (!a->workspace || a->workspace_bytes < (uint64_t)p.total_tiles * FA_COUNT * sizeof(float))
    return fail(UGL_EWORKSPACE, "flow_loss_forward: workspace too small (%llu bytes given)", (unsigned long long)a->workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  using Tile = FlowFwdTile<kFTW, kFTH>;
  constexpr size_t smem = Tile::kSmemFloats * sizeof(float);
  auto kern = flow_loss_fwd_kernel<kFTW, kFTH, kFNT>;
  static_assert(smem <= 227 * 1024, "forward tile does not fit in shared memory");
  if ((rc = opt_in_smem(kern, smem))) return rc;
  kern<<<p.total_tiles, kFNT, smem, st>>>(p);
  if ((rc = check_launch("flow_loss_fwd_kernel"))) return rc;
  flow_loss_finalize_kernel<<<p.B, 32 * kMaxLevels, 0, st>>>(p);
  return check_launch("flow_loss_finalize_kernel");
}

extern "C" int ugl_flow_loss_forward_grad(const UglFlowLossArgs* a) {
  FlowGradParams gp;
  int rc = build_params<kBTW, kBTH>(a, false, gp.base);
  if (rc) return rc;
  if (!a->loss) return fail(UGL_EINVAL, "flow_loss_forward_grad: null loss");
  for (int l = 0; l < a->scales; ++l) {
    if (!a->basis[l]) return fail(UGL_EINVAL, "flow_loss_forward_grad: null basis pointer at level %d", l);
    if (reinterpret_cast<uintptr_t>(a->basis[l]) & 7u) return fail(UGL_EALIGN, "flow_loss_forward_grad: basis not 8-byte aligned");
    gp.basis[l] = a->basis[l];
  }
  if (!a->workspace || a->workspace_bytes < (uint64_t)gp.base.total_tiles * FA_COUNT * sizeof(float))
    return fail(UGL_EWORKSPACE, "flow_loss_forward_grad: workspace too small (%llu bytes given)", (unsigned long long)a->workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  using Tile = FlowGradTile<kBTW, kBTH, kBNT>;
  constexpr size_t smem = Tile::kSmemFloats * sizeof(float);
  static_assert(smem <= 227 * 1024, "single-pass tile does not fit in shared memory");
  auto kern = flow_loss_fwdgrad_kernel<kBTW, kBTH, kBNT>;
  if ((rc = opt_in_smem(kern, smem))) return rc;
  kern<<<gp.base.total_tiles, kBNT, smem, st>>>(gp);
  if ((rc = check_launch("flow_loss_fwdgrad_kernel"))) return rc;
  flow_loss_finalize_kernel<<<gp.base.B, 32 * kMaxLevels, 0, st>>>(gp.base);
  return check_launch("flow_loss_finalize_kernel");
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
  flow_combine_kernel<<<dim3(chunks, gp.base.B, gp.base.scales), 256, 0, static_cast<cudaStream_t>(a->stream)>>>(gp);
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
