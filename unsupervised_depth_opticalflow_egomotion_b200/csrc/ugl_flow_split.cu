// CUDA kernels + host launchers of the split single-pass flow-mode / geom-mode loss (ugl_flow_split.cuh): photometry kernel ->
// TMA-staged stencil kernel -> finalize (ugl_flow_loss.cu).  Called by ugl_flow_loss_forward_grad / ugl_geom_flow_forward_grad.
#include <cuda.h>
#include <string.h>

#include "ugl_flow_split.cuh"
#include "ugl_flow_split_host.cuh"

#ifndef UGL_STENCIL_REVERSE
#define UGL_STENCIL_REVERSE 1
#endif

namespace ugl {

// ---- TMA / mbarrier primitives (PTX; sm_90+) --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  const uint32_t a = smem_u32(bar);
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}
// generic-proxy accesses (the threads' loads / stores) before, async-proxy accesses (TMA writes) after
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// one box of a rank-3 tensor (x = innermost element, y = row, z = plane) -> shared memory; out-of-range elements arrive as zeros
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

// ---- photometry kernel: one thread per pixel ------------------------------------------------------------------------------
// Own tile grid (kPhotoTW x kPhotoTH pixels per CTA, independent of the stencil tiles).  A warp covers a kPatchW x kPatchH
// patch, not a 32 x 1 row segment: the gathers of a warp then fall into a compact window of the source frame (the flow varies
// less across 8 x 4 pixels than along 32), fewer distinct 32-byte sectors per request; the coalesced loads / stores of a
// patch row are still whole sectors (8 floats = 32 bytes, 8 pairs = 64 bytes).
// A CTA of NT threads covers its TW x TH tile in TH / PASS_H passes of TW x PASS_H pixels, top to bottom: consecutive passes gather
// from overlapping rows of the source frames (L1 reuse), the coalesced loads of pass k + 1 are in flight under the work of pass k,
// and the tile's sums are reduced once.
template <int TW, int TH, int NT, int PW_, int PH_, bool kGeom>
__global__ void __launch_bounds__(NT, kPhotoMinBlocks) flow_photo_kernel(const __grid_constant__ FlowGradParams gp) {
  constexpr int PASS_H = NT / TW;
  static_assert(NT % 32 == 0 && PW_ * PH_ == 32 && TW % PW_ == 0 && PASS_H % PH_ == 0 && TH % PASS_H == 0 && PASS_H * TW == NT,
                "whole warps, one patch per warp, whole passes");
  using Px = FlowPhotoPixel<kGeom>;
  constexpr int NA = Px::kAcc;
  __shared__ float red[(NT / 32) * NA];
  __shared__ float mats[kGeom ? 33 : 1];
  // tile decode: blockIdx.x counts this sample's photometry tiles over all levels
  int r = blockIdx.x, lv = 0;
  while (lv + 1 < gp.base.scales && r >= gp.photo.per_img[lv]) { r -= gp.photo.per_img[lv]; ++lv; }
  const int b = blockIdx.y;
  const int tile = gp.photo.tile_begin[lv] + b * gp.photo.per_img[lv] + r;
  const int tyi = r / gp.photo.tiles_x[lv], txi = r - tyi * gp.photo.tiles_x[lv];
  if (kGeom) {
    if (threadIdx.x < 9) mats[threadIdx.x] = gp.Kinv[lv][b * 9 + threadIdx.x];
    else if (threadIdx.x < 21) mats[threadIdx.x] = gp.P[0][lv][b * 12 + threadIdx.x - 9];
    else if (threadIdx.x < 33) mats[threadIdx.x] = gp.P[1][lv][b * 12 + threadIdx.x - 21];
    __syncthreads();
  }
  const FlowLevelDesc& L = gp.base.lv[lv];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int PPR = TW / PW_;                      // patches per tile row
  const int py = warp / PPR, px = warp - py * PPR;
  const int ly = lane / PW_, lx = lane - ly * PW_;
  const int i0 = tyi * TH + py * PH_ + ly, j = txi * TW + px * PW_ + lx;
  float acc[NA];
#pragma unroll
  for (int k = 0; k < NA; ++k) acc[k] = 0.f;
  DirectLoads cur = {};
  if (i0 < L.h && j < L.w) cur = Px::load(gp, lv, b, i0, j);
#pragma unroll 1
  for (int pass = 0; pass < TH / PASS_H; ++pass) {
    const int i = i0 + pass * PASS_H;
    DirectLoads nxt = cur;
    if (pass + 1 < TH / PASS_H && i + PASS_H < L.h && j < L.w) nxt = Px::load(gp, lv, b, i + PASS_H, j);
    if (i < L.h && j < L.w) Px::run(gp, lv, b, i, j, cur, acc, mats);
    cur = nxt;
  }
  const float v = block_reduce_n<NT, NA>(acc, red);
  if (threadIdx.x < NA) gp.photo.partials[(long)tile * NA + threadIdx.x] = v;
}

// ---- stencil kernel -------------------------------------------------------------------------------------------------------
// ---- step mode: the photometry kernel's weight sums -> stats columns, before the stencil kernel needs the normalisers ----------
// Same summation order as flow_loss_finalize_kernel (ugl_flow_loss.cu), which later rewrites the same columns with the same bits.
template <bool kGeom>
__global__ void __launch_bounds__(256) flow_photo_norm_kernel(const __grid_constant__ FlowGradParams gp) {
  using Px = FlowPhotoPixel<kGeom>;
  constexpr int NA = Px::kAcc;
  constexpr int ROW = kGeom ? (int)GA_COUNT : (int)FA_COUNT;
  __shared__ double red[256 / 32][NA];
  const int l = blockIdx.x, b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  griddep_wait();                 // the photometry kernel is complete
  griddep_launch_dependents();    // the stencil kernel may start: it needs this kernel's output only for its last phase
  const float* pb = gp.photo.partials + ((long)gp.photo.tile_begin[l] + (long)b * gp.photo.per_img[l]) * NA;
  double s[NA];
#pragma unroll
  for (int k = 0; k < NA; ++k) s[k] = 0.0;
  for (int t = threadIdx.x; t < gp.photo.per_img[l]; t += 256) {
#pragma unroll
    for (int k = 0; k < NA; ++k) s[k] += (double)pb[(long)t * NA + k];
  }
#pragma unroll
  for (int k = 0; k < NA; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_down_sync(0xffffffffu, s[k], o);
    if (lane == 0) red[warp][k] = s[k];
  }
  __syncthreads();
  if (threadIdx.x < NA) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 256 / 32; ++w) v += red[w][threadIdx.x];
    gp.base.stats[((long)b * gp.base.scales + l) * ROW + Px::column(threadIdx.x)] = (float)v;
  }
  if (gp.step_scales) {   // the closing divisions once per (sample, level) instead of once per stencil thread; layout: FlowStencilTile::phase4b_step
    __syncthreads();
    if (threadIdx.x == 0) {
      const FlowLevelDesc& L = gp.base.lv[l];
      const float* S = gp.base.stats + ((long)b * gp.base.scales + l) * ROW;
      float* o = gp.step_scales + ((long)b * gp.base.scales + l) * 8;
      if (kGeom) {
        const GeomCombineScales k = geom_combine_scales(S, L.h, L.w, gp.base.gloss, gp.base.B, b);
        o[0] = k.pix_r[0]; o[1] = k.pix_r[1]; o[2] = k.pix_d[0]; o[3] = k.pix_d[1]; o[4] = k.ssim[0]; o[5] = k.ssim[1]; o[6] = k.sm; o[7] = k.cons;
      } else {
        const FlowCombineScales k = flow_combine_scales(S, L.h, L.w, gp.base.gloss, gp.base.B, b);
        o[0] = k.pix[0]; o[1] = k.pix[1]; o[2] = k.pix[0]; o[3] = k.pix[1]; o[4] = k.ssim[0]; o[5] = k.ssim[1]; o[6] = k.sm; o[7] = k.cons;
      }
    }
  }
}

template <int TW, int TH, int NT, bool kGeom, bool kStep>
__global__ void __launch_bounds__(NT, kStencilMinBlocks)
flow_stencil_kernel(const __grid_constant__ FlowGradParams gp, const __grid_constant__ FlowTmaMaps tm) {
  using Tile = FlowStencilTile<TW, TH, NT, kGeom>;
  constexpr int NA = Tile::kAcc;
  constexpr int ROW = kGeom ? (int)GA_COUNT : (int)FA_COUNT;
  extern __shared__ float sm_raw[];
  // TMA destinations must be 128-byte aligned.  The offset is added to the array (not to an integer) so the compiler keeps
  // treating `sm` as a shared-memory pointer (LDS / STS instead of generic LD / ST).
  float* sm = sm_raw + (((128u - (smem_u32(sm_raw) & 127u)) & 127u) >> 2);
  __shared__ float red[(NT / 32) * NA];
  __shared__ __align__(8) uint64_t bars[4];
  int tile;
#if UGL_STENCIL_REVERSE
  // last sample / last tiles first: what the photometry kernel wrote last is what the L2 still holds
  const TileCoord tc = decode_tile_2d<TW, TH>(gp.base, gridDim.x - 1 - blockIdx.x, gridDim.y - 1 - blockIdx.y, tile);
#else
  const TileCoord tc = decode_tile_2d<TW, TH>(gp.base, blockIdx.x, blockIdx.y, tile);
#endif
  const int tid = threadIdx.x;
  const bool kTma = tm.use_tma[tc.level] != 0;   // per level (CTA-uniform): TMA staging where the level's strides allow it
  // Step mode: launched as a programmatic dependent of the weight-sum kernel, which releases it as soon as it has itself seen the
  // photometry kernel complete: the photometry planes are final, the scale factors are not (waited for before phase 4b).
  if (!kStep) griddep_wait();

  // copy group g (see FlowStencilTile::load_group_plain): issued by one thread, completion counted in bytes on bars[g]
  auto issue = [&](int g) {
    if (kTma) {
      if (tid == 0) {
        constexpr uint32_t kHaloPair = 2 * Tile::SPN * 4, kHaloScal = Tile::SPN * 4, kTilePair = 2 * Tile::TN * 4;
        fence_proxy_async();
        const int y0 = tc.y0 - Tile::R;
        const int xs = tc.x0 - Tile::R - Tile::kScalX;           // halo planes start at column x0 - 4 (16-byte aligned rows of the box)
        const int zs = tc.b * kPhotoPairs;
        uint64_t* bar = &bars[g];
        if (g == 0) {
          mbar_expect_tx(bar, 2 * kHaloPair + 3 * kHaloScal + 2 * kTilePair);
          for (int c = 0; c < 3; ++c) tma_load_3d(sm + Tile::kOffI + c * Tile::kScalP, &tm.img[tc.level], xs, y0, tc.b * 3 + c, bar);
        } else if (g < 3) {
          mbar_expect_tx(bar, 2 * kHaloPair + 2 * kTilePair);
        } else {
          mbar_expect_tx(bar, 4 * kHaloScal);
        }
        if (g < 3) {
          float* st = Tile::stage(sm, g);
          tma_load_3d(st, &tm.scr_halo[tc.level], 2 * xs, y0, zs + PP_X0 + g, bar);
          tma_load_3d(st + Tile::kSlotY, &tm.scr_halo[tc.level], 2 * xs, y0, zs + PP_Y0 + g, bar);
          tma_load_3d(st + Tile::kSlotDu, &tm.scr_tile[tc.level], 2 * tc.x0, tc.y0, zs + PP_DW0 + 2 * g, bar);
          tma_load_3d(st + Tile::kSlotDv, &tm.scr_tile[tc.level], 2 * tc.x0, tc.y0, zs + PP_DW0 + 2 * g + 1, bar);
        } else {
          float* raw = sm + Tile::kOffRawFlow;
          tma_load_3d(raw, &tm.flow_f[tc.level], xs, y0, tc.b * 2, bar);
          tma_load_3d(raw + Tile::kScalP, &tm.flow_f[tc.level], xs, y0, tc.b * 2 + 1, bar);
          tma_load_3d(raw + 2 * Tile::kScalP, &tm.flow_b[tc.level], xs, y0, tc.b * 2, bar);
          tma_load_3d(raw + 3 * Tile::kScalP, &tm.flow_b[tc.level], xs, y0, tc.b * 2 + 1, bar);
        }
      }
    } else {
      Tile::load_group_plain(gp, tc, g, tid, NT, sm);
    }
  };
  auto arrived = [&](int g) {
    if (kTma) mbar_wait(&bars[g], 0);   // every barrier is used for exactly one phase
  };

  if (kTma) {
    if (tid == 0) {
#pragma unroll
      for (int g = 0; g < 4; ++g) mbar_init(&bars[g], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  issue(0);
  issue(1);
  float acc[NA];
#pragma unroll
  for (int k = 0; k < NA; ++k) acc[k] = 0.f;
  float2 g3[Tile::kP3][4];
#pragma unroll
  for (int n = 0; n < Tile::kP3; ++n)
#pragma unroll
    for (int k = 0; k < 4; ++k) g3[n][k] = make_float2(0.f, 0.f);
  unsigned geo2[Tile::kP2], geo3[Tile::kP3];
  Tile::strip_geometry(gp, tc, tid, NT, geo2, geo3);
  if (!kTma) __syncthreads();
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    arrived(c);
    Tile::phase2(gp, tc, c, tid, NT, sm, acc, geo2);
    __syncthreads();
    Tile::phase3_accumulate(gp, tc, c, tid, NT, sm, g3, geo3);
    __syncthreads();                   // ring slot c & 1, the x plane and the coefficient planes are free again
    if (c < 2) issue(c + 2);           // channel 2 -> slot 0; the raw flow planes -> slot 1
  }
  float4 pre[kStep ? Tile::kP3 : 1][3];
  if (kStep) Tile::prefetch_step(gp, tc, tid, NT, pre);   // in flight under the smoothness phases
  else Tile::phase3_store(gp, tc, tid, NT, g3);
  arrived(3);
  Tile::convert_flows(tid, NT, sm);
  __syncthreads();
  Tile::phase2(gp, tc, 3, tid, NT, sm, acc, geo2);   // smoothness edge weights, over the raw flow planes (converted above)
  __syncthreads();
  Tile::phase4a(gp, tc, tid, NT, sm, acc);
  __syncthreads();
  if (kStep) {
    griddep_wait();                  // the weight-sum kernel is complete: scale factors final
    const float4* ks = reinterpret_cast<const float4*>(gp.step_scales + ((long)tc.b * gp.base.scales + tc.level) * 8);
    const float4 k0 = __ldcg(ks), k1 = __ldcg(ks + 1);
    const float k[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
    Tile::phase4b_step(gp, tc, k, tid, NT, sm, g3, pre);
  } else {
    Tile::phase4b(gp, tc, tid, NT, sm);
  }
  const float v = block_reduce_n<NT, NA>(acc, red);
  if (tid < NA) gp.base.partials[(long)tile * ROW + Tile::column(tid)] = v;
}

// ---- host side ----------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  // resolved through the runtime (no link-time dependency on libcuda); the pointer is process-wide and immutable once set
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// rank-3 fp32 tensor (inner, rows, planes) with box (box_inner, box_rows, 1); zero fill outside
static bool encode3(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint64_t planes, uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t dims[3] = {inner, rows, planes};
  const cuuint64_t strides[2] = {inner * 4, inner * rows * 4};
  const cuuint32_t box[3] = {box_inner, box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// TMA needs 16-byte aligned bases, 16-byte multiples for every global stride (scalar planes -> width % 4 == 0) and a box that starts on
// a 16-byte boundary of its row (probed on B200: a start column of x0 - 2 floats raises `illegal instruction`; scratch/tma_probe.cu).
// Decided per level; returns the number of levels staged by TMA.
template <int TW, int TH, int NT, bool kGeom>
static int build_tma_maps(const FlowGradParams& gp, FlowTmaMaps& tm) {
  using Tile = FlowStencilTile<TW, TH, NT, kGeom>;
  static_assert(2 * Tile::SPW <= 256 && Tile::PH <= 256, "TMA box dimensions are limited to 256 elements");
  const FlowLossParams& p = gp.base;
  int n = 0;
  for (int l = 0; l < p.scales; ++l) {
    const FlowLevelDesc& L = p.lv[l];
    tm.use_tma[l] = 0;
    if ((L.w & 3) != 0) continue;
    if (!aligned16(gp.scratch[l]) || !aligned16(L.img) || !aligned16(L.flow_f) || !aligned16(L.flow_b)) continue;
    const uint64_t w = (uint64_t)L.w, h = (uint64_t)L.h, B = (uint64_t)p.B;
    if (!encode3(&tm.scr_halo[l], gp.scratch[l], 2 * w, h, kPhotoPairs * B, 2 * Tile::SPW, Tile::PH)) continue;
    if (!encode3(&tm.scr_tile[l], gp.scratch[l], 2 * w, h, kPhotoPairs * B, 2 * TW, TH)) continue;
    if (!encode3(&tm.img[l], L.img, w, h, 3 * B, Tile::SPW, Tile::PH)) continue;
    if (!encode3(&tm.flow_f[l], L.flow_f, w, h, 2 * B, Tile::SPW, Tile::PH)) continue;
    if (!encode3(&tm.flow_b[l], L.flow_b, w, h, 2 * B, Tile::SPW, Tile::PH)) continue;
    tm.use_tma[l] = 1;
    ++n;
  }
  return n;
}

uint64_t flow_split_scratch_bytes(const int32_t* height, const int32_t* width, int scales, int batch) {
  uint64_t n = 0;
  for (int l = 0; l < scales && l < kMaxLevels; ++l) {
    n += ((uint64_t)batch * kPhotoFloats * height[l] * width[l] * sizeof(float) + 255) & ~(uint64_t)255;
  }
  return n + 256;
}

void flow_split_assign_scratch(FlowGradParams& gp, void* base) {
  uintptr_t p = (reinterpret_cast<uintptr_t>(base) + 255) & ~(uintptr_t)255;
  for (int l = 0; l < gp.base.scales; ++l) {
    gp.scratch[l] = reinterpret_cast<float*>(p);
    p += ((uint64_t)gp.base.B * kPhotoFloats * gp.base.lv[l].h * gp.base.lv[l].w * sizeof(float) + 255) & ~(uint64_t)255;
  }
}

uint64_t flow_split_photo_partials_bytes(const int32_t* height, const int32_t* width, int scales, int batch) {
  uint64_t tiles = 0;
  for (int l = 0; l < scales && l < kMaxLevels; ++l)
    tiles += (uint64_t)((width[l] + kPhotoTW - 1) / kPhotoTW) * ((height[l] + kPhotoTH - 1) / kPhotoTH) * batch;
  // + [B][scales][8] step-mode scale factors in front of the rows
  return (((uint64_t)batch * scales * 8 * sizeof(float) + 255) & ~(uint64_t)255) + (((tiles * 10 * sizeof(float)) + 255) & ~(uint64_t)255);
}

// fills gp.photo (tiling + partial rows at `partials`)
template <bool kGeom>
static void assign_photo_tiling(FlowGradParams& gp, void* partials) {
  int tiles = 0, per_sample = 0;
  for (int l = 0; l < gp.base.scales; ++l) {
    const FlowLevelDesc& L = gp.base.lv[l];
    const int tx = (L.w + kPhotoTW - 1) / kPhotoTW, ty = (L.h + kPhotoTH - 1) / kPhotoTH;
    gp.photo.tiles_x[l] = tx;
    gp.photo.per_img[l] = tx * ty;
    gp.photo.tile_begin[l] = tiles;
    tiles += tx * ty * gp.base.B;
    per_sample += tx * ty;
  }
  gp.photo.per_sample = per_sample;
  gp.photo.nacc = FlowPhotoPixel<kGeom>::kAcc;
  gp.step_scales = static_cast<float*>(partials);
  gp.photo.partials = reinterpret_cast<float*>(static_cast<char*>(partials) + (((uint64_t)gp.base.B * gp.base.scales * 8 * sizeof(float) + 255) & ~(uint64_t)255));
}

template <bool kGeom>
int launch_flow_split(FlowGradParams& gp, void* photo_partials, cudaStream_t st, int tma_mode, int parts) {
  constexpr int TW = kBTW, TH = kBTH, NT = kSplitNT;
  using Tile = FlowStencilTile<TW, TH, NT, kGeom>;
  assign_photo_tiling<kGeom>(gp, photo_partials);
  const dim3 grid(gp.base.total_tiles / gp.base.B, gp.base.B);
  int rc = UGL_OK;
  // programmatic dependent launches between consecutive kernels of THIS call (a kernel launched alone orders by the stream as usual)
  const bool pdl_norm = (parts & 1) && (parts & 2);
  const bool pdl_stencil = gp.step ? ((parts & 2) && (parts & 4)) : ((parts & 1) && (parts & 4));
  if (parts & 1) {
    if ((rc = launch_kernel("flow_photo_kernel", flow_photo_kernel<kPhotoTW, kPhotoTH, kPhotoNT, kPatchW, kPatchH, kGeom>,
                            dim3(gp.photo.per_sample, gp.base.B), dim3(kPhotoNT), 0, st, false, gp))) return rc;
  }
  constexpr size_t smem = Tile::kSmemFloats * sizeof(float) + 128;
  static_assert(smem <= 227 * 1024, "stencil tile does not fit in shared memory");
  if (gp.step && (parts & 2)) {
    if ((rc = launch_kernel("flow_photo_norm_kernel", flow_photo_norm_kernel<kGeom>, dim3(gp.base.scales, gp.base.B), dim3(256), 0, st, pdl_norm, gp))) return rc;
  }
  if (!(parts & 4)) return UGL_OK;
  FlowTmaMaps tm;
  memset(&tm, 0, sizeof(tm));
  const int n_tma = tma_mode != 0 ? build_tma_maps<TW, TH, NT, kGeom>(gp, tm) : 0;
  if (tma_mode == 2 && n_tma != gp.base.scales)
    return fail(UGL_EUNSUPPORTED, "flow_loss: TMA staging requested but only %d of %d levels allow it (width %% 4, 16-byte alignment, driver entry point)",
                n_tma, gp.base.scales);
  if (gp.step) {
    auto kern = flow_stencil_kernel<TW, TH, NT, kGeom, true>;
    if ((rc = opt_in_smem(kern, smem))) return rc;
    return launch_kernel("flow_stencil_kernel", kern, grid, dim3(NT), smem, st, pdl_stencil, gp, tm);
  }
  auto kern = flow_stencil_kernel<TW, TH, NT, kGeom, false>;
  if ((rc = opt_in_smem(kern, smem))) return rc;
  return launch_kernel("flow_stencil_kernel", kern, grid, dim3(NT), smem, st, pdl_stencil, gp, tm);
}

template int launch_flow_split<false>(FlowGradParams&, void*, cudaStream_t, int, int);
template int launch_flow_split<true>(FlowGradParams&, void*, cudaStream_t, int, int);

}  // namespace ugl
