// Host-side interface of the split single-pass kernels (ugl_flow_split.cu) for the C-ABI translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "ugl_flow_grad.cuh"
#include "ugl_host.cuh"

namespace ugl {

#ifndef UGL_SPLIT_NT
#define UGL_SPLIT_NT 256
#endif
#ifndef UGL_STENCIL_MINB
#define UGL_STENCIL_MINB 4
#endif
#ifndef UGL_PHOTO_MINB
#define UGL_PHOTO_MINB 3
#endif
constexpr int kSplitNT = UGL_SPLIT_NT;                 // threads per CTA of the stencil kernel
constexpr int kStencilMinBlocks = UGL_STENCIL_MINB;    // resident CTAs per SM the register allocation aims at
constexpr int kPhotoMinBlocks = UGL_PHOTO_MINB;
#ifndef UGL_PHOTO_TW
#define UGL_PHOTO_TW 32
#endif
#ifndef UGL_PHOTO_TH
#define UGL_PHOTO_TH 32
#endif
#ifndef UGL_PHOTO_NT
#define UGL_PHOTO_NT 256
#endif
#ifndef UGL_PATCH_W
#define UGL_PATCH_W 32
#endif
constexpr int kPhotoTW = UGL_PHOTO_TW, kPhotoTH = UGL_PHOTO_TH;   // tile of a photometry-kernel CTA
constexpr int kPhotoNT = UGL_PHOTO_NT;                            // its threads: one pixel each per pass (tile height / (NT / width) passes)
constexpr int kPatchW = UGL_PATCH_W, kPatchH = 32 / UGL_PATCH_W;  // pixels a warp covers

#if defined(CUDA_VERSION) || defined(__cuda_cuda_h__)
// tensor maps of the stencil kernel's TMA copies, one set per level (kernel parameter, __grid_constant__)
struct FlowTmaMaps {
  CUtensorMap scr_halo[kMaxLevels];   // photometry pair planes (2w, h, kPhotoPairs B), box = tile + 2-pixel halo
  CUtensorMap scr_tile[kMaxLevels];   // same tensor, box = tile
  CUtensorMap img[kMaxLevels];        // centre frame (w, h, 3 B), halo box
  CUtensorMap flow_f[kMaxLevels];     // (w, h, 2 B), halo box
  CUtensorMap flow_b[kMaxLevels];
  int use_tma[kMaxLevels];            // 0: this level is staged with plain loads (strides / alignment do not allow TMA)
};
#endif

// bytes of the photometry planes for these levels (what ugl_*_workspace_bytes adds behind the tile partials)
uint64_t flow_split_scratch_bytes(const int32_t* height, const int32_t* width, int scales, int batch);
// gp.scratch[l] <- 256-byte aligned slices of `base`
void flow_split_assign_scratch(FlowGradParams& gp, void* base);
// bytes of the photometry kernel's partial-sum rows
uint64_t flow_split_photo_partials_bytes(const int32_t* height, const int32_t* width, int scales, int batch);
// photometry kernel + stencil kernel on `st`; fills gp.photo (the finalize kernel reads it).
// tma_mode: 0 = plain-load staging, 1 = TMA where the shapes allow it, 2 = TMA or fail
// parts: bit 0 photometry kernel, bit 1 weight sums (step mode), bit 2 stencil kernel (per-kernel timing launches a subset)
template <bool kGeom>
int launch_flow_split(FlowGradParams& gp, void* photo_partials, cudaStream_t st, int tma_mode, int parts = 7);

}  // namespace ugl
