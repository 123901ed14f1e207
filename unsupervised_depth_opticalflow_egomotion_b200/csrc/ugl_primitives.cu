// Primitive ops behind the reference's free functions: image pyramid and warp_flow (forward,
// d/d flow, deterministic d/d x).  See include/ugl.h for the C-ABI contract.
#include "ugl_host.cuh"
#include "ugl_primitives.cuh"
#include "ugl_scatter.cuh"

namespace ugl {

constexpr int kPrimThreads = 256;

__global__ void __launch_bounds__(kPrimThreads)
pyramid_kernel(const float* __restrict__ img, int planes, int H, int W, int l, int mode, float* __restrict__ out) {
  const int oh = H >> l, ow = W >> l;
  const long n = (long)planes * oh * ow;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
    const int oj = (int)(idx % ow);
    const long r = idx / ow;
    const int oi = (int)(r % oh);
    const long pl = r / oh;
    out[idx] = pyramid_pixel(img + pl * (long)H * W, W, l, mode, oi, oj);
  }
}

__global__ void __launch_bounds__(kPrimThreads)
warp_fwd_kernel(const float* __restrict__ x, const float* __restrict__ flow, int B, int C, int H, int W, int use_mask,
                float* __restrict__ out, float* __restrict__ mask) {
  const long n = (long)B * H * W;
  const WarpGeom geom = make_warp_geom(W, H);
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % W);
    const long r = idx / W;
    const int i = (int)(r % H), b = (int)(r / H);
    const float keep = warp_pixel_forward(x, flow, C, geom, b, i, j, use_mask, out);
    if (mask) mask[idx] = keep;
  }
}

__global__ void __launch_bounds__(kPrimThreads)
warp_bwd_flow_kernel(const float* __restrict__ x, const float* __restrict__ flow, const float* __restrict__ gout, int B,
                     int C, int H, int W, int use_mask, float* __restrict__ gflow) {
  const long n = (long)B * H * W;
  const WarpGeom geom = make_warp_geom(W, H);
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % W);
    const long r = idx / W;
    const int i = (int)(r % H), b = (int)(r / H);
    warp_pixel_backward_flow(x, flow, gout, C, geom, b, i, j, use_mask, gflow);
  }
}

__global__ void __launch_bounds__(kPrimThreads)
warp_bwd_scatter_kernel(const float* __restrict__ flow, const float* __restrict__ gout, int B, int C, int H, int W,
                        int use_mask, const unsigned* __restrict__ maxbits, unsigned long long* __restrict__ acc) {
  const long n = (long)B * H * W, plane = (long)H * W;
  const int e = fixed_point_exponent(__uint_as_float(*maxbits), plane);
  const WarpGeom geom = make_warp_geom(W, H);
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % W);
    const long r = idx / W;
    const int i = (int)(r % H), b = (int)(r / H);
    const long pix = (long)i * W + j;
    const float u = flow[((long)b * 2) * plane + pix], v = flow[((long)b * 2 + 1) * plane + pix];
    const Tap t = flow_tap(j, i, u, v, geom);
    const float keep = use_mask ? tap_keep(t) : 1.0f;
    if (keep == 0.f || t.inb == 0u) continue;
    for (int c = 0; c < C; ++c)
      scatter_tap(acc + ((long)b * C + c) * plane, W, t, gout[((long)b * C + c) * plane + pix] * keep, e);
  }
}

// EXTENSION (parity unpinned): forward splat = the `transformerFwd` that Model_flow.get_occlusion_mask_from_flow
// (model_flow.py:33-39) calls but the reference never defines.  Semantics of upstream TrianFlow: every source pixel (j,i)
// adds x[b,c,i,j] * bilinear weight to the four integer neighbours of (j+u, i+v) in pixel units; corners outside the image
// are dropped.  Deterministic (64-bit fixed-point accumulation, order independent).
__global__ void __launch_bounds__(kPrimThreads)
splat_scatter_kernel(const float* __restrict__ x, const float* __restrict__ flow, int B, int C, int H, int W,
                     const unsigned* __restrict__ maxbits, unsigned long long* __restrict__ acc) {
  const long n = (long)B * H * W, plane = (long)H * W;
  const int e = fixed_point_exponent(__uint_as_float(*maxbits), plane);
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % W);
    const long r = idx / W;
    const int i = (int)(r % H), b = (int)(r / H);
    const long pix = (long)i * W + j;
    const float tx = add_rn((float)j, flow[((long)b * 2) * plane + pix]), ty = add_rn((float)i, flow[((long)b * 2 + 1) * plane + pix]);
    const Tap t = make_tap(tx, ty, W, H);
    if (t.inb == 0u) continue;
    for (int c = 0; c < C; ++c) scatter_tap(acc + ((long)b * C + c) * plane, W, t, x[((long)b * C + c) * plane + pix], e);
  }
}

// Tile-local form of the two scatters above (ugl_scatter.cuh: ScatterWindow): a CTA takes a 32 x 8 tile of source pixels of one
// sample and a chunk of channels; four channels share one pass over the window.  kSplat: the forward splat (values x, tap at the
// pixel-unit target) instead of the warp backward (values grad_out * keep, tap of the normalised grid).  Bit-identical to the global
// form: integer sums.
constexpr int kScTW = 32, kScTH = 8, kScM = 8, kScCP = 4;
using WarpScatterWindow = ScatterWindow<kScTW, kScTH, kScM, kScCP>;

template <bool kSplat>
__global__ void __launch_bounds__(kScTW* kScTH)
scatter_tiled_kernel(const float* __restrict__ flow, const float* __restrict__ val, int B, int C, int H, int W, int use_mask,
                     const unsigned* __restrict__ maxbits, unsigned long long* __restrict__ acc, int tiles_x, int tiles_y, int cchunk) {
  __shared__ unsigned win[WarpScatterWindow::kWords];
  const long plane = (long)H * W;
  const int e = fixed_point_exponent(__uint_as_float(*maxbits), plane);
  const int tid = threadIdx.x, nt = kScTW * kScTH;
  int tile = blockIdx.x;
  const int b = tile / (tiles_x * tiles_y);
  tile -= b * tiles_x * tiles_y;
  const int x0 = (tile % tiles_x) * kScTW, y0 = (tile / tiles_x) * kScTH;
  const int wx0 = x0 - kScM, wy0 = y0 - kScM;
  const int c_begin = blockIdx.y * cchunk, c_end = min(C, c_begin + cchunk);
  const int j = x0 + (tid & (kScTW - 1)), i = y0 + tid / kScTW;
  const long pix = (long)i * W + j;
  Tap t;
  t.inb = 0u;
  float keep = 0.f;
  if (j < W && i < H) {
    const float u = flow[((long)b * 2) * plane + pix], v = flow[((long)b * 2 + 1) * plane + pix];
    if (kSplat) {
      t = make_tap(add_rn((float)j, u), add_rn((float)i, v), W, H);
      keep = 1.0f;
    } else {
      t = flow_tap(j, i, u, v, make_warp_geom(W, H));
      keep = use_mask ? tap_keep(t) : 1.0f;
    }
  }
  const bool active = keep != 0.f && t.inb != 0u;
  const bool local = WarpScatterWindow::local(t, wx0, wy0);
  WarpScatterWindow::clear(win, tid, nt);
  __syncthreads();
  for (int c0 = c_begin; c0 < c_end; c0 += kScCP) {
    const int ncp = min(kScCP, c_end - c0);
    unsigned long long* plane0 = acc + ((long)b * C + c0) * plane;
    if (active) {
      for (int k = 0; k < ncp; ++k) {
        const float g = val[((long)b * C + c0 + k) * plane + pix] * keep;
        if (local) WarpScatterWindow::add_local(win, k, wx0, wy0, t, g, e);
        else scatter_tap(plane0 + k * plane, W, t, g, e);
      }
    }
    __syncthreads();
    WarpScatterWindow::flush(win, ncp, wx0, wy0, plane0, plane, W, tid, nt);
    __syncthreads();
  }
}

// grid of the tiled scatter: tiles x channel chunks, at least ~8 CTAs per SM when the channels allow it
static void scatter_tiled_grid(int B, int C, int H, int W, dim3& grid, int& tiles_x, int& tiles_y, int& cchunk) {
  tiles_x = (W + kScTW - 1) / kScTW;
  tiles_y = (H + kScTH - 1) / kScTH;
  const long tiles = (long)B * tiles_x * tiles_y;
  long chunks = (148L * 8 + tiles - 1) / tiles;
  const long max_chunks = (C + kScCP - 1) / kScCP;
  chunks = chunks < 1 ? 1 : (chunks > max_chunks ? max_chunks : chunks);
  cchunk = (int)(((C + chunks - 1) / chunks + kScCP - 1) / kScCP * kScCP);
  grid = dim3((unsigned)tiles, (unsigned)((C + cchunk - 1) / cchunk), 1);
}

__global__ void __launch_bounds__(kPrimThreads)
fixed_to_float_clamp_kernel(const unsigned long long* __restrict__ acc, long n, long n_contrib, const unsigned* __restrict__ maxbits,
                            int clamp01, float* __restrict__ out) {
  const int e = fixed_point_exponent(__uint_as_float(*maxbits), n_contrib);
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
    float v = (float)ldexp((double)(long long)acc[idx], -e);
    if (clamp01) v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
    out[idx] = v;
  }
}

// All levels, both modes and up to three images in one launch: one thread per F x F block of a full-resolution plane
// (F = 2^(levels-1)) keeps the block in registers and emits every level's outputs that fall inside it, with the same
// per-output arithmetic as pyramid_pixel (row-major box sums / horizontal-then-vertical lerp).
struct PyramidMultiParams {
  int planes, H, W, levels, images;
  const float* img[UGL_PYRAMID_MAX_IMAGES];
  float* box[UGL_PYRAMID_MAX_IMAGES][kMaxLevels];
  float* bil[UGL_PYRAMID_MAX_IMAGES][kMaxLevels];
};

template <int F>
__global__ void __launch_bounds__(kPrimThreads) pyramid_multi_kernel(const __grid_constant__ PyramidMultiParams p) {
  const int im = blockIdx.y;
  const int ch = p.H / F, cw = p.W / F;
  const long n = (long)p.planes * ch * cw;
  const float* img = p.img[im];
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
    const int cj = (int)(idx % cw);
    const long r = idx / cw;
    const int ci = (int)(r % ch);
    const long pl = r / ch;
    const float* src = img + pl * (long)p.H * p.W + (long)ci * F * p.W + (long)cj * F;
    float v[F][F];
#pragma unroll
    for (int y = 0; y < F; ++y) {
      if (F >= 4) {
#pragma unroll
        for (int x = 0; x < F; x += 4) {
          const float4 q = *reinterpret_cast<const float4*>(src + (long)y * p.W + x);
          v[y][x] = q.x; v[y][x + 1] = q.y; v[y][x + 2] = q.z; v[y][x + 3] = q.w;
        }
      } else {
        const float2 q = *reinterpret_cast<const float2*>(src + (long)y * p.W);
        v[y][0] = q.x; v[y][1] = q.y;
      }
    }
#pragma unroll
    for (int l = 1; l < 4; ++l) {
      const int g = 1 << l;
      if (g > F) break;
      const int ow = p.W >> l;
      const long obase = pl * (long)(p.H >> l) * ow;
      float* ob = p.box[im][l];
      float* ol = p.bil[im][l];
#pragma unroll
      for (int by = 0; by < F / g; ++by)
#pragma unroll
        for (int bx = 0; bx < F / g; ++bx) {
          const long o = obase + (long)(ci * (F / g) + by) * ow + (cj * (F / g) + bx);
          if (ob) {
            float sum = 0.f;
#pragma unroll
            for (int dy = 0; dy < g; ++dy)
#pragma unroll
              for (int dx = 0; dx < g; ++dx) sum = add_rn(sum, v[by * g + dy][bx * g + dx]);
            ob[o] = div_rn(sum, (float)(g * g));
          }
          if (ol) {
            const int y = by * g + g / 2 - 1, x = bx * g + g / 2 - 1;
            const float top = add_rn(mul_rn(0.5f, v[y][x]), mul_rn(0.5f, v[y][x + 1]));
            const float bot = add_rn(mul_rn(0.5f, v[y + 1][x]), mul_rn(0.5f, v[y + 1][x + 1]));
            ol[o] = add_rn(mul_rn(0.5f, top), mul_rn(0.5f, bot));
          }
        }
    }
  }
}

static int grid_for(long n) {
  long g = (n + kPrimThreads - 1) / kPrimThreads;
  const long cap = 148L * 16;   // 148 SMs x a few resident CTAs, grid-stride beyond that
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace ugl

using namespace ugl;

extern "C" int ugl_image_pyramid(const float* img, int32_t B, int32_t C, int32_t H, int32_t W, int32_t levels,
                                 int32_t mode, float* const* out, void* stream) {
  if (!img || !out) return fail(UGL_EINVAL, "image_pyramid: null pointer");
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || levels < 1 || levels > UGL_MAX_LEVELS || (mode != 0 && mode != 1))
    return fail(UGL_EINVAL, "image_pyramid: bad arguments");
  const int f = 1 << (levels - 1);
  if (H % f || W % f) return fail(UGL_EUNSUPPORTED, "image_pyramid: %dx%d not divisible by 2^%d", H, W, levels - 1);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int l = 1; l < levels; ++l) {
    if (!out[l]) return fail(UGL_EINVAL, "image_pyramid: null output for level %d", l);
    const long n = (long)B * C * (H >> l) * (W >> l);
    pyramid_kernel<<<grid_for(n), kPrimThreads, 0, st>>>(img, B * C, H, W, l, mode, out[l]);
    int rc = check_launch("pyramid_kernel");
    if (rc) return rc;
  }
  return UGL_OK;
}

// uint8 frames -> fp32 frames in [0,1] with the reference dataset's arithmetic (kitti_prepared.py:89 `img / 255.0`, then `.float()`):
// the correctly rounded fp32 quotient (identical for all 256 byte values; tested).  16 bytes in, four 128-bit stores out per thread.
constexpr int kU8MaxFrames = 4;
struct U8Params { const unsigned char* src[kU8MaxFrames]; float* dst[kU8MaxFrames]; long n; };

__global__ void __launch_bounds__(kPrimThreads) u8_frames_kernel(const __grid_constant__ U8Params p) {
  const unsigned char* __restrict__ src = p.src[blockIdx.y];
  float* __restrict__ dst = p.dst[blockIdx.y];
  constexpr float r255 = 1.0f / 255.0f;
  const long n16 = p.n >> 4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long)gridDim.x * blockDim.x) {
    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(src) + i);
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 o;
      o.x = div_c((float)(w[k] & 0xffu), 255.0f, r255);
      o.y = div_c((float)((w[k] >> 8) & 0xffu), 255.0f, r255);
      o.z = div_c((float)((w[k] >> 16) & 0xffu), 255.0f, r255);
      o.w = div_c((float)(w[k] >> 24), 255.0f, r255);
      reinterpret_cast<float4*>(dst)[i * 4 + k] = o;
    }
  }
  if (blockIdx.x == 0)      // tail (n not a multiple of 16)
    for (long i = (n16 << 4) + threadIdx.x; i < p.n; i += blockDim.x) dst[i] = div_c((float)src[i], 255.0f, r255);
}

extern "C" int ugl_frames_u8_to_float(const void* const* src, void* const* dst, int32_t frames, uint64_t elements, void* stream) {
  if (!src || !dst) return fail(UGL_EINVAL, "frames_u8_to_float: null pointer array");
  if (frames < 1 || frames > kU8MaxFrames || elements == 0) return fail(UGL_EINVAL, "frames_u8_to_float: bad arguments");
  U8Params p;
  p.n = (long)elements;
  for (int i = 0; i < frames; ++i) {
    if (!src[i] || !dst[i]) return fail(UGL_EINVAL, "frames_u8_to_float: null frame %d", i);
    if ((reinterpret_cast<uintptr_t>(src[i]) | reinterpret_cast<uintptr_t>(dst[i])) & 15u)
      return fail(UGL_EALIGN, "frames_u8_to_float: frame %d not 16-byte aligned", i);
    p.src[i] = static_cast<const unsigned char*>(src[i]);
    p.dst[i] = static_cast<float*>(dst[i]);
  }
  const dim3 grid(grid_for((long)(elements >> 4) + 1), frames);
  u8_frames_kernel<<<grid, kPrimThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("u8_frames_kernel");
}

extern "C" int ugl_image_pyramid_multi(const UglPyramidArgs* a) {
  if (!a) return fail(UGL_EINVAL, "image_pyramid_multi: null args");
  if (a->batch <= 0 || a->channels <= 0 || a->height <= 0 || a->width <= 0 || a->images < 1 || a->images > UGL_PYRAMID_MAX_IMAGES)
    return fail(UGL_EINVAL, "image_pyramid_multi: bad arguments");
  if (a->levels < 2 || a->levels > 4) return fail(UGL_EUNSUPPORTED, "image_pyramid_multi: levels must be 2..4 (got %d)", a->levels);
  const int f = 1 << (a->levels - 1);
  if (a->height % f || a->width % f) return fail(UGL_EUNSUPPORTED, "image_pyramid_multi: %dx%d not divisible by %d", a->height, a->width, f);
  PyramidMultiParams p;
  p.planes = a->batch * a->channels; p.H = a->height; p.W = a->width; p.levels = a->levels; p.images = a->images;
  for (int i = 0; i < a->images; ++i) {
    if (!a->img[i]) return fail(UGL_EINVAL, "image_pyramid_multi: null image %d", i);
    if (reinterpret_cast<uintptr_t>(a->img[i]) & 15u) return fail(UGL_EALIGN, "image_pyramid_multi: image %d not 16-byte aligned", i);
    p.img[i] = a->img[i];
    for (int l = 0; l < kMaxLevels; ++l) {
      p.box[i][l] = (l >= 1 && l < a->levels) ? a->box[i][l] : nullptr;
      p.bil[i][l] = (l >= 1 && l < a->levels) ? a->bil[i][l] : nullptr;
    }
  }
  const long n = (long)p.planes * (p.H / f) * (p.W / f);
  const dim3 grid(grid_for(n), a->images);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  if (f == 2) pyramid_multi_kernel<2><<<grid, kPrimThreads, 0, st>>>(p);
  else if (f == 4) pyramid_multi_kernel<4><<<grid, kPrimThreads, 0, st>>>(p);
  else pyramid_multi_kernel<8><<<grid, kPrimThreads, 0, st>>>(p);
  return check_launch("pyramid_multi_kernel");
}

extern "C" int ugl_warp_flow_forward(const float* x, const float* flow, int32_t B, int32_t C, int32_t H, int32_t W,
                                     int32_t use_mask, float* out, float* mask, void* stream) {
  if (!x || !flow || !out) return fail(UGL_EINVAL, "warp_flow_forward: null pointer");
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return fail(UGL_EINVAL, "warp_flow_forward: bad shape");
  const long n = (long)B * H * W;
  warp_fwd_kernel<<<grid_for(n), kPrimThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, flow, B, C, H, W, use_mask, out, mask);
  return check_launch("warp_fwd_kernel");
}

extern "C" uint64_t ugl_warp_flow_backward_workspace_bytes(int32_t B, int32_t C, int32_t H, int32_t W, int32_t need_grad_x) {
  if (!need_grad_x) return 0;
  return (uint64_t)B * C * H * W * sizeof(unsigned long long) + 256;
}

extern "C" int ugl_warp_flow_backward(const float* x, const float* flow, const float* grad_out, int32_t B, int32_t C,
                                      int32_t H, int32_t W, int32_t use_mask, float* grad_flow, float* grad_x,
                                      void* workspace, uint64_t workspace_bytes, void* stream) {
  return ugl_warp_flow_backward_ex(x, flow, grad_out, B, C, H, W, use_mask, grad_flow, grad_x, workspace, workspace_bytes,
                                   UGL_SCATTER_TILE_LOCAL, stream);
}

extern "C" int ugl_warp_flow_backward_ex(const float* x, const float* flow, const float* grad_out, int32_t B, int32_t C,
                                         int32_t H, int32_t W, int32_t use_mask, float* grad_flow, float* grad_x,
                                         void* workspace, uint64_t workspace_bytes, int32_t scatter, void* stream) {
  if (!x || !flow || !grad_out) return fail(UGL_EINVAL, "warp_flow_backward: null pointer");
  if (scatter != UGL_SCATTER_TILE_LOCAL && scatter != UGL_SCATTER_GLOBAL) return fail(UGL_EINVAL, "warp_flow_backward: unknown scatter form %d", scatter);
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return fail(UGL_EINVAL, "warp_flow_backward: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long n = (long)B * H * W;
  int rc;
  if (grad_flow) {
    warp_bwd_flow_kernel<<<grid_for(n), kPrimThreads, 0, st>>>(x, flow, grad_out, B, C, H, W, use_mask, grad_flow);
    if ((rc = check_launch("warp_bwd_flow_kernel"))) return rc;
  }
  if (grad_x) {
    const uint64_t need = ugl_warp_flow_backward_workspace_bytes(B, C, H, W, 1);
    if (!workspace || workspace_bytes < need) return fail(UGL_EWORKSPACE, "warp_flow_backward: workspace too small");
    if (reinterpret_cast<uintptr_t>(workspace) & 7u) return fail(UGL_EALIGN, "warp_flow_backward: workspace not 8-byte aligned");
    const long nx = (long)B * C * H * W;
    unsigned long long* acc = static_cast<unsigned long long*>(workspace);
    unsigned* maxbits = reinterpret_cast<unsigned*>(acc + nx);
    cudaError_t e = cudaMemsetAsync(workspace, 0, need, st);
    if (e != cudaSuccess) return fail((int)e, "warp_flow_backward: memset: %s", cudaGetErrorString(e));
    absmax_kernel<<<grid_for(nx), kPrimThreads, 0, st>>>(grad_out, nx, maxbits);
    if ((rc = check_launch("absmax_kernel"))) return rc;
    if (scatter == UGL_SCATTER_GLOBAL) {
      warp_bwd_scatter_kernel<<<grid_for(n), kPrimThreads, 0, st>>>(flow, grad_out, B, C, H, W, use_mask, maxbits, acc);
      if ((rc = check_launch("warp_bwd_scatter_kernel"))) return rc;
    } else {
      dim3 grid;
      int tx, ty, cchunk;
      scatter_tiled_grid(B, C, H, W, grid, tx, ty, cchunk);
      scatter_tiled_kernel<false><<<grid, kScTW * kScTH, 0, st>>>(flow, grad_out, B, C, H, W, use_mask, maxbits, acc, tx, ty, cchunk);
      if ((rc = check_launch("scatter_tiled_kernel<warp backward>"))) return rc;
    }
    fixed_to_float_kernel<<<grid_for(nx), kPrimThreads, 0, st>>>(acc, nx, (long)H * W, maxbits, grad_x);
    if ((rc = check_launch("fixed_to_float_kernel"))) return rc;
  }
  return UGL_OK;
}

extern "C" uint64_t ugl_forward_splat_workspace_bytes(int32_t B, int32_t C, int32_t H, int32_t W) {
  return (uint64_t)B * C * H * W * sizeof(unsigned long long) + 256;
}

extern "C" int ugl_forward_splat(const float* x, const float* flow, int32_t B, int32_t C, int32_t H, int32_t W, int32_t clamp01,
                                 float* out, void* workspace, uint64_t workspace_bytes, void* stream) {
  return ugl_forward_splat_ex(x, flow, B, C, H, W, clamp01, out, workspace, workspace_bytes, UGL_SCATTER_TILE_LOCAL, stream);
}

extern "C" int ugl_forward_splat_ex(const float* x, const float* flow, int32_t B, int32_t C, int32_t H, int32_t W, int32_t clamp01,
                                    float* out, void* workspace, uint64_t workspace_bytes, int32_t scatter, void* stream) {
  if (!x || !flow || !out) return fail(UGL_EINVAL, "forward_splat: null pointer");
  if (scatter != UGL_SCATTER_TILE_LOCAL && scatter != UGL_SCATTER_GLOBAL) return fail(UGL_EINVAL, "forward_splat: unknown scatter form %d", scatter);
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return fail(UGL_EINVAL, "forward_splat: bad shape");
  const uint64_t need = ugl_forward_splat_workspace_bytes(B, C, H, W);
  if (!workspace || workspace_bytes < need) return fail(UGL_EWORKSPACE, "forward_splat: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 7u) return fail(UGL_EALIGN, "forward_splat: workspace not 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long nx = (long)B * C * H * W, n = (long)B * H * W;
  unsigned long long* acc = static_cast<unsigned long long*>(workspace);
  unsigned* maxbits = reinterpret_cast<unsigned*>(acc + nx);
  cudaError_t e = cudaMemsetAsync(workspace, 0, need, st);
  if (e != cudaSuccess) return fail((int)e, "forward_splat: memset: %s", cudaGetErrorString(e));
  int rc;
  absmax_kernel<<<grid_for(nx), kPrimThreads, 0, st>>>(x, nx, maxbits);
  if ((rc = check_launch("absmax_kernel"))) return rc;
  if (scatter == UGL_SCATTER_GLOBAL) {
    splat_scatter_kernel<<<grid_for(n), kPrimThreads, 0, st>>>(x, flow, B, C, H, W, maxbits, acc);
    if ((rc = check_launch("splat_scatter_kernel"))) return rc;
  } else {
    dim3 grid;
    int tx, ty, cchunk;
    scatter_tiled_grid(B, C, H, W, grid, tx, ty, cchunk);
    scatter_tiled_kernel<true><<<grid, kScTW * kScTH, 0, st>>>(flow, x, B, C, H, W, 0, maxbits, acc, tx, ty, cchunk);
    if ((rc = check_launch("scatter_tiled_kernel<splat>"))) return rc;
  }
  fixed_to_float_clamp_kernel<<<grid_for(nx), kPrimThreads, 0, st>>>(acc, nx, (long)H * W, maxbits, clamp01, out);
  return check_launch("fixed_to_float_clamp_kernel");
}
