// Single-pass edge-aware disparity smoothness (compute_smooth_loss: model_geometry.py:225-252, model_depth.py:220-247)
// for up to three (image, disparity pyramid) lists at once — the three calls of model_geometry.py:938-940.
//
//   forward_grad  one tiled kernel: the image tile and its edge weights exp(-mean_c |dI|) are formed ONCE in shared memory
//                 and shared by every level; per level the bilinearly up-sampled disparity tile (halo 1) is formed once,
//                 the |first difference| * weight sums are accumulated, and G_l = d loss / d up_l(Y,X) (un-normalised:
//                 without the upstream gradient) is written at full resolution.  + the fixed-order fp64 finalize.
//   combine       grad_disp_l = grad_out[b] * (bilinear up-sampling)^T G_l, gather form (deterministic).
//
// The per-method kernels in ugl_terms.cu (ugl_disp_smooth_forward / _backward) recompute instead of saving G.
#include "ugl_common.cuh"
#include "ugl_reduce.cuh"

namespace ugl {

constexpr int kDsTW = 32, kDsTH = 16, kDsNT = 256;
constexpr int kDsPW = kDsTW + 2, kDsPH = kDsTH + 2, kDsPN = kDsPW * kDsPH;
constexpr int kDsLists = UGL_DISP_SMOOTH_MAX_LISTS;
constexpr int kDsPatch = (kDsPW / 2 + 3) * (kDsPH / 2 + 3);   // low-resolution patch of a >= 2x level: 20 x 12

struct DsParams {
  int B, lists, levels, H, W, tiles_x, tiles_y;
  int h[kMaxLevels], w[kMaxLevels];
  const float* img[kDsLists];
  const float* disp[kDsLists][kMaxLevels];
  float* G[kDsLists][kMaxLevels];
  float* partials;              // [lists*B][tiles][2]
  const float* gout;            // (lists,B)
  float* gdisp[kDsLists][kMaxLevels];
};

struct DsTap { int i0, i1; float l0, l1; };
// ATen upsample_bilinear2d, align_corners=False: src = scale*(dst+0.5)-0.5 clamped at 0
__device__ __forceinline__ DsTap ds_tap(int dst, int n_in, float scale) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  DsTap t;
  t.i0 = (int)src;
  t.i1 = t.i0 + (t.i0 < n_in - 1 ? 1 : 0);
  t.l1 = src - (float)t.i0;
  t.l0 = 1.0f - t.l1;
  return t;
}

__global__ void __launch_bounds__(kDsNT) disp_smooth_fwdgrad_kernel(const __grid_constant__ DsParams p) {
  __shared__ float sI[3][kDsPN], sWx[kDsPN], sWy[kDsPN], sU[kDsPN];
  __shared__ float red[(kDsNT / 32) * 2];
  __shared__ int sXi[2][kDsPW], sYo[2][kDsPH];        // column indices / row offsets of the current level's up-sampling taps
  __shared__ float sXl[2][kDsPW], sYl[2][kDsPH];      // and their weights
  __shared__ float sD[kDsPatch];                      // the low-resolution disparity patch under the tile
  const int tile = blockIdx.x, b = blockIdx.y, li = blockIdx.z;
  const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
  const int x0 = tx * kDsTW, y0 = ty * kDsTH, H = p.H, W = p.W;
  const long plane = (long)H * W;
  const float* img = p.img[li] + (long)b * 3 * plane;
  // image tile (halo 1); out-of-image positions are never used
  for (int idx = threadIdx.x; idx < kDsPN; idx += kDsNT) {
    const int ly = idx / kDsPW, lx = idx - ly * kDsPW;
    const int Y = y0 - 1 + ly, X = x0 - 1 + lx;
    const bool in = (Y >= 0 && Y < H && X >= 0 && X < W);
    const long o = (long)Y * W + X;
#pragma unroll
    for (int c = 0; c < 3; ++c) sI[c][idx] = in ? img[c * plane + o] : 0.f;
  }
  __syncthreads();
  // edge weights: sWx[idx] between (Y,X) and (Y,X+1), sWy[idx] between (Y,X) and (Y+1,X)
  for (int idx = threadIdx.x; idx < kDsPN; idx += kDsNT) {
    const int ly = idx / kDsPW, lx = idx - ly * kDsPW;
    const int Y = y0 - 1 + ly, X = x0 - 1 + lx;
    float wx = 0.f, wy = 0.f;
    if (Y >= 0 && Y < H && X >= 0 && X < W) {
      const float r3 = 1.0f / 3.0f;
      if (X + 1 < W && lx + 1 < kDsPW) {
        const float s = add_rn(add_rn(fabsf(sub_rn(sI[0][idx], sI[0][idx + 1])), fabsf(sub_rn(sI[1][idx], sI[1][idx + 1]))),
                               fabsf(sub_rn(sI[2][idx], sI[2][idx + 1])));
        wx = expf(-div_c(s, 3.0f, r3));
      }
      if (Y + 1 < H && ly + 1 < kDsPH) {
        const float s = add_rn(add_rn(fabsf(sub_rn(sI[0][idx], sI[0][idx + kDsPW])), fabsf(sub_rn(sI[1][idx], sI[1][idx + kDsPW]))),
                               fabsf(sub_rn(sI[2][idx], sI[2][idx + kDsPW])));
        wy = expf(-div_c(s, 3.0f, r3));
      }
    }
    sWx[idx] = wx; sWy[idx] = wy;
  }
  float acc[2] = {0.f, 0.f};
  const float inx = 1.0f / ((float)H * (float)(W - 1)), iny = 1.0f / ((float)(H - 1) * (float)W);
  for (int l = 0; l < p.levels; ++l) {
    const int h = p.h[l], w = p.w[l];
    const float* d = p.disp[li][l] + (long)b * h * w;
    const bool full = (h == H && w == W);
    const float sy = (float)h / (float)H, sx = (float)w / (float)W;
    // the up-sampling taps are separable: the 34 column taps and 18 row taps of the tile are formed once per level (the previous
    // level's taps are dead since its mid barrier) instead of twice per halo pixel.  The low-resolution patch they address (at most
    // 19 x 11 values for a 2x level) is staged in shared memory with coalesced loads, so the four taps of a pixel are shared-memory
    // reads instead of four scattered global loads (the up-sampling line carried most of the kernel's long-scoreboard stalls).
    int cx0 = 0, ry0 = 0, pw = 0, ph = 0;
    if (!full) {
      const int Xa = x0 - 1 < 0 ? 0 : x0 - 1, Xb = x0 - 2 + kDsPW >= W ? W - 1 : x0 - 2 + kDsPW;
      const int Ya = y0 - 1 < 0 ? 0 : y0 - 1, Yb = y0 - 2 + kDsPH >= H ? H - 1 : y0 - 2 + kDsPH;
      cx0 = ds_tap(Xa, w, sx).i0; pw = ds_tap(Xb, w, sx).i1 - cx0 + 1;
      ry0 = ds_tap(Ya, h, sy).i0; ph = ds_tap(Yb, h, sy).i1 - ry0 + 1;
    }
    const bool patch = !full && pw * ph <= kDsPatch;          // uniform over the CTA
    if (!full) {
      if (threadIdx.x < kDsPW) {
        const int X = x0 - 1 + (int)threadIdx.x;
        const DsTap t = ds_tap(X < 0 ? 0 : (X >= W ? W - 1 : X), w, sx);
        sXi[0][threadIdx.x] = t.i0 - (patch ? cx0 : 0); sXi[1][threadIdx.x] = t.i1 - (patch ? cx0 : 0);
        sXl[0][threadIdx.x] = t.l0; sXl[1][threadIdx.x] = t.l1;
      } else if (threadIdx.x >= 64 && threadIdx.x < 64 + kDsPH) {
        const int r = (int)threadIdx.x - 64, Y = y0 - 1 + r;
        const DsTap t = ds_tap(Y < 0 ? 0 : (Y >= H ? H - 1 : Y), h, sy);
        sYo[0][r] = patch ? (t.i0 - ry0) * pw : t.i0 * w; sYo[1][r] = patch ? (t.i1 - ry0) * pw : t.i1 * w;
        sYl[0][r] = t.l0; sYl[1][r] = t.l1;
      }
    }
    __syncthreads();                       // weights ready (first level) / previous level's sU and patch consumed / taps ready
    if (patch) {
      for (int k = threadIdx.x; k < pw * ph; k += kDsNT) {
        const int r = k / pw, c = k - r * pw;
        sD[k] = d[(long)(ry0 + r) * w + cx0 + c];
      }
      __syncthreads();
    }
    const float* src = patch ? sD : d;
    for (int idx = threadIdx.x; idx < kDsPN; idx += kDsNT) {
      const int ly = idx / kDsPW, lx = idx - ly * kDsPW;
      const int Y = y0 - 1 + ly, X = x0 - 1 + lx;
      float v = 0.f;
      if (Y >= 0 && Y < H && X >= 0 && X < W) {
        if (full) {
          v = d[(long)Y * W + X];
        } else {
          const float* r0 = src + sYo[0][ly];
          const float* r1 = src + sYo[1][ly];
          const int i0 = sXi[0][lx], i1 = sXi[1][lx];
          const float xl0 = sXl[0][lx], xl1 = sXl[1][lx];
          v = sYl[0][ly] * (xl0 * r0[i0] + xl1 * r0[i1]) + sYl[1][ly] * (xl0 * r1[i0] + xl1 * r1[i1]);
        }
      }
      sU[idx] = v;
    }
    __syncthreads();
    float* G = p.G[li][l] ? p.G[li][l] + (long)b * plane : nullptr;
    for (int t = threadIdx.x; t < kDsTW * kDsTH; t += kDsNT) {
      const int ly = t / kDsTW + 1, lx = t % kDsTW + 1;
      const int Y = y0 - 1 + ly, X = x0 - 1 + lx;
      if (Y >= H || X >= W) continue;
      const int idx = ly * kDsPW + lx;
      const float c = sU[idx];
      float gx = 0.f, gy = 0.f;
      if (X <= W - 2) { const float a = c - sU[idx + 1]; acc[0] += fabsf(a) * sWx[idx]; gx += sWx[idx] * sgnf(a); }
      if (X >= 1) gx -= sWx[idx - 1] * sgnf(sU[idx - 1] - c);
      if (Y <= H - 2) { const float a = c - sU[idx + kDsPW]; acc[1] += fabsf(a) * sWy[idx]; gy += sWy[idx] * sgnf(a); }
      if (Y >= 1) gy -= sWy[idx - kDsPW] * sgnf(sU[idx - kDsPW] - c);
      if (G) G[(long)Y * W + X] = gx * inx + gy * iny;
    }
  }
  __syncthreads();
  const float v = block_reduce_n<kDsNT, 2>(acc, red);
  if (threadIdx.x < 2) p.partials[(((long)li * p.B + b) * gridDim.x + tile) * 2 + threadIdx.x] = v;
}

struct DsFinal {
  float* out;
  float nx, ny;
  __device__ void operator()(int s, const double* S) const { out[s] = (float)(S[0] / nx) + (float)(S[1] / ny); }
};

// transpose of the bilinear up-sampling by an integer factor, gather form: low-res pixel (y, x) collects G over the <= 2F + 2 full-res
// rows / columns whose taps touch it.  The per-row and per-column weights are separable and are formed ONCE per thread (the first
// version re-derived both taps for every one of the (2F+2)^2 elements: 80 us -> 25 us per geom step).
template <int F>
__device__ __forceinline__ float ds_transpose_gather(const float* __restrict__ G, int y, int x, int h, int w, int H, int W) {
  constexpr int K = 2 * F + 2;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const int Y0 = F * y - F / 2 - 1, X0 = F * x - F / 2 - 1;
  float wy[K], wx[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int Y = Y0 + k, X = X0 + k;
    wy[k] = 0.f; wx[k] = 0.f;
    if (Y >= 0 && Y < H) { const DsTap t = ds_tap(Y, h, sy); wy[k] = (t.i0 == y ? t.l0 : 0.f) + (t.i1 == y ? t.l1 : 0.f); }
    if (X >= 0 && X < W) { const DsTap t = ds_tap(X, w, sx); wx[k] = (t.i0 == x ? t.l0 : 0.f) + (t.i1 == x ? t.l1 : 0.f); }
  }
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < K; ++a) {
    if (wy[a] == 0.f) continue;                       // also skips rows outside the image (weight 0 by construction)
    const float* row = G + (long)(Y0 + a) * W + X0;
    float r = 0.f;
#pragma unroll
    for (int c = 0; c < K; ++c)
      if (wx[c] != 0.f) r += wx[c] * row[c];
    acc += wy[a] * r;
  }
  return acc;
}

// grid (chunks, lists*B, levels)
__global__ void __launch_bounds__(256) disp_smooth_combine_kernel(const __grid_constant__ DsParams p) {
  const int l = blockIdx.z, li = blockIdx.y / p.B, b = blockIdx.y - li * p.B;
  const int h = p.h[l], w = p.w[l], H = p.H, W = p.W;
  const float* G = p.G[li][l] + (long)b * H * W;
  float* gd = p.gdisp[li][l] + (long)b * h * w;
  const float g = p.gout[li * p.B + b];
  const int n = h * w;
  if (h == H && w == W) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) gd[i] = g * G[i];
    return;
  }
  const int fy = H / h, fx = W / w;
  if (fy == fx && (fy == 2 || fy == 4 || fy == 8)) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const int y = i / w, x = i - y * w;
      const float acc = fy == 2 ? ds_transpose_gather<2>(G, y, x, h, w, H, W)
                                : (fy == 4 ? ds_transpose_gather<4>(G, y, x, h, w, H, W) : ds_transpose_gather<8>(G, y, x, h, w, H, W));
      gd[i] = g * acc;
    }
    return;
  }
  // generic integer factors
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int y = i / w, x = i - y * w;
    const int Y0 = max(0, fy * y - fy / 2 - 1), Y1 = min(H, fy * y + (3 * fy) / 2 + 1);
    const int X0 = max(0, fx * x - fx / 2 - 1), X1 = min(W, fx * x + (3 * fx) / 2 + 1);
    float acc = 0.f;
    for (int Y = Y0; Y < Y1; ++Y) {
      const DsTap ty = ds_tap(Y, h, sy);
      const float wy = (ty.i0 == y ? ty.l0 : 0.f) + (ty.i1 == y ? ty.l1 : 0.f);
      if (wy == 0.f) continue;
      float row = 0.f;
      for (int X = X0; X < X1; ++X) {
        const DsTap tx = ds_tap(X, w, sx);
        const float wx = (tx.i0 == x ? tx.l0 : 0.f) + (tx.i1 == x ? tx.l1 : 0.f);
        row += wx * G[(long)Y * W + X];
      }
      acc += wy * row;
    }
    gd[i] = g * acc;
  }
}

static int ds_fill(const UglDispSmoothArgs* a, DsParams& p) {
  if (!a) return fail(UGL_EINVAL, "disp_smooth: null args");
  if (a->batch <= 0 || a->lists <= 0 || a->lists > kDsLists || a->levels <= 0 || a->levels > UGL_MAX_LEVELS || a->height < 2 || a->width < 2)
    return fail(UGL_EINVAL, "disp_smooth: bad batch/lists/levels/size (%d/%d/%d/%dx%d)", a->batch, a->lists, a->levels, a->height, a->width);
  if ((long)a->batch * a->lists > 65535) return fail(UGL_EUNSUPPORTED, "disp_smooth: batch * lists > 65535");
  p.B = a->batch; p.lists = a->lists; p.levels = a->levels; p.H = a->height; p.W = a->width;
  p.tiles_x = (p.W + kDsTW - 1) / kDsTW; p.tiles_y = (p.H + kDsTH - 1) / kDsTH;
  for (int l = 0; l < a->levels; ++l) {
    p.h[l] = a->lheight[l]; p.w[l] = a->lwidth[l];
    if (p.h[l] <= 0 || p.w[l] <= 0 || p.H % p.h[l] || p.W % p.w[l])
      return fail(UGL_EUNSUPPORTED, "disp_smooth: level %d (%dx%d) does not divide %dx%d", l, p.h[l], p.w[l], p.H, p.W);
  }
  return UGL_OK;
}

}  // namespace ugl

using namespace ugl;

extern "C" uint64_t ugl_disp_smooth_fused_workspace_bytes(const UglDispSmoothArgs* a) {
  if (!a) return 0;
  const uint64_t tiles = (uint64_t)((a->width + kDsTW - 1) / kDsTW) * ((a->height + kDsTH - 1) / kDsTH);
  return (uint64_t)a->batch * a->lists * tiles * 2 * sizeof(float);
}

extern "C" int ugl_disp_smooth_forward_grad(const UglDispSmoothArgs* a) {
  DsParams p;
  int rc = ds_fill(a, p);
  if (rc) return rc;
  if (!a->out) return fail(UGL_EINVAL, "disp_smooth_forward_grad: null out");
  for (int li = 0; li < a->lists; ++li) {
    if (!a->img[li]) return fail(UGL_EINVAL, "disp_smooth_forward_grad: null image %d", li);
    p.img[li] = a->img[li];
    for (int l = 0; l < a->levels; ++l) {
      if (!a->disp[li][l]) return fail(UGL_EINVAL, "disp_smooth_forward_grad: null disparity (list %d, level %d)", li, l);
      p.disp[li][l] = a->disp[li][l];
      p.G[li][l] = a->G[li][l];               // may be null: loss only
    }
  }
  if (!a->workspace || a->workspace_bytes < ugl_disp_smooth_fused_workspace_bytes(a))
    return fail(UGL_EWORKSPACE, "disp_smooth_forward_grad: workspace too small (%llu bytes given)", (unsigned long long)a->workspace_bytes);
  p.partials = static_cast<float*>(a->workspace);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  const int tiles = p.tiles_x * p.tiles_y, S = p.B * p.lists;
  disp_smooth_fwdgrad_kernel<<<dim3(tiles, p.B, p.lists), kDsNT, 0, st>>>(p);
  if ((rc = check_launch("disp_smooth_fwdgrad_kernel"))) return rc;
  DsFinal fin{a->out, (float)p.H * (float)(p.W - 1), (float)(p.H - 1) * (float)p.W};
  sample_finalize_kernel<2><<<(S + 3) / 4, 128, 0, st>>>(p.partials, tiles, S, fin);
  return check_launch("disp_smooth finalize");
}

extern "C" int ugl_disp_smooth_combine(const UglDispSmoothArgs* a) {
  DsParams p;
  int rc = ds_fill(a, p);
  if (rc) return rc;
  if (!a->grad_out) return fail(UGL_EINVAL, "disp_smooth_combine: null grad_out");
  p.gout = a->grad_out;
  for (int li = 0; li < a->lists; ++li)
    for (int l = 0; l < a->levels; ++l) {
      if (!a->G[li][l] || !a->grad_disp[li][l]) return fail(UGL_EINVAL, "disp_smooth_combine: null G / grad_disp (list %d, level %d)", li, l);
      p.G[li][l] = a->G[li][l]; p.gdisp[li][l] = a->grad_disp[li][l];
    }
  int chunks = (p.h[0] * p.w[0] + 256 * 4 - 1) / (256 * 4);
  chunks = chunks < 1 ? 1 : (chunks > 64 ? 64 : chunks);
  disp_smooth_combine_kernel<<<dim3(chunks, p.B * p.lists, p.levels), 256, 0, static_cast<cudaStream_t>(a->stream)>>>(p);
  return check_launch("disp_smooth_combine_kernel");
}
