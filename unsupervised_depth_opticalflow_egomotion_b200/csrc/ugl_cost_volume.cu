// PWC-Net cost volume — replaces PWC_tf.corr_naive (core/networks/structures/pwc_tf.py:97-106):
//   out[b, i*(2d+1)+j, y, x] = mean_c( f1[b,c,y,x] * f2pad[b,c,y+i,x+j] ),   f2pad = f2 zero-padded by d on every side
// which the reference evaluates as (2d+1)^2 = 81 separate multiply / mean / unsqueeze launches plus a cat per call (10 calls per
// step: 5 pyramid levels x 2 directions), and as many again in autograd's backward.  Here: one forward kernel and two backward
// kernels (gather form, no atomics: bit-reproducible).  SURVEY 8(f) rank 2; CUDA cores only (at most ~1 GMAC per step).
#include "ugl_common.cuh"
#include "ugl_host.cuh"

namespace ugl {

constexpr int kCvMaxD = 4;

// One thread per OUTPUT element keeps every pyramid level parallel (the top PWC level is 4x13 pixels with 196 channels: a
// per-pixel mapping would leave the GPU empty).  x is the fastest index: all loads of a warp are coalesced row segments, the
// 81-fold re-reads of the feature rows are L1 / L2 hits (a whole level fits in L2).
// forward: idx over (b, k = i*n + j, y, x)
__global__ void __launch_bounds__(256) cost_volume_fwd_kernel(const float* __restrict__ f1, const float* __restrict__ f2, int B, int C, int H,
                                                              int W, int d, float* __restrict__ out) {
  const int n = 2 * d + 1;
  const long plane = (long)H * W, total = (long)B * n * n * plane;
  const float fc = (float)C;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    long r = idx / W;
    const int y = (int)(r % H); r /= H;
    const int k = (int)(r % (n * n)), b = (int)(r / (n * n));
    const int yy = y + k / n - d, xx = x + k % n - d;
    float acc = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const float* p1 = f1 + (long)b * C * plane + (long)y * W + x;
      const float* p2 = f2 + (long)b * C * plane + (long)yy * W + xx;
      int c = 0;
      for (; c + 4 <= C; c += 4) {             // four independent loads in flight per operand
        const float a0 = __ldg(p1 + (c + 0) * plane), a1 = __ldg(p1 + (c + 1) * plane), a2 = __ldg(p1 + (c + 2) * plane), a3 = __ldg(p1 + (c + 3) * plane);
        const float b0 = __ldg(p2 + (c + 0) * plane), b1 = __ldg(p2 + (c + 1) * plane), b2 = __ldg(p2 + (c + 2) * plane), b3 = __ldg(p2 + (c + 3) * plane);
        acc = fmaf(a0, b0, acc); acc = fmaf(a1, b1, acc); acc = fmaf(a2, b2, acc); acc = fmaf(a3, b3, acc);
      }
      for (; c < C; ++c) acc = fmaf(__ldg(p1 + c * plane), __ldg(p2 + c * plane), acc);
    }
    out[idx] = div_rn(acc, fc);
  }
}

// backward, gather form (no atomics): idx over (b, c, y, x)
//   which == 0: grad_f1[c,y,x]   = (1/C) sum_ij g[ij, y, x]             * f2[c, y+i-d, x+j-d]
//   which == 1: grad_f2[c,y',x'] = (1/C) sum_ij g[ij, y'-i+d, x'-j+d]   * f1[c, y'-i+d, x'-j+d]
template <int D>
__global__ void __launch_bounds__(256) cost_volume_bwd_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                                                              const float* __restrict__ gout, int B, int C, int H, int W,
                                                              float* __restrict__ g1, float* __restrict__ g2) {
  constexpr int N = 2 * D + 1;
  const long plane = (long)H * W, total = (long)B * C * plane;
  const float rc = 1.0f / (float)C;
  for (long id2 = blockIdx.x * (long)blockDim.x + threadIdx.x; id2 < 2 * total; id2 += (long)gridDim.x * blockDim.x) {
    const int which = id2 >= total ? 1 : 0;             // first half of the index space: grad_f1, second half: grad_f2
    const long idx = id2 - (which ? total : 0);
    float* grad = which ? g2 : g1;
    if (!grad) continue;
    const float* other = which ? f1 : f2;
    const int x = (int)(idx % W);
    long r = idx / W;
    const int y = (int)(r % H); r /= H;
    const int c = (int)(r % C), b = (int)(r / C);
    const float* gb = gout + (long)b * N * N * plane;
    const float* p = other + ((long)b * C + c) * plane;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int sy = which == 0 ? y + i - D : y - i + D;
      if (sy < 0 || sy >= H) continue;
      const float* grow = gb + (long)(i * N) * plane + (which == 0 ? (long)y * W + x : (long)sy * W);
      const float* prow = p + (long)sy * W;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const int sx = which == 0 ? x + j - D : x - j + D;
        if (sx >= 0 && sx < W) s = fmaf(__ldg(grow + (long)j * plane + (which == 0 ? 0 : sx)), __ldg(prow + sx), s);
      }
    }
    grad[idx] = s * rc;
  }
}

// Tiled kernels for the large levels.  A CTA owns a 32x8 pixel tile of one sample; the "window" operand (f2 for the forward and
// grad_f1, f1 for grad_f2) is staged per chunk of 8 channels in shared memory with a halo of D.  The one-pixel-per-thread form
// of this kernel is bound by shared-memory bandwidth (one 4-byte LDS per FMA), so a thread owns TWO horizontally adjacent pixels
// and THREE displacement rows: per channel and row it reads the 2D+2 window values both pixels need as 64-bit loads and issues
// 2(2D+1) FMAs on them (0.55 LDS bytes-units per FMA instead of 1), with 6(2D+1) accumulators in registers.
// block (16 pixel pairs, 8 rows, RG row groups); MODE 0: forward; 1: grad_f1; 2: grad_f2 (window offsets mirrored).
constexpr int kCvTW = 32, kCvTH = 8, kCvCh = 8;
template <int D> struct CvCfg {
  static constexpr int N = 2 * D + 1;
  static constexpr int RG = 3;                       // row groups (threads per pixel pair)
  static constexpr int RPT = (N + RG - 1) / RG;      // displacement rows per thread
  static constexpr int PW = kCvTW + 2 * D;           // even: 64-bit window loads stay aligned
  static constexpr int PH = kCvTH + 2 * D, PN = PW * PH;
  static constexpr int NT = (kCvTW / 2) * kCvTH * RG;
};

template <int D, int MODE>
__device__ __forceinline__ void cost_volume_tile(const float* __restrict__ f1, const float* __restrict__ f2, const float* __restrict__ gout,
                                                 int C, int H, int W, int b, float* __restrict__ out, float (*tile)[CvCfg<D>::PN],
                                                 float* part) {
  using Cfg = CvCfg<D>;
  constexpr int N = Cfg::N, RPT = Cfg::RPT, PW = Cfg::PW, PN = Cfg::PN, NT = Cfg::NT;
  const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
  const int tid = (tz * kCvTH + ty) * (kCvTW / 2) + tx;
  const int x0 = blockIdx.x * kCvTW, y0 = blockIdx.y * kCvTH;
  const int x = x0 + 2 * tx, y = y0 + ty;                   // left pixel of the pair
  const bool row_live = (y < H), live0 = row_live && x < W, live1 = row_live && x + 1 < W;
  const long plane = (long)H * W;
  const float* win = (MODE == 2 ? f1 : f2) + (long)b * C * plane;
  const int i0 = tz * RPT;                                   // first displacement row of this thread
  // forward: accumulators; backward: upstream gradients of the two pixels for this thread's displacement rows
  float r[2][RPT][N];
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int a = 0; a < RPT; ++a)
#pragma unroll
      for (int j = 0; j < N; ++j) r[p][a][j] = 0.f;
  if (MODE != 0) {
    const float* gb = gout + (long)b * N * N * plane;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int a = 0; a < RPT; ++a)
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const int i = i0 + a;
          if (i < N && (p == 0 ? live0 : live1)) {
            if (MODE == 1) {
              r[p][a][j] = __ldg(gb + (long)(i * N + j) * plane + (long)y * W + x + p);
            } else {                   // upstream gradient at the partner pixel (y - (i - D), x - (j - D))
              const int sy = y - i + D, sx = x + p - j + D;
              if (sy >= 0 && sy < H && sx >= 0 && sx < W) r[p][a][j] = __ldg(gb + (long)(i * N + j) * plane + (long)sy * W + sx);
            }
          }
        }
  }
  const float rc = 1.0f / (float)C;
  // the tile elements this thread stages are the same for every channel chunk: decode them once (-1 = outside the image)
  constexpr int kSlots = (PN + NT - 1) / NT;
  int slot_off[kSlots];
#pragma unroll
  for (int sl = 0; sl < kSlots; ++sl) {
    const int q = tid + sl * NT;
    const int ly = q / PW, lx = q - ly * PW;
    const int gy = y0 - D + ly, gx = x0 - D + lx;
    slot_off[sl] = (q < PN && gy >= 0 && gy < H && gx >= 0 && gx < W) ? gy * W + gx : -1;
  }
  for (int c0 = 0; c0 < C; c0 += kCvCh) {
    const int nc = C - c0 < kCvCh ? C - c0 : kCvCh;
    __syncthreads();
#pragma unroll
    for (int sl = 0; sl < kSlots; ++sl) {
      const int q = tid + sl * NT;
      if (q < PN) {
        const float* src = win + (long)c0 * plane + slot_off[sl];
        for (int cc = 0; cc < nc; ++cc) tile[cc][q] = slot_off[sl] >= 0 ? __ldg(src + (long)cc * plane) : 0.f;
      }
    }
    __syncthreads();
    float ps[kCvCh][2];
    for (int cc = 0; cc < nc; ++cc) {
      const float* t = tile[cc];
      float a0 = 0.f, a1 = 0.f;
      if (MODE == 0) {
        const float* pf = f1 + ((long)b * C + c0 + cc) * plane + (long)y * W + x;
        a0 = live0 ? __ldg(pf) : 0.f;
        a1 = live1 ? __ldg(pf + 1) : 0.f;
      }
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int a = 0; a < RPT; ++a) {
        const int i = i0 + a;
        if (i >= N) break;
        // window values of this row for both pixels: tile columns 2 tx .. 2 tx + 2D + 1, 64-bit loads
        const int trow = (MODE == 2 ? ty + 2 * D - i : ty + i) * PW + 2 * tx;
        float w[2 * D + 2];
#pragma unroll
        for (int q = 0; q < D + 1; ++q) {
          const float2 v = *reinterpret_cast<const float2*>(t + trow + 2 * q);
          w[2 * q] = v.x; w[2 * q + 1] = v.y;
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
          // MODE 0/1: pixel p pairs with column 2 tx + p + j; MODE 2: with column 2 tx + p + 2D - j
          const float w0 = MODE == 2 ? w[2 * D - j] : w[j], w1 = MODE == 2 ? w[2 * D - j + 1] : w[j + 1];
          if (MODE == 0) {
            r[0][a][j] = fmaf(a0, w0, r[0][a][j]);
            r[1][a][j] = fmaf(a1, w1, r[1][a][j]);
          } else {
            s0 = fmaf(r[0][a][j], w0, s0);
            s1 = fmaf(r[1][a][j], w1, s1);
          }
        }
      }
      if (MODE != 0) { ps[cc][0] = s0; ps[cc][1] = s1; }
    }
    if (MODE != 0) {
      // combine the row groups of a pixel in a fixed order (tz = 0, 1, 2): deterministic
      float* mine = part + (tz * kCvCh) * (kCvTW * kCvTH) + ty * kCvTW + 2 * tx;
      for (int cc = 0; cc < nc; ++cc) *reinterpret_cast<float2*>(mine + cc * (kCvTW * kCvTH)) = make_float2(ps[cc][0], ps[cc][1]);
      __syncthreads();
      for (int e = tid; e < nc * kCvTW * kCvTH; e += NT) {
        const int cc = e / (kCvTW * kCvTH), q = e - cc * (kCvTW * kCvTH);
        const int py = q / kCvTW, px = q - py * kCvTW;
        const int gy = y0 + py, gx = x0 + px;
        if (gy < H && gx < W) {
          float sum = part[(0 * kCvCh + cc) * (kCvTW * kCvTH) + q];
#pragma unroll
          for (int g = 1; g < Cfg::RG; ++g) sum += part[(g * kCvCh + cc) * (kCvTW * kCvTH) + q];
          out[((long)b * C + c0 + cc) * plane + (long)gy * W + gx] = sum * rc;
        }
      }
    }
  }
  if (MODE == 0) {
    const float fc = (float)C;
#pragma unroll
    for (int a = 0; a < RPT; ++a) {
      const int i = i0 + a;
      if (i >= N) break;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        float* o = out + ((long)b * N * N + i * N + j) * plane + (long)y * W + x;
        if (live0) o[0] = div_rn(r[0][a][j], fc);
        if (live1) o[1] = div_rn(r[1][a][j], fc);
      }
    }
  }
}

template <int D>
__global__ void __launch_bounds__(CvCfg<D>::NT) cost_volume_tile_fwd_kernel(const float* __restrict__ f1, const float* __restrict__ f2, int C,
                                                                             int H, int W, float* __restrict__ out) {
  __shared__ __align__(8) float tile[kCvCh][CvCfg<D>::PN];
  cost_volume_tile<D, 0>(f1, f2, nullptr, C, H, W, blockIdx.z, out, tile, nullptr);
}
// both input gradients in one launch (grid.z = 2 B): the two are independent and each alone leaves the GPU half empty
template <int D>
__global__ void __launch_bounds__(CvCfg<D>::NT) cost_volume_tile_bwd_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                                                                             const float* __restrict__ gout, int B, int C, int H, int W,
                                                                             float* __restrict__ g1, float* __restrict__ g2) {
  __shared__ __align__(8) float tile[kCvCh][CvCfg<D>::PN];
  __shared__ float part[CvCfg<D>::RG * kCvCh * kCvTW * kCvTH];      // per-row-group partial sums of a chunk
  const int z = blockIdx.z;
  if (z < B) {
    if (g1) cost_volume_tile<D, 1>(f1, f2, gout, C, H, W, z, g1, tile, part);
  } else {
    if (g2) cost_volume_tile<D, 2>(f1, f2, gout, C, H, W, z - B, g2, tile, part);
  }
}

static int cv_grid(long n) {
  long g = (n + 255) / 256;
  const long cap = 148L * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// small levels (a few hundred pixels): the per-output kernels keep the GPU populated; otherwise the tiled kernels
static bool cv_use_tiles(int H, int W) { return (long)H * W >= 2048; }

template <int D>
static int launch_bwd(const float* f1, const float* f2, const float* go, int B, int C, int H, int W, float* g1, float* g2, cudaStream_t st) {
  if (cv_use_tiles(H, W)) {
    const dim3 grid((W + kCvTW - 1) / kCvTW, (H + kCvTH - 1) / kCvTH, 2 * B), block(kCvTW / 2, kCvTH, CvCfg<D>::RG);
    cost_volume_tile_bwd_kernel<D><<<grid, block, 0, st>>>(f1, f2, go, B, C, H, W, g1, g2);
    return check_launch("cost_volume_tile_bwd_kernel");
  }
  cost_volume_bwd_kernel<D><<<cv_grid(2L * B * C * H * W), 256, 0, st>>>(f1, f2, go, B, C, H, W, g1, g2);
  return check_launch("cost_volume_bwd_kernel");
}

template <int D>
static int launch_fwd_tiles(const float* f1, const float* f2, int B, int C, int H, int W, float* out, cudaStream_t st) {
  const dim3 grid((W + kCvTW - 1) / kCvTW, (H + kCvTH - 1) / kCvTH, B), block(kCvTW / 2, kCvTH, CvCfg<D>::RG);
  cost_volume_tile_fwd_kernel<D><<<grid, block, 0, st>>>(f1, f2, C, H, W, out);
  return check_launch("cost_volume_tile_fwd_kernel");
}

}  // namespace ugl

using namespace ugl;

static int cv_check(const char* what, const void* a, const void* b, const void* c, int B, int C, int H, int W, int d) {
  if (!a || !b || !c) return fail(UGL_EINVAL, "%s: null pointer", what);
  if (B <= 0 || B > 32767 || C <= 0 || H <= 0 || H > 65535 || W <= 0) return fail(UGL_EINVAL, "%s: bad shape (%d,%d,%d,%d)", what, B, C, H, W);
  if (d < 1 || d > kCvMaxD) return fail(UGL_EUNSUPPORTED, "%s: max displacement d=%d outside [1, %d]", what, d, kCvMaxD);
  return UGL_OK;
}

extern "C" int ugl_cost_volume_forward(const float* f1, const float* f2, int32_t B, int32_t C, int32_t H, int32_t W, int32_t d, float* out,
                                       void* stream) {
  int rc = cv_check("cost_volume_forward", f1, f2, out, B, C, H, W, d);
  if (rc) return rc;
  const int n = 2 * d + 1;
  if (cv_use_tiles(H, W)) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (d) {
      case 1: return launch_fwd_tiles<1>(f1, f2, B, C, H, W, out, st);
      case 2: return launch_fwd_tiles<2>(f1, f2, B, C, H, W, out, st);
      case 3: return launch_fwd_tiles<3>(f1, f2, B, C, H, W, out, st);
      default: return launch_fwd_tiles<4>(f1, f2, B, C, H, W, out, st);
    }
  }
  cost_volume_fwd_kernel<<<cv_grid((long)B * n * n * H * W), 256, 0, static_cast<cudaStream_t>(stream)>>>(f1, f2, B, C, H, W, d, out);
  return check_launch("cost_volume_fwd_kernel");
}

extern "C" int ugl_cost_volume_backward(const float* f1, const float* f2, const float* grad_out, int32_t B, int32_t C, int32_t H, int32_t W,
                                        int32_t d, float* grad_f1, float* grad_f2, void* stream) {
  int rc = cv_check("cost_volume_backward", f1, f2, grad_out, B, C, H, W, d);
  if (rc) return rc;
  if (!grad_f1 && !grad_f2) return fail(UGL_EINVAL, "cost_volume_backward: no gradient requested");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (d) {
    case 1: return launch_bwd<1>(f1, f2, grad_out, B, C, H, W, grad_f1, grad_f2, st);
    case 2: return launch_bwd<2>(f1, f2, grad_out, B, C, H, W, grad_f1, grad_f2, st);
    case 3: return launch_bwd<3>(f1, f2, grad_out, B, C, H, W, grad_f1, grad_f2, st);
    default: return launch_bwd<4>(f1, f2, grad_out, B, C, H, W, grad_f1, grad_f2, st);
  }
}
