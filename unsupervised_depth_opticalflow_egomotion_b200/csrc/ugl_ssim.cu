// 3x3 box-moment SSIM (pytorch_ssim/ssim.py:4-19) as a tiled stencil:
//   * the primitive map SSIM(x, y) with its backward w.r.t. x and y, and
//   * the masked SSIM loss of compute_ssim_loss / compute_loss_ssim (model_geometry.py:212-223,
//     model_flow.py:141-152): mean(clamp((1 - SSIM(img*m, warped*m)) / 2, 0, 1)) / (mean(m) + 1e-12).
// One CTA = one 32x8 tile of one (sample, channel) plane; inputs are staged in shared memory with a
// 1-pixel (forward) or 2-pixel (backward) zero-padded halo.
#include "ugl_common.cuh"
#include "ugl_reduce.cuh"

namespace ugl {

constexpr int kSTW = 32, kSTH = 8, kSNT = 256;

struct SsimArgs {
  const float *x, *y, *mask;      // (B,C,H,W), (B,C,H,W), (B,1,H,W) or null
  int B, C, H, W, tiles_x, tiles_y;
};

template <int R>
__device__ __forceinline__ void load_tile(const SsimArgs& a, int b, int c, int x0, int y0, float* sx, float* sy, float* sm_) {
  constexpr int PW = kSTW + 2 * R, PH = kSTH + 2 * R;
  const long plane = (long)a.H * a.W;
  const float* xp = a.x + ((long)b * a.C + c) * plane;
  const float* yp = a.y + ((long)b * a.C + c) * plane;
  const float* mp = a.mask ? a.mask + (long)b * plane : nullptr;
  for (int idx = threadIdx.x; idx < PW * PH; idx += kSNT) {
    const int ly = idx / PW, lx = idx - ly * PW;
    const int i = y0 - R + ly, j = x0 - R + lx;
    float vx = 0.f, vy = 0.f, vm = 0.f;
    if (i >= 0 && i < a.H && j >= 0 && j < a.W) {
      const long p = (long)i * a.W + j;
      vm = mp ? mp[p] : 1.0f;
      vx = mp ? mul_rn(xp[p], vm) : xp[p];
      vy = mp ? mul_rn(yp[p], vm) : yp[p];
    }
    sx[idx] = vx; sy[idx] = vy;
    if (sm_) sm_[idx] = vm;
  }
}

template <int PW>
__device__ __forceinline__ Moments tile_moments(const float* sx, const float* sy, int c0) {
  Moments m = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) moments_add(m, sx[c0 + dy * PW + dx], sy[c0 + dy * PW + dx]);
  return m;
}

// forward: kLoss == false writes the map; kLoss == true accumulates the clamped loss and the mask sum
template <bool kLoss>
__global__ void __launch_bounds__(kSNT) ssim_fwd_kernel(SsimArgs a, float* __restrict__ out, float* __restrict__ partials) {
  constexpr int R = 1, PW = kSTW + 2 * R, PH = kSTH + 2 * R;
  __shared__ float sx[PW * PH], sy[PW * PH], smk[PW * PH];
  __shared__ float red[(kSNT / 32) * 2];
  const int tile = blockIdx.x, c = blockIdx.y, b = blockIdx.z;
  const int x0 = (tile % a.tiles_x) * kSTW, y0 = (tile / a.tiles_x) * kSTH;
  load_tile<R>(a, b, c, x0, y0, sx, sy, kLoss ? smk : nullptr);
  __syncthreads();
  const int ty = threadIdx.x / kSTW, tx = threadIdx.x % kSTW;
  const int i = y0 + ty, j = x0 + tx;
  float acc[2] = {0.f, 0.f};
  if (i < a.H && j < a.W) {
    const int c0 = (ty + R) * PW + tx + R;
    const float S = ssim_from_sums(tile_moments<PW>(sx, sy, c0));
    if (kLoss) {
      acc[0] = ssim_loss_value(S);
      if (c == 0) acc[1] = smk[c0];
    } else {
      out[(((long)b * a.C + c) * a.H + i) * a.W + j] = S;
    }
  }
  if (kLoss) {
    const float v = block_reduce_n<kSNT, 2>(acc, red);
    const int chunks = a.C * a.tiles_x * a.tiles_y;
    if (threadIdx.x < 2) partials[((long)b * chunks + c * a.tiles_x * a.tiles_y + tile) * 2 + threadIdx.x] = v;
  }
}

struct SsimLossFinal {
  float *out, *den;
  float n_num, n_den;
  __device__ void operator()(int b, const double* S) const {
    const float d = (float)(S[1] / n_den) + 1e-12f;
    den[b] = d;
    out[b] = (float)(S[0] / n_num) / d;
  }
};

// backward: g(p) = grad_out(p) (map) or the clamp derivative times the per-sample scale (loss);
// grad_x(q) = m(q)/9 [sum g ax + 2 x(q) sum g b + y(q) sum g c], grad_y likewise with ay.
template <bool kLoss>
__global__ void __launch_bounds__(kSNT)
ssim_bwd_kernel(SsimArgs a, const float* __restrict__ gout, const float* __restrict__ den, float scale, float* __restrict__ gx,
                float* __restrict__ gy) {
  constexpr int R = 2, PW = kSTW + 2 * R, PH = kSTH + 2 * R, CW = kSTW + 2, CH = kSTH + 2;
  __shared__ float sx[PW * PH], sy[PW * PH], smk[PW * PH];
  __shared__ float cax[CW * CH], cay[CW * CH], cb[CW * CH], cc[CW * CH];
  const int tile = blockIdx.x, c = blockIdx.y, b = blockIdx.z;
  const int x0 = (tile % a.tiles_x) * kSTW, y0 = (tile / a.tiles_x) * kSTH;
  load_tile<R>(a, b, c, x0, y0, sx, sy, smk);
  __syncthreads();
  const long plane = (long)a.H * a.W;
  const float k = kLoss ? gout[b] * scale / den[b] : 0.f;
  for (int idx = threadIdx.x; idx < CW * CH; idx += kSNT) {
    const int ly = idx / CW, lx = idx - ly * CW;
    const int i = y0 - 1 + ly, j = x0 - 1 + lx;
    float ax = 0.f, bx = 0.f, ay = 0.f, by = 0.f, cxy = 0.f;
    if (i >= 0 && i < a.H && j >= 0 && j < a.W) {
      const SsimTerms t = ssim_terms<true>(tile_moments<PW>(sx, sy, (ly + 1) * PW + lx + 1));
      float g;
      if (kLoss) {
        const float v = mul_rn(sub_rn(1.0f, t.S), 0.5f);
        g = (v >= 0.f && v <= 1.f) ? -0.5f * k : 0.f;
      } else {
        g = gout[((long)b * a.C + c) * plane + (long)i * a.W + j];
      }
      ssim_partials(t, g, ax, bx, ay, by, cxy);
    }
    cax[idx] = ax; cay[idx] = ay; cb[idx] = bx; cc[idx] = cxy;
  }
  __syncthreads();
  const int ty = threadIdx.x / kSTW, tx = threadIdx.x % kSTW;
  const int i = y0 + ty, j = x0 + tx;
  if (i < a.H && j < a.W) {
    const int q0 = (ty + 1) * CW + tx + 1, c0 = (ty + R) * PW + tx + R;
    float sAx = 0.f, sAy = 0.f, sB = 0.f, sC = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int q = q0 + dy * CW + dx;
        sAx += cax[q]; sAy += cay[q]; sB += cb[q]; sC += cc[q];
      }
    const float xv = sx[c0], yv = sy[c0], m = smk[c0] * (1.0f / 9.0f);
    const long o = ((long)b * a.C + c) * plane + (long)i * a.W + j;
    if (gx) gx[o] = (sAx + 2.0f * xv * sB + yv * sC) * m;
    if (gy) gy[o] = (sAy + 2.0f * yv * sB + xv * sC) * m;
  }
}

static int fill_args(SsimArgs& a, const float* x, const float* y, const float* mask, int B, int C, int H, int W) {
  if (!x || !y) return fail(UGL_EINVAL, "ssim: null pointer");
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || B > 65535 || C > 65535) return fail(UGL_EINVAL, "ssim: bad shape");
  a.x = x; a.y = y; a.mask = mask; a.B = B; a.C = C; a.H = H; a.W = W;
  a.tiles_x = (W + kSTW - 1) / kSTW; a.tiles_y = (H + kSTH - 1) / kSTH;
  return UGL_OK;
}

}  // namespace ugl

using namespace ugl;

extern "C" int ugl_ssim_forward(const float* x, const float* y, int32_t B, int32_t C, int32_t H, int32_t W, float* out, void* stream) {
  SsimArgs a;
  int rc = fill_args(a, x, y, nullptr, B, C, H, W);
  if (rc) return rc;
  if (!out) return fail(UGL_EINVAL, "ssim_forward: null output");
  ssim_fwd_kernel<false><<<dim3(a.tiles_x * a.tiles_y, C, B), kSNT, 0, static_cast<cudaStream_t>(stream)>>>(a, out, nullptr);
  return check_launch("ssim_fwd_kernel");
}

extern "C" int ugl_ssim_backward(const float* x, const float* y, const float* grad_out, int32_t B, int32_t C, int32_t H, int32_t W,
                                 float* grad_x, float* grad_y, void* stream) {
  SsimArgs a;
  int rc = fill_args(a, x, y, nullptr, B, C, H, W);
  if (rc) return rc;
  if (!grad_out || (!grad_x && !grad_y)) return fail(UGL_EINVAL, "ssim_backward: null pointer");
  ssim_bwd_kernel<false><<<dim3(a.tiles_x * a.tiles_y, C, B), kSNT, 0, static_cast<cudaStream_t>(stream)>>>(a, grad_out, nullptr, 0.f, grad_x, grad_y);
  return check_launch("ssim_bwd_kernel");
}

extern "C" uint64_t ugl_ssim_loss_workspace_bytes(int32_t B, int32_t C, int32_t H, int32_t W) {
  return (uint64_t)B * C * ((W + kSTW - 1) / kSTW) * ((H + kSTH - 1) / kSTH) * 2 * sizeof(float);
}

extern "C" int ugl_ssim_loss_forward(const float* img, const float* warped, const float* mask, int32_t B, int32_t C, int32_t H, int32_t W,
                                     float* out, float* den, void* ws, uint64_t ws_bytes, void* stream) {
  SsimArgs a;
  int rc = fill_args(a, img, warped, mask, B, C, H, W);
  if (rc) return rc;
  if (!out || !den) return fail(UGL_EINVAL, "ssim_loss_forward: null output");
  if (!ws || ws_bytes < ugl_ssim_loss_workspace_bytes(B, C, H, W)) return fail(UGL_EWORKSPACE, "ssim_loss_forward: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partials = static_cast<float*>(ws);
  ssim_fwd_kernel<true><<<dim3(a.tiles_x * a.tiles_y, C, B), kSNT, 0, st>>>(a, nullptr, partials);
  if ((rc = check_launch("ssim_fwd_kernel<loss>"))) return rc;
  const float plane = (float)H * (float)W;
  SsimLossFinal fin{out, den, (float)C * plane, plane};
  sample_finalize_kernel<2><<<(B + 3) / 4, 128, 0, st>>>(partials, C * a.tiles_x * a.tiles_y, B, fin);
  return check_launch("ssim_loss_finalize");
}

extern "C" int ugl_ssim_loss_backward(const float* img, const float* warped, const float* mask, const float* den, const float* grad_out,
                                      int32_t B, int32_t C, int32_t H, int32_t W, float* grad_img, float* grad_warped, void* stream) {
  SsimArgs a;
  int rc = fill_args(a, img, warped, mask, B, C, H, W);
  if (rc) return rc;
  if (!den || !grad_out || (!grad_img && !grad_warped)) return fail(UGL_EINVAL, "ssim_loss_backward: null pointer");
  const float scale = 1.0f / ((float)C * (float)H * (float)W);
  ssim_bwd_kernel<true><<<dim3(a.tiles_x * a.tiles_y, C, B), kSNT, 0, static_cast<cudaStream_t>(stream)>>>(a, grad_out, den, scale, grad_img, grad_warped);
  return check_launch("ssim_bwd_kernel<loss>");
}
