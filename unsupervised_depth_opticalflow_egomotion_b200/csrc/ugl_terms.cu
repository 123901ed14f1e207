// Stand-alone loss terms and masks of the reference's loss methods (one C-ABI entry per method):
//   masked means P(d,m)            compute_photometric_loss / compute_loss_with_mask / compute_depth_flow_consis_loss
//   occlusion weights + valid      compute_occ_weight (hard) / compute_diff_weight (soft)
//   texture, dynamic, rigid masks  compute_texture_mask / compute_dynamic_mask / get_rigid_mask, fusion_mask*
//   flow smoothness / consistency  compute_loss_flow_smooth / compute_loss_flow_consis
//   disparity smoothness           compute_smooth_loss
//   depth consistency map          compute_consis_loss
// Reference file:line citations are in include/ugl.h next to each entry point.
#include "ugl_common.cuh"
#include "ugl_reduce.cuh"

namespace ugl {

// ================================================================================================
// masked mean  P(d, m) = mean_{c,h,w}(d * m) / (mean_{h,w}(m) + 1e-12)
// ================================================================================================
struct MaskedMeanPixel {
  const float *a, *b, *mask;
  int C, mode;
  long plane;
  __device__ void operator()(int bi, long p, float* acc) const {
    const float m = mask ? mask[(long)bi * plane + p] : 1.0f;
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      const long o = ((long)bi * C + c) * plane + p;
      s += (mode == 0) ? fabsf(a[o] - b[o]) : a[o];
    }
    acc[0] += s * m;
    acc[1] += m;
  }
};
struct MaskedMeanFinal {
  float *out, *den;
  float n_num, n_den;
  __device__ void operator()(int b, const double* S) const {
    const float d = (float)(S[1] / n_den) + 1e-12f;
    den[b] = d;
    out[b] = (float)(S[0] / n_num) / d;
  }
};
struct MaskedMeanGrad {
  const float *a, *b, *mask, *den, *gout;
  float *ga, *gb;
  int C, mode;
  long plane;
  float n_num;
  __device__ void operator()(long idx) const {   // idx over B*C*plane
    const long p = idx % plane;
    const long bc = idx / plane;
    const int bi = (int)(bc / C);
    const float m = mask ? mask[(long)bi * plane + p] : 1.0f;
    const float k = gout[bi] / n_num / den[bi] * m;
    if (mode == 0) {
      const float s = sgnf(b[idx] - a[idx]) * k;
      if (gb) gb[idx] = s;
      if (ga) ga[idx] = -s;
    } else {
      ga[idx] = k;
    }
  }
};

// ================================================================================================
// occlusion weights (M2 + M3) and the channel-mean absolute difference
// ================================================================================================
struct OccWeights {
  const float *from_l, *img, *from_r;
  float *w_bwd, *w_fwd, *valid_bwd, *valid_fwd, *diff_bwd, *diff_fwd;
  long plane;
  int soft;
  __device__ void operator()(long idx) const {   // idx over B*plane
    const long p = idx % plane;
    const long b = idx / plane;
    float I[3], L[3], R[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long o = (b * 3 + c) * plane + p;
      I[c] = img[o]; L[c] = from_l[o]; R[c] = from_r[o];
    }
    const float vb = (L[0] == 0.f && L[1] == 0.f && L[2] == 0.f) ? 0.f : 1.f;
    const float vf = (R[0] == 0.f && R[1] == 0.f && R[2] == 0.f) ? 0.f : 1.f;
    const float s3 = 1.0f / 3.0f;
    const float dl = div_c(add_rn(add_rn(fabsf(sub_rn(I[0], L[0])), fabsf(sub_rn(I[1], L[1]))), fabsf(sub_rn(I[2], L[2]))), 3.0f, s3);
    const float dr = div_c(add_rn(add_rn(fabsf(sub_rn(I[0], R[0])), fabsf(sub_rn(I[1], R[1]))), fabsf(sub_rn(I[2], R[2]))), 3.0f, s3);
    float wl, wr;
    one_minus_softmax2(dl, dr, wl, wr);
    if (soft) {
      w_bwd[idx] = soft_occ_weight(wl) * vb;
      w_fwd[idx] = soft_occ_weight(wr) * vf;
    } else {
      w_bwd[idx] = wl > 0.48f ? 1.f : 0.f;
      w_fwd[idx] = wr > 0.48f ? 1.f : 0.f;
    }
    valid_bwd[idx] = vb;
    valid_fwd[idx] = vf;
    if (diff_bwd) diff_bwd[idx] = dl;
    if (diff_fwd) diff_fwd[idx] = dr;
  }
};

// d = mean_c |img - warped|  ->  d d / d warped_c = sign(warped_c - img_c) / C
struct MeanAbsDiffGrad {
  const float *img, *warped, *gdiff;
  float* gw;
  int C;
  long plane;
  __device__ void operator()(long idx) const {   // idx over B*C*plane
    const long p = idx % plane;
    const long b = idx / plane / C;
    gw[idx] = sgnf(warped[idx] - img[idx]) * gdiff[b * plane + p] / (float)C;
  }
};

struct TextureMask {
  const float *img, *rec, *src;
  float* out;
  long plane;
  __device__ void operator()(long idx) const {
    const long p = idx % plane;
    const long b = idx / plane;
    float a = 0.f, s = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long o = (b * 3 + c) * plane + p;
      a = add_rn(a, fabsf(sub_rn(img[o], rec[o])));
      s = add_rn(s, fabsf(sub_rn(img[o], src[o])));
    }
    const float r3 = 1.0f / 3.0f;
    out[idx] = div_c(a, 3.0f, r3) < div_c(s, 3.0f, r3) ? 1.f : 0.f;
  }
};

// ================================================================================================
// dynamic mask (M6) : n(x) = sqrt(x0^2 + x1^2) + 1e-12 ; dyn = [n(fd)^2 < alpha (n(f)^2 + n(rf)^2) + beta]
// ================================================================================================

struct DynamicMask {
  const float *flow, *rflow;
  float *fd, *dyn, *score;
  long plane;
  float alpha, beta;
  __device__ void operator()(long idx) const {   // idx over B*plane
    const long p = idx % plane;
    const long b = idx / plane;
    const long o0 = (b * 2) * plane + p, o1 = o0 + plane;
    const float fu = flow[o0], fv = flow[o1], ru = rflow[o0], rv = rflow[o1];
    const float nf = norm2_eps(fu, fv), nr = norm2_eps(ru, rv);
    const float bound = add_rn(mul_rn(alpha, add_rn(mul_rn(nf, nf), mul_rn(nr, nr))), beta);
    const float du = fabsf(sub_rn(ru, fu)), dv = fabsf(sub_rn(rv, fv));
    const float nd = norm2_eps(du, dv);
    fd[o0] = du; fd[o1] = dv;
    dyn[idx] = mul_rn(nd, nd) < bound ? 1.f : 0.f;
    if (score) score[idx] = div_rn(1.0f, add_rn(1e-4f, nd));
  }
};

struct AbsDiffGrad {   // d = |a - b|
  const float *a, *b, *g;
  float *ga, *gb;
  __device__ void operator()(long idx) const {
    const float s = sgnf(a[idx] - b[idx]) * g[idx];
    if (ga) ga[idx] = s;
    if (gb) gb[idx] = -s;
  }
};

struct MaskProduct {
  const float* m[4];
  int inv[4];
  int n;
  float* out;
  __device__ void operator()(long idx) const {
    float v = 1.f;
    for (int k = 0; k < n; ++k) {
      const float x = m[k][idx];
      v = mul_rn(v, inv[k] ? sub_rn(1.0f, x) : x);
    }
    out[idx] = v;
  }
};

struct RigidMask {   // get_rigid_mask
  const float* dist;
  float *rigid, *inlier, *score;
  float rigid_thres, inlier_thres;
  __device__ void operator()(long idx) const {
    const float d = dist[idx];
    const float r = d < rigid_thres ? 1.f : 0.f;
    rigid[idx] = r;
    inlier[idx] = d < inlier_thres ? 1.f : 0.f;
    score[idx] = div_rn(mul_rn(r, 1.0f), add_rn(1.0f, d));
  }
};

// ================================================================================================
// second-order flow smoothness, one level (cal_grad2_error(flow / 20, img))
// ================================================================================================
__device__ __forceinline__ float edge_w10(const float* img, long base, long plane, long p, long q) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) s = add_rn(s, fabsf(sub_rn(img[base + c * plane + q], img[base + c * plane + p])));
  return expf(-10.0f * div_c(s, 3.0f, 1.0f / 3.0f));
}
__device__ __forceinline__ float sdiff20(const float* f, long p, long s) {
  constexpr float r20 = 1.0f / 20.0f;
  const float a = div_c(f[p - s], 20.0f, r20), m = div_c(f[p], 20.0f, r20), b = div_c(f[p + s], 20.0f, r20);
  return sub_rn(sub_rn(b, m), sub_rn(m, a));
}

struct FlowSmoothPixel {
  const float *flow, *img;
  int H, W;
  __device__ void operator()(int b, long p, float* acc) const {
    const long plane = (long)H * W;
    const int i = (int)(p / W), j = (int)(p % W);
    const long ib = (long)b * 3 * plane;
    if (j >= 1 && j <= W - 2) {
      const float wx = edge_w10(img, ib, plane, p, p + 1);
      acc[0] += wx * (fabsf(sdiff20(flow + ((long)b * 2) * plane, p, 1)) + fabsf(sdiff20(flow + ((long)b * 2 + 1) * plane, p, 1)));
    }
    if (i >= 1 && i <= H - 2) {
      const float wy = edge_w10(img, ib, plane, p, p + W);
      acc[1] += wy * (fabsf(sdiff20(flow + ((long)b * 2) * plane, p, W)) + fabsf(sdiff20(flow + ((long)b * 2 + 1) * plane, p, W)));
    }
  }
};
struct FlowSmoothFinal {
  float* out;
  float nx, ny;
  __device__ void operator()(int b, const double* S) const { out[b] = ((float)(S[0] / nx) + (float)(S[1] / ny)) * 0.5f; }
};
struct FlowSmoothGrad {
  const float *flow, *img, *gout;
  float* gflow;
  int H, W;
  float nx, ny;
  __device__ void operator()(long idx) const {   // idx over B*plane
    const long plane = (long)H * W;
    const long p = idx % plane;
    const int b = (int)(idx / plane);
    const int i = (int)(p / W), j = (int)(p % W);
    const long ib = (long)b * 3 * plane;
    const float kx = gout[b] * 0.5f / nx / 20.0f, ky = gout[b] * 0.5f / ny / 20.0f;
    float g[2] = {0.f, 0.f};
#pragma unroll
    for (int t = -1; t <= 1; ++t) {
      const float coef = t == 0 ? -2.f : 1.f;
      const int jc = j + t, ic = i + t;
      if (jc >= 1 && jc <= W - 2) {
        const float w = edge_w10(img, ib, plane, p + t, p + t + 1) * coef * kx;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) g[ch] += w * sgnf(sdiff20(flow + ((long)b * 2 + ch) * plane, p + t, 1));
      }
      if (ic >= 1 && ic <= H - 2) {
        const float w = edge_w10(img, ib, plane, p + (long)t * W, p + (long)t * W + W) * coef * ky;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) g[ch] += w * sgnf(sdiff20(flow + ((long)b * 2 + ch) * plane, p + (long)t * W, W));
      }
    }
    gflow[((long)b * 2) * plane + p] = g[0];
    gflow[((long)b * 2 + 1) * plane + p] = g[1];
  }
};

// ================================================================================================
// forward/backward direction consistency, one level (compute_loss_flow_consis)
// ================================================================================================
struct FlowConsisPixel {
  const float *fwd, *bwd, *occ;
  long plane;
  __device__ void operator()(int b, long p, float* acc) const {
    const long o = ((long)b * 2) * plane + p;
    const float uf = fwd[o], vf = fwd[o + plane], ub = bwd[o], vb = bwd[o + plane];
    const float inf_ = fast_div(1.0f, sqrt_rn(uf * uf + vf * vf) + 1e-12f), inb_ = fast_div(1.0f, sqrt_rn(ub * ub + vb * vb) + 1e-12f);
    const float om = 1.0f - occ[(long)b * plane + p];
    acc[0] += (fabsf(uf * inf_ + ub * inb_) + fabsf(vf * inf_ + vb * inb_)) * om;
    acc[1] += om;
  }
};
struct FlowConsisGrad {
  const float *fwd, *bwd, *occ, *den, *gout;
  float* gfwd;
  long plane;
  __device__ void operator()(long idx) const {   // idx over B*plane
    const long p = idx % plane;
    const int b = (int)(idx / plane);
    const long o = ((long)b * 2) * plane + p;
    const float uf = fwd[o], vf = fwd[o + plane], ub = bwd[o], vb = bwd[o + plane];
    const float rf = sqrt_rn(uf * uf + vf * vf), rb = sqrt_rn(ub * ub + vb * vb);
    const float inf_ = fast_div(1.0f, rf + 1e-12f), inb_ = fast_div(1.0f, rb + 1e-12f);
    const float om = (1.0f - occ[idx]) * gout[b] / (2.0f * (float)plane) / den[b];
    const float su = sgnf(uf * inf_ + ub * inb_) * om, sv = sgnf(vf * inf_ + vb * inb_) * om;
    const float gn = -(su * uf + sv * vf) * inf_ * inf_;
    const float ir = rf > 0.f ? fast_div(1.0f, rf) : 0.f;
    gfwd[o] = su * inf_ + gn * uf * ir;
    gfwd[o + plane] = sv * inf_ + gn * vf * ir;
  }
};

// ================================================================================================
// depth consistency map (compute_consis_loss): clamp(|c - p| / |c + p|, 0, 1)
// ================================================================================================
struct DepthDiff {
  const float *comp, *proj;
  float* out;
  __device__ void operator()(long idx) const {
    const float v = div_rn(fabsf(sub_rn(comp[idx], proj[idx])), fabsf(add_rn(comp[idx], proj[idx])));
    out[idx] = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
  }
};
struct DepthDiffGrad {
  const float *comp, *proj, *g;
  float *gc, *gp;
  __device__ void operator()(long idx) const {
    const float c = comp[idx], p = proj[idx];
    const float n = c - p, d = c + p;
    const float an = fabsf(n), ad = fabsf(d);
    const float v = an / ad;
    const float go = (v >= 0.f && v <= 1.f) ? g[idx] : 0.f;
    const float dn = go * sgnf(n) / ad;              // d/d n
    const float dd = -go * an / (ad * ad) * sgnf(d); // d/d d
    if (gc) gc[idx] = dn + dd;
    if (gp) gp[idx] = -dn + dd;
  }
};

// ================================================================================================
// disparity smoothness (compute_smooth_loss): every level bilinearly up-sampled to full resolution
// ================================================================================================
struct UpTap { int i0, i1; float l0, l1; };
// ATen upsample_bilinear2d, align_corners=False: src = scale*(dst+0.5)-0.5 clamped at 0
__device__ __forceinline__ UpTap up_tap(int dst, int n_in, float scale) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  UpTap t;
  t.i0 = (int)src;
  t.i1 = t.i0 + (t.i0 < n_in - 1 ? 1 : 0);
  t.l1 = src - (float)t.i0;
  t.l0 = 1.0f - t.l1;
  return t;
}
__device__ __forceinline__ float up_value(const float* __restrict__ d, int h, int w, int H, int W, int Y, int X) {
  if (h == H && w == W) return d[(long)Y * W + X];
  const UpTap ty = up_tap(Y, h, (float)h / (float)H), tx = up_tap(X, w, (float)w / (float)W);
  const float* r0 = d + (long)ty.i0 * w;
  const float* r1 = d + (long)ty.i1 * w;
  return ty.l0 * (tx.l0 * r0[tx.i0] + tx.l1 * r0[tx.i1]) + ty.l1 * (tx.l0 * r1[tx.i0] + tx.l1 * r1[tx.i1]);
}
__device__ __forceinline__ float edge_w1(const float* img, long base, long plane, long p, long q) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) s = add_rn(s, fabsf(sub_rn(img[base + c * plane + p], img[base + c * plane + q])));
  return expf(-div_c(s, 3.0f, 1.0f / 3.0f));
}

struct DispLevels {
  const float* d[kMaxLevels];
  int h[kMaxLevels], w[kMaxLevels];
  int n;
};

struct DispSmoothPixel {
  const float* img;
  DispLevels lv;
  int H, W;
  __device__ void operator()(int b, long p, float* acc) const {
    const long plane = (long)H * W;
    const int Y = (int)(p / W), X = (int)(p % W);
    const long ib = (long)b * 3 * plane;
    const bool hx = X <= W - 2, hy = Y <= H - 2;
    const float wx = hx ? edge_w1(img, ib, plane, p, p + 1) : 0.f;
    const float wy = hy ? edge_w1(img, ib, plane, p, p + W) : 0.f;
    for (int l = 0; l < lv.n; ++l) {
      const float* d = lv.d[l] + (long)b * lv.h[l] * lv.w[l];
      const float c = up_value(d, lv.h[l], lv.w[l], H, W, Y, X);
      if (hx) acc[0] += fabsf(c - up_value(d, lv.h[l], lv.w[l], H, W, Y, X + 1)) * wx;
      if (hy) acc[1] += fabsf(c - up_value(d, lv.h[l], lv.w[l], H, W, Y + 1, X)) * wy;
    }
  }
};
struct DispSmoothFinal {
  float* out;
  float nx, ny;
  __device__ void operator()(int b, const double* S) const { out[b] = (float)(S[0] / nx) + (float)(S[1] / ny); }
};
// G = d loss / d up(X,Y) for one level, written at full resolution
struct DispSmoothG {
  const float *img, *d, *gout;
  float* G;
  int h, w, H, W;
  float nx, ny;
  __device__ void operator()(long idx) const {   // idx over B*H*W
    const long plane = (long)H * W;
    const long p = idx % plane;
    const int b = (int)(idx / plane);
    const int Y = (int)(p / W), X = (int)(p % W);
    const long ib = (long)b * 3 * plane;
    const float* db = d + (long)b * h * w;
    const float kx = gout[b] / nx, ky = gout[b] / ny;
    const float c = up_value(db, h, w, H, W, Y, X);
    float g = 0.f;
    if (X <= W - 2) g += kx * edge_w1(img, ib, plane, p, p + 1) * sgnf(c - up_value(db, h, w, H, W, Y, X + 1));
    if (X >= 1) g -= kx * edge_w1(img, ib, plane, p - 1, p) * sgnf(up_value(db, h, w, H, W, Y, X - 1) - c);
    if (Y <= H - 2) g += ky * edge_w1(img, ib, plane, p, p + W) * sgnf(c - up_value(db, h, w, H, W, Y + 1, X));
    if (Y >= 1) g -= ky * edge_w1(img, ib, plane, p - W, p) * sgnf(up_value(db, h, w, H, W, Y - 1, X) - c);
    G[idx] = g;
  }
};
// transpose of the bilinear up-sampling, gather form (deterministic): each low-res pixel collects the
// full-resolution G values whose taps touch it
struct UpsampleTranspose {
  const float* G;
  float* gd;
  int h, w, H, W;
  __device__ void operator()(long idx) const {   // idx over B*h*w
    const long lp = (long)h * w;
    const int b = (int)(idx / lp);
    const int y = (int)((idx % lp) / w), x = (int)(idx % w);
    const int fy = H / h, fx = W / w;
    const float sy = (float)h / (float)H, sx = (float)w / (float)W;
    const float* Gb = G + (long)b * H * W;
    float acc = 0.f;
    const int Y0 = max(0, fy * (y - 1)), Y1 = min(H, fy * (y + 2));
    const int X0 = max(0, fx * (x - 1)), X1 = min(W, fx * (x + 2));
    for (int Y = Y0; Y < Y1; ++Y) {
      const UpTap ty = up_tap(Y, h, sy);
      const float wy = (ty.i0 == y ? ty.l0 : 0.f) + (ty.i1 == y ? ty.l1 : 0.f);
      if (wy == 0.f) continue;
      float row = 0.f;
      for (int X = X0; X < X1; ++X) {
        const UpTap tx = up_tap(X, w, sx);
        const float wx = (tx.i0 == x ? tx.l0 : 0.f) + (tx.i1 == x ? tx.l1 : 0.f);
        row += wx * Gb[(long)Y * W + X];
      }
      acc += wy * row;
    }
    gd[idx] = acc;
  }
};

}  // namespace ugl

using namespace ugl;

#define UGL_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) return fail(code, __VA_ARGS__); \
  } while (0)

// ---- masked mean -----------------------------------------------------------------------------------
extern "C" uint64_t ugl_reduce_workspace_bytes(int32_t B, int32_t H, int32_t W) { return reduce_workspace_bytes(B, (long)H * W, 16); }

extern "C" int ugl_masked_mean_forward(const float* a, const float* b, const float* mask, int32_t B, int32_t C, int32_t H,
                                       int32_t W, int32_t mode, float* out, float* den, void* ws, uint64_t ws_bytes, void* stream) {
  UGL_REQUIRE(a && out && den && (mode == 1 || b), UGL_EINVAL, "masked_mean_forward: null pointer");
  UGL_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && (mode == 0 || mode == 1), UGL_EINVAL, "masked_mean_forward: bad arguments");
  const long plane = (long)H * W;
  MaskedMeanPixel px{a, b, mask, C, mode, plane};
  MaskedMeanFinal fin{out, den, (float)C * (float)plane, (float)plane};
  return launch_sample_reduce<2>(px, fin, B, plane, ws, ws_bytes, static_cast<cudaStream_t>(stream), "masked_mean_forward");
}

extern "C" int ugl_masked_mean_backward(const float* a, const float* b, const float* mask, const float* den, const float* grad_out,
                                        int32_t B, int32_t C, int32_t H, int32_t W, int32_t mode, float* grad_a, float* grad_b,
                                        void* stream) {
  UGL_REQUIRE(a && den && grad_out && (mode == 1 || b), UGL_EINVAL, "masked_mean_backward: null pointer");
  UGL_REQUIRE(mode == 0 ? (grad_a || grad_b) : grad_a != nullptr, UGL_EINVAL, "masked_mean_backward: no gradient requested");
  const long plane = (long)H * W;
  MaskedMeanGrad g{a, b, mask, den, grad_out, grad_a, grad_b, C, mode, plane, (float)C * (float)plane};
  return launch_pointwise(g, (long)B * C * plane, static_cast<cudaStream_t>(stream), "masked_mean_backward");
}

// ---- masks -----------------------------------------------------------------------------------------
extern "C" int ugl_occlusion_weights(const float* from_l, const float* img, const float* from_r, int32_t B, int32_t H, int32_t W,
                                     int32_t soft, float* w_bwd, float* w_fwd, float* valid_bwd, float* valid_fwd,
                                     float* diff_bwd, float* diff_fwd, void* stream) {
  UGL_REQUIRE(from_l && img && from_r && w_bwd && w_fwd && valid_bwd && valid_fwd, UGL_EINVAL, "occlusion_weights: null pointer");
  const long plane = (long)H * W;
  OccWeights f{from_l, img, from_r, w_bwd, w_fwd, valid_bwd, valid_fwd, diff_bwd, diff_fwd, plane, soft};
  return launch_pointwise(f, (long)B * plane, static_cast<cudaStream_t>(stream), "occlusion_weights");
}

extern "C" int ugl_channel_mean_abs_diff_backward(const float* img, const float* warped, const float* grad_diff, int32_t B,
                                                  int32_t C, int32_t H, int32_t W, float* grad_warped, void* stream) {
  UGL_REQUIRE(img && warped && grad_diff && grad_warped, UGL_EINVAL, "channel_mean_abs_diff_backward: null pointer");
  const long plane = (long)H * W;
  MeanAbsDiffGrad f{img, warped, grad_diff, grad_warped, C, plane};
  return launch_pointwise(f, (long)B * C * plane, static_cast<cudaStream_t>(stream), "channel_mean_abs_diff_backward");
}

extern "C" int ugl_texture_mask(const float* img, const float* rec, const float* src, int32_t B, int32_t H, int32_t W, float* mask,
                                void* stream) {
  UGL_REQUIRE(img && rec && src && mask, UGL_EINVAL, "texture_mask: null pointer");
  const long plane = (long)H * W;
  TextureMask f{img, rec, src, mask, plane};
  return launch_pointwise(f, (long)B * plane, static_cast<cudaStream_t>(stream), "texture_mask");
}

extern "C" int ugl_dynamic_mask_forward(const float* flow, const float* rigid_flow, int32_t B, int32_t H, int32_t W, float alpha,
                                        float beta, float* flow_diff, float* dyn_mask, float* score, void* stream) {
  UGL_REQUIRE(flow && rigid_flow && flow_diff && dyn_mask, UGL_EINVAL, "dynamic_mask_forward: null pointer");
  const long plane = (long)H * W;
  DynamicMask f{flow, rigid_flow, flow_diff, dyn_mask, score, plane, alpha, beta};
  return launch_pointwise(f, (long)B * plane, static_cast<cudaStream_t>(stream), "dynamic_mask_forward");
}

extern "C" int ugl_abs_diff_backward(const float* a, const float* b, const float* grad_out, int64_t n, float* grad_a, float* grad_b,
                                     void* stream) {
  UGL_REQUIRE(a && b && grad_out && (grad_a || grad_b), UGL_EINVAL, "abs_diff_backward: null pointer");
  AbsDiffGrad f{a, b, grad_out, grad_a, grad_b};
  return launch_pointwise(f, (long)n, static_cast<cudaStream_t>(stream), "abs_diff_backward");
}

extern "C" int ugl_mask_product(const float* const* masks, const int32_t* invert, int32_t n_masks, int64_t n, float* out, void* stream) {
  UGL_REQUIRE(masks && out && n_masks >= 1 && n_masks <= 4, UGL_EINVAL, "mask_product: 1..4 masks");
  MaskProduct f;
  for (int k = 0; k < 4; ++k) { f.m[k] = k < n_masks ? masks[k] : nullptr; f.inv[k] = (k < n_masks && invert) ? invert[k] : 0; }
  for (int k = 0; k < n_masks; ++k) UGL_REQUIRE(f.m[k], UGL_EINVAL, "mask_product: null mask %d", k);
  f.n = n_masks; f.out = out;
  return launch_pointwise(f, (long)n, static_cast<cudaStream_t>(stream), "mask_product");
}

extern "C" int ugl_rigid_mask(const float* dist, int64_t n, float rigid_thres, float inlier_thres, float* rigid, float* inlier,
                              float* score, void* stream) {
  UGL_REQUIRE(dist && rigid && inlier && score, UGL_EINVAL, "rigid_mask: null pointer");
  RigidMask f{dist, rigid, inlier, score, rigid_thres, inlier_thres};
  return launch_pointwise(f, (long)n, static_cast<cudaStream_t>(stream), "rigid_mask");
}

// ---- step glue: gradient accumulation over several tensors, weighted total -----------------------------------------------
// dst[i] += src[i][0] (+ src[i][1] + src[i][2]) for up to kAccumulateMax destinations in ONE launch (grid.y = destination): what
// autograd's engine does with one `add` launch per contribution when a tensor feeds several loss terms.  Each destination appears
// once (its contributions are added in the order given: deterministic).
constexpr int kAccumulateMax = 24, kAccumulateSrc = 3;
struct AccumulateParams { float* dst[kAccumulateMax]; const float* src[kAccumulateMax][kAccumulateSrc]; long n[kAccumulateMax]; };

__global__ void __launch_bounds__(256) accumulate_multi_kernel(const __grid_constant__ AccumulateParams p) {
  const int k = blockIdx.y;
  float* __restrict__ d = p.dst[k];
  const float* __restrict__ s0 = p.src[k][0];
  const float* __restrict__ s1 = p.src[k][1];
  const float* __restrict__ s2 = p.src[k][2];
  const long n = p.n[k];
  const long stride = (long)gridDim.x * blockDim.x, t0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const uintptr_t bits = reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(s0) | reinterpret_cast<uintptr_t>(s1) | reinterpret_cast<uintptr_t>(s2);
  if ((n & 3) == 0 && (bits & 15u) == 0) {
    float4* d4 = reinterpret_cast<float4*>(d);
    for (long i = t0; i < (n >> 2); i += stride) {
      float4 a = d4[i];
      const float4 b = reinterpret_cast<const float4*>(s0)[i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      if (s1) { const float4 c = reinterpret_cast<const float4*>(s1)[i]; a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w; }
      if (s2) { const float4 c = reinterpret_cast<const float4*>(s2)[i]; a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w; }
      d4[i] = a;
    }
  } else {
    for (long i = t0; i < n; i += stride) {
      float a = d[i] + s0[i];
      if (s1) a += s1[i];
      if (s2) a += s2[i];
      d[i] = a;
    }
  }
}

extern "C" int ugl_accumulate_multi(float* const* dst, const float* const* src, const int64_t* numel, int32_t n, void* stream) {
  UGL_REQUIRE(dst && src && numel && n >= 1 && n <= kAccumulateMax, UGL_EINVAL, "accumulate_multi: 1..%d destinations", kAccumulateMax);
  AccumulateParams p;
  long nmax = 0;
  for (int k = 0; k < n; ++k) {
    UGL_REQUIRE(dst[k] && src[kAccumulateSrc * k] && numel[k] >= 0, UGL_EINVAL, "accumulate_multi: null pointer / negative size at %d", k);
    for (int j = 0; j < k; ++j) UGL_REQUIRE(dst[j] != dst[k], UGL_EINVAL, "accumulate_multi: destination %d listed twice", k);
    p.dst[k] = dst[k]; p.n[k] = (long)numel[k];
    for (int j = 0; j < kAccumulateSrc; ++j) p.src[k][j] = src[kAccumulateSrc * k + j];
    nmax = numel[k] > nmax ? (long)numel[k] : nmax;
  }
  long chunks = (nmax / 4 + 255) / 256;
  chunks = chunks < 1 ? 1 : (chunks > 148 * 4 ? 148 * 4 : chunks);
  accumulate_multi_kernel<<<dim3((unsigned)chunks, (unsigned)n), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("accumulate_multi_kernel");
}

// dst[i][b] = sum_{r < nsum[i]} src[i][r * B + b]: the (B,) outputs of the per-term kernels gathered into the rows of one (n,B) loss
// matrix (nsum[i] > 1: e.g. the three compute_smooth_loss calls of model_geometry.py:938-940 summed in call order).
constexpr int kAssembleMax = 16;
struct AssembleParams { const float* src[kAssembleMax]; int nsum[kAssembleMax]; int n, B; float* dst; };
__global__ void assemble_rows_kernel(const __grid_constant__ AssembleParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n * p.B) return;
  const int k = i / p.B, b = i - k * p.B;
  float v = p.src[k][b];
  for (int r = 1; r < p.nsum[k]; ++r) v += p.src[k][r * p.B + b];
  p.dst[i] = v;
}
extern "C" int ugl_assemble_rows(const float* const* src, const int32_t* nsum, int32_t n, int32_t batch, float* dst, void* stream) {
  UGL_REQUIRE(src && nsum && dst && n >= 1 && n <= kAssembleMax && batch >= 1, UGL_EINVAL, "assemble_rows: 1..%d rows", kAssembleMax);
  AssembleParams p;
  for (int k = 0; k < n; ++k) {
    UGL_REQUIRE(src[k] && nsum[k] >= 1, UGL_EINVAL, "assemble_rows: null row / empty sum at %d", k);
    p.src[k] = src[k]; p.nsum[k] = nsum[k];
  }
  p.n = n; p.B = batch; p.dst = dst;
  assemble_rows_kernel<<<(n * batch + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("assemble_rows_kernel");
}

// total = sum_k w[k] * mean_b loss[k][b] (train.py:211-214) in one launch, fixed order (deterministic); and its backward
// grad[k][b] = g * w[k] / B.
__global__ void weighted_total_kernel(const float* __restrict__ loss, const float* __restrict__ w, int K, int B, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float tot = 0.f;
    for (int k = 0; k < K; ++k) {
      float s = 0.f;
      for (int b = 0; b < B; ++b) s += loss[k * B + b];
      tot += w[k] * (s / (float)B);
    }
    out[0] = tot;
  }
}
__global__ void weighted_total_grad_kernel(const float* __restrict__ g, const float* __restrict__ w, int K, int B, float* __restrict__ grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K * B) grad[i] = g[0] * (w[i / B] / (float)B);
}

extern "C" int ugl_weighted_total_forward(const float* loss, const float* weights, int32_t terms, int32_t batch, float* out, void* stream) {
  UGL_REQUIRE(loss && weights && out && terms >= 1 && batch >= 1, UGL_EINVAL, "weighted_total_forward: null pointer / empty matrix");
  weighted_total_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(loss, weights, terms, batch, out);
  return check_launch("weighted_total_kernel");
}
extern "C" int ugl_weighted_total_backward(const float* grad_out, const float* weights, int32_t terms, int32_t batch, float* grad_loss, void* stream) {
  UGL_REQUIRE(grad_out && weights && grad_loss && terms >= 1 && batch >= 1, UGL_EINVAL, "weighted_total_backward: null pointer / empty matrix");
  weighted_total_grad_kernel<<<(terms * batch + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(grad_out, weights, terms, batch, grad_loss);
  return check_launch("weighted_total_grad_kernel");
}

// ---- flow regularisers --------------------------------------------------------------------------------
extern "C" int ugl_flow_smooth_forward(const float* flow, const float* img, int32_t B, int32_t H, int32_t W, float* out, void* ws,
                                       uint64_t ws_bytes, void* stream) {
  UGL_REQUIRE(flow && img && out, UGL_EINVAL, "flow_smooth_forward: null pointer");
  UGL_REQUIRE(H >= 3 && W >= 3, UGL_EUNSUPPORTED, "flow_smooth_forward: need at least 3x3");
  FlowSmoothPixel px{flow, img, H, W};
  FlowSmoothFinal fin{out, 2.0f * (float)H * (float)(W - 2), 2.0f * (float)(H - 2) * (float)W};
  return launch_sample_reduce<2>(px, fin, B, (long)H * W, ws, ws_bytes, static_cast<cudaStream_t>(stream), "flow_smooth_forward");
}

extern "C" int ugl_flow_smooth_backward(const float* flow, const float* img, const float* grad_out, int32_t B, int32_t H, int32_t W,
                                        float* grad_flow, void* stream) {
  UGL_REQUIRE(flow && img && grad_out && grad_flow, UGL_EINVAL, "flow_smooth_backward: null pointer");
  FlowSmoothGrad g{flow, img, grad_out, grad_flow, H, W, 2.0f * (float)H * (float)(W - 2), 2.0f * (float)(H - 2) * (float)W};
  return launch_pointwise(g, (long)B * H * W, static_cast<cudaStream_t>(stream), "flow_smooth_backward");
}

extern "C" int ugl_flow_consis_forward(const float* fwd, const float* bwd, const float* occ, int32_t B, int32_t H, int32_t W,
                                       float* out, float* den, void* ws, uint64_t ws_bytes, void* stream) {
  UGL_REQUIRE(fwd && bwd && occ && out && den, UGL_EINVAL, "flow_consis_forward: null pointer");
  const long plane = (long)H * W;
  FlowConsisPixel px{fwd, bwd, occ, plane};
  MaskedMeanFinal fin{out, den, 2.0f * (float)plane, (float)plane};
  return launch_sample_reduce<2>(px, fin, B, plane, ws, ws_bytes, static_cast<cudaStream_t>(stream), "flow_consis_forward");
}

extern "C" int ugl_flow_consis_backward(const float* fwd, const float* bwd, const float* occ, const float* den, const float* grad_out,
                                        int32_t B, int32_t H, int32_t W, float* grad_fwd, void* stream) {
  UGL_REQUIRE(fwd && bwd && occ && den && grad_out && grad_fwd, UGL_EINVAL, "flow_consis_backward: null pointer");
  const long plane = (long)H * W;
  FlowConsisGrad g{fwd, bwd, occ, den, grad_out, grad_fwd, plane};
  return launch_pointwise(g, (long)B * plane, static_cast<cudaStream_t>(stream), "flow_consis_backward");
}

// ---- depth consistency map -----------------------------------------------------------------------------
extern "C" int ugl_depth_diff_forward(const float* comp, const float* proj, int64_t n, float* out, void* stream) {
  UGL_REQUIRE(comp && proj && out, UGL_EINVAL, "depth_diff_forward: null pointer");
  DepthDiff f{comp, proj, out};
  return launch_pointwise(f, (long)n, static_cast<cudaStream_t>(stream), "depth_diff_forward");
}
extern "C" int ugl_depth_diff_backward(const float* comp, const float* proj, const float* grad_out, int64_t n, float* grad_comp,
                                       float* grad_proj, void* stream) {
  UGL_REQUIRE(comp && proj && grad_out && (grad_comp || grad_proj), UGL_EINVAL, "depth_diff_backward: null pointer");
  DepthDiffGrad f{comp, proj, grad_out, grad_comp, grad_proj};
  return launch_pointwise(f, (long)n, static_cast<cudaStream_t>(stream), "depth_diff_backward");
}

// ---- disparity smoothness --------------------------------------------------------------------------------
static int fill_disp_levels(const float* const* disps, const int32_t* hs, const int32_t* ws, int levels, int H, int W, DispLevels& lv) {
  UGL_REQUIRE(disps && hs && ws && levels >= 1 && levels <= UGL_MAX_LEVELS, UGL_EINVAL, "disp_smooth: bad level arguments");
  lv.n = levels;
  for (int l = 0; l < levels; ++l) {
    UGL_REQUIRE(disps[l], UGL_EINVAL, "disp_smooth: null disparity at level %d", l);
    UGL_REQUIRE(hs[l] > 0 && ws[l] > 0 && H % hs[l] == 0 && W % ws[l] == 0, UGL_EUNSUPPORTED,
                "disp_smooth: level %d (%dx%d) does not divide %dx%d", l, hs[l], ws[l], H, W);
    lv.d[l] = disps[l]; lv.h[l] = hs[l]; lv.w[l] = ws[l];
  }
  return UGL_OK;
}

extern "C" int ugl_disp_smooth_forward(const float* img, const float* const* disps, const int32_t* hs, const int32_t* wsz,
                                       int32_t levels, int32_t B, int32_t H, int32_t W, float* out, void* ws, uint64_t ws_bytes,
                                       void* stream) {
  UGL_REQUIRE(img && out && H >= 2 && W >= 2, UGL_EINVAL, "disp_smooth_forward: bad arguments");
  DispSmoothPixel px;
  px.img = img; px.H = H; px.W = W;
  int rc = fill_disp_levels(disps, hs, wsz, levels, H, W, px.lv);
  if (rc) return rc;
  DispSmoothFinal fin{out, (float)H * (float)(W - 1), (float)(H - 1) * (float)W};
  return launch_sample_reduce<2>(px, fin, B, (long)H * W, ws, ws_bytes, static_cast<cudaStream_t>(stream), "disp_smooth_forward");
}

extern "C" uint64_t ugl_disp_smooth_backward_workspace_bytes(int32_t B, int32_t H, int32_t W) { return (uint64_t)B * H * W * sizeof(float); }

extern "C" int ugl_disp_smooth_backward(const float* img, const float* const* disps, const int32_t* hs, const int32_t* wsz,
                                        int32_t levels, const float* grad_out, int32_t B, int32_t H, int32_t W,
                                        float* const* grad_disps, void* ws, uint64_t ws_bytes, void* stream) {
  UGL_REQUIRE(img && grad_out && grad_disps, UGL_EINVAL, "disp_smooth_backward: null pointer");
  DispLevels lv;
  int rc = fill_disp_levels(disps, hs, wsz, levels, H, W, lv);
  if (rc) return rc;
  UGL_REQUIRE(ws && ws_bytes >= ugl_disp_smooth_backward_workspace_bytes(B, H, W), UGL_EWORKSPACE, "disp_smooth_backward: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float nx = (float)H * (float)(W - 1), ny = (float)(H - 1) * (float)W;
  for (int l = 0; l < levels; ++l) {
    UGL_REQUIRE(grad_disps[l], UGL_EINVAL, "disp_smooth_backward: null grad at level %d", l);
    const bool full = lv.h[l] == H && lv.w[l] == W;
    float* G = full ? grad_disps[l] : static_cast<float*>(ws);
    DispSmoothG g{img, lv.d[l], grad_out, G, lv.h[l], lv.w[l], H, W, nx, ny};
    if ((rc = launch_pointwise(g, (long)B * H * W, st, "disp_smooth_backward(G)"))) return rc;
    if (!full) {
      UpsampleTranspose t{G, grad_disps[l], lv.h[l], lv.w[l], H, W};
      if ((rc = launch_pointwise(t, (long)B * lv.h[l] * lv.w[l], st, "disp_smooth_backward(transpose)"))) return rc;
    }
  }
  return UGL_OK;
}
