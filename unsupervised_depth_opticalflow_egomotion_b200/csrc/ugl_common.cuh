// Shared device/host math for the photometric view-synthesis loss kernels (sm_100a).
//
// Everything here is `__host__ __device__` so that the per-tile logic of every kernel can also be
// compiled by g++ into the host emulator under tests/hostemu/ (test infrastructure that checks the
// kernels' indexing and arithmetic on the CPU-only build box; it is never linked into the shipped
// library and the product has no CPU path).
//
// Arithmetic conventions (verified against the oracle):
//  * sampling = ATen grid_sampler_2d, bilinear, zeros padding, align_corners=False
//    (torch/include/ATen/native/GridSampler.h:27-36): ix = ((gx+1)*W-1)/2, which both the CPU
//    (vectorised fma) and the CUDA (nvcc-contracted) builds of torch evaluate with a single
//    rounding as fma(gx+1, W/2, -0.5).  We pin exactly that sequence with *_rn intrinsics so the
//    bilinear cell (floor) is bit-identical to the oracle's.
//  * coordinates that feed masks or floor() are computed with explicitly rounded operations
//    (no compiler FMA contraction); plain loss arithmetic may be contracted.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define UGL_HD __host__ __device__ __forceinline__
#define UGL_D __device__ __forceinline__
#else
#define UGL_HD inline
#define UGL_D inline
#endif

#if !defined(__CUDACC__)
struct float2 { float x, y; };
inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
struct alignas(16) float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#endif

namespace ugl {

constexpr int kMaxLevels = 6;

UGL_HD int imin(int a, int b) { return a < b ? a : b; }
UGL_HD int imax(int a, int b) { return a > b ? a : b; }

// ---- explicitly rounded fp32 ops (identical results on device and in the host emulator, which is
// ---- compiled with -ffp-contract=off) --------------------------------------------------------
UGL_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}
UGL_HD float sub_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  volatile float r = a - b; return r;
#endif
}
UGL_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}
UGL_HD float div_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b; return r;
#endif
}
UGL_HD float fma_rn(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}
UGL_HD float sqrt_rn(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}
// sign(v) in {-1, 0, +1}: two float-valued compares (FSET.BF) and a subtraction: 3 instructions instead of the 6 of the
// integer form (float)((v > 0) - (v < 0))
UGL_HD float sgnf(float v) { return (v > 0.f ? 1.0f : 0.0f) - (v < 0.f ? 1.0f : 0.0f); }

// ---- packed fp32 pairs (sm_100a FADD2 / FMUL2 / FFMA2 = add/mul/fma.rn.f32x2) ---------------------------------------
// One instruction produces two individually IEEE-rounded fp32 results from 64-bit register pairs.  The fp32 pipe still
// needs two passes, so a pure fp32 stream gains nothing (profiles/microbench/f32x2_issue.cu: 123 vs 126 results/clk/SM), but
// an ISSUE-bound mix does: the packed form frees every second issue slot for the integer / shared-memory instructions
// around it (51 -> 66 results/clk/SM with two integer ops per pair).  The stencil phases of the single-pass kernel carry
// the two warp directions of a pixel as such a pair (.x = forward-flow / frame-0 direction, .y = the other), which keeps
// every rounding of the scalar formulation and halves the fp32 instruction count.  The host emulator gets the scalar ops.
UGL_HD float2 add2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  return __fadd2_rn(a, b);
#else
  return make_float2(add_rn(a.x, b.x), add_rn(a.y, b.y));
#endif
}
UGL_HD float2 mul2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  return __fmul2_rn(a, b);
#else
  return make_float2(mul_rn(a.x, b.x), mul_rn(a.y, b.y));
#endif
}
UGL_HD float2 fma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__)
  return __ffma2_rn(a, b, c);
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
UGL_HD float2 splat2(float v) { return make_float2(v, v); }
// acc + prod where prod is the (separately rounded) result of a mul2.  ptxas 12.9 contracts `mul.rn.f32x2` followed by
// `add.rn.f32x2` into one FFMA2 although both carry .rn (it does not do that to the scalar forms; -fmad=false, a volatile asm and
// a literal 1.0 do not stop it; checked in the SASS), which would drop the product's rounding and break the bit-identity of the
// SSIM moments with ATen's.  fma(prod, one, acc) with a 1.0 the compiler cannot see (a kernel parameter) is the same single
// rounding of prod + acc, costs the same one instruction and cannot be contracted (it would need two multiplies).
UGL_HD float2 acc2_rn(float2 acc, float2 prod, float2 one) {
#if defined(__CUDA_ARCH__)
  return __ffma2_rn(prod, one, acc);
#else
  (void)one;
  return make_float2(add_rn(acc.x, prod.x), add_rn(acc.y, prod.y));
#endif
}
// a - b where b is the (separately rounded) result of a mul2: the opaque form of acc2_rn
UGL_HD float2 sub2_rn(float2 a, float2 b, float2 one);
// a - b with one rounding: fma(b, -1, a) (the product is exact).  NOT for a b that is itself a bare mul2 result (ptxas would
// contract it): use sub2_rn there.
UGL_HD float2 sub2(float2 a, float2 b) { return fma2(b, splat2(-1.0f), a); }
UGL_HD float2 sub2_rn(float2 a, float2 b, float2 one) {
#if defined(__CUDA_ARCH__)
  return __ffma2_rn(b, make_float2(-one.x, -one.y), a);
#else
  (void)one;
  return make_float2(sub_rn(a.x, b.x), sub_rn(a.y, b.y));
#endif
}
UGL_HD float2 div_c2(float2 a, float c, float rc) {
  const float2 q = mul2(a, splat2(rc));
  return fma2(fma2(splat2(-c), q, a), splat2(rc), q);
}
UGL_HD float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
UGL_HD float2 hi2(const float4& v) { return make_float2(v.z, v.w); }

// a / c for a divisor whose correctly rounded reciprocal rc = RN(1/c) is known: multiply, one FMA
// residual, one FMA correction.  Gives the correctly rounded quotient (same bits as IEEE division;
// checked against a/c on 1.3e9 random operands for c in {1..4096, 0.03}) in 3 instructions instead
// of the ~15 of the IEEE division subroutine.
UGL_HD float div_c(float a, float c, float rc) {
  const float q = mul_rn(a, rc);
  return fma_rn(fma_rn(-c, q, a), rc, q);
}
// approximate quotient / reciprocal (MUFU.RCP, ~1 ulp) for values that do not feed masks, floor() or
// cancellation-prone differences
// One MUFU.RCP (+ one FMUL): __fdividef wraps the same reciprocal in a range fix-up for |b| > 2^126 that costs four more
// instructions per use; no divisor on this path gets near that (they are sums of squares of image-scale numbers).
UGL_HD float fast_rcp(float b) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
  return r;
#else
  return 1.0f / b;
#endif
}
UGL_HD float fast_div(float a, float b) { return a * fast_rcp(b); }

// Fixed-point scale exponent of the deterministic scatter (ugl_scatter.cuh): |sum| * 2^e < 2^61 for up
// to n_contrib contributions of magnitude <= max_abs.
UGL_HD int fixed_point_exponent(float max_abs, long n_contrib) {
  if (!(max_abs > 0.f)) return 0;
  int e_max, e_cnt;
  frexpf(max_abs, &e_max);
  frexpf((float)n_contrib, &e_cnt);
  return 61 - e_max - e_cnt;
}

// ---- bilinear footprint ------------------------------------------------------------------------
// One backward-warp lookup: the nw corner, the four weights and which corners are inside the image.
struct Tap {
  int x0, y0;                 // nw corner (may be out of range)
  float wnw, wne, wsw, wse;   // (x1-ix)(y1-iy), (ix-x0)(y1-iy), (x1-ix)(iy-y0), (ix-x0)(iy-y0)
  float tx, ty;               // ix-x0, iy-y0
  unsigned inb;               // bit0 nw, bit1 ne, bit2 sw, bit3 se inside the image
};

// un-normalise a [-1,1] coordinate exactly like ATen (see header comment)
UGL_HD float unnormalize(float g, int size) { return fma_rn(add_rn(g, 1.0f), 0.5f * (float)size, -0.5f); }

UGL_HD Tap make_tap(float ix, float iy, int W, int H) {
  Tap t;
  // guard the float->int conversion (ATen's is UB for huge values; everything out there is zeros)
  const bool sane = (ix > -2.0f) && (ix < (float)W + 1.0f) && (iy > -2.0f) && (iy < (float)H + 1.0f);
  const float fx = sane ? floorf(ix) : -2.0f;
  const float fy = sane ? floorf(iy) : -2.0f;
  t.x0 = (int)fx;
  t.y0 = (int)fy;
  const float ex = sane ? ix : -2.0f, ey = sane ? iy : -2.0f;
  t.tx = ex - fx;
  t.ty = ey - fy;
  const float ox = (fx + 1.0f) - ex, oy = (fy + 1.0f) - ey;
  t.wnw = ox * oy;
  t.wne = t.tx * oy;
  t.wsw = ox * t.ty;
  t.wse = t.tx * t.ty;
  const bool xl = (t.x0 >= 0) && (t.x0 < W), xr = (t.x0 + 1 >= 0) && (t.x0 + 1 < W);
  const bool yt = (t.y0 >= 0) && (t.y0 < H), yb = (t.y0 + 1 >= 0) && (t.y0 + 1 < H);
  t.inb = (unsigned)(xl && yt) | ((unsigned)(xr && yt) << 1) | ((unsigned)(xl && yb) << 2) | ((unsigned)(xr && yb) << 3);
  return t;
}

// per-image constants of the flow warp: the (W-1)/(H-1) normalisers of net_utils.py:42-43 with
// their reciprocals, and d ix / d u = W/(W-1), d iy / d v = H/(H-1)
struct WarpGeom { int W, H; float dw, dh, rdw, rdh, sx, sy; };

UGL_HD WarpGeom make_warp_geom(int W, int H) {
  WarpGeom g;
  g.W = W; g.H = H;
  g.dw = (float)(W - 1 > 1 ? W - 1 : 1);
  g.dh = (float)(H - 1 > 1 ? H - 1 : 1);
  g.rdw = div_rn(1.0f, g.dw);
  g.rdh = div_rn(1.0f, g.dh);
  g.sx = (float)W / g.dw;
  g.sy = (float)H / g.dh;
  return g;
}

// structures/net_utils.py:39-46: target (j+u, i+v) -> normalised with (W-1) -> un-normalised with W.
// Each Python-level op of the reference rounds to fp32, so every step here is an explicit *_rn op.
UGL_HD Tap flow_tap(int j, int i, float u, float v, const WarpGeom& g) {
  const float gx = sub_rn(div_c(mul_rn(2.0f, add_rn((float)j, u)), g.dw, g.rdw), 1.0f);
  const float gy = sub_rn(div_c(mul_rn(2.0f, add_rn((float)i, v)), g.dh, g.rdh), 1.0f);
  return make_tap(unnormalize(gx, g.W), unnormalize(gy, g.H), g.W, g.H);
}

// coverage of a ones-image = sum of the in-bounds weights in corner order nw, ne, sw, se
UGL_HD float tap_coverage(const Tap& t) {
  float s = 0.f;
  if (t.inb & 1u) s = add_rn(s, t.wnw);
  if (t.inb & 2u) s = add_rn(s, t.wne);
  if (t.inb & 4u) s = add_rn(s, t.wsw);
  if (t.inb & 8u) s = add_rn(s, t.wse);
  return s;
}

// structures/net_utils.py:47-51: keep = [coverage >= 0.9999]
UGL_HD float tap_keep(const Tap& t) { return tap_coverage(t) >= 0.9999f ? 1.0f : 0.0f; }

struct Corners { float nw, ne, sw, se; };

UGL_HD Corners tap_fetch(const float* __restrict__ plane, int W, const Tap& t) {
  Corners c;
  const float* p = plane + (long)t.y0 * W + t.x0;
  c.nw = (t.inb & 1u) ? p[0] : 0.f;
  c.ne = (t.inb & 2u) ? p[1] : 0.f;
  c.sw = (t.inb & 4u) ? p[W] : 0.f;
  c.se = (t.inb & 8u) ? p[W + 1] : 0.f;
  return c;
}

// ATen's CUDA sampler accumulates `out += value * weight` corner by corner (nw, ne, sw, se), which nvcc contracts into
// RN(nw w_nw) -> fma(ne, w_ne, .) -> fma(sw, w_sw, .) -> fma(se, w_se, .).  Written out explicitly: left to the compiler, the
// scalar expression was contracted as fma(nw, w_nw, RN(ne w_ne)) instead (seen in the SASS), one ulp off on some pixels.
UGL_HD float corners_value(const Corners& c, const Tap& t) {
  return fma_rn(c.se, t.wse, fma_rn(c.sw, t.wsw, fma_rn(c.ne, t.wne, mul_rn(c.nw, t.wnw))));
}
// d value / d ix and d value / d iy (ATen grid_sampler_2d_backward: out-of-range corners count as 0)
UGL_HD float corners_ddx(const Corners& c, const Tap& t) { return (c.ne - c.nw) * (1.0f - t.ty) + (c.se - c.sw) * t.ty; }
UGL_HD float corners_ddy(const Corners& c, const Tap& t) { return (c.sw - c.nw) * (1.0f - t.tx) + (c.se - c.ne) * t.tx; }

// ---- SSIM (pytorch_ssim/ssim.py:4-19) ----------------------------------------------------------
constexpr float kC1 = 0.0001f;   // 0.01^2
constexpr float kC2 = 0.0009f;   // 0.03^2

struct Moments { float sx, sy, sxx, syy, sxy; };   // 3x3 window sums (not yet divided by 9)

// One window tap.  Every step is explicitly rounded so the sums equal ATen's avg_pool2d
// accumulation (row-major over the window, fp32) of the separately rounded x*x, y*y, x*y tensors.
UGL_HD void moments_add(Moments& m, float x, float y) {
  m.sx = add_rn(m.sx, x);
  m.sy = add_rn(m.sy, y);
  m.sxx = add_rn(m.sxx, mul_rn(x, x));
  m.syy = add_rn(m.syy, mul_rn(y, y));
  m.sxy = add_rn(m.sxy, mul_rn(x, y));
}

// The SSIM terms in the reference's own operation order (ssim.py:8-18), each op rounded to fp32:
// cancellation in E[x^2]-mu^2 amplifies any re-association, so this is pinned, not contracted.
struct SsimTerms { float mx, my, n1, n2, d1, d2, S, rD; };   // rD = 1 / (d1 d2), shared by the value and its derivatives

template <bool kExactDiv = true>
UGL_HD SsimTerms ssim_terms(const Moments& m) {
  SsimTerms t;
  constexpr float r9 = 1.0f / 9.0f;
  t.mx = div_c(m.sx, 9.0f, r9);
  t.my = div_c(m.sy, 9.0f, r9);
  const float mxx = mul_rn(t.mx, t.mx), myy = mul_rn(t.my, t.my), mxy = mul_rn(t.mx, t.my);
  const float vx = sub_rn(div_c(m.sxx, 9.0f, r9), mxx);
  const float vy = sub_rn(div_c(m.syy, 9.0f, r9), myy);
  const float cxy = sub_rn(div_c(m.sxy, 9.0f, r9), mxy);
  t.n1 = add_rn(mul_rn(mul_rn(2.0f, t.mx), t.my), kC1);
  t.n2 = add_rn(mul_rn(2.0f, cxy), kC2);
  t.d1 = add_rn(add_rn(mxx, myy), kC1);
  t.d2 = add_rn(add_rn(vx, vy), kC2);
  // kExactDiv = false: the moments stay bit-identical to ATen's, only the last quotient uses the ~1-ulp
  // reciprocal (2 instructions instead of the ~15 of the IEEE division subroutine)
  const float dd = mul_rn(t.d1, t.d2);
  t.rD = fast_rcp(dd);
  t.S = kExactDiv ? div_rn(mul_rn(t.n1, t.n2), dd) : mul_rn(mul_rn(t.n1, t.n2), t.rD);
  return t;
}

UGL_HD float ssim_from_sums(const Moments& m) { return ssim_terms<true>(m).S; }

// loss value clamp((1-S)/2, 0, 1)
UGL_HD float ssim_loss_value(float S) {
  const float v = mul_rn(sub_rn(1.0f, S), 0.5f);   // /2 is exact
  return v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
}
// d S / d(mu_x), d(E[x^2]), d(mu_y), d(E[y^2]), d(E[xy]) of one window, scaled by g.
UGL_HD void ssim_partials(const SsimTerms& t, float g, float& ax, float& bx, float& ay, float& by, float& cxy) {
  // explicit operation order (no compiler contraction): the packed variant ssim_partials2 below performs the same roundings,
  // so the single-pass, recompute and per-method kernels agree on these cancellation-prone coefficients
  const float invD = mul_rn(g, t.rD), invD2 = add_rn(invD, invD);
  const float nS = -t.S, dn = sub_rn(t.n2, t.n1), dd = sub_rn(t.d2, t.d1);
  ax = mul_rn(fma_rn(t.my, dn, mul_rn(mul_rn(nS, t.mx), dd)), invD2);
  ay = mul_rn(fma_rn(t.mx, dn, mul_rn(mul_rn(nS, t.my), dd)), invD2);
  bx = by = mul_rn(mul_rn(nS, t.d1), invD);
  cxy = mul_rn(t.n1, invD2);
}
// Given window sums, produce g * dS/d(mu_y), g * dS/d(E[y^2]), g * dS/d(E[xy]) where g = d loss / dS
// (= -1/2 inside the clamp range [0,1] inclusive, 0 outside — torch.clamp backward).
UGL_HD void ssim_backward_coeffs(const Moments& m, float& cA, float& cB, float& cC) {
  const SsimTerms t = ssim_terms<true>(m);
  const float v = mul_rn(sub_rn(1.0f, t.S), 0.5f);
  const float g = (v >= 0.f && v <= 1.f) ? -0.5f : 0.f;
  float ax, bx;
  ssim_partials(t, g, ax, bx, cA, cB, cC);
}

// ---- the same SSIM terms for a pair of windows (the two warp directions of one window position): identical roundings,
// packed instructions.  n1 = 2 mx my + C1 and n2 = 2 cxy + C2 use one FMA each: the doubling is exact, so the single
// rounding equals the scalar add_rn(mul_rn(2, .), C).
struct Moments2 { float2 sx, sy, sxx, syy, sxy; };
struct SsimTerms2 { float2 mx, my, n1, n2, d1, d2, S, rD; };

UGL_HD SsimTerms2 ssim_terms2(const Moments2& m, float2 one) {
  SsimTerms2 t;
  constexpr float r9 = 1.0f / 9.0f;
  t.mx = div_c2(m.sx, 9.0f, r9);
  t.my = div_c2(m.sy, 9.0f, r9);
  const float2 mxx = mul2(t.mx, t.mx), myy = mul2(t.my, t.my), mxy = mul2(t.mx, t.my);
  // differences and sums whose operand is a bare product go through the opaque 1.0 (see acc2_rn): E - mxx must not become
  // fma(-mx, mx, E)
  const float2 vx = sub2_rn(div_c2(m.sxx, 9.0f, r9), mxx, one);
  const float2 vy = sub2_rn(div_c2(m.syy, 9.0f, r9), myy, one);
  const float2 cxy = sub2_rn(div_c2(m.sxy, 9.0f, r9), mxy, one);
  t.n1 = acc2_rn(splat2(kC1), mul2(splat2(2.0f), mxy), one);
  t.n2 = fma2(splat2(2.0f), cxy, splat2(kC2));
  t.d1 = add2(acc2_rn(myy, mxx, one), splat2(kC1));     // mxx + myy: a sum of products
  t.d2 = add2(add2(vx, vy), splat2(kC2));
  const float2 dd = mul2(t.d1, t.d2);
  t.rD = make_float2(fast_rcp(dd.x), fast_rcp(dd.y));
  t.S = mul2(mul2(t.n1, t.n2), t.rD);
  return t;
}
// loss value before the clamp, (1 - S) / 2, of both windows (S is a bare product: opaque subtraction)
UGL_HD float2 ssim_half_one_minus2(float2 S, float2 one) { return mul2(sub2_rn(splat2(1.0f), S, one), splat2(0.5f)); }
// g * dS/d(mu_y), g * dS/d(E[y^2]), g * dS/d(E[xy]) of both windows (ssim_partials' ay, by, cxy)
UGL_HD void ssim_partials2(const SsimTerms2& t, float2 g, float2& cA, float2& cB, float2& cC) {
  const float2 invD = mul2(g, t.rD), invD2 = mul2(invD, splat2(2.0f));
  const float2 nS = mul2(t.S, splat2(-1.0f));
  const float2 e = fma2(t.mx, sub2(t.n2, t.n1), mul2(mul2(nS, t.my), sub2(t.d2, t.d1)));
  cA = mul2(e, invD2);
  cB = mul2(mul2(nS, t.d1), invD);
  cC = mul2(t.n1, invD2);
}

// ---- dynamic mask (model_geometry.py:698-707): n(x) = sqrt(x0^2 + x1^2) + 1e-12 ; dyn = [n(|rf-f|)^2 < alpha (n(f)^2 + n(rf)^2) + beta]
UGL_HD float norm2_eps(float u, float v) { return add_rn(sqrt_rn(add_rn(mul_rn(u, u), mul_rn(v, v))), 1e-12f); }
UGL_HD float dynamic_mask_value(float fu, float fv, float ru, float rv, float alpha, float beta) {
  const float nf = norm2_eps(fu, fv), nr = norm2_eps(ru, rv);
  const float bound = add_rn(mul_rn(alpha, add_rn(mul_rn(nf, nf), mul_rn(nr, nr))), beta);
  const float nd = norm2_eps(fabsf(sub_rn(ru, fu)), fabsf(sub_rn(rv, fv)));
  return mul_rn(nd, nd) < bound ? 1.f : 0.f;
}

// ---- soft / hard occlusion weights (model_flow.py:105-138, model_geometry.py:105-132) ----------
// wgt = 1 - softmax([d_l, d_r]) evaluated like ATen's softmax: exp(x - max) / sum.
UGL_HD void one_minus_softmax2(float d_l, float d_r, float& wl, float& wr) {
  const float m = d_l > d_r ? d_l : d_r;
  const float el = expf(sub_rn(d_l, m)), er = expf(sub_rn(d_r, m));
  const float s = add_rn(el, er);
  wl = sub_rn(1.0f, div_rn(el, s));
  wr = sub_rn(1.0f, div_rn(er, s));
}
UGL_HD float soft_occ_weight(float wgt) {   // 2*exp(-(wgt-0.5)^2/0.03)
  const float c = sub_rn(wgt, 0.5f);
  return 2.0f * expf(div_c(-mul_rn(c, c), 0.03f, 1.0f / 0.03f));
}

}  // namespace ugl
