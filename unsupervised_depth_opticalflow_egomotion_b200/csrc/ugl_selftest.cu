// Device self-test: the packed fp32 pair forms of the SSIM arithmetic against the scalar forms, bit for bit.
#include "ugl_common.cuh"
#include "ugl_host.cuh"

namespace ugl {

__device__ __forceinline__ float st_uniform(unsigned& s) {
  s = s * 1664525u + 1013904223u;
  return (float)(s >> 8) * (1.0f / 16777216.0f);
}

__global__ void __launch_bounds__(128) selftest_packed_pairs_kernel(unsigned long long* __restrict__ mism, int iters, float onev) {
  const float2 one = splat2(onev);      // the opaque 1.0 of acc2_rn / sub2_rn (a kernel parameter)
  unsigned s = blockIdx.x * blockDim.x + threadIdx.x + 12345u;
  unsigned long long bad[2][14];
#pragma unroll
  for (int d = 0; d < 2; ++d)
#pragma unroll
    for (int f = 0; f < 14; ++f) bad[d][f] = 0ull;
  for (int it = 0; it < iters; ++it) {
    float x[2][9], y[2][9];
    const float base = st_uniform(s), w0 = st_uniform(s) * 2.f, w1 = st_uniform(s) * 2.f;
    for (int t = 0; t < 9; ++t) {
      const float I = base + 0.05f * st_uniform(s), W0 = I + 0.02f * (st_uniform(s) - 0.5f), W1 = I + 0.1f * (st_uniform(s) - 0.5f);
      x[0][t] = mul_rn(I, w0); y[0][t] = mul_rn(W0, w0); x[1][t] = mul_rn(I, w1); y[1][t] = mul_rn(W1, w1);
    }
    Moments m[2];
    Moments2 m2;
    for (int d = 0; d < 2; ++d) {
      m[d] = Moments{0.f, 0.f, 0.f, 0.f, 0.f};
      for (int t = 0; t < 9; ++t) moments_add(m[d], x[d][t], y[d][t]);
    }
    for (int t = 0; t < 9; ++t) {
      const float2 X = make_float2(x[0][t], x[1][t]), Y = make_float2(y[0][t], y[1][t]);
      const float2 XX = mul2(X, X), YY = mul2(Y, Y), XY = mul2(X, Y);
      if (t == 0) { m2.sx = X; m2.sy = Y; m2.sxx = XX; m2.syy = YY; m2.sxy = XY; }
      else {
        m2.sx = add2(m2.sx, X); m2.sy = add2(m2.sy, Y);
        m2.sxx = acc2_rn(m2.sxx, XX, one); m2.syy = acc2_rn(m2.syy, YY, one); m2.sxy = acc2_rn(m2.sxy, XY, one);
      }
    }
    const SsimTerms a0 = ssim_terms<false>(m[0]), a1 = ssim_terms<false>(m[1]);
    const SsimTerms2 b = ssim_terms2(m2, one);
    float ax, bx, ay0, by0, c0, ay1, by1, c1;
    float2 cA, cB, cC;
    ssim_partials(a0, -0.5f, ax, bx, ay0, by0, c0);
    ssim_partials(a1, -0.5f, ax, bx, ay1, by1, c1);
    ssim_partials2(b, splat2(-0.5f), cA, cB, cC);
    const float sc[2][14] = {{m[0].sx, m[0].sy, m[0].sxx, m[0].syy, m[0].sxy, a0.mx, a0.n1, a0.n2, a0.d1, a0.d2, a0.S, ay0, by0, c0},
                             {m[1].sx, m[1].sy, m[1].sxx, m[1].syy, m[1].sxy, a1.mx, a1.n1, a1.n2, a1.d1, a1.d2, a1.S, ay1, by1, c1}};
    const float pk[2][14] = {{m2.sx.x, m2.sy.x, m2.sxx.x, m2.syy.x, m2.sxy.x, b.mx.x, b.n1.x, b.n2.x, b.d1.x, b.d2.x, b.S.x, cA.x, cB.x, cC.x},
                             {m2.sx.y, m2.sy.y, m2.sxx.y, m2.syy.y, m2.sxy.y, b.mx.y, b.n1.y, b.n2.y, b.d1.y, b.d2.y, b.S.y, cA.y, cB.y, cC.y}};
#pragma unroll
    for (int d = 0; d < 2; ++d)
#pragma unroll
      for (int f = 0; f < 14; ++f) bad[d][f] += (__float_as_uint(sc[d][f]) != __float_as_uint(pk[d][f])) ? 1ull : 0ull;
  }
#pragma unroll
  for (int d = 0; d < 2; ++d)
#pragma unroll
    for (int f = 0; f < 14; ++f)
      if (bad[d][f]) atomicAdd(&mism[d * 14 + f], bad[d][f]);      // integer atomics: order-independent
}

}  // namespace ugl

extern "C" int ugl_selftest_packed_pairs(uint64_t* mismatch, int32_t blocks, int32_t windows_per_thread, void* stream) {
  using namespace ugl;
  if (!mismatch) return fail(UGL_EINVAL, "selftest_packed_pairs: null mismatch buffer");
  if (blocks <= 0 || windows_per_thread <= 0) return fail(UGL_EINVAL, "selftest_packed_pairs: blocks and windows_per_thread must be positive");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cudaError_t e = cudaMemsetAsync(mismatch, 0, 28 * sizeof(uint64_t), st);
  if (e != cudaSuccess) return fail((int)e, "selftest_packed_pairs: memset: %s", cudaGetErrorString(e));
  selftest_packed_pairs_kernel<<<blocks, 128, 0, st>>>(reinterpret_cast<unsigned long long*>(mismatch), windows_per_thread, 1.0f);
  return check_launch("selftest_packed_pairs_kernel");
}
