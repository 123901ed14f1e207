// Deterministic scatter-accumulate for the "gradient w.r.t. the sampled image" of bilinear sampling.
// ATen's grid_sampler_2d_backward uses floating-point atomicAdd (order dependent => run-to-run
// differences).  Here every contribution is converted to 64-bit fixed point with a scale 2^e chosen
// from max|grad| and the worst-case number of contributions, and accumulated with integer atomics:
// integer addition is associative, so the result is independent of the arrival order.
#pragma once

#include "ugl_common.cuh"
#include "ugl_host.cuh"

namespace ugl {

#if defined(__CUDACC__)
// max |g| as the bit pattern of a non-negative float (unsigned order == float order)
static __global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ g, long n, unsigned* __restrict__ out) {
  float m = 0.f;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
    const float a = fabsf(g[idx]);
    m = (a > m && a <= 3.0e38f) ? a : m;   // ignores NaN/inf (they poison the result anyway)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

static __global__ void __launch_bounds__(256)
fixed_to_float_kernel(const unsigned long long* __restrict__ acc, long n, long n_contrib, const unsigned* __restrict__ maxbits,
                      float* __restrict__ out) {
  const int e = fixed_point_exponent(__uint_as_float(*maxbits), n_contrib);
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x)
    out[idx] = (float)ldexp((double)(long long)acc[idx], -e);
}

// add g * (bilinear weights) to the four corners of a tap in one fixed-point plane
__device__ __forceinline__ void scatter_tap(unsigned long long* __restrict__ plane, int W, const Tap& t, float g, int e) {
  const float wgt[4] = {t.wnw, t.wne, t.wsw, t.wse};
  const long off[4] = {0, 1, W, (long)W + 1};
  const long base = (long)t.y0 * W + t.x0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (t.inb & (1u << k)) {
      const long long q = __double2ll_rn(ldexp((double)(g * wgt[k]), e));
      atomicAdd(plane + base + off[k], (unsigned long long)q);
    }
  }
}

inline int scatter_grid(long n) {
  long g = (n + 255) / 256;
  const long cap = 148L * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
#endif

}  // namespace ugl
