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

// ---- tile-local form --------------------------------------------------------------------------------------------------------
// A CTA owns a TW x TH tile of SOURCE pixels.  Warps are local: most taps land within a few pixels of their source, so the CTA
// accumulates them in a shared-memory window (the tile plus a margin of M pixels) and sends each touched window cell to the global
// 64-bit plane ONCE; only taps that leave the window use a global atomic of their own.  The arithmetic is the same 64-bit fixed point
// as scatter_tap (integer sums: the split into window and global adds cannot change the result, bit for bit).  Shared memory has no
// native 64-bit add (ATOMS.CAST.SPIN loop), so a cell is two 32-bit words: `lo` wraps, and the ONE add that observes the wrap carries
// into `hi`: sum = hi * 2^32 + lo exactly, whatever the order.
template <int TW, int TH, int M, int CP>
struct ScatterWindow {
  static constexpr int WW = TW + 2 * M, WH = TH + 2 * M, N = WW * WH;   // window cells per channel
  static constexpr int kWords = 2 * CP * N;                             // lo[CP][N] then hi[CP][N]

  __device__ static __forceinline__ long long to_fixed(float v, int e) { return __double2ll_rn(ldexp((double)v, e)); }

  // all four corners of the tap inside the window whose cell (0,0) is image pixel (wx0, wy0)?
  __device__ static __forceinline__ bool local(const Tap& t, int wx0, int wy0) {
    const int wx = t.x0 - wx0, wy = t.y0 - wy0;
    return wx >= 0 && wx < WW - 1 && wy >= 0 && wy < WH - 1;
  }

  __device__ static __forceinline__ void add_local(unsigned* __restrict__ win, int k, int wx0, int wy0, const Tap& t, float g, int e) {
    const float wgt[4] = {t.wnw, t.wne, t.wsw, t.wse};
    const int off[4] = {0, 1, WW, WW + 1};
    unsigned* lo = win + k * N + (t.y0 - wy0) * WW + (t.x0 - wx0);
    unsigned* hi = lo + CP * N;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (t.inb & (1u << c)) {
        const long long q = to_fixed(g * wgt[c], e);
        if (q == 0) continue;
        const unsigned ql = (unsigned)q, qh = (unsigned)((unsigned long long)q >> 32);
        const unsigned old = atomicAdd(lo + off[c], ql);
        atomicAdd(hi + off[c], qh + ((unsigned)(old + ql) < ql ? 1u : 0u));
      }
    }
  }

  __device__ static __forceinline__ void clear(unsigned* __restrict__ win, int tid, int nt) {
    for (int idx = tid; idx < kWords; idx += nt) win[idx] = 0u;
  }

  // send the touched cells of channels [0, ncp) to their global planes (plane k = plane0 + k * plane_stride) and clear them
  __device__ static __forceinline__ void flush(unsigned* __restrict__ win, int ncp, int wx0, int wy0, unsigned long long* __restrict__ plane0,
                                               long plane_stride, int W, int tid, int nt) {
    for (int idx = tid; idx < ncp * N; idx += nt) {
      const unsigned lo = win[idx], hi = win[CP * N + idx];
      if ((lo | hi) == 0u) continue;
      const int k = idx / N, r = idx - k * N, wy = r / WW, wx = r - wy * WW;
      atomicAdd(plane0 + k * plane_stride + (long)(wy0 + wy) * W + (wx0 + wx), ((unsigned long long)hi << 32) | lo);
      win[idx] = 0u;
      win[CP * N + idx] = 0u;
    }
  }
};

inline int scatter_grid(long n) {
  long g = (n + 255) / 256;
  const long cap = 148L * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
#endif

}  // namespace ugl
