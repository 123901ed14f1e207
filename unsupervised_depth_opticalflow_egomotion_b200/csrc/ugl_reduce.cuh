// Deterministic per-sample reductions shared by the loss-term kernels.
//
// Every loss of the reference is a per-sample mean (`mean((1,2,3))`), so the pattern is always:
//   pass 1: grid (chunks, B); each CTA folds a strided slice of one sample's pixels into NACC
//           per-thread accumulators, reduces them in a fixed order (shuffle tree + shared memory)
//           and writes one partial row -> partials[b][chunk][NACC]
//   pass 2: one warp per sample sums the chunk partials in fp64 in a fixed order and applies the
//           term's closing formula.
// No floating-point atomics anywhere: results are bit-reproducible run to run.
#pragma once

#include "ugl_host.cuh"

namespace ugl {

constexpr int kRedThreads = 256;
constexpr int kMaxChunks = 148;   // one wave of CTAs per sample row is plenty for 2e5 pixels

inline int reduce_chunks(long npix) {
  long c = (npix + (long)kRedThreads * 4 - 1) / ((long)kRedThreads * 4);
  return (int)(c < 1 ? 1 : (c > kMaxChunks ? kMaxChunks : c));
}

template <int NACC, class PixelFn>
__global__ void __launch_bounds__(kRedThreads) sample_reduce_kernel(PixelFn fn, long npix, float* __restrict__ partials) {
  __shared__ float red[(kRedThreads / 32) * NACC];
  const int b = blockIdx.y;
  float acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
  for (long p = blockIdx.x * (long)kRedThreads + threadIdx.x; p < npix; p += (long)gridDim.x * kRedThreads) fn(b, p, acc);
  const float v = block_reduce_n<kRedThreads, NACC>(acc, red);
  if (threadIdx.x < NACC) partials[((long)b * gridDim.x + blockIdx.x) * NACC + threadIdx.x] = v;
}

// one warp per sample; FinFn(b, const double* S) runs on lane 0
template <int NACC, class FinFn>
__global__ void sample_finalize_kernel(const float* __restrict__ partials, int chunks, int B, FinFn fin) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  double s[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) s[k] = 0.0;
  for (int c = lane; c < chunks; c += 32) {
#pragma unroll
    for (int k = 0; k < NACC; ++k) s[k] += (double)partials[((long)b * chunks + c) * NACC + k];
  }
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_down_sync(0xffffffffu, s[k], o);
  }
  if (lane == 0) fin(b, s);
}

template <int NACC, class PixelFn, class FinFn>
int launch_sample_reduce(PixelFn fn, FinFn fin, int B, long npix, void* workspace, uint64_t workspace_bytes, cudaStream_t st,
                         const char* what) {
  const int chunks = reduce_chunks(npix);
  const uint64_t need = (uint64_t)B * chunks * NACC * sizeof(float);
  if (!workspace || workspace_bytes < need) return fail(UGL_EWORKSPACE, "%s: workspace too small (%llu < %llu)", what,
                                                        (unsigned long long)workspace_bytes, (unsigned long long)need);
  float* partials = static_cast<float*>(workspace);
  sample_reduce_kernel<NACC><<<dim3(chunks, B), kRedThreads, 0, st>>>(fn, npix, partials);
  int rc = check_launch(what);
  if (rc) return rc;
  sample_finalize_kernel<NACC><<<(B + 3) / 4, 128, 0, st>>>(partials, chunks, B, fin);
  return check_launch(what);
}

inline uint64_t reduce_workspace_bytes(int B, long npix, int nacc) {
  return (uint64_t)B * reduce_chunks(npix) * nacc * sizeof(float);
}

// plain grid-stride element-wise launch
template <class Fn>
__global__ void __launch_bounds__(kRedThreads) pointwise_kernel(Fn fn, long n) {
  for (long p = blockIdx.x * (long)kRedThreads + threadIdx.x; p < n; p += (long)gridDim.x * kRedThreads) fn(p);
}

template <class Fn>
int launch_pointwise(Fn fn, long n, cudaStream_t st, const char* what) {
  long g = (n + kRedThreads - 1) / kRedThreads;
  const long cap = 148L * 8;
  pointwise_kernel<<<(int)(g < 1 ? 1 : (g > cap ? cap : g)), kRedThreads, 0, st>>>(fn, n);
  return check_launch(what);
}

}  // namespace ugl
