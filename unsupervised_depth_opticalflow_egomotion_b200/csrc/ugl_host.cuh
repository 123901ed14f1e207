// Host-side helpers shared by the C-ABI translation units: thread-local error text, launch
// checks, argument validation.  No allocation, no global mutable state besides the error string.
#pragma once

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/ugl.h"

namespace ugl {

char* last_error_buffer();   // thread-local, 512 bytes (defined in ugl_api.cu)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

inline int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
  return UGL_OK;
}

// Opt a kernel in to more than 48 KB of dynamic shared memory.  Idempotent and cheap; it is not a stream
// operation, so it is legal while the stream is being captured into a CUDA graph.
template <typename K>
inline int opt_in_smem(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return UGL_OK;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(%zu bytes): %s", bytes, cudaGetErrorString(e));
  return UGL_OK;
}

inline bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

// warp-shuffle + shared-memory block reduction of N per-thread accumulators, fixed order
// (deterministic).  Result valid in threads 0..N-1 of the block (each holds one value).
template <int NT, int N>
__device__ __forceinline__ float block_reduce_n(float (&acc)[N], float* red /* [NT/32][N] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp * N + k] = v;
  }
  __syncthreads();
  float out = 0.f;
  if (threadIdx.x < N) {
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) out += red[w * N + threadIdx.x];
  }
  return out;
}

}  // namespace ugl
