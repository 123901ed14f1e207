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

// Programmatic dependent launch (griddepcontrol.*; sm_90+).  A kernel launched with `pdl` may start while its stream predecessor is still
// running; it must call griddep_wait() before it touches anything the predecessor (or, transitively, an earlier kernel of the chain)
// wrote.  Without the attribute both calls are no-ops.
#ifndef UGL_PDL
#define UGL_PDL 1
#endif
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline int launch_kernel(const char* what, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (pdl && UGL_PDL) ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
  return UGL_OK;
}

inline bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

// Warp reduction of N per-thread accumulators by recursive halving: at shuffle distance o the lanes with bit o clear keep
// the first half of the values and the lanes with bit o set the second half, each adding what its partner sends, so N values
// need about N shuffles in total (N/2 + N/4 + ...) instead of 5 N (12 accumulators: 13 instead of 60; SHFL issues at a quarter
// of the ALU rate).  An unpaired middle value is reduced by a plain butterfly step.  After the five steps every lane holds the
// complete warp sum of ONE value, whose index `which` is tracked alongside; lanes holding the same index hold the same bits
// (a + b and b + a are the same rounding).  Fixed order: deterministic.
template <int N>
__device__ __forceinline__ void warp_halving_step(float (&v)[N], int (&idx)[N], int o, bool upper) {
  constexpr int half = (N + 1) / 2;
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const float send = upper ? v[i] : v[i + half];
    const float keep = upper ? v[i + half] : v[i];
    idx[i] = upper ? idx[i + half] : idx[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
  }
  if (N & 1) v[half - 1] += __shfl_xor_sync(0xffffffffu, v[half - 1], o);
}

template <int N>
__device__ __forceinline__ float warp_reduce_n(float (&acc)[N], int& which) {
  const int lane = threadIdx.x & 31;
  int idx[N];
#pragma unroll
  for (int k = 0; k < N; ++k) idx[k] = k;
  constexpr int n1 = (N + 1) / 2, n2 = (n1 + 1) / 2, n3 = (n2 + 1) / 2, n4 = (n3 + 1) / 2;
  warp_halving_step<N>(acc, idx, 16, (lane & 16) != 0);
  float (&a1)[n1] = reinterpret_cast<float (&)[n1]>(acc); int (&i1)[n1] = reinterpret_cast<int (&)[n1]>(idx);
  warp_halving_step<n1>(a1, i1, 8, (lane & 8) != 0);
  float (&a2)[n2] = reinterpret_cast<float (&)[n2]>(acc); int (&i2)[n2] = reinterpret_cast<int (&)[n2]>(idx);
  warp_halving_step<n2>(a2, i2, 4, (lane & 4) != 0);
  float (&a3)[n3] = reinterpret_cast<float (&)[n3]>(acc); int (&i3)[n3] = reinterpret_cast<int (&)[n3]>(idx);
  warp_halving_step<n3>(a3, i3, 2, (lane & 2) != 0);
  float (&a4)[n4] = reinterpret_cast<float (&)[n4]>(acc); int (&i4)[n4] = reinterpret_cast<int (&)[n4]>(idx);
  warp_halving_step<n4>(a4, i4, 1, (lane & 1) != 0);
  static_assert(N <= 32, "one value per lane at most");
  which = idx[0];
  return acc[0];
}

// warp-halving + shared-memory block reduction of N per-thread accumulators, fixed order (deterministic).
// Result valid in threads 0..N-1 of the block (each holds one value).  acc is clobbered.
template <int NT, int N>
__device__ __forceinline__ float block_reduce_n(float (&acc)[N], float* red /* [NT/32][N] */) {
  const int warp = threadIdx.x >> 5;
  int which;
  const float v = warp_reduce_n<N>(acc, which);
  red[warp * N + which] = v;          // lanes holding the same value index write identical bits
  __syncthreads();
  float out = 0.f;
  if (threadIdx.x < N) {
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) out += red[w * N + threadIdx.x];
  }
  return out;
}

}  // namespace ugl
