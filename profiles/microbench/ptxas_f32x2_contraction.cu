// Evidence for DESIGN.md 4.1: ptxas 12.9 contracts `mul.rn.f32x2` + `add.rn.f32x2` into one FFMA2 although both carry .rn.
// Five ways of writing acc += x * y on packed pairs; compile and count the packed opcodes per variant:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cubin -o f.cubin ptxas_f32x2_contraction.cu
//   for v in 1 2 3 4 5; do cuobjdump -sass -fun "_Z1kILi${v}EEvPK6float2PS0_S0_" f.cubin | grep -E "F(ADD|MUL|FMA)2? " | awk '{print $2}' | sort | uniq -c; done
// Result (CUDA 12.9.86):  V1 intrinsics            3 FADD2 4 FFMA2            <- product contracted into the sum
//                         V2 fma(prod, one, acc), `one` a kernel parameter  3 FADD2 4 FFMA2 4 FMUL2   <- rounding of the product kept
//                         V3 volatile asm mul.rn.f32x2                     3 FADD2 4 FFMA2            <- contracted
//                         V4 scalar products, packed add                   7 FADD2 8 FMUL             <- kept (but twice the multiplies)
//                         V5 fma(prod, literal 1.0, acc)                   3 FADD2 4 FFMA2            <- simplified and contracted
// The kernels use V2 (ugl_common.cuh: acc2_rn / sub2_rn); `-Xptxas -fmad=false` does not change V1.
#include <cuda_runtime.h>
__device__ __forceinline__ float2 mul2v(float2 a, float2 b) {
  unsigned long long r, x = *reinterpret_cast<unsigned long long*>(&a), y = *reinterpret_cast<unsigned long long*>(&b);
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
  return *reinterpret_cast<float2*>(&r);
}
template <int V>
__global__ void k(const float2* a, float2* o, float2 one) {
  float2 acc = a[threadIdx.x], x = a[threadIdx.x + 32], y = a[threadIdx.x + 64];
  for (int i = 0; i < 4; ++i) {
    if (V == 1) acc = __fadd2_rn(acc, __fmul2_rn(x, y));
    if (V == 2) acc = __ffma2_rn(__fmul2_rn(x, y), one, acc);
    if (V == 3) acc = __fadd2_rn(acc, mul2v(x, y));
    if (V == 4) { float2 p = make_float2(__fmul_rn(x.x, y.x), __fmul_rn(x.y, y.y)); acc = __fadd2_rn(acc, p); }
    if (V == 5) acc = __ffma2_rn(__fmul2_rn(x, y), make_float2(1.0f, 1.0f), acc);
    x = __fadd2_rn(x, y);
  }
  o[threadIdx.x] = acc;
}
template __global__ void k<1>(const float2*, float2*, float2);
template __global__ void k<2>(const float2*, float2*, float2);
template __global__ void k<3>(const float2*, float2*, float2);
template __global__ void k<4>(const float2*, float2*, float2);
template __global__ void k<5>(const float2*, float2*, float2);
