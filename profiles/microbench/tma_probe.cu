// standalone probe: which rank-3 fp32 TMA boxes are legal on sm_100a
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#include <stdlib.h>
#include <string.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct Maps { CUtensorMap m[8]; };
__global__ void probe(const __grid_constant__ Maps maps, int which, int x, int y, int z, int nbytes, int dst_off, float* out, int n) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < n + dst_off; i += blockDim.x) sm[i] = -7.f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nbytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(sm + dst_off)), "l"(reinterpret_cast<uint64_t>(&maps.m[which])), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  } while (!ok);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = sm[dst_off + i];
}
int main(int argc, char** argv) {
  int only = argc > 1 ? atoi(argv[1]) : -1; int ci = -1;
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const int W = 208, H = 64, P = 6;
  std::vector<float> h((size_t)W * H * P);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100000);
  float *d, *out; cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 1 << 16);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  struct Cfg { int bw, bh, x, y, off; } cfgs[] = {{72, 17, -4, -2, 0}, {36, 17, 0, 0, 0}, {40, 17, -2, -2, 0}, {44, 17, -2, -2, 0}, {48, 17, -2, -2, 0}, {64, 13, 0, 0, 0}, {40, 17, -2, -2, 1248}, {40,17,-3,-2,32}, {36,17,-2,-2,0}, {68,17,-2,-2,0}, {8,17,-2,-2,0},{4,17,-2,-2,0},{12,17,-2,-2,0}};
  for (auto c : cfgs) {
    ++ci; if (only >= 0 && ci != only) continue;
    Maps maps; memset(&maps, 0, sizeof(maps));
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P}; cuuint64_t str[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)c.bw, (cuuint32_t)c.bh, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&maps.m[3], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int n = c.bw * c.bh;
    probe<<<1, 128, (n + c.off) * 4 + 256>>>(maps, 3, c.x, c.y, 1, n * 4, c.off, out, n);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(n);
    cudaMemcpy(o.data(), out, n * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int yy = 0; yy < c.bh; ++yy) for (int xx = 0; xx < c.bw; ++xx) {
      const int gx = c.x + xx, gy = c.y + yy;
      const float want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[(size_t)1 * W * H + (size_t)gy * W + gx] : 0.f;
      if (o[yy * c.bw + xx] != want) ++bad;
    }
    printf("box %dx%d at (%d,%d) dst_off %d: encode %d, run %s, mismatches %d\n", c.bw, c.bh, c.x, c.y, c.off, (int)r, cudaGetErrorString(e), bad);
    if (e != cudaSuccess) break;
  }
  return 0;
}
