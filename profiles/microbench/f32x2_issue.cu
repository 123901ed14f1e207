// Micro-benchmark: does the packed fp32 pipe of sm_100a (FADD2 / FMUL2 / FFMA2, add.rn.f32x2 ...) relieve an
// instruction-issue-bound kernel?  Four loops with the same number of IEEE-rounded fp32 results per thread:
//   scalar      : N x FADD                          (8 independent chains)
//   packed      : N/2 x FADD2
//   scalar+int  : N x FADD  + N x integer ALU (one LOP3 + one IADD per two fp results) independent of the fp chains
//   packed+int  : N/2 x FADD2 + N x integer ALU
// Prints fp32 results per clock per SM and warp-instructions per clock per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_issue f32x2_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, unsigned* iout, int iters, float seed) {
  float2 a[4];
  unsigned q[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { a[i] = make_float2(seed + i + threadIdx.x, seed * 2 + i); q[i] = threadIdx.x * 7 + i; }
  const float2 inc = make_float2(seed * 0.5f, seed * 0.25f);
  const unsigned kx = (unsigned)iters * 0x55u, ky = (unsigned)iters + 0x9e3779b9u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (MODE == 0 || MODE == 2) { a[i].x = __fadd_rn(a[i].x, inc.x); a[i].y = __fadd_rn(a[i].y, inc.y); }
        else a[i] = __fadd2_rn(a[i], inc);
        if (MODE >= 2) { q[i] = (q[i] ^ kx) + ky; }
      }
    }
  }
  float s = 0.f; unsigned t = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) { s += a[i].x + a[i].y; t += q[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE>
void run(const char* name, float* out, unsigned* iout, int sms) {
  const int iters = 4096, blocks = sms * 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(out, iout, 16, 1.0f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, iout, iters, 1.0f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double clk = ms * 1e-3 * khz * 1e3;                      // at the maximum clock (an upper bound of cycles)
  const double results = (double)blocks * 256 * iters * 8 * 4 * 2;  // fp32 results
  printf("%-12s %8.3f ms  fp32 results/clk/SM (at max clock) %7.1f\n", name, ms, results / clk / sms);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; unsigned* iout;
  cudaMalloc(&out, sms * 8 * 256 * 4); cudaMalloc(&iout, sms * 8 * 256 * 4);
  run<0>("scalar", out, iout, sms);
  run<1>("packed", out, iout, sms);
  run<2>("scalar+int", out, iout, sms);
  run<3>("packed+int", out, iout, sms);
  return 0;
}
