__global__ void k(unsigned long long* out, const int* idx, const long long* q) {
  __shared__ unsigned long long w[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) w[i] = 0;
  __syncthreads();
  atomicAdd(&w[idx[threadIdx.x]], (unsigned long long)q[threadIdx.x]);
  __syncthreads();
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) if (w[i]) atomicAdd(out + i, w[i]);
}
