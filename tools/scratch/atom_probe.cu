// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/scratch/atom_probe tools/scratch/atom_probe.cu
// probe: shared-memory histogram update cost, float CAS-loop atomics vs native int32 atomics, attention-bwd access pattern
#include <cstdio>
#include <cuda_runtime.h>
constexpr int W = 47, NB = W * W;
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(long long* out, float* sink) {
  __shared__ float hf[NB + 8];
  __shared__ int hi[NB + 8];
  __shared__ int kinfo[128];
  const int t = threadIdx.x;
  for (int e = t; e < NB + 8; e += 512) { hf[e] = 0.f; hi[e] = 0; }
  if (t < 128) { const int pid = 256 + t; kinfo[t] = (pid / 24) * W + (pid % 24); }
  __syncthreads();
  const int r = (t & 127), col0 = (t >> 7) * 32;
  float dsv[32];
  for (int j = 0; j < 32; ++j) dsv[j] = 1e-3f * ((t * 37 + j * 11) % 97 - 48);
  long long t0 = clock64();
  for (int rep = 0; rep < 8; ++rep) {
    const int pid = (rep * 128 + r) % 576;
    const int rowbase = (pid / 24 + 23) * W + (pid % 24 + 23);
    if (MODE == 0) {
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) atomicAdd(&hf[rowbase - kinfo[col0 + jj]], dsv[jj]);
    } else {
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) atomicAdd(&hi[rowbase - kinfo[col0 + jj]], __float2int_rn(dsv[jj] * 1048576.f));
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (t == 0) out[blockIdx.x] = t1 - t0;
  float s = 0.f;
  for (int e = t; e < NB; e += 512) s += hf[e] + (float)hi[e];
  sink[blockIdx.x * 512 + t] = s;
}
int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
  long long h[148];
  for (int m = 0; m < 2; ++m) {
    for (int it = 0; it < 2; ++it) {
      if (m == 0) k<0><<<148, 512>>>(out, sink); else k<1><<<148, 512>>>(out, sink);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%s: %.0f cycles per 128x128 tile (%s)\n", m == 0 ? "float CAS" : "int native", h[0] / 8.0, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
