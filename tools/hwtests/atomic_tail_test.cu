// How long does the "every CTA adds its per-column partial sums to the same global addresses" tail of the statistics
// kernels take?  G CTAs, each adds C values to the same C addresses at the end (after a grid-wide-ish delay so that
// they arrive together).  Variants: fp64 atomicAdd, fp32 atomicAdd, fp32 red.v4.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomic_tail_test atomic_tail_test.cu && ./atomic_tail_test
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void tail_kernel(double* d64, float* d32, int C, int spin) {
  // busy phase so that all CTAs are resident and finish at about the same time
  float x = threadIdx.x;
  for (int i = 0; i < spin; ++i) x = x * 1.0001f + 0.5f;
  if (x == 123.456f) d32[0] = x;
  __syncthreads();
  if (MODE == 0) { for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(d64 + c, 1.0); }
  if (MODE == 1) { for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(d32 + c, 1.0f); }
  if (MODE == 2) {
    for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d32 + c), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
  }
  if (MODE == 3) { /* no tail */ }
}

template <int MODE>
float run(int G, int C, double* d64, float* d32) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  tail_kernel<MODE><<<G, 256>>>(d64, d32, C, 2000);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 20; ++r) tail_kernel<MODE><<<G, 256>>>(d64, d32, C, 2000);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 20 * 1e3f;
}

int main() {
  double* d64; float* d32;
  cudaMalloc(&d64, 8 * 4096); cudaMalloc(&d32, 4 * 4096);
  cudaMemset(d64, 0, 8 * 4096); cudaMemset(d32, 0, 4 * 4096);
  const int Gs[] = {148, 296, 592, 1184};
  const int Cs[] = {16, 64, 256, 1024};
  for (int G : Gs)
    for (int C : Cs) {
      const float t3 = run<3>(G, C, d64, d32);
      printf("G=%4d C=%4d  base %6.1f us | fp64 +%6.1f | fp32 +%6.1f | red.v4.f32 +%6.1f\n", G, C, t3, run<0>(G, C, d64, d32) - t3,
             run<1>(G, C, d64, d32) - t3, run<2>(G, C, d64, d32) - t3);
    }
  return 0;
}
