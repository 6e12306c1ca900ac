// mma.sync m16n8k16 bf16 issue interval / latency on sm_100a: W warps per CTA (one CTA per SM), each running a chain
// of dependent MMAs on A independent accumulators.  Prints cycles per MMA per warp.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int A>
__global__ void probe(int iters, long long* out, float* sink) {
  float acc[A][4];
  uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
  uint32_t b0 = 0x3c003c00u + threadIdx.x, b1 = 0x3c003c00u;
#pragma unroll
  for (int i = 0; i < A; ++i)
    for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < A; ++i)
      asm volatile(
          "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
          : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
          : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < A; ++i)
    for (int q = 0; q < 4; ++q) s += acc[i][q];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int A>
void run(int warps, int ctas) {
  long long* d;
  float* sink;
  cudaMalloc(&d, 8);
  cudaMalloc(&sink, sizeof(float) * ctas * warps * 32);
  const int iters = 2000;
  probe<A><<<ctas, warps * 32>>>(iters, d, sink);
  probe<A><<<ctas, warps * 32>>>(iters, d, sink);
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("accumulators %d  warps/CTA %2d  ctas %3d : %.1f cycles per MMA per warp, %.2f cycles per MMA per SM\n", A, warps,
         ctas, (double)h / (iters * A), (double)h / (iters * A * warps));
  cudaFree(d);
  cudaFree(sink);
}

int main() {
  for (int warps : {1, 4, 8, 16}) {
    run<1>(warps, 148);
    run<2>(warps, 148);
    run<4>(warps, 148);
    run<8>(warps, 148);
  }
  run<4>(8, 32);
  return 0;
}
