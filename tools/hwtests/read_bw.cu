// pure-read / copy bandwidth probe: grid-stride 16-byte loads, UR loads in flight per thread
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int UR, bool WRITE>
__global__ void __launch_bounds__(256) rd(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ o,
                                          int64_t n, unsigned* sink) {
  unsigned acc = 0;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += UR * step) {
    uint4 va[UR], vb[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const int64_t j = i + u * step;
      va[u] = j < n ? a[j] : make_uint4(0, 0, 0, 0);
      if (b) vb[u] = j < n ? b[j] : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      acc += va[u].x ^ va[u].y ^ va[u].z ^ va[u].w;
      if (b) acc += vb[u].x ^ vb[u].y ^ vb[u].z ^ vb[u].w;
      if (WRITE) { const int64_t j = i + u * step; if (j < n) o[j] = va[u]; }
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}
template <int UR, bool WRITE>
void run(const char* name, int ctas_per_sm, bool two, const uint4* a, const uint4* b, uint4* o, int64_t n, unsigned* sink) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    rd<UR, WRITE><<<148 * ctas_per_sm, 256>>>(a, two ? b : nullptr, o, n, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  }
  double bytes = (double)n * 16 * ((two ? 2 : 1) + (WRITE ? 1 : 0));
  printf("%-22s UR=%d ctas/SM=%2d : %.3f ms  %.2f TB/s\n", name, UR, ctas_per_sm, ms, bytes / ms / 1e9);
}
int main() {
  const int64_t n = (int64_t)1 << 26;   // 1 GiB per array
  uint4 *a, *b, *o; unsigned* sink;
  cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMalloc(&o, n * 16); cudaMalloc(&sink, 4);
  cudaMemset(a, 1, n * 16); cudaMemset(b, 2, n * 16);
  for (int c : {2, 4, 8}) {
    run<1, false>("read 1 array", c, false, a, b, o, n, sink);
    run<4, false>("read 1 array", c, false, a, b, o, n, sink);
    run<8, false>("read 1 array", c, false, a, b, o, n, sink);
    run<4, false>("read 2 arrays", c, true, a, b, o, n, sink);
    run<4, true>("copy", c, false, a, b, o, n, sink);
    run<4, true>("read 2 write 1", c, true, a, b, o, n, sink);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
