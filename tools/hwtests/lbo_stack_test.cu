// Hardware experiment (not product code): can the leading-dimension byte offset (LBO) of an MN-major tcgen05 shared
// memory descriptor be ONE ROW, so that consecutive MN sub-blocks of one MMA read the same tile shifted by one row each
// ("tap stacking": several convolution taps as extra N columns / M rows of a single instruction)?
//   mode 0: B MN-major, 64-byte rows (32 bf16, SW64), N = 96 = 3 sub-blocks, LBO = 64 B:
//           D[m][s*32 + n] = sum_k A[k][m] * Bp[k + s][n]          (weight gradient: taps of dY as extra N columns)
//   mode 1: A MN-major, 32-byte rows (16 bf16, SW32), M = 128 = 8 sub-blocks, LBO = 32 B:
//           D[s*16 + c][n] = sum_k Ap[k + s][c] * B[k][n]          (small-channel weight gradient: taps as extra M rows)
// Tiles are written with the canonical absolute-address swizzle (what TMA produces).  Prints wrong accumulator counts.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o lbo_stack_test lbo_stack_test.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint32_t lo = ((saddr >> 4) & 0x3FFFu) | ((lbo & 0x3FFFu) << 16);
  uint32_t hi = (sbo & 0x3FFFu) | (1u << 14) | (layout << 29);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ float aval(int row, int c) { return (float)(((row * 7 + c * 3) % 17) - 8); }
__device__ __forceinline__ float bval(int row, int n) { return (float)(((n * 5 + row) % 13) - 6); }
__device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t col_bytes, uint32_t pitch) {
  uint32_t lin = row * pitch + col_bytes;
  uint32_t x = pitch == 128 ? ((lin >> 7) & 7u) : pitch == 64 ? ((lin >> 7) & 3u) : ((lin >> 7) & 1u);
  return lin ^ (x << 4);
}

constexpr int KROWS = 64;     // contraction rows per test (4 MMAs of K = 16)

__global__ void __launch_bounds__(128, 1) test_kernel(int mode, int* wrong, float* maxerr) {
  extern __shared__ __align__(1024) uint8_t dyn[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* a_s = base;
  uint8_t* b_s = base + 32768;
  const int tid = threadIdx.x;
  const int N = mode == 0 ? 96 : 32;
  if (mode == 0) {
    // A[k][c], c < 128 in two 64-channel groups (SW128), group stride 96*128 bytes; B patch [k + halo][32] SW64
    for (int i = tid; i < 96 * 128; i += 128) {
      int row = i / 128, c = i % 128;
      *reinterpret_cast<__nv_bfloat16*>(a_s + (c / 64) * (96 * 128) + swz(row, (c % 64) * 2, 128)) = __float2bfloat16(aval(row, c));
    }
    for (int i = tid; i < 96 * 32; i += 128) {
      int row = i / 32, n = i % 32;
      *reinterpret_cast<__nv_bfloat16*>(b_s + swz(row, n * 2, 64)) = __float2bfloat16(bval(row, n));
    }
  } else {
    // A patch [k + halo][16] SW32 ; B [k][32] SW64
    for (int i = tid; i < 96 * 16; i += 128) {
      int row = i / 16, c = i % 16;
      *reinterpret_cast<__nv_bfloat16*>(a_s + swz(row, c * 2, 32)) = __float2bfloat16(aval(row, c));
    }
    for (int i = tid; i < 96 * 32; i += 128) {
      int row = i / 32, n = i % 32;
      *reinterpret_cast<__nv_bfloat16*>(b_s + swz(row, n * 2, 64)) = __float2bfloat16(bval(row, n));
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    for (int k = 0; k < KROWS / 16; ++k) {
      uint64_t ad, bd;
      if (mode == 0) {
        ad = mk_desc(smem_u32(a_s) + k * 16 * 128, (96 * 128) >> 4, 1024 >> 4, 2);      // SW128, 2 channel groups
        bd = mk_desc(smem_u32(b_s) + k * 16 * 64, 64 >> 4, 512 >> 4, 4);                // SW64, LBO = ONE ROW
      } else {
        ad = mk_desc(smem_u32(a_s) + k * 16 * 32, 32 >> 4, 256 >> 4, 6);                // SW32, LBO = ONE ROW
        bd = mk_desc(smem_u32(b_s) + k * 16 * 64, (96 * 64) >> 4, 512 >> 4, 4);         // SW64 (one sub-block)
      }
      umma(tmem, ad, bd, idesc, k ? 1u : 0u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(done)
                   : "r"(smem_u32(&bar))
                   : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = tid >> 5, lane = tid & 31;
  const int m = warp * 32 + lane;
  int bad = 0;
  float me = 0.f;
  for (int c = 0; c < N; c += 16) {
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int e = 0; e < 16; ++e) {
      const int n = c + e;
      float ref = 0.f;
      if (mode == 0) {
        const int s = n / 32, nn = n % 32;
        for (int k = 0; k < KROWS; ++k) ref += aval(k, m) * bval(k + s, nn);
      } else {
        const int s = m / 16, cc = m % 16;
        for (int k = 0; k < KROWS; ++k) ref += aval(k + s, cc) * bval(k, n);
      }
      const float got = __uint_as_float(v[e]);
      const float er = fabsf(got - ref);
      if (er > 0.5f) ++bad;
      me = fmaxf(me, er);
    }
  }
  atomicAdd(wrong, bad);
  atomicMax(reinterpret_cast<int*>(maxerr), __float_as_int(me));
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

int main() {
  int* wrong;
  float* maxerr;
  cudaMalloc(&wrong, 4);
  cudaMalloc(&maxerr, 4);
  cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(wrong, 0, 4);
    cudaMemset(maxerr, 0, 4);
    test_kernel<<<1, 128, 70 * 1024>>>(mode, wrong, maxerr);
    cudaError_t e = cudaDeviceSynchronize();
    int w = -1;
    float me = -1.f;
    cudaMemcpy(&w, wrong, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&me, maxerr, 4, cudaMemcpyDeviceToHost);
    printf("mode=%d (%s) wrong=%d/%d maxerr=%g %s\n", mode, mode ? "A stacked along M, SW32, LBO = 32 B" : "B stacked along N, SW64, LBO = 64 B",
           w, 128 * (mode ? 32 : 96), me, e == cudaSuccess ? "" : cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
