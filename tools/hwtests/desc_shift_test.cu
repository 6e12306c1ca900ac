// Hardware experiment (not product code): does a tcgen05 shared-memory matrix descriptor whose start address is
// shifted by r rows (r * 128 bytes, i.e. NOT aligned to the 1024-byte swizzle pattern) read rows r..r+127 of a
// contiguous 128-byte-swizzled tile that was written with the canonical (absolute-address) swizzle?  And does the
// descriptor's "matrix base offset" field (bits 49-51) have to be set to (start >> 7) & 7 for that?
//   case K : A is K-major (rows = M, 64 bf16 = 128 B per row), shift along M        (tap shift of the forward conv)
//   case MN: A is MN-major (rows = K, 64 channels = 128 B per row), shift along K   (tap shift of the weight gradient)
// Prints, per shift r and per variant (base_offset = 0 / (start>>7)&7), the number of wrong accumulator values.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o desc_shift_test desc_shift_test.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  uint32_t lo = ((saddr >> 4) & 0x3FFFu) | ((lbo & 0x3FFFu) << 16);
  uint32_t hi = (sbo & 0x3FFFu) | (1u << 14) | ((base_off & 7u) << 17) | (layout << 29);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ float aval(int row, int k) { return (float)(((row * 7 + k * 3) % 17) - 8); }
__device__ __forceinline__ float bval(int n, int k) { return (float)(((n * 5 + k) % 13) - 6); }

// swizzled byte offset of element (row, col_elem) in a tile of `pitch`-byte rows (pitch = 128 -> SW128, 64 -> SW64)
__device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t col_bytes, uint32_t pitch) {
  uint32_t lin = row * pitch + col_bytes;
  uint32_t x = pitch == 128 ? ((lin >> 7) & 7u) : pitch == 64 ? ((lin >> 7) & 3u) : ((lin >> 7) & 1u);   // absolute-address swizzle
  return lin ^ (x << 4);
}

constexpr int N = 32;
// mode 0: K-major A (rows=M), shift r rows along M.   mode 1: MN-major A and B, shift r rows along K.
__global__ void __launch_bounds__(128, 1) test_kernel(int mode, int r, int variant, int* wrong, float* maxerr) {
  extern __shared__ __align__(1024) uint8_t dyn[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* a_s = base;                 // mode 0: 160 rows x 128 B ; mode 1: 2 groups x (96 rows x 128 B)
  uint8_t* b_s = base + 32768;         // mode 0: 32 rows x 128 B (K-major) ; mode 1: 96 rows x 64 B (MN-major, SW64)
  const int tid = threadIdx.x;
  if (mode == 0) {
    for (int i = tid; i < 160 * 64; i += 128) {
      int row = i / 64, k = i % 64;
      *reinterpret_cast<__nv_bfloat16*>(a_s + swz(row, k * 2, 128)) = __float2bfloat16(aval(row, k));
    }
    for (int i = tid; i < N * 64; i += 128) {
      int n = i / 64, k = i % 64;
      *reinterpret_cast<__nv_bfloat16*>(b_s + swz(n, k * 2, 128)) = __float2bfloat16(bval(n, k));
    }
  } else {
    // A[k_row][c], c < 128 in two 64-channel groups, group stride 96*128 bytes
    for (int i = tid; i < 96 * 128; i += 128) {
      int row = i / 128, c = i % 128;
      *reinterpret_cast<__nv_bfloat16*>(a_s + (c / 64) * (96 * 128) + swz(row, (c % 64) * 2, 128)) = __float2bfloat16(aval(row, c));
    }
    for (int i = tid; i < 96 * N; i += 128) {
      int row = i / N, n = i % N;
      *reinterpret_cast<__nv_bfloat16*>(b_s + swz(row, n * 2, 64)) = __float2bfloat16(bval(n, row));
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    if (mode == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t a_addr = smem_u32(a_s) + (uint32_t)r * 128u, b_addr = smem_u32(b_s);
      const uint32_t bo = variant ? ((a_addr >> 7) & 7u) : 0u;
      for (int k = 0; k < 4; ++k) {
        uint64_t ad = mk_desc(a_addr + k * 32, 1, 1024 >> 4, 2, bo);
        uint64_t bd = mk_desc(b_addr + k * 32, 1, 1024 >> 4, 2, 0);
        umma(tmem, ad, bd, idesc, k ? 1u : 0u);
      }
    } else {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      const uint32_t a_addr = smem_u32(a_s) + (uint32_t)r * 128u, b_addr = smem_u32(b_s) + (uint32_t)r * 64u;
      const uint32_t boa = variant ? ((a_addr >> 7) & 7u) : 0u;
      const uint32_t bob = variant ? ((b_addr >> 7) & 3u) : 0u;       // SW64 pattern repeats every 512 B
      for (int k = 0; k < 4; ++k) {       // K = 64 rows, 16 per MMA
        uint64_t ad = mk_desc(a_addr + k * 16 * 128, (96 * 128) >> 4, 1024 >> 4, 2, boa);
        uint64_t bd = mk_desc(b_addr + k * 16 * 64, (96 * 64) >> 4, 512 >> 4, 4, bob);
        umma(tmem, ad, bd, idesc, k ? 1u : 0u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait
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
        for (int k = 0; k < 64; ++k) ref += aval(r + m, k) * bval(n, k);
      } else {
        for (int k = 0; k < 64; ++k) ref += aval(r + k, m) * bval(n, r + k);
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
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main() {
  int* wrong;
  float* maxerr;
  cudaMalloc(&wrong, 4);
  cudaMalloc(&maxerr, 4);
  cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  for (int mode = 0; mode < 2; ++mode)
    for (int variant = 0; variant < 2; ++variant)
      for (int r = 0; r <= 9; ++r) {
        cudaMemset(wrong, 0, 4);
        cudaMemset(maxerr, 0, 4);
        test_kernel<<<1, 128, 70 * 1024>>>(mode, r, variant, wrong, maxerr);
        cudaError_t e = cudaDeviceSynchronize();
        int w = -1;
        float me = -1.f;
        cudaMemcpy(&w, wrong, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&me, maxerr, 4, cudaMemcpyDeviceToHost);
        printf("mode=%s variant=%s r=%d wrong=%d/4096 maxerr=%g %s\n", mode ? "MN-major(K shift)" : "K-major(M shift)",
               variant ? "base_offset=(addr>>7)&7" : "base_offset=0", r, w, me, e == cudaSuccess ? "" : cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
      }
  return 0;
}
