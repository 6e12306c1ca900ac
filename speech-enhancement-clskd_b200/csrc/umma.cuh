// tcgen05 / TMEM / TMA / mbarrier PTX wrappers and the host-side tensor-map encoder shared by the
// tensor-core kernels (tapconv_umma.cu, tapconv_wgrad_umma.cu, gram_umma.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace clskd {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// warp-convergent forms of the producer's instructions (see umma_bf16_warp)
__device__ __forceinline__ void mbar_expect_tx_warp(uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n"
      "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n}\n" ::"r"(smem_u32(bar)),
      "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_warp(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2,
                                                 int c3, int c4) {
  asm volatile(
      "{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n"
      "@e cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n}\n" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_warp(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2,
                                                 int c3) {
  asm volatile(
      "{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n"
      "@e cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n}\n" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_warp(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n"
      "@e cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n}\n" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_warp(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n"
      "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n}\n" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// warp-convergent forms: executed by ALL lanes of the issuing warp, one elected lane issues (election and the
// predicated instruction in one asm block, no divergent branch around them)
__device__ __forceinline__ void umma_bf16_warp(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_warp(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
      "%12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// the same load without the wait: several loads in flight, one tmem_ld_wait() before the registers are read
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
      "%12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo, uint32_t layout) {
  uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);            // start address, LBO = 1 (unused)
  uint32_t hi = (sbo & 0x3FFFu) | (1u << 14) | (layout << 29);    // SBO, version = 1, layout type
  return ((uint64_t)hi << 32) | lo;
}


// One elected lane of a fully active warp.  The MMA-issuing warp runs its loops on all 32 lanes (warp-uniform control
// flow and operands) and guards only the tcgen05 instructions with this: inside an `if (lane == 0)` region the compiler
// cannot prove the uniform-register operands of UTCHMMA uniform and wraps EVERY issue in an ELECT / BRA.U.ANY retry
// loop (about 50 cycles per MMA - more than a narrow-N instruction takes to execute).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}

// shared-memory matrix descriptor with an explicit leading-dimension byte offset (MN-major operands)
__device__ __forceinline__ uint64_t make_smem_desc_lbo(uint32_t saddr, uint32_t lbo, uint32_t sbo,
                                                       uint32_t layout) {
  uint32_t lo = ((saddr >> 4) & 0x3FFFu) | ((lbo & 0x3FFFu) << 16);
  uint32_t hi = (sbo & 0x3FFFu) | (1u << 14) | (layout << 29);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols)
               : "memory");
}
__device__ __forceinline__ void fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// ---- host side
// cuTensorMapEncodeTiled is a driver-API call: make sure this thread (e.g. an autograd worker) has
// the primary context bound before the first one.
inline void ensure_context() {
  static thread_local bool done = false;
  if (!done) {
    cudaFree(nullptr);
    done = true;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  ensure_context();
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}


inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

inline CUtensorMapSwizzle swizzle_for_bytes(int inner_bytes) {
  return inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
         : inner_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                             : CU_TENSOR_MAP_SWIZZLE_32B;
}
// UMMA shared-memory layout type for a swizzle span in bytes (2 = SW128, 4 = SW64, 6 = SW32)
inline uint32_t layout_for_bytes(int inner_bytes) {
  return inner_bytes == 128 ? 2u : inner_bytes == 64 ? 4u : 6u;
}

// channels-last activation [B,Ti,Fi,C] as the 5-D tensor (C, f-parity, Fi/sf, Ti, B); box = [block_k, 1, fo_tile, t_tile, 1]
inline int encode_act(EncodeTiledFn enc, CUtensorMap* tm, const void* x, int C, int sf, int Fi, int Ti,
               int B, int64_t sB, int64_t sT, int64_t sF, int block_k, int fo_tile, int t_tile,
               CUtensorMapSwizzle sw) {
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)sf, (cuuint64_t)(Fi / sf), (cuuint64_t)Ti,
                        (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)sF * 2, (cuuint64_t)sF * 2 * sf, (cuuint64_t)sT * 2,
                           (cuuint64_t)sB * 2};
  cuuint32_t box[5] = {(cuuint32_t)block_k, 1, (cuuint32_t)fo_tile, (cuuint32_t)t_tile, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}



// fp32 reduction of 16 consecutive accumulator columns into global memory (split-K / split-row epilogues): four-wide
// vector reductions (red.global.add.v4.f32) when the destination is 16-byte aligned and all 16 columns exist - a quarter
// of the atomic operations that hundreds of CTAs send to the same addresses at the end of a kernel (measured on the Gram
// forward: a fixed ~35 us tail per launch with scalar reductions)
__device__ __forceinline__ void red_add16(float* dst, const uint32_t* v, int nvalid) {
  if (nvalid >= 16 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
    for (int e = 0; e < 16; e += 4)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + e), "f"(__uint_as_float(v[e])),
                   "f"(__uint_as_float(v[e + 1])), "f"(__uint_as_float(v[e + 2])), "f"(__uint_as_float(v[e + 3]))
                   : "memory");
  } else {
#pragma unroll
    for (int e = 0; e < 16; ++e)
      if (e < nvalid) atomicAdd(dst + e, __uint_as_float(v[e]));
  }
}

}  // namespace umma
}  // namespace clskd
