// SPKD Gram matrix and its gradient on the tcgen05 tensor cores (framework.py:157-158).
//
// Forward  G[B,B] += Z Z^T over a split of the feature axis K.  Z is [B, K] bf16 with K contiguous,
// so a [128 rows x 64 k] TMA box (rows >= B are zero-filled by TMA) is at once the A operand
// (M = 128) and - its first Npad rows - the B operand of a K-major tcgen05.mma: every feature
// element is read from HBM exactly once and never re-staged.  The partial Gram of a CTA stays in
// TMEM for its whole K range and is added to G with fp32 reductions at the end; the full Gram is
// never materialised per split in HBM.  The kernel is HBM-bound (64 FLOP/B at B = 64).
//
// Backward  dZ[i,k] = g * sum_j S[i,j] Z[j,k],  S = dG + dG^T.  Per 128-wide k tile:
// D[k, i] = sum_j Z^T[k, j] S[j, i] with the A operand MN-major straight from the same TMA boxes and
// S held in shared memory as bf16 hi + lo parts (two MMAs, ~16 mantissa bits).  Accumulators are
// double-buffered in TMEM so the epilogue (transposed, coalesced store of dZ) overlaps the next
// tile's loads and MMAs.
#include "umma.cuh"

namespace clskd {
namespace {
using namespace umma;

constexpr int kThreads = 192;
constexpr int GF_STAGES = 6;                     // 96 KB of ring per CTA, two CTAs per SM: 6 one-box stages (one row block:
                                                 // 12 x 8-16 KB in flight per SM - the kernel is HBM-latency bound) or
                                                 // 3 two-box stages (cross blocks of a batch > 128)
constexpr uint32_t GF_STAGE_BYTES = 2 * 128 * 128;   // two 128-row x 64-bf16 boxes (A block, B block)

struct GramFwdParams {
  int B, npad;                      // rows of the A block (M side) / padded rows of the B block (UMMA N)
  int Bb;                           // rows of the B block
  int cross;                        // 1: the B operand is a DIFFERENT row block (second tensor map); the result is
                                    // also written mirrored (G[j][i]) - blocked Gram for batches > 128
  int ldg, ro_a, ro_b;              // leading dimension of G and the row offsets of the two blocks in it
  int64_t chunks, chunks_per_cta;   // 64-element K chunks
  float* G;
};

__global__ void __launch_bounds__(kThreads, 2)
gram_fwd_umma_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmZb,
                     const GramFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[GF_STAGES], empty_bar[GF_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t c_beg = (int64_t)blockIdx.x * p.chunks_per_cta;
  const int64_t c_end = min(p.chunks, c_beg + p.chunks_per_cta);
  const int nch = (int)(c_end - c_beg);
  const uint32_t cols = p.npad < 32 ? 32u : (p.npad <= 64 ? 64u : 128u);
  const int nstages = p.cross ? GF_STAGES / 2 : GF_STAGES;
  const uint32_t stage_bytes = p.cross ? GF_STAGE_BYTES : GF_STAGE_BYTES / 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < GF_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, cols);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    {                  // all lanes run the loop, one elected lane issues
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < nch; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_expect_tx_warp(&full_bar[stage], stage_bytes);
        tma_load_2d_warp(ring + (size_t)stage * stage_bytes, &tmZ, &full_bar[stage], (int)((c_beg + it) * 64), 0);
        if (p.cross)
          tma_load_2d_warp(ring + (size_t)stage * stage_bytes + GF_STAGE_BYTES / 2, &tmZb, &full_bar[stage],
                      (int)((c_beg + it) * 64), 0);
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (nch > 0) {     // all lanes run the loop, one elected lane issues
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.npad >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < nch; ++it) {
        mbar_wait(&full_bar[stage], phase);
        fence_after();
        const uint32_t addr = smem_u32(ring + (size_t)stage * stage_bytes);
        const uint64_t desc = make_smem_desc(addr, 1024u >> 4, 2u);   // K-major, 128-byte swizzle
        const uint64_t descb = p.cross ? make_smem_desc(addr + GF_STAGE_BYTES / 2, 1024u >> 4, 2u) : desc;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_warp(tmem_base, desc + (uint64_t)(k * 2), descb + (uint64_t)(k * 2), idesc, (it | k) ? 1u : 0u);
        umma_commit_warp(&empty_bar[stage]);
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit_warp(&tmem_full_bar);
    }
  } else if (nch > 0) {
    const int q = warp & 3;
    const int i = q * 32 + lane;
    mbar_wait(&tmem_full_bar, 0);
    fence_after();
    if (q * 32 < p.B) {     // warp-uniform: rows of this lane quarter exist
      for (int c = 0; c < p.npad; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
        if (i < p.B) {
          float* grow = p.G + (int64_t)(p.ro_a + i) * p.ldg + p.ro_b + c;
          if (!p.cross) {
            // (vector reductions: see red_add16)
            red_add16(grow, v, p.Bb - c);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (c + e < p.Bb) {
                atomicAdd(grow + e, __uint_as_float(v[e]));
                if (p.cross) atomicAdd(p.G + (int64_t)(p.ro_b + c + e) * p.ldg + p.ro_a + i, __uint_as_float(v[e]));
              }
          }
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, cols);
}

// ------------------------------------------------------------------------------------------ backward
constexpr int GB_STAGES = 4;

struct GramBwdParams {
  int B, jb, npad;            // jb = rows of the contraction (64 or 128), npad = UMMA N
  int64_t K, tiles, tiles_per_cta;   // 128-wide k tiles
  const float* dG;
  const float* gout;
  void* dz;
  int dz_dtype;
  int64_t lddz;
  uint32_t sub_bytes;         // jb * 128 : one 64-k sub-tile of Z^T
  uint32_t s_bytes;           // npad * 128 : one 64-j chunk of S (hi or lo)
  int ldg, i0, j0, Bj;        // blocked Gram: dG leading dimension, row offsets of the output (i) / contraction (j) blocks,
                              // rows of the contraction block
  int accumulate;             // dz += (second contraction block of a batch > 128)
  int stages;                 // ring depth (<= GB_STAGES)
  uint32_t stg_bytes;         // TMA-store staging: one [npad rows][64 k] bf16 box (two per tile, two tiles buffered)
};

__device__ __forceinline__ uint32_t sw128_offset(int row, int col_elem) {
  // byte offset of bf16 element (row, col) in a K-major 128-byte-swizzled tile (64 elements per row)
  int chunk = col_elem >> 3;
  return (uint32_t)(row * 128 + (((chunk ^ (row & 7)) & 7) << 4) + ((col_elem & 7) << 1));
}

// TMAO: the bf16 dZ tile goes out through shared memory and a TMA store (reduce-add when ACC).  A thread of the direct
// epilogue walks down the rows of dZ, one 2 MB page per store instruction; at 128 rows that pattern measured 146 GB/s.
template <bool ACC, bool TMAO>
__global__ void __launch_bounds__(kThreads, 1)
gram_bwd_umma_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmDZ,
                     const GramBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[GB_STAGES], empty_bar[GB_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_smem;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int jchunks = p.jb / 64;
  uint8_t* s_hi = base;                                     // [jchunks][npad rows][128 B]
  uint8_t* s_lo = base + (size_t)jchunks * p.s_bytes;
  uint8_t* ring = base + 2 * (size_t)jchunks * p.s_bytes;   // p.stages x (2 sub-tiles)
  const uint32_t stage_bytes = 2 * p.sub_bytes;
  uint8_t* stg_base = ring + (size_t)p.stages * stage_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t t_beg = (int64_t)blockIdx.x * p.tiles_per_cta;
  const int64_t t_end = min(p.tiles, t_beg + p.tiles_per_cta);
  const int nt = (int)(t_end - t_beg);
  const uint32_t cols = 2 * p.npad <= 32 ? 32u : (2 * p.npad <= 64 ? 64u : (2 * p.npad <= 128 ? 128u : 256u));

  if (threadIdx.x == 0) {
    for (int s = 0; s < GB_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 128);
    }
    fence_barrier_init();
  }
  // S = dG + dG^T as bf16 hi/lo, element (i = N row, j = K column), zero padded
  for (int idx = threadIdx.x; idx < p.npad * p.jb; idx += blockDim.x) {
    const int i = idx / p.jb, j = idx - i * p.jb;
    float s = 0.f;
    if (i < p.B && j < p.Bj)
      s = p.dG[(int64_t)(p.i0 + i) * p.ldg + p.j0 + j] + p.dG[(int64_t)(p.j0 + j) * p.ldg + p.i0 + i];
    const __nv_bfloat16 hi = __float2bfloat16_rn(s);
    const __nv_bfloat16 lo = __float2bfloat16_rn(s - __bfloat162float(hi));
    const uint32_t off = (uint32_t)(j >> 6) * p.s_bytes + sw128_offset(i, j & 63);
    *reinterpret_cast<__nv_bfloat16*>(s_hi + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(s_lo + off) = lo;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) tmem_alloc(&tmem_base_smem, cols);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    {                  // all lanes run the loop, one elected lane issues
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < nt; ++it) {
        const int64_t k0 = (t_beg + it) * 128;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_expect_tx_warp(&full_bar[stage], stage_bytes);
        uint8_t* dst = ring + (size_t)stage * stage_bytes;
        tma_load_2d_warp(dst, &tmZ, &full_bar[stage], (int)k0, 0);
        tma_load_2d_warp(dst + p.sub_bytes, &tmZ, &full_bar[stage], (int)(k0 + 64), 0);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    {                  // all lanes run the loop, one elected lane issues
      // A = Z^T tile, MN-major (bit 15); B = S, K-major; N = npad, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) |
                             ((uint32_t)(p.npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < nt; ++it) {
        const int as = it & 1;
        mbar_wait(&acc_empty[as], ((it >> 1) & 1) ^ 1u);
        mbar_wait(&full_bar[stage], phase);
        fence_after();
        const uint32_t a_addr = smem_u32(ring + (size_t)stage * stage_bytes);
        const uint32_t d_addr = tmem_base + (uint32_t)(as * p.npad);
        int first = 1;
        for (int part = 0; part < 2; ++part) {
          const uint32_t s_addr = smem_u32(part ? s_lo : s_hi);
          for (int k = 0; k < p.jb / 16; ++k) {
            // A: 16 contraction rows (j) per MMA -> +16 rows of 128 B; groups of 64 k at LBO = sub_bytes
            const uint64_t adesc = make_smem_desc_lbo(a_addr + (uint32_t)k * 2048u, p.sub_bytes >> 4, 1024u >> 4, 2u);
            // B: K-major S chunk (64 j per 128-byte row); +32 B per 16 j inside a chunk
            const uint32_t sb = s_addr + (uint32_t)(k >> 2) * p.s_bytes + (uint32_t)(k & 3) * 32u;
            const uint64_t bdesc = make_smem_desc(sb, 1024u >> 4, 2u);
            umma_bf16_warp(d_addr, adesc, bdesc, idesc, first ? 0u : 1u);
            first = 0;
          }
        }
        umma_commit_warp(&empty_bar[stage]);
        umma_commit_warp(&acc_full[as]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    const int q = warp & 3;
    const float g = p.gout ? p.gout[0] : 1.f;
    const bool issuer = threadIdx.x == 64;
    for (int it = 0; it < nt; ++it) {
      const int as = it & 1;
      mbar_wait(&acc_full[as], (it >> 1) & 1);
      fence_after();
      if (TMAO) {
        // the store that last read this staging buffer (two tiles ago) must be done with it
        uint8_t* stg = stg_base + (size_t)(it & 1) * 2 * p.stg_bytes;
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int kl = q * 32 + lane, kk = kl & 63;
        uint8_t* colp = stg + (size_t)(kl >> 6) * p.stg_bytes + ((kk & 7) << 1);
        const int chunk = kk >> 3;
        for (int c = 0; c < p.npad; c += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.npad + c), v);
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int i = c + e;
            *reinterpret_cast<__nv_bfloat16*>(colp + i * 128 + (((chunk ^ (i & 7)) & 7) << 4)) =
                __float2bfloat16_rn(g * __uint_as_float(v[e]));
          }
        }
        fence_before();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[as])) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          const int64_t k0 = (t_beg + it) * 128;
          for (int h = 0; h < 2; ++h) {
            if (k0 + 64 * h >= p.K) break;
            if (ACC)
              asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::
                               "l"(&tmDZ), "r"((int)(k0 + 64 * h)), "r"(0), "r"(smem_u32(stg + (size_t)h * p.stg_bytes))
                           : "memory");
            else
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::
                               "l"(&tmDZ), "r"((int)(k0 + 64 * h)), "r"(0), "r"(smem_u32(stg + (size_t)h * p.stg_bytes))
                           : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        continue;
      }
      const int64_t k = (t_beg + it) * 128 + q * 32 + lane;
      const bool kv = k < p.K;
      for (int c = 0; c < p.npad; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.npad + c), v);
        if (kv) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int i = c + e;
            if (i < p.B) {
              float val = g * __uint_as_float(v[e]);
              if (p.dz_dtype == CLSKD_BF16) {
                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.dz) + (int64_t)i * p.lddz + k;
                if (ACC) val += __bfloat162float(*o);
                *o = __float2bfloat16_rn(val);
              } else {
                float* o = reinterpret_cast<float*>(p.dz) + (int64_t)i * p.lddz + k;
                if (ACC) val += *o;
                *o = val;
              }
            }
          }
        }
      }
      fence_before();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[as])) : "memory");
    }
    if (TMAO && issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
  }
  fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, cols);
}

const char* gram_unsupported(const void* z, int dtype, int B, int64_t K, int64_t ldz) {
  if (dtype != CLSKD_BF16) return "z must be bf16";
  if (B < 1 || B > 512) return "B must be <= 512";      // > 128: blocked into 128-row blocks
  if ((uintptr_t)z % 16 || (ldz * 2) % 16) return "z not 16-byte aligned";
  if (K < 64 * 64) return "K too small";
  if (K >= 2147483647LL) return "K too large for a TMA coordinate";
  if (!get_encode()) return "cuTensorMapEncodeTiled unavailable";
  return nullptr;
}

int encode_z(CUtensorMap* tm, const void* z, int B, int64_t K, int64_t ldz, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)B};
  cuuint64_t strides[1] = {(cuuint64_t)ldz * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return (int)get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(z), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace
}  // namespace clskd

using namespace clskd;

extern "C" int clskd_gram_umma_supported(const void* z, int dtype, int B, int64_t K, int64_t ldz) {
  return (z && gram_unsupported(z, dtype, B, K, ldz) == nullptr) ? 1 : 0;
}

extern "C" int clskd_gram_fwd_umma(const void* z, int dtype, int B, int64_t K, int64_t ldz, float* G,
                                   int accumulate, void* stream) {
  CLSKD_CHECK_ARG(z && G, "clskd_gram_fwd_umma: null pointer");
  if (const char* why = gram_unsupported(z, dtype, B, K, ldz)) {
    set_error("clskd_gram_fwd_umma: unsupported: %s", why);
    return CLSKD_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(G, 0, sizeof(float) * (size_t)B * B, st);
    if (e != cudaSuccess) { set_error("clskd_gram_fwd_umma: memset: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
  }
  const size_t smem = (size_t)(GF_STAGES / 2) * GF_STAGE_BYTES + 1024;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gram_fwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("clskd_gram_fwd_umma: smem attr: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
    attr = true;
  }
  // batches > 128 rows: 128-row blocks; block pair (a, b >= a) is one launch (the cross blocks read both row blocks and
  // write the result and its mirror image)
  const __nv_bfloat16* zb = reinterpret_cast<const __nv_bfloat16*>(z);
  for (int ra = 0; ra < B; ra += 128)
    for (int rb = ra; rb < B; rb += 128) {
      const int Ba = B - ra < 128 ? B - ra : 128, Bb = B - rb < 128 ? B - rb : 128;
      CUtensorMap tmZ, tmZb;
      int rc = encode_z(&tmZ, zb + (int64_t)ra * ldz, Ba, K, ldz, 128);
      if (!rc) rc = encode_z(&tmZb, zb + (int64_t)rb * ldz, Bb, K, ldz, 128);
      if (rc) { set_error("clskd_gram_fwd_umma: cuTensorMapEncodeTiled failed: %d", rc); return CLSKD_ERR_CUDA; }
      GramFwdParams p;
      p.B = Ba;
      p.Bb = Bb;
      p.npad = (Bb + 15) / 16 * 16;
      p.cross = rb != ra ? 1 : 0;
      p.ldg = B; p.ro_a = ra; p.ro_b = rb;
      p.chunks = (K + 63) / 64;
      int64_t ctas = 2 * (int64_t)sm_count();
      if (ctas > p.chunks / 16) ctas = p.chunks / 16;
      if (ctas < 1) ctas = 1;
      p.chunks_per_cta = (p.chunks + ctas - 1) / ctas;
      ctas = (p.chunks + p.chunks_per_cta - 1) / p.chunks_per_cta;
      p.G = G;
      gram_fwd_umma_kernel<<<(unsigned)ctas, kThreads, smem, st>>>(tmZ, tmZb, p);
      CLSKD_CHECK_LAUNCH("clskd_gram_fwd_umma");
    }
  return CLSKD_OK;
}

extern "C" int clskd_gram_bwd_umma(const void* z, int dtype, int B, int64_t K, int64_t ldz, const float* dG,
                                   const float* gout, void* dz, int dz_dtype, int64_t lddz, void* stream) {
  CLSKD_CHECK_ARG(z && dG && dz, "clskd_gram_bwd_umma: null pointer");
  if (const char* why = gram_unsupported(z, dtype, B, K, ldz)) {
    set_error("clskd_gram_bwd_umma: unsupported: %s", why);
    return CLSKD_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* zb = reinterpret_cast<const __nv_bfloat16*>(z);
  const size_t dze = dz_dtype == CLSKD_BF16 ? 2 : 4;
  // batches > 128 rows: output block i (128 rows of dZ) = sum over contraction blocks j of S[i, j] Z_j: one launch per
  // (i, j), the second and later j accumulate into dz
  for (int i0 = 0; i0 < B; i0 += 128)
    for (int j0 = 0; j0 < B; j0 += 128) {
      const int Bi = B - i0 < 128 ? B - i0 : 128, Bj = B - j0 < 128 ? B - j0 : 128;
      GramBwdParams p;
      p.B = Bi;
      p.Bj = Bj;
      p.jb = Bj <= 64 ? 64 : 128;
      p.npad = (Bi + 15) / 16 * 16;
      p.K = K;
      p.tiles = (K + 127) / 128;
      int64_t ctas = 2 * (int64_t)sm_count();
      if (ctas > p.tiles) ctas = p.tiles;
      p.tiles_per_cta = (p.tiles + ctas - 1) / ctas;
      ctas = (p.tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
      p.dG = dG; p.gout = gout;
      p.dz = reinterpret_cast<uint8_t*>(dz) + (size_t)i0 * lddz * dze;
      p.dz_dtype = dz_dtype; p.lddz = lddz;
      p.ldg = B; p.i0 = i0; p.j0 = j0;
      p.accumulate = j0 > 0 ? 1 : 0;
      p.sub_bytes = (uint32_t)p.jb * 128u;
      p.s_bytes = ((uint32_t)p.npad * 128u + 1023u) & ~1023u;
      CUtensorMap tmZ, tmDZ;
      int rc = encode_z(&tmZ, zb + (int64_t)j0 * ldz, Bj, K, ldz, p.jb);
      if (rc) { set_error("clskd_gram_bwd_umma: cuTensorMapEncodeTiled failed: %d", rc); return CLSKD_ERR_CUDA; }
      const bool tmao = dz_dtype == CLSKD_BF16 && (uintptr_t)p.dz % 16 == 0 && (lddz * 2) % 16 == 0;
      p.stg_bytes = tmao ? (((uint32_t)p.npad * 128u + 1023u) & ~1023u) : 0u;
      if (tmao) {
        rc = encode_z(&tmDZ, p.dz, Bi, K, lddz, p.npad);
        if (rc) { set_error("clskd_gram_bwd_umma: cuTensorMapEncodeTiled (dz) failed: %d", rc); return CLSKD_ERR_CUDA; }
      } else {
        tmDZ = tmZ;
      }
      // two CTAs per SM while the tile is small (64 contraction rows), one at 128
      const size_t fixed = 2 * (size_t)(p.jb / 64) * p.s_bytes + 4 * (size_t)p.stg_bytes + 1024;
      const size_t budget = p.jb == 64 ? 110u * 1024u : 226u * 1024u;
      p.stages = GB_STAGES;
      while (p.stages > 2 && fixed + (size_t)p.stages * 2 * p.sub_bytes > budget) --p.stages;
      const size_t smem = fixed + (size_t)p.stages * 2 * p.sub_bytes;
      static size_t smem_set = 0;
      if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(gram_bwd_umma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gram_bwd_umma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gram_bwd_umma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gram_bwd_umma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("clskd_gram_bwd_umma: smem attr: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
        smem_set = smem;
      }
      const unsigned grid = (unsigned)ctas;
      if (tmao) {
        if (p.accumulate) gram_bwd_umma_kernel<true, true><<<grid, kThreads, smem, st>>>(tmZ, tmDZ, p);
        else gram_bwd_umma_kernel<false, true><<<grid, kThreads, smem, st>>>(tmZ, tmDZ, p);
      } else {
        if (p.accumulate) gram_bwd_umma_kernel<true, false><<<grid, kThreads, smem, st>>>(tmZ, tmDZ, p);
        else gram_bwd_umma_kernel<false, false><<<grid, kThreads, smem, st>>>(tmZ, tmDZ, p);
      }
      CLSKD_CHECK_LAUNCH("clskd_gram_bwd_umma");
    }
  return CLSKD_OK;
}
