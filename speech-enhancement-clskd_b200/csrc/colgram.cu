// Channel Gram matrix and column sums of a dense bf16 map in ONE pass (clskd_colgram):
//     G[i][j] = sum_m x[m][i] x[m][j]   (fp64 [C][C]),      sx[j] = sum_m x[m][j]   (fp64 [C])
// for x [M][C], C in {16, 32, 64, 128}.  Used by the folded BatchNorm backward of the ABF's 1x1 conv
// (framework.py:209; AbfFoldFn): with z1 = W1 x the batch terms of dW1 need  z1^T x = W1 (x^T x)  and  sum x  instead of a
// second pass over the C_mid-channel map z1.  The map is read once (M*C*2 bytes - the floor of this kernel).
//
// mma.sync m16n8k16 (bf16, fp32 accumulate): D[i][j] += A[i][k] B[k][j] with k = 16 rows of x; both operands are the
// SAME shared-memory tile [row][channel] read through ldmatrix.trans (A = x^T needs the transpose, and B is stored
// k-major).  A CTA's 8 warps split as WI warps over the C/16 row blocks of G x WR warps over the rows of a tile, so a
// warp keeps (C/8) x 4 accumulators; the column sums ride along as one extra MMA per column block with a constant A
// fragment (a row of ones).  Tiles of 16384 / C rows arrive through a double-buffered cp.async pipeline (rows beyond M are
// zero-filled); per-CTA partial sums are combined in shared memory, then added to the fp64 outputs.
#include "common.cuh"

namespace clskd {
namespace {

__device__ __forceinline__ uint32_t cg_s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int CG_T = 256;        // threads per CTA
// rows per tile: 16 K elements (32 KB) per stage whatever the channel count - with a fixed 128-row tile the narrow maps
// (C = 16: 4 KB per stage) had far too few bytes in flight per SM to cover the HBM latency
constexpr int cg_rows(int C) { return 16384 / C; }

template <int C>
__global__ void __launch_bounds__(CG_T, 2) colgram_kernel(const __nv_bfloat16* __restrict__ x, int64_t M,
                                                          double* __restrict__ G, double* __restrict__ sx) {
  constexpr int CG_R = cg_rows(C);
  constexpr int PITCH = C * 2 + 16;                 // bytes per staged row (+16: conflict-free ldmatrix rows)
  constexpr int NIT = C / 16, NJT = C / 8;
  constexpr int WI = NIT < 8 ? NIT : 8, WR = 8 / WI;
  constexpr int ROWS_W = CG_R / WR;                  // rows of a tile per warp
  constexpr int NONE = (NJT + WI - 1) / WI;          // column blocks whose sums this warp carries
  constexpr int STAGE = CG_R * PITCH;
  constexpr int CPR = C / 8;                         // 16-byte chunks per row
  extern __shared__ __align__(16) uint8_t cg_sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int a = warp % WI, r = warp / WI;

  float acc[NJT][4];
  float oacc[NONE][4];
#pragma unroll
  for (int j = 0; j < NJT; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
#pragma unroll
  for (int j = 0; j < NONE; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) oacc[j][e] = 0.f;
  // constant A fragment: row 0 of the 16 x 16 tile is all ones (bf16 1.0 = 0x3f80)
  const uint32_t one2 = (lane >> 2) == 0 ? 0x3f803f80u : 0u;

  const int64_t ntiles = (M + CG_R - 1) / CG_R;
  auto issue = [&](int64_t tile, int stage) {
    uint8_t* dst = cg_sm + stage * STAGE;
    const int64_t row0 = tile * CG_R;
    for (int i = tid; i < CG_R * CPR; i += CG_T) {
      const int rr = i / CPR, ch = i - rr * CPR;
      const int64_t m = row0 + rr;
      const bool ok = m < M;
      const __nv_bfloat16* src = x + (ok ? m : 0) * C + ch * 8;
      const int sz = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(cg_s_u32(dst + rr * PITCH + ch * 16)), "l"(src),
                   "r"(sz)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int stage = 0;
  if ((int64_t)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t nxt = tile + gridDim.x;
    if (nxt < ntiles) {
      issue(nxt, stage ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint32_t base = cg_s_u32(cg_sm + stage * STAGE);
    const int mi = lane >> 3, lr = lane & 7;
#pragma unroll 1
    for (int k0 = r * ROWS_W; k0 < (r + 1) * ROWS_W; k0 += 16) {
      uint32_t a0, a1, a2, a3;
      // A = x^T: matrices (k 0-7, i 0-7), (k 0-7, i 8-15), (k 8-15, i 0-7), (k 8-15, i 8-15) -> a0..a3
      ldsm_x4_t(base + (uint32_t)(k0 + lr + (mi >> 1) * 8) * PITCH + (uint32_t)(a * 16 + (mi & 1) * 8) * 2, a0, a1, a2, a3);
#pragma unroll
      for (int jt = 0; jt < NJT; jt += 2) {
        uint32_t b0, b1, b2, b3;
        // B: matrices (k 0-7, j 0-7), (k 8-15, j 0-7), (k 0-7, j 8-15), (k 8-15, j 8-15)
        ldsm_x4_t(base + (uint32_t)(k0 + lr + (mi & 1) * 8) * PITCH + (uint32_t)(jt * 8 + (mi >> 1) * 8) * 2, b0, b1, b2, b3);
        mma_bf16(acc[jt], a0, a1, a2, a3, b0, b1);
        mma_bf16(acc[jt + 1], a0, a1, a2, a3, b2, b3);
        if (jt % WI == a) mma_bf16(oacc[jt / WI], one2, 0u, one2, 0u, b0, b1);
        if ((jt + 1) % WI == a) mma_bf16(oacc[(jt + 1) / WI], one2, 0u, one2, 0u, b2, b3);
      }
    }
    __syncthreads();
    stage ^= 1;
  }

  // ---- per-CTA reduction in shared memory (the WR row groups hold the same entries), then fp64 atomics
  float* red = reinterpret_cast<float*>(cg_sm);            // [C*C] + [C]
  for (int i = tid; i < C * C + C; i += CG_T) red[i] = 0.f;
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int jt = 0; jt < NJT; ++jt) {
    const int i0 = a * 16 + g, j0 = jt * 8 + 2 * t;
    atomicAdd(&red[i0 * C + j0], acc[jt][0]);
    atomicAdd(&red[i0 * C + j0 + 1], acc[jt][1]);
    atomicAdd(&red[(i0 + 8) * C + j0], acc[jt][2]);
    atomicAdd(&red[(i0 + 8) * C + j0 + 1], acc[jt][3]);
    if (jt % WI == a && g == 0) {
      atomicAdd(&red[C * C + j0], oacc[jt / WI][0]);
      atomicAdd(&red[C * C + j0 + 1], oacc[jt / WI][1]);
    }
  }
  __syncthreads();
  for (int i = tid; i < C * C; i += CG_T) atomicAdd(G + i, (double)red[i]);
  for (int i = tid; i < C; i += CG_T) atomicAdd(sx + i, (double)red[C * C + i]);
}

template <int C>
int colgram_launch(const void* x, int64_t M, double* G, double* sx, cudaStream_t st) {
  constexpr int PITCH = C * 2 + 16;
  constexpr int CG_R = cg_rows(C);
  size_t sh = (size_t)2 * CG_R * PITCH;
  const size_t red = sizeof(float) * (C * C + C);
  if (red > sh) sh = red;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(colgram_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr = true;
  }
  const int64_t ntiles = (M + CG_R - 1) / CG_R;
  const int64_t cap = (int64_t)sm_count() * 2;
  const int grid = (int)(ntiles < cap ? ntiles : cap);
  colgram_kernel<C><<<grid, CG_T, sh, st>>>((const __nv_bfloat16*)x, M, G, sx);
  CLSKD_CHECK_LAUNCH("clskd_colgram");
  return CLSKD_OK;
}

}  // namespace
}  // namespace clskd

using namespace clskd;

extern "C" int clskd_colgram_supported(int dtype, int C) {
  return (dtype == CLSKD_BF16 && (C == 16 || C == 32 || C == 64 || C == 128)) ? 1 : 0;
}

extern "C" int clskd_colgram(const void* x, int dtype, int64_t M, int C, double* G, double* sx, void* stream) {
  CLSKD_CHECK_ARG(x && G && sx, "clskd_colgram: null pointer");
  CLSKD_CHECK_ARG(M >= 0, "clskd_colgram: M");
  if (!clskd_colgram_supported(dtype, C)) {
    set_error("clskd_colgram: unsupported: bf16 maps with 16 / 32 / 64 / 128 channels only");
    return CLSKD_ERR_UNSUPPORTED;
  }
  CLSKD_CHECK_ARG(((uintptr_t)x % 16) == 0, "clskd_colgram: x must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = zero_spans(st, G, sizeof(double) * C * C, sx, sizeof(double) * C);
  if (e != cudaSuccess) { set_error("clskd_colgram: memset: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
  if (M == 0) return CLSKD_OK;
  switch (C) {
    case 16: return colgram_launch<16>(x, M, G, sx, st);
    case 32: return colgram_launch<32>(x, M, G, sx, st);
    case 64: return colgram_launch<64>(x, M, G, sx, st);
    default: return colgram_launch<128>(x, M, G, sx, st);
  }
}
