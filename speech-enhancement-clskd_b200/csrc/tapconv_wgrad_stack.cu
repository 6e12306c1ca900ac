// Weight gradient of a stride-1 "same" convolution with the FREQUENCY TAPS STACKED ALONG THE MMA's N DIMENSION.
//
//   dW[(dt,df)][c][n] = sum_{b,t,f} X[b,t+dt,f+df,c] dY[b,t,f,n] = sum_{b,t',f'} X[b,t',f',c] dY[b,t'-dt,f'-df,n]
//
// The per-tap kernels (tapconv_wgrad_umma*.cu) shift the X tile per tap and issue one M128 x N x K16 instruction chain
// per tap; for narrow N (32 / 64 output channels: the ABF conv2 layers) those instructions are bound by streaming the
// 4 KB X operand from shared memory, once per tap.  Here the X tile is the FIXED operand and dY carries the shift: one
// dY patch per time tap holds the tile's rows plus the frequency halo, and the nf frequency taps are nf sub-blocks of
// ONE instruction's N dimension whose leading-dimension byte offset is a single patch row, so sub-block s reads the
// patch shifted by s rows (measured to work on B200 with the canonical absolute-address swizzle:
// tools/hwtests/lbo_stack_test.cu).  Per K step: one X read per time tap instead of one per tap (3 x 4 KB + 3 x 3 KB
// instead of 9 x 5 KB at N = 32).
//
// Tiles: 128 consecutive frequencies of one time row when Fo >= 128; otherwise t_tile = 128 / Fo time rows, each
// padded to a pitch of Fo + 8 rows in BOTH operands (halo columns are zero-filled by TMA, so the shifted reads never
// pair a live X row with a row of another time line).
// One CTA owns a 128-channel tile, a chunk of the time taps (those whose nf * Np accumulator columns fit TMEM
// together) and a contiguous range of row tiles; accumulators stay in TMEM over the whole range; fp32 reductions into
// dW at the end (split-K over CTAs).  Warp roles as in the other tcgen05 kernels: TMA producer, MMA issuer, 4 epilogue
// warps.
#include "umma.cuh"

namespace clskd {
extern int g_wgrad_mode;
namespace {
using namespace umma;

constexpr int kThreads = 192;
constexpr int MAX_STAGES = 4;
constexpr int MAX_DT = 4, MAX_DF = 8;

struct StackParams {
  int B, T, F;                   // output = input extents (stride 1, same padding)
  int t_tile, f_tiles, t_tiles;
  int n_row_tiles, tiles_per_cta;
  int a_rows;                    // K rows of one tile (multiple of 16)
  int a_fbox_start;              // frequency coordinate of the X box relative to f0 (0 or -1)
  int b_fbox_start;              // ... of the dY box: a_fbox_start - df_max
  int ndt, nf, Np, N, Ctot;      // time taps, stacked frequency taps, padded / real N, real channels (c0 + c1)
  int c0, c0p, c1r, Ctot_p;      // source 0 real / padded to 16, real channels of source 1, padded total
  int gw_a;                      // channels per swizzle group of the X tile (64 / 32 / 16)
  uint32_t pitch_a, layout_a;
  int gd, ngd;                   // time taps per CTA, number of such chunks
  int c_tiles;
  int dts[MAX_DT];
  int tap_of[MAX_DT][MAX_DF];    // index into the caller's tap list of (time tap g, stack slot s); -1: no such tap
  uint32_t a_sub_bytes;          // one 64-channel group of the X tile
  uint32_t a_stage_bytes, b_patch_bytes, stage_bytes;
  uint32_t b_tx;                 // bytes TMA writes into one dY patch (the rest of the slot is zeroed once)
  uint32_t pitch_b, layout_b;
  int stages;
  uint32_t tmem_cols;
  float* dw;
};

__global__ void __launch_bounds__(kThreads, 1)
tapconv_wgrad_stack_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmX1,
                           const __grid_constant__ CUtensorMap tmDY, const StackParams p) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int w = blockIdx.y;
  const int c_t = w % p.c_tiles;
  const int gchunk = w / p.c_tiles;
  const int g0 = gchunk * p.gd;
  const int gcur = min(p.gd, p.ndt - g0);
  const int cbase = c_t * 128;
  const int cvalid = min(128, p.Ctot_p - cbase);      // padded channel axis: [c0 -> c0p | c1 -> c1p]
  const int nsub_a = cvalid / p.gw_a;
  const int tile_beg = blockIdx.x * p.tiles_per_cta;
  const int tile_end = min(p.n_row_tiles, tile_beg + p.tiles_per_cta);
  const int ntile_cta = tile_end - tile_beg;
  const int ncols = p.nf * p.Np;           // accumulator columns (= MMA N) of one time tap

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  // The shifted sub-blocks read up to nf - 1 rows past a dY patch: those bytes are never written by TMA and only ever
  // multiply zero-filled X rows, but they must be finite.  Zero the tail of every patch slot once.
  for (int s = 0; s < p.stages; ++s)
    for (int g = 0; g < p.gd; ++g) {
      uint8_t* patch = ring + (size_t)s * p.stage_bytes + p.a_stage_bytes + (size_t)g * p.b_patch_bytes;
      for (uint32_t i = p.b_tx + 4u * threadIdx.x; i < p.b_patch_bytes; i += 4u * kThreads)
        *reinterpret_cast<uint32_t*>(patch + i) = 0u;
    }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && ntile_cta > 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < ntile_cta; ++it) {
        int r = tile_beg + it;
        const int f_blk = r % p.f_tiles;
        r /= p.f_tiles;
        const int t_blk = r % p.t_tiles;
        const int b = r / p.t_tiles;
        const int t0 = t_blk * p.t_tile, f0 = f_blk * 128;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_expect_tx(&full_bar[stage], (uint32_t)nsub_a * p.a_sub_bytes + (uint32_t)gcur * p.b_tx);
        uint8_t* a_dst = ring + (size_t)stage * p.stage_bytes;
        for (int s = 0; s < nsub_a; ++s) {
          const int cc = cbase + s * p.gw_a;
          const bool src0 = cc < p.c0p;        // channels beyond a source's real extent are zero-filled by TMA
          tma_load_4d(a_dst + (size_t)s * p.a_sub_bytes, src0 ? &tmX : &tmX1, &full_bar[stage], src0 ? cc : cc - p.c0p,
                      f0 + p.a_fbox_start, t0, b);
        }
        for (int g = 0; g < gcur; ++g)
          tma_load_4d(a_dst + p.a_stage_bytes + (size_t)g * p.b_patch_bytes, &tmDY, &full_bar[stage], 0,
                      f0 + p.b_fbox_start, t0 - p.dts[g0 + g], b);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && ntile_cta > 0) {
      // D = f32, A = B = bf16, both MN-major, N = nf * Np, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(ncols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      const int ksteps = p.a_rows / 16;
      for (int it = 0; it < ntile_cta; ++it) {
        mbar_wait(&full_bar[stage], phase);
        fence_after();
        const uint32_t a_addr = smem_u32(ring + (size_t)stage * p.stage_bytes);
        for (int g = 0; g < gcur; ++g) {
          const uint32_t b_addr = a_addr + p.a_stage_bytes + (uint32_t)g * p.b_patch_bytes;
          for (int k = 0; k < ksteps; ++k) {
            // X: channel groups at LBO = a_sub_bytes, 8-row groups at SBO = 8 rows (sub-blocks past the tile's channels
            // read stale shared memory: their accumulator rows are never stored)
            const uint64_t adesc = make_smem_desc_lbo(a_addr + (uint32_t)k * 16u * p.pitch_a, p.a_sub_bytes >> 4,
                                                      (8u * p.pitch_a) >> 4, p.layout_a);
            // dY: frequency taps at LBO = ONE ROW of the patch
            const uint64_t bdesc = make_smem_desc_lbo(b_addr + (uint32_t)k * 16u * p.pitch_b, p.pitch_b >> 4,
                                                      (8u * p.pitch_b) >> 4, p.layout_b);
            umma_bf16(tmem_base + (uint32_t)(g * ncols), adesc, bdesc, idesc, (it | k) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(&tmem_full_bar);
    }
  } else if (ntile_cta > 0) {
    // ===================== epilogue (warps 2..5): TMEM -> fp32 reductions into dW =====================
    const int q = warp & 3;
    const int c_pad = cbase + q * 32 + lane;                 // index on the padded channel axis
    const int c = c_pad < p.c0p ? c_pad : p.c0 + (c_pad - p.c0p);
    const bool valid = (q * 32 + lane) < cvalid && (c_pad < p.c0p ? c_pad < p.c0 : (c_pad - p.c0p) < p.c1r);
    mbar_wait(&tmem_full_bar, 0);
    fence_after();
    for (int g = 0; g < gcur; ++g)
      for (int s = 0; s < p.nf; ++s) {
        const int tap = p.tap_of[g0 + g][s];
        for (int cc = 0; cc < p.Np; cc += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * ncols + s * p.Np + cc), v);
          if (valid && tap >= 0) {
            float* dst = p.dw + ((int64_t)tap * p.Ctot + c) * p.N + cc;
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (cc + e < p.N) atomicAdd(dst + e, __uint_as_float(v[e]));
          }
        }
      }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// geometry the stacked kernel covers; fills the tap grid.  nullptr = supported
const char* stack_unsupported(const ClskdTapConv* d, StackParams* out) {
  if (d->x_dtype != CLSKD_BF16 || d->y_dtype != CLSKD_BF16) return "x and dy must be bf16";
  if (d->sf != 1) return "stride 1 only";
  if (d->Ti != d->To || d->Fi != d->Fo) return "same-size convolution only";
  if (d->c0 % 8 || d->c1 % 8 || d->c0 < 8) return "channels must be multiples of 8";
  const int Np = (d->N + 15) & ~15;
  if (Np != 16 && Np != 32 && Np != 64) return "N must pad to 16, 32 or 64";
  if (d->N % 8) return "N must be a multiple of 8";
  if (!is_pow2(d->Fo) || d->Fo < 16 || (d->Fo > 128 && d->Fo % 128)) return "Fo must be a power of two >= 16";
  if (d->accumulate) return "accumulate unsupported";
  // the taps must be a full (time x frequency) grid with consecutive frequency offsets
  int dts[MAX_DT], ndt = 0, dfmin = 1 << 30, dfmax = -(1 << 30);
  for (int j = 0; j < d->ntaps; ++j) {
    bool seen = false;
    for (int g = 0; g < ndt; ++g) seen = seen || dts[g] == d->dt[j];
    if (!seen) {
      if (ndt == MAX_DT) return "too many time taps";
      dts[ndt++] = d->dt[j];
    }
    dfmin = d->df[j] < dfmin ? d->df[j] : dfmin;
    dfmax = d->df[j] > dfmax ? d->df[j] : dfmax;
  }
  const int nf = dfmax - dfmin + 1;
  if (nf < 2 || nf > MAX_DF || ndt * nf != d->ntaps) return "taps are not a full grid";
  if (nf * Np > 256) return "stacked N exceeds 256";
  if (nf - 1 > 7) return "frequency span too wide for the padded pitch";
  StackParams p;
  memset(&p, 0, sizeof(p));
  for (int g = 0; g < ndt; ++g) {
    p.dts[g] = dts[g];
    for (int s = 0; s < nf; ++s) {
      p.tap_of[g][s] = -1;
      for (int j = 0; j < d->ntaps; ++j)
        if (d->dt[j] == dts[g] && d->df[j] == dfmax - s) p.tap_of[g][s] = j;     // slot s reads the patch shifted by s rows
      if (p.tap_of[g][s] < 0) return "taps are not a full grid";
    }
  }
  auto chk = [&](const void* x, int64_t sB, int64_t sT, int64_t sF) -> bool {
    return (uintptr_t)x % 16 == 0 && (sB * 2) % 16 == 0 && (sT * 2) % 16 == 0 && (sF * 2) % 16 == 0;
  };
  if (!chk(d->x0, d->x0_sB, d->x0_sT, d->x0_sF) || !chk(d->y, d->y_sB, d->y_sT, d->y_sF)) return "alignment";
  if (d->c1 && !chk(d->x1, d->x1_sB, d->x1_sT, d->x1_sF)) return "alignment";
  if ((int64_t)d->B * d->To * d->Fo < 65536) return "too few rows";
  if (!get_encode()) return "cuTensorMapEncodeTiled unavailable";
  p.B = d->B; p.T = d->To; p.F = d->Fo;
  p.ndt = ndt; p.nf = nf; p.Np = Np; p.N = d->N; p.Ctot = d->c0 + d->c1;
  p.c0 = d->c0; p.c0p = (d->c0 + 15) & ~15; p.c1r = d->c1; p.Ctot_p = p.c0p + ((d->c1 + 15) & ~15);
  p.gw_a = (p.c0p % 64 == 0 && (p.Ctot_p - p.c0p) % 64 == 0) ? 64 : ((p.c0p % 32 == 0 && (p.Ctot_p - p.c0p) % 32 == 0) ? 32 : 16);
  p.pitch_a = (uint32_t)p.gw_a * 2u;
  p.layout_a = layout_for_bytes((int)p.pitch_a);
  if (d->Fo >= 128) {
    p.t_tile = 1;
    p.f_tiles = d->Fo / 128;
    p.a_rows = 128;
    p.a_fbox_start = 0;
  } else {
    p.t_tile = 128 / d->Fo;
    p.f_tiles = 1;
    p.a_rows = p.t_tile * (d->Fo + 8);
    p.a_fbox_start = -1;
    if (p.a_rows % 16) return "padded tile is not a multiple of 16 rows";
  }
  p.b_fbox_start = p.a_fbox_start - dfmax;
  p.t_tiles = cdiv(d->To, p.t_tile);
  const int64_t nrt = (int64_t)d->B * p.t_tiles * p.f_tiles;
  if (nrt > 2147483647LL) return "too many row tiles";
  p.n_row_tiles = (int)nrt;
  p.gd = 512 / (nf * Np);
  if (p.gd > ndt) p.gd = ndt;
  p.ngd = cdiv(ndt, p.gd);
  p.c_tiles = cdiv(p.Ctot_p, 128);
  p.pitch_b = (uint32_t)Np * 2u;
  p.layout_b = layout_for_bytes((int)p.pitch_b);
  p.a_sub_bytes = (uint32_t)p.a_rows * p.pitch_a;
  p.a_stage_bytes = (uint32_t)(128 / p.gw_a) * p.a_sub_bytes;            // 128 channel rows of the MMA
  p.b_tx = (uint32_t)(p.a_rows + (p.t_tile == 1 ? nf - 1 : 0)) * p.pitch_b;
  p.b_patch_bytes = (p.b_tx + (uint32_t)(nf - 1) * p.pitch_b + 1023u) & ~1023u;           // shifted reads stay inside
  p.stage_bytes = p.a_stage_bytes + (uint32_t)p.gd * p.b_patch_bytes;
  int stages = (int)((200u * 1024u) / p.stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return "stage too large";
  p.stages = stages;
  int cols = 32;
  while (cols < p.gd * nf * Np) cols <<= 1;
  p.tmem_cols = (uint32_t)cols;
  if (out) *out = p;
  return nullptr;
}

}  // namespace
}  // namespace clskd

using namespace clskd;

extern "C" int clskd_tapconv_wgrad_umma_stacked_supported(const ClskdTapConv* d) {
  if (!d || !d->x0 || !d->y || d->ntaps < 1 || d->ntaps > CLSKD_MAX_TAPS) return 0;
  return stack_unsupported(d, nullptr) == nullptr ? 1 : 0;
}

extern "C" int clskd_tapconv_wgrad_umma_stacked(const ClskdTapConv* d, void* stream) {
  CLSKD_CHECK_ARG(d && d->x0 && d->w && d->y, "clskd_tapconv_wgrad_umma_stacked: null pointer");
  CLSKD_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= CLSKD_MAX_TAPS, "clskd_tapconv_wgrad_umma_stacked: ntaps");
  StackParams p;
  if (const char* why = stack_unsupported(d, &p)) {
    set_error("clskd_tapconv_wgrad_umma_stacked: unsupported: %s", why);
    return CLSKD_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(const_cast<void*>(d->w), 0, sizeof(float) * (size_t)d->ntaps * (d->c0 + d->c1) * d->N, st);
  if (e != cudaSuccess) { set_error("clskd_tapconv_wgrad_umma_stacked: memset: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
  p.dw = reinterpret_cast<float*>(const_cast<void*>(d->w));
  const int ycount = p.ngd * p.c_tiles;
  int nsplit = (sm_count() + ycount - 1) / ycount;
  if (nsplit > p.n_row_tiles) nsplit = p.n_row_tiles;
  if (nsplit < 1) nsplit = 1;
  p.tiles_per_cta = cdiv(p.n_row_tiles, nsplit);
  nsplit = cdiv(p.n_row_tiles, p.tiles_per_cta);

  EncodeTiledFn enc = get_encode();
  CUtensorMap tmX, tmX1, tmDY;
  const int frows = p.t_tile == 1 ? 128 : d->Fo + 8;          // box rows along f per time line (X)
  auto enc_x = [&](CUtensorMap* tm, const void* x, int C, int64_t sB, int64_t sT, int64_t sF) -> int {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)d->Fi, (cuuint64_t)d->Ti, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)sF * 2, (cuuint64_t)sT * 2, (cuuint64_t)sB * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.gw_a, (cuuint32_t)frows, (cuuint32_t)p.t_tile, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return (int)enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes((int)p.pitch_a), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  int rc = enc_x(&tmX, d->x0, d->c0, d->x0_sB, d->x0_sT, d->x0_sF);
  if (!rc && d->c1) rc = enc_x(&tmX1, d->x1, d->c1, d->x1_sB, d->x1_sT, d->x1_sF);
  if (rc) { set_error("clskd_tapconv_wgrad_umma_stacked: cuTensorMapEncodeTiled(x) failed: %d", rc); return CLSKD_ERR_CUDA; }
  if (!d->c1) tmX1 = tmX;
  {
    const int brows = p.t_tile == 1 ? 128 + p.nf - 1 : d->Fo + 8;
    cuuint64_t dims[4] = {(cuuint64_t)d->N, (cuuint64_t)d->Fo, (cuuint64_t)d->To, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->y_sF * 2, (cuuint64_t)d->y_sT * 2, (cuuint64_t)d->y_sB * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.Np, (cuuint32_t)brows, (cuuint32_t)p.t_tile, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmDY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->y), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes((int)p.pitch_b), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { set_error("clskd_tapconv_wgrad_umma_stacked: cuTensorMapEncodeTiled(dy) failed: %d", (int)r); return CLSKD_ERR_CUDA; }
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    e = cudaFuncSetAttribute(tapconv_wgrad_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("clskd_tapconv_wgrad_umma_stacked: smem attr: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
    smem_set = smem;
  }
  dim3 grid((unsigned)nsplit, (unsigned)ycount);
  tapconv_wgrad_stack_kernel<<<grid, kThreads, smem, st>>>(tmX, tmX1, tmDY, p);
  CLSKD_CHECK_LAUNCH("clskd_tapconv_wgrad_umma_stacked");
  return CLSKD_OK;
}
