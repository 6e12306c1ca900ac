// Weight gradient of the tap-list implicit GEMM on the tcgen05 tensor cores.
//
//   dW[j][c][n] = sum_{b,to,fo} X[b, to+dt[j], fo*sf+df[j], c] * dY[b,to,fo,n]        (fp32 out)
//
// GEMM view per tap j: D_j [C x N] = A_j^T [C x rows] * dY [rows x N].  The contraction runs over
// the OUTPUT ROWS, so both operands are "MN-major" for the tensor core: A_j has the channel axis
// contiguous (UMMA M = channels), dY has n contiguous (UMMA N = n).  Both come straight from the
// channels-last tensors by TMA: a 128-row (t_tile x fo_tile) patch of one utterance is one box per
// 64/32/16-channel swizzle group, the tap only shifts the box coordinates (zero fill outside the
// tensor = conv padding), the skip connection is a second tensor map (no materialised concat).
//
// Operand reuse: taps are grouped and a group loads ONE activation patch per row tile; each tap's A tile is that
// patch read through a shared-memory descriptor shifted by whole rows (the swizzle XOR is a function of the absolute
// address: profiles/r02_hw_desc_shift.log) - "full" halo patch for 128-wide frequency tiles at stride 1, time-grouped
// patches otherwise, one box per tap as the fallback.
//
// One CTA owns (tap group g, 128-channel tile, n tile) and a contiguous range of row patches: the
// accumulators D_j of the G = 512/n_tile taps of its group live in TMEM for the whole range (the
// dY patch is loaded once per row patch and shared by the G taps), and are added to dW with fp32
// reductions at the end (split-K over CTAs).
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-5
// epilogue.
#include "umma.cuh"

namespace clskd {
int g_wgrad_mode = 0;                  // clskd_set_tuning key 6: 1 one box per tap, 2 time-grouped patches at most
namespace {
using namespace umma;

constexpr int kThreads = 192;
constexpr int ROWS = 128;              // rows (GEMM K) per smem patch
constexpr int MAX_A_STAGES = 4;

struct WgradParams {
  int B, To, Fo;
  int t_tile, fo_tile, f_tiles, t_tiles;
  int n_row_tiles, tiles_per_cta;
  int ntaps, G, ngroups, c_tiles, n_tiles, n_tile;
  // taps in patch-group order: tap_w = tap index in dW, tap_pg = patch group, tap_roff = first row of the tap's
  // 128-row tile inside the patch; patch group coordinates (parity, f shift, t shift) relative to the tile origin
  int tap_w[CLSKD_MAX_TAPS], tap_pg[CLSKD_MAX_TAPS], tap_roff[CLSKD_MAX_TAPS];
  int pg_p[CLSKD_MAX_TAPS], pg_f[CLSKD_MAX_TAPS], pg_t[CLSKD_MAX_TAPS];
  uint32_t a_stage_bytes;              // one patch: nsub_a sub-blocks of a_sub_bytes
  int gw_a, gw_b;                      // swizzle group widths in elements (64 / 32 / 16)
  int c0, Ctot, N;
  uint32_t a_sub_bytes, b_sub_bytes, b_stage_bytes;
  uint32_t layout_a, layout_b;
  int a_stages;
  uint32_t tmem_cols;
  float* dw;
};

__global__ void __launch_bounds__(kThreads, 1)
tapconv_wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                          const __grid_constant__ CUtensorMap tmDY, const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t a_full[MAX_A_STAGES], a_empty[MAX_A_STAGES];
  __shared__ __align__(8) uint64_t b_full[2], b_empty[2];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint8_t* b_buf = base;                                   // 2 x b_stage_bytes
  uint8_t* a_buf = base + 2 * (size_t)p.b_stage_bytes;     // a_stages x 32 KB

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work decode: blockIdx.y -> (tap group, channel tile, n tile); blockIdx.x -> row-patch range
  int w = blockIdx.y;
  const int n_t = w % p.n_tiles;
  w /= p.n_tiles;
  const int c_t = w % p.c_tiles;
  const int grp = w / p.c_tiles;
  const int tap0 = grp * p.G;
  const int gcur = min(p.G, p.ntaps - tap0);
  const int n0 = n_t * p.n_tile;
  const int cbase = c_t * 128;
  const int cvalid = min(128, p.Ctot - cbase);
  const int nsub_a = cvalid / p.gw_a;
  const int nsub_b = p.n_tile / p.gw_b;
  const int tile_beg = blockIdx.x * p.tiles_per_cta;
  const int tile_end = min(p.n_row_tiles, tile_beg + p.tiles_per_cta);
  const int ntile_cta = tile_end - tile_beg;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer (all lanes run the loops, one elected lane issues) =====================
    if (ntile_cta > 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < ntile_cta; ++it) {
        int r = tile_beg + it;
        const int f_blk = r % p.f_tiles;
        r /= p.f_tiles;
        const int t_blk = r % p.t_tiles;
        const int b = r / p.t_tiles;
        const int t0 = t_blk * p.t_tile, f0 = f_blk * p.fo_tile;
        const int bs = it & 1;
        mbar_wait(&b_empty[bs], ((it >> 1) & 1) ^ 1u);
        mbar_expect_tx_warp(&b_full[bs], (uint32_t)nsub_b * p.b_sub_bytes);
        for (int s = 0; s < nsub_b; ++s)
          tma_load_4d_warp(b_buf + (size_t)bs * p.b_stage_bytes + (size_t)s * p.b_sub_bytes, &tmDY, &b_full[bs],
                      n0 + s * p.gw_b, f0, t0, b);
        for (int g = 0; g < gcur; ++g) {
          const int j = tap0 + g;
          if (g > 0 && p.tap_pg[j] == p.tap_pg[j - 1]) continue;        // same patch as the previous tap
          const int pg = p.tap_pg[j];
          mbar_wait(&a_empty[stage], phase ^ 1u);
          mbar_expect_tx_warp(&a_full[stage], (uint32_t)nsub_a * p.a_sub_bytes);
          for (int s = 0; s < nsub_a; ++s) {
            const int cc = cbase + s * p.gw_a;
            const bool src0 = cc < p.c0;
            tma_load_5d_warp(a_buf + (size_t)stage * p.a_stage_bytes + (size_t)s * p.a_sub_bytes, src0 ? &tmA0 : &tmA1,
                        &a_full[stage], src0 ? cc : cc - p.c0, p.pg_p[pg], f0 + p.pg_f[pg], t0 + p.pg_t[pg], b);
          }
          if (++stage == p.a_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (all lanes run the loops, one elected lane issues) =====================
    if (ntile_cta > 0) {
      // D=f32, A=B=bf16, both operands MN-major (bits 15/16), N = n_tile, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t pitch_a = (uint32_t)p.gw_a * 2u, pitch_b = (uint32_t)p.gw_b * 2u;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < ntile_cta; ++it) {
        const int bs = it & 1;
        mbar_wait(&b_full[bs], (it >> 1) & 1);
        const uint32_t b_addr = smem_u32(b_buf + (size_t)bs * p.b_stage_bytes);
        for (int g = 0; g < gcur; ++g) {
          const int j = tap0 + g;
          if (g == 0 || p.tap_pg[j] != p.tap_pg[j - 1]) {               // first tap of a patch: wait for it
            mbar_wait(&a_full[stage], phase);
            fence_after();
          }
          const uint32_t a_addr = smem_u32(a_buf + (size_t)stage * p.a_stage_bytes) + (uint32_t)p.tap_roff[j] * pitch_a;
#pragma unroll
          for (int k = 0; k < ROWS / 16; ++k) {
            const uint64_t adesc = make_smem_desc_lbo(a_addr + (uint32_t)k * 16u * pitch_a, p.a_sub_bytes >> 4,
                                                      (8u * pitch_a) >> 4, p.layout_a);
            const uint64_t bdesc = make_smem_desc_lbo(b_addr + (uint32_t)k * 16u * pitch_b, p.b_sub_bytes >> 4,
                                                      (8u * pitch_b) >> 4, p.layout_b);
            umma_bf16_warp(tmem_base + (uint32_t)(g * p.n_tile), adesc, bdesc, idesc, (it | k) ? 1u : 0u);
          }
          if (g == gcur - 1 || p.tap_pg[j + 1] != p.tap_pg[j]) {        // last tap of the patch: release the stage
            umma_commit_warp(&a_empty[stage]);
            if (++stage == p.a_stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        umma_commit_warp(&b_empty[bs]);
      }
      umma_commit_warp(&tmem_full_bar);
    }
  } else if (ntile_cta > 0) {
    // ===================== epilogue (warps 2..5): TMEM -> fp32 reductions into dW =====================
    const int q = warp & 3;
    const int c_glob = cbase + q * 32 + lane;
    const bool valid = (q * 32 + lane) < cvalid;
    mbar_wait(&tmem_full_bar, 0);
    fence_after();
    for (int g = 0; g < gcur; ++g) {
      float* dst = p.dw + ((int64_t)p.tap_w[tap0 + g] * p.Ctot + c_glob) * p.N + n0;
      for (int c = 0; c < p.n_tile; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * p.n_tile + c), v);
        if (valid) red_add16(dst + c, v, 16);
      }
    }
  }

  fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

int pick_gw(int a, int b) {
  if (a % 64 == 0 && b % 64 == 0) return 64;
  if (a % 32 == 0 && b % 32 == 0) return 32;
  return 16;
}

const char* wgrad_unsupported(const ClskdTapConv* d) {
  if (d->x_dtype != CLSKD_BF16 || d->y_dtype != CLSKD_BF16) return "x and dy must be bf16";
  if (d->c0 % 8 || d->c1 % 8 || d->c0 < 8) return "channels must be multiples of 8";      // padded to 16 by the round-1 kernel
  if (d->N % 8 || d->N < 8) return "N must be a multiple of 8";
  if (d->N > 256 && d->N % 128) return "N > 256 must be a multiple of 128";
  if (d->sf != 1 && d->sf != 2) return "sf must be 1 or 2";
  if (!is_pow2(d->Fo) || (d->Fo > 128 && d->Fo % 128)) return "Fo must be a power of two";
  if (d->Fi % d->sf) return "Fi must be a multiple of sf";
  auto chk = [&](const void* x, int64_t sB, int64_t sT, int64_t sF) -> const char* {
    if ((uintptr_t)x % 16) return "tensor not 16-byte aligned";
    if ((sB * 2) % 16 || (sT * 2) % 16 || (sF * 2) % 16) return "strides not 16-byte multiples";
    return nullptr;
  };
  if (const char* r = chk(d->x0, d->x0_sB, d->x0_sT, d->x0_sF)) return r;
  if (d->c1)
    if (const char* r = chk(d->x1, d->x1_sB, d->x1_sT, d->x1_sF)) return r;
  if (const char* r = chk(d->y, d->y_sB, d->y_sT, d->y_sF)) return r;
  if ((int64_t)d->B * d->To * d->Fo < 512) return "too few rows to amortise the split-K epilogue";
  if (!get_encode()) return "cuTensorMapEncodeTiled unavailable";
  return nullptr;
}

}  // namespace
}  // namespace clskd

using namespace clskd;

extern "C" int clskd_tapconv_wgrad_umma_supported(const ClskdTapConv* d) {
  if (!d || !d->x0 || !d->y || !d->w) return 0;
  return wgrad_unsupported(d) == nullptr ? 1 : 0;
}

extern "C" int clskd_tapconv_wgrad_umma(const ClskdTapConv* d, void* stream) {
  CLSKD_CHECK_ARG(d && d->x0 && d->w && d->y, "clskd_tapconv_wgrad_umma: null pointer");
  CLSKD_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= CLSKD_MAX_TAPS, "clskd_tapconv_wgrad_umma: ntaps");
  if (const char* why = wgrad_unsupported(d)) {
    set_error("clskd_tapconv_wgrad_umma: unsupported: %s", why);
    return CLSKD_ERR_UNSUPPORTED;
  }
  // narrow-N stride-1 grids of taps: frequency taps stacked along the MMA's N (tapconv_wgrad_stack.cu)
  if (g_wgrad_mode == 0 && clskd_tapconv_wgrad_umma_stacked_supported(d)) return clskd_tapconv_wgrad_umma_stacked(d, stream);
  cudaStream_t st = (cudaStream_t)stream;
  const int Ctot = d->c0 + d->c1;
  if (!d->accumulate) {
    cudaError_t e = cudaMemsetAsync(const_cast<void*>(d->w), 0, sizeof(float) * (size_t)d->ntaps * Ctot * d->N, st);
    if (e != cudaSuccess) { set_error("clskd_tapconv_wgrad_umma: memset: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
  }
  EncodeTiledFn enc = get_encode();
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.To = d->To; p.Fo = d->Fo;
  p.fo_tile = d->Fo < ROWS ? d->Fo : ROWS;
  p.t_tile = ROWS / p.fo_tile;
  p.f_tiles = d->Fo / p.fo_tile;
  p.t_tiles = cdiv(d->To, p.t_tile);
  const int64_t nrt = (int64_t)d->B * p.t_tiles * p.f_tiles;
  CLSKD_CHECK_ARG(nrt <= 2147483647LL, "clskd_tapconv_wgrad_umma: too many row patches");
  p.n_row_tiles = (int)nrt;
  p.ntaps = d->ntaps;
  int tt[CLSKD_MAX_TAPS], tpar[CLSKD_MAX_TAPS], tf[CLSKD_MAX_TAPS];
  int tmin = 1 << 30, tmax = -(1 << 30), fmin = 1 << 30, fmax = -(1 << 30);
  for (int j = 0; j < d->ntaps; ++j) {
    int df = d->df[j];
    int fl = df >= 0 ? df / d->sf : -((-df + d->sf - 1) / d->sf);
    tt[j] = d->dt[j];
    tf[j] = fl;
    tpar[j] = df - fl * d->sf;
    tmin = tt[j] < tmin ? tt[j] : tmin; tmax = tt[j] > tmax ? tt[j] : tmax;
    fmin = fl < fmin ? fl : fmin; fmax = fl > fmax ? fl : fmax;
  }
  p.n_tile = d->N <= 256 ? d->N : (d->N % 256 == 0 ? 256 : 128);
  p.n_tiles = d->N / p.n_tile;
  p.G = 512 / p.n_tile;
  if (p.G > d->ntaps) p.G = d->ntaps;
  p.c_tiles = cdiv(Ctot, 128);
  p.gw_a = pick_gw(d->c0, d->c1);
  p.gw_b = pick_gw(p.n_tile, 0);
  p.c0 = d->c0; p.Ctot = Ctot; p.N = d->N;
  p.b_sub_bytes = (uint32_t)ROWS * p.gw_b * 2;
  p.b_stage_bytes = ((uint32_t)ROWS * p.n_tile * 2 + 1023u) & ~1023u;
  p.layout_a = layout_for_bytes(p.gw_a * 2);
  p.layout_b = layout_for_bytes(p.gw_b * 2);
  // the MMA always reads M = 128 channels = 128 / gw_a sub-blocks (rows of absent channels are discarded by the
  // epilogue), so a stage must span all of them even when fewer are loaded
  const int nsub_max = 128 / p.gw_a;
  const uint32_t kBudget = 222u * 1024u;
  // ---- patch grouping: 3 full halo patch (all taps in one CTA), 2 time-grouped patches, 1 one box per tap
  const bool full_ok = d->sf == 1 && p.fo_tile == ROWS && d->ntaps > 1 && p.G >= d->ntaps &&
                       ((ROWS + (fmax - fmin) + 7) & ~7) <= 256 && (p.t_tile + (tmax - tmin)) <= 16;
  const bool time_ok = d->ntaps > 1 && p.fo_tile % 8 == 0 && p.t_tile > 1 && tmax > tmin && (p.t_tile + (tmax - tmin)) <= 256;
  int mode = full_ok ? 3 : (time_ok ? 2 : 1);
  // measured (profiles/r02_step_breakdown_d/e.json): patches pay for wide tiles (N >= 128 with >= 64 channels: -12 %);
  // narrow-N launches are bound by the shared-memory operand reads of their MMAs and by stage depth, where the larger
  // patch stages lose (+4 .. +15 %): those keep one box per tap.  g_wgrad_mode = 3 forces patches for tests.
  if ((g_wgrad_mode == 0 || g_wgrad_mode == 5) && !(p.n_tile >= 128 && Ctot >= 64)) mode = 1;
  if (g_wgrad_mode == 1) mode = 1;
  const bool padded = (d->c0 % 16) || (d->c1 % 16) || (d->N % 16);
  if (padded) mode = 1;
  if ((mode == 1 && g_wgrad_mode != 4) || padded)
    return clskd_tapconv_wgrad_umma_v1(d, stream);     // (dW already zeroed: harmless)
  if (g_wgrad_mode == 2) mode = time_ok ? 2 : 1;
  int box_f = p.fo_tile, box_t = p.t_tile;
  int stages = 0;
  for (;;) {
    box_f = p.fo_tile;
    box_t = p.t_tile;
    int npg = 0;
    if (mode == 3) {
      box_f = (ROWS + (fmax - fmin) + 7) & ~7;
      box_t = p.t_tile + (tmax - tmin);
      npg = 1;
      p.pg_p[0] = 0; p.pg_f[0] = fmin; p.pg_t[0] = tmin;
      for (int j = 0; j < d->ntaps; ++j) {
        p.tap_w[j] = j;
        p.tap_pg[j] = 0;
        p.tap_roff[j] = (tt[j] - tmin) * box_f + (tf[j] - fmin);
      }
    } else if (mode == 2) {
      box_t = p.t_tile + (tmax - tmin);
      bool used[CLSKD_MAX_TAPS] = {false};
      int n = 0;
      for (int j = 0; j < d->ntaps; ++j) {
        if (used[j]) continue;
        p.pg_p[npg] = tpar[j]; p.pg_f[npg] = tf[j]; p.pg_t[npg] = tmin;
        for (int i = j; i < d->ntaps; ++i)
          if (!used[i] && tpar[i] == tpar[j] && tf[i] == tf[j]) {
            used[i] = true;
            p.tap_w[n] = i;
            p.tap_pg[n] = npg;
            p.tap_roff[n] = (tt[i] - tmin) * box_f;
            ++n;
          }
        ++npg;
      }
      // taps per CTA: whole patch groups where possible (a group cut by the CTA boundary is loaded twice)
      const int gsz = tmax - tmin + 1;
      if (p.G < d->ntaps && p.G > gsz && d->ntaps % gsz == 0) p.G = p.G / gsz * gsz;
    } else {
      npg = d->ntaps;
      for (int j = 0; j < d->ntaps; ++j) {
        p.tap_w[j] = j;
        p.tap_pg[j] = j;
        p.tap_roff[j] = 0;
        p.pg_p[j] = tpar[j]; p.pg_f[j] = tf[j]; p.pg_t[j] = tt[j];
      }
    }
    p.a_sub_bytes = (uint32_t)box_f * box_t * p.gw_a * 2;
    p.a_stage_bytes = ((uint32_t)nsub_max * p.a_sub_bytes + 1023u) & ~1023u;
    stages = (int)((kBudget - 1024u - 2u * p.b_stage_bytes) / p.a_stage_bytes);
    if (stages > MAX_A_STAGES) stages = MAX_A_STAGES;
    if (stages >= 2 || mode == 1) break;
    mode = mode == 3 && time_ok ? 2 : 1;          // the patch does not fit twice: smaller patches
  }
  if (stages < 2) stages = 2;
  p.a_stages = stages;
  p.ngroups = cdiv(d->ntaps, p.G);
  int cols = 32;
  while (cols < p.G * p.n_tile) cols <<= 1;
  p.tmem_cols = (uint32_t)cols;
  p.dw = reinterpret_cast<float*>(const_cast<void*>(d->w));

  const int ycount = p.ngroups * p.c_tiles * p.n_tiles;
  int nsplit = (2 * sm_count() + ycount - 1) / ycount;
  if (nsplit > p.n_row_tiles) nsplit = p.n_row_tiles;
  if (nsplit < 1) nsplit = 1;
  p.tiles_per_cta = cdiv(p.n_row_tiles, nsplit);
  nsplit = cdiv(p.n_row_tiles, p.tiles_per_cta);

  CUtensorMap tmA0, tmA1, tmDY;
  CUtensorMapSwizzle swa = swizzle_for_bytes(p.gw_a * 2), swb = swizzle_for_bytes(p.gw_b * 2);
  int rc = encode_act(enc, &tmA0, d->x0, d->c0, d->sf, d->Fi, d->Ti, d->B, d->x0_sB, d->x0_sT, d->x0_sF, p.gw_a,
                      box_f, box_t, swa);
  if (rc) { set_error("clskd_tapconv_wgrad_umma: cuTensorMapEncodeTiled(x0) failed: %d", rc); return CLSKD_ERR_CUDA; }
  if (d->c1) {
    rc = encode_act(enc, &tmA1, d->x1, d->c1, d->sf, d->Fi, d->Ti, d->B, d->x1_sB, d->x1_sT, d->x1_sF, p.gw_a,
                    box_f, box_t, swa);
    if (rc) { set_error("clskd_tapconv_wgrad_umma: cuTensorMapEncodeTiled(x1) failed: %d", rc); return CLSKD_ERR_CUDA; }
  } else {
    tmA1 = tmA0;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->N, (cuuint64_t)d->Fo, (cuuint64_t)d->To, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->y_sF * 2, (cuuint64_t)d->y_sT * 2, (cuuint64_t)d->y_sB * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.gw_b, (cuuint32_t)p.fo_tile, (cuuint32_t)p.t_tile, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmDY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->y), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swb, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { set_error("clskd_tapconv_wgrad_umma: cuTensorMapEncodeTiled(dy) failed: %d", (int)r); return CLSKD_ERR_CUDA; }
  }
  size_t smem = 2 * (size_t)p.b_stage_bytes + (size_t)p.a_stages * p.a_stage_bytes + 1024;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(tapconv_wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("clskd_tapconv_wgrad_umma: smem attr: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
    smem_set = smem;
  }
  dim3 grid((unsigned)nsplit, (unsigned)ycount);
  tapconv_wgrad_umma_kernel<<<grid, kThreads, smem, st>>>(tmA0, tmA1, tmDY, p);
  CLSKD_CHECK_LAUNCH("clskd_tapconv_wgrad_umma");
  return CLSKD_OK;
}
