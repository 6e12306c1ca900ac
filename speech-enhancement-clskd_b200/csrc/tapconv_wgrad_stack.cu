// Weight gradient of a convolution with the FREQUENCY TAPS STACKED ALONG THE MMA's N DIMENSION.
//
//   dW[(dt,df)][c][n] = sum_{b,t,f} X[b,t+dt,f+df,c] dY[b,t,f,n] = sum_{b,t',f'} X[b,t',f',c] dY[b,t'-dt,f'-df,n]
//
// (stride 2 along f - the encoder layers - splits the input frequencies by parity: for input f'' = 2j + par only the
// taps with df = par (mod 2) contribute, with dY[t'-dt, j - fl], fl = (df - par) / 2: the same problem per parity class
// on the sub-sampled input, which TMA delivers through a parity coordinate.)
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
constexpr int MAX_DT = 4, MAX_DF = 8, MAX_GRP = 8;

struct StackParams {
  int B, T, F;                   // output extents (F = input frequencies / sf)
  int t_tile, f_tiles, t_tiles;
  int n_row_tiles, tiles_per_cta;
  int a_rows;                    // K rows of one tile (multiple of 16)
  int a_fbox_start;              // frequency coordinate of the X box relative to f0 (0 or -1)
  int b_fbox_start;              // ... of the dY box: a_fbox_start - fl_max
  int npar;                      // parity classes of the input frequency axis (= sf)
  int Np, N, Ctot;               // padded / real N, real channels (c0 + c1)
  int c0, c0p, c1r, Ctot_p;      // source 0 real / padded to 16, real channels of source 1, padded total
  int gw_a;                      // channels per swizzle group of the X tile (64 / 32 / 16)
  uint32_t pitch_a, layout_a;
  int ndt, gd, ngd;              // time taps, time taps per CTA, number of such chunks
  int c_tiles;
  int dts[MAX_DT];
  // group g = (time tap g / npar, parity g % npar): its taps are nf consecutive frequency shifts
  int grp_nf[MAX_GRP];
  int grp_shift[MAX_GRP];        // patch row of the group's first slot (slot s reads the patch shifted by shift + s rows)
  int grp_col[MAX_GRP];          // accumulator column inside the CTA's chunk
  int tap_of[MAX_GRP][MAX_DF];   // index into the caller's tap list of (group, slot)
  uint32_t a_sub_bytes;          // one channel group of one X tile
  uint32_t a_tile_bytes;         // the 128 channel rows of the MMA (one parity class)
  uint32_t b_patch_bytes, stage_bytes;
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
  const int g0 = gchunk * p.gd;                         // first time tap of this CTA
  const int gcur = min(p.gd, p.ndt - g0);
  const int cbase = c_t * 128;
  const int cvalid = min(128, p.Ctot_p - cbase);        // padded channel axis: [c0 -> c0p | c1 -> c1p]
  const int nsub_a = cvalid / p.gw_a;
  const int tile_beg = blockIdx.x * p.tiles_per_cta;
  const int tile_end = min(p.n_row_tiles, tile_beg + p.tiles_per_cta);
  const int ntile_cta = tile_end - tile_beg;
  const uint32_t b_off = (uint32_t)p.npar * p.a_tile_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  // The shifted sub-blocks read a few rows past what TMA writes into a dY patch: those bytes only ever multiply
  // zero-filled X rows, but they must be finite.  Zero the tail of every patch slot once.
  for (int s = 0; s < p.stages; ++s)
    for (int g = 0; g < p.gd; ++g) {
      uint8_t* patch = ring + (size_t)s * p.stage_bytes + b_off + (size_t)g * p.b_patch_bytes;
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
        const int t0 = t_blk * p.t_tile, f0 = f_blk * 128;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_expect_tx_warp(&full_bar[stage], (uint32_t)(p.npar * nsub_a) * p.a_sub_bytes + (uint32_t)gcur * p.b_tx);
        uint8_t* a_dst = ring + (size_t)stage * p.stage_bytes;
        for (int par = 0; par < p.npar; ++par)
          for (int s = 0; s < nsub_a; ++s) {
            const int cc = cbase + s * p.gw_a;
            const bool src0 = cc < p.c0p;        // channels beyond a source's real extent are zero-filled by TMA
            uint8_t* dst = a_dst + (size_t)par * p.a_tile_bytes + (size_t)s * p.a_sub_bytes;
            if (p.npar == 1)       // 4-d map at stride 1 (the 5-d form with a unit parity axis measured 20 % slower)
              tma_load_4d_warp(dst, src0 ? &tmX : &tmX1, &full_bar[stage], src0 ? cc : cc - p.c0p, f0 + p.a_fbox_start, t0, b);
            else
              tma_load_5d_warp(dst, src0 ? &tmX : &tmX1, &full_bar[stage], src0 ? cc : cc - p.c0p, par, f0 + p.a_fbox_start, t0, b);
          }
        for (int g = 0; g < gcur; ++g)
          tma_load_4d_warp(a_dst + b_off + (size_t)g * p.b_patch_bytes, &tmDY, &full_bar[stage], 0, f0 + p.b_fbox_start,
                      t0 - p.dts[g0 + g], b);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (all lanes run the loops; one elected lane issues) =====================
    if (ntile_cta > 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int ksteps = p.a_rows / 16;
      // per-group constants in registers (static indexing: the issue loop of this single thread paces the tensor pipe)
      const int ng = gcur * p.npar, gbase = g0 * p.npar;
      uint32_t g_idesc[MAX_GRP], g_a[MAX_GRP], g_b[MAX_GRP], g_col[MAX_GRP];
#pragma unroll
      for (int u = 0; u < MAX_GRP; ++u) {
        const int gi = gbase + (u < ng ? u : 0);
        const int ncols = p.grp_nf[gi] * p.Np;
        // D = f32, A = B = bf16, both MN-major, N = nf * Np, M = 128
        g_idesc[u] = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(ncols >> 3) << 17) |
                     ((uint32_t)(128 >> 4) << 24);
        g_a[u] = (uint32_t)(gi % p.npar) * p.a_tile_bytes;
        g_b[u] = b_off + (uint32_t)(gi / p.npar - g0) * p.b_patch_bytes + (uint32_t)p.grp_shift[gi] * p.pitch_b;
        g_col[u] = (uint32_t)p.grp_col[gi];
      }
      const uint32_t a_lbo = p.a_sub_bytes >> 4, a_sbo = (8u * p.pitch_a) >> 4, b_lbo = p.pitch_b >> 4, b_sbo = (8u * p.pitch_b) >> 4;
      const uint32_t a_kstep = 16u * p.pitch_a, b_kstep = 16u * p.pitch_b;
      for (int it = 0; it < ntile_cta; ++it) {
        mbar_wait(&full_bar[stage], phase);
        fence_after();
        const uint32_t st_addr = smem_u32(ring + (size_t)stage * p.stage_bytes);
#pragma unroll
        for (int u = 0; u < MAX_GRP; ++u) {
          if (u < ng) {
            const uint32_t a_addr = st_addr + g_a[u], b_addr = st_addr + g_b[u];
            for (int k = 0; k < ksteps; ++k) {
              // X: channel groups at LBO = a_sub_bytes, 8-row groups at SBO = 8 rows (sub-blocks past the tile's
              // channels read stale shared memory: their accumulator rows are never stored)
              const uint64_t adesc = make_smem_desc_lbo(a_addr + (uint32_t)k * a_kstep, a_lbo, a_sbo, p.layout_a);
              // dY: frequency taps at LBO = ONE ROW of the patch
              const uint64_t bdesc = make_smem_desc_lbo(b_addr + (uint32_t)k * b_kstep, b_lbo, b_sbo, p.layout_b);
              umma_bf16_warp(tmem_base + g_col[u], adesc, bdesc, g_idesc[u], (it | k) ? 1u : 0u);
            }
          }
        }
        umma_commit_warp(&empty_bar[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit_warp(&tmem_full_bar);
    }
  } else if (ntile_cta > 0) {
    // ===================== epilogue (warps 2..5): TMEM -> fp32 reductions into dW =====================
    const int q = warp & 3;
    const int c_pad = cbase + q * 32 + lane;                 // index on the padded channel axis
    const int c = c_pad < p.c0p ? c_pad : p.c0 + (c_pad - p.c0p);
    const bool valid = (q * 32 + lane) < cvalid && (c_pad < p.c0p ? c_pad < p.c0 : (c_pad - p.c0p) < p.c1r);
    mbar_wait(&tmem_full_bar, 0);
    fence_after();
    for (int gi = g0 * p.npar; gi < (g0 + gcur) * p.npar; ++gi)
      for (int s = 0; s < p.grp_nf[gi]; ++s) {
        const int tap = p.tap_of[gi][s];
        for (int cc = 0; cc < p.Np; cc += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.grp_col[gi] + s * p.Np + cc), v);
          if (valid && tap >= 0) {
            float* dst = p.dw + ((int64_t)tap * p.Ctot + c) * p.N + cc;
            red_add16(dst, v, p.N - cc);
          }
        }
      }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// geometry the stacked kernel covers; fills the tap groups.  nullptr = supported
const char* stack_unsupported(const ClskdTapConv* d, StackParams* out) {
  if (d->x_dtype != CLSKD_BF16 || d->y_dtype != CLSKD_BF16) return "x and dy must be bf16";
  if (d->sf != 1 && d->sf != 2) return "stride 1 or 2 only";
  // (To may differ from Ti - the decoder's sub-pixel phases produce Ti + 1 rows: the tiles walk the INPUT rows, and dY
  // rows outside [0, To) are zero-filled by its tensor map)
  if (d->Fi != d->Fo * d->sf) return "Fi = sf * Fo only";
  if (d->To < d->Ti - 4 || d->To > d->Ti + 4) return "To too far from Ti";
  if (d->c0 % 8 || d->c1 % 8 || d->c0 < 8) return "channels must be multiples of 8";
  const int Np = (d->N + 15) & ~15;
  if (Np != 16 && Np != 32 && Np != 64) return "N must pad to 16, 32 or 64";
  if (d->N % 8) return "N must be a multiple of 8";
  if (!is_pow2(d->Fo) || d->Fo < 16 || (d->Fo > 128 && d->Fo % 128)) return "Fo must be a power of two >= 16";
  if (d->accumulate) return "accumulate unsupported";
  StackParams p;
  memset(&p, 0, sizeof(p));
  const int sf = d->sf;
  // time taps; frequency taps as (fl, parity): df = fl * sf + parity
  int fl[CLSKD_MAX_TAPS], par[CLSKD_MAX_TAPS], tg[CLSKD_MAX_TAPS];
  int ndt = 0, flmin = 1 << 30, flmax = -(1 << 30);
  for (int j = 0; j < d->ntaps; ++j) {
    const int df = d->df[j];
    fl[j] = df >= 0 ? df / sf : -((-df + sf - 1) / sf);
    par[j] = df - fl[j] * sf;
    int g = 0;
    while (g < ndt && p.dts[g] != d->dt[j]) ++g;
    if (g == ndt) {
      if (ndt == MAX_DT) return "too many time taps";
      p.dts[ndt++] = d->dt[j];
    }
    tg[j] = g;
    flmin = fl[j] < flmin ? fl[j] : flmin;
    flmax = fl[j] > flmax ? fl[j] : flmax;
  }
  const int nfl = flmax - flmin + 1;
  if (nfl > MAX_DF) return "frequency span too wide";
  if (ndt * sf > MAX_GRP) return "too many tap groups";
  int cols_per_dt = 0, stacked = 0;
  for (int g = 0; g < ndt * sf; ++g) {
    const int gt = g / sf, gp = g % sf;
    int hi = -(1 << 30), lo = 1 << 30, cnt = 0;
    for (int j = 0; j < d->ntaps; ++j)
      if (tg[j] == gt && par[j] == gp) {
        hi = fl[j] > hi ? fl[j] : hi;
        lo = fl[j] < lo ? fl[j] : lo;
        ++cnt;
      }
    if (cnt == 0 || hi - lo + 1 != cnt) return "taps are not a grid of consecutive frequency shifts";
    p.grp_nf[g] = cnt;
    p.grp_shift[g] = flmax - hi;
    for (int s = 0; s < cnt; ++s) {
      p.tap_of[g][s] = -1;
      for (int j = 0; j < d->ntaps; ++j)
        if (tg[j] == gt && par[j] == gp && fl[j] == hi - s) {
          if (p.tap_of[g][s] >= 0) return "duplicate tap";
          p.tap_of[g][s] = j;
        }
      if (p.tap_of[g][s] < 0) return "taps are not a grid of consecutive frequency shifts";
    }
    if (cnt * Np > 256) return "stacked N exceeds 256";
    stacked = cnt > stacked ? cnt : stacked;
    if (gt == 0) cols_per_dt += cnt * Np;
  }
  for (int g = sf; g < ndt * sf; ++g)
    if (p.grp_nf[g] != p.grp_nf[g % sf]) return "time taps differ in their frequency taps";
  if (stacked < 2) return "nothing to stack";
  if (cols_per_dt > 512) return "accumulators of one time tap exceed TMEM";
  auto chk = [&](const void* x, int64_t sB, int64_t sT, int64_t sF) -> bool {
    return (uintptr_t)x % 16 == 0 && (sB * 2) % 16 == 0 && (sT * 2) % 16 == 0 && (sF * 2) % 16 == 0;
  };
  if (!chk(d->x0, d->x0_sB, d->x0_sT, d->x0_sF) || !chk(d->y, d->y_sB, d->y_sT, d->y_sF)) return "alignment";
  if (d->c1 && !chk(d->x1, d->x1_sB, d->x1_sT, d->x1_sF)) return "alignment";
  if ((int64_t)d->B * d->Ti * d->Fo < 65536) return "too few rows";
  if (!get_encode()) return "cuTensorMapEncodeTiled unavailable";
  p.B = d->B; p.T = d->Ti; p.F = d->Fo;
  p.npar = sf;
  p.ndt = ndt; p.Np = Np; p.N = d->N; p.Ctot = d->c0 + d->c1;
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
  p.b_fbox_start = p.a_fbox_start - flmax;
  p.t_tiles = cdiv(d->Ti, p.t_tile);
  const int64_t nrt = (int64_t)d->B * p.t_tiles * p.f_tiles;
  if (nrt > 2147483647LL) return "too many row tiles";
  p.n_row_tiles = (int)nrt;
  p.gd = 512 / cols_per_dt;
  if (p.gd > ndt) p.gd = ndt;
  p.ngd = cdiv(ndt, p.gd);
  for (int g = 0; g < ndt * sf; ++g) {
    const int first = (g / sf / p.gd) * p.gd * sf;            // first group of this group's CTA chunk
    int col = 0;
    for (int h = first; h < g; ++h) col += p.grp_nf[h] * Np;
    p.grp_col[g] = col;
  }
  p.c_tiles = cdiv(p.Ctot_p, 128);
  p.pitch_b = (uint32_t)Np * 2u;
  p.layout_b = layout_for_bytes((int)p.pitch_b);
  p.a_sub_bytes = (uint32_t)p.a_rows * p.pitch_a;
  p.a_tile_bytes = (uint32_t)(128 / p.gw_a) * p.a_sub_bytes;            // 128 channel rows of the MMA
  p.b_tx = (uint32_t)(p.a_rows + (p.t_tile == 1 ? nfl - 1 : 0)) * p.pitch_b;
  p.b_patch_bytes = (p.b_tx + (uint32_t)(nfl - 1) * p.pitch_b + 1023u) & ~1023u;           // shifted reads stay inside
  p.stage_bytes = (uint32_t)sf * p.a_tile_bytes + (uint32_t)p.gd * p.b_patch_bytes;
  int stages = (int)((220u * 1024u) / p.stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return "stage too large";
  p.stages = stages;
  int cols = 32;
  while (cols < p.gd * cols_per_dt) cols <<= 1;
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
  CUtensorMapSwizzle swa = swizzle_for_bytes((int)p.pitch_a);
  auto enc_x = [&](CUtensorMap* tm, const void* x, int C, int64_t sB, int64_t sT, int64_t sF) -> int {
    if (d->sf != 1) return encode_act(enc, tm, x, C, d->sf, d->Fi, d->Ti, d->B, sB, sT, sF, p.gw_a, frows, p.t_tile, swa);
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)d->Fi, (cuuint64_t)d->Ti, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)sF * 2, (cuuint64_t)sT * 2, (cuuint64_t)sB * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.gw_a, (cuuint32_t)frows, (cuuint32_t)p.t_tile, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return (int)enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swa, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  int rc = enc_x(&tmX, d->x0, d->c0, d->x0_sB, d->x0_sT, d->x0_sF);
  if (!rc && d->c1) rc = enc_x(&tmX1, d->x1, d->c1, d->x1_sB, d->x1_sT, d->x1_sF);
  if (rc) { set_error("clskd_tapconv_wgrad_umma_stacked: cuTensorMapEncodeTiled(x) failed: %d", rc); return CLSKD_ERR_CUDA; }
  if (!d->c1) tmX1 = tmX;
  {
    const int brows = (int)(p.b_tx / p.pitch_b) / p.t_tile;
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
