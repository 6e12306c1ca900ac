// Tap-list implicit GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), bf16
// operands, fp32 accumulation in tensor memory.
//
//   Y[b,to,fo,n] = bias[n] + sum_j sum_c X[b, to+dt[j], fo*sf+df[j], c] * W[j][n][c]
//
// One CTA computes a 128 x BLOCK_N output tile: the 128 rows are a (t_tile x fo_tile) patch of one
// utterance, so that for every tap the A operand is ONE TMA box of the channels-last activation
// tensor viewed as the 5-D tensor (C, f-parity, F/sf, T, B); the tap only shifts the box
// coordinates, and frequency/time zero padding (and the ragged last time tile) come for free from
// TMA out-of-bounds zero fill.  The skip connection of the decoder is a second tensor map that
// supplies the upper K range (no materialised concat).  B is the packed block weight
// [tap][n][c] (K-major).  Both operands land in 128/64/32-byte swizzled K-major shared memory and
// are consumed by tcgen05.mma (M=128, N=BLOCK_N, K=16 per instruction) issued by one thread.
//
// Persistent: one CTA per SM loops over its tiles.  Warp roles (192 threads): warp 0 = TMA
// producer (runs ahead across tiles through the shared-memory ring), warp 1 = TMEM allocator + MMA
// issuer, warps 2-5 = epilogue (tcgen05.ld -> bias -> bf16/fp32 -> global).  The accumulator is
// double-buffered in TMEM (2 x BLOCK_N <= 512 columns), so the epilogue of tile i overlaps the main
// loop of tile i+1, and barrier/TMEM set-up is paid once per CTA instead of once per tile.
#include "umma.cuh"

namespace clskd {
namespace {
using namespace umma;

constexpr int UM = 128;       // UMMA M
constexpr int kThreads = 192;

struct UmmaParamsV1 {
  int B, To, Fo;
  int t_tile, fo_tile, f_tiles, t_tiles, tiles_n;
  int block_n, block_k;
  int chunks0, chunks_tot;  // K chunks of source 0 / total per tap
  int ntaps;
  int tap_t[CLSKD_MAX_TAPS];   // time offset
  int tap_p[CLSKD_MAX_TAPS];   // parity coordinate (df mod sf)
  int tap_f[CLSKD_MAX_TAPS];   // floor(df / sf)
  int stages;
  uint32_t a_bytes, b_bytes;   // per stage, padded to 1024
  uint32_t tx_bytes;           // bytes actually delivered per stage
  uint32_t sbo;                // stride byte offset >> 4
  uint32_t layout_type;        // UMMA smem layout type (2 = SW128, 4 = SW64, 6 = SW32)
  uint32_t tmem_cols;
  void* y;
  int64_t y_sB, y_sT, y_sF;
  int y_dtype;
  const float* bias;
  int N;
  int num_tiles;
  // epilogue staging for the TMA store: sub-tiles of gw_y columns, [128 rows][gw_y] each, swizzled
  int gw_y, es;                // columns per sub-tile, bytes per output element
  uint32_t y_sub_bytes;        // 128 * gw_y * es
  uint32_t stage_region;       // bytes of the operand ring (staging buffers follow it)
  uint32_t staging_bytes;      // one staging buffer: 128 rows x block_n x es
  int nstg;                    // 1 or 2 staging buffers (double-buffered TMA stores)
  int ecols;                   // columns staged per TMA-store round (<= 128): block_n / ecols rounds per tile
  // fused epilogue (see ClskdTapConv): folded eval BatchNorm, PReLU, batch statistics of the stored outputs
  const float* ep_scale;
  const float* ep_shift;
  const float* ep_slope;
  double* stats_sum;
  double* stats_sumsq;
};

// EP = false: plain contraction (+bias); EP = true: the fused epilogue variants (kept out of the plain
// instantiation so that it stays at its lean register count)
template <bool EP>
__global__ void __launch_bounds__(kThreads, 2)
tapconv_umma_v1_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmY,
                    const UmmaParamsV1 p) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  // 1024-byte aligned operand ring
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                             ~(uintptr_t)1023);
  const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 128);     // every epilogue thread arrives
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int num_k = p.ntaps * p.chunks_tot;

  // persistent: this CTA owns tiles blockIdx.x, blockIdx.x + gridDim.x, ...  (n tile fastest, so
  // consecutive CTAs share the activation patch in L2)
  if (warp == 0) {
    // ===================== TMA producer (all lanes run the loops, one elected lane issues) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.tiles_n;
        int r = tile / p.tiles_n;
        const int f_blk = r % p.f_tiles;
        r /= p.f_tiles;
        const int t_blk = r % p.t_tiles;
        const int b = r / p.t_tiles;
        const int t0 = t_blk * p.t_tile, f0 = f_blk * p.fo_tile, n0 = n_tile * p.block_n;
        for (int it = 0; it < num_k; ++it) {
          const int tap = it / p.chunks_tot;
          const int ch = it - tap * p.chunks_tot;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_expect_tx_warp(&full_bar[stage], p.tx_bytes);
          uint8_t* a_dst = ring + (size_t)stage * stage_bytes;
          uint8_t* b_dst = a_dst + p.a_bytes;
          const bool src0 = ch < p.chunks0;
          const int cc = (src0 ? ch : ch - p.chunks0) * p.block_k;
          tma_load_5d_warp(a_dst, src0 ? &tmA0 : &tmA1, &full_bar[stage], cc, p.tap_p[tap],
                      f0 + p.tap_f[tap], t0 + p.tap_t[tap], b);
          tma_load_3d_warp(b_dst, &tmB, &full_bar[stage], ch * p.block_k, n0, tap);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (all lanes run the loops, one elected lane issues) =====================
    {
      // instruction descriptor: D=f32, A=B=bf16, K-major both, N, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) |
                             ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      const int ksteps = p.block_k / 16;
      int local = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
        const int as = local & 1;
        mbar_wait(&tmem_empty_bar[as], ((local >> 1) & 1) ^ 1u);   // epilogue drained this accumulator
        fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t)(as * p.block_n);
        for (int it = 0; it < num_k; ++it) {
          mbar_wait(&full_bar[stage], phase);
          fence_after();
          const uint32_t a_addr = smem_u32(ring + (size_t)stage * stage_bytes);
          const uint32_t b_addr = a_addr + p.a_bytes;
          const uint64_t adesc = make_smem_desc(a_addr, p.sbo, p.layout_type);
          const uint64_t bdesc = make_smem_desc(b_addr, p.sbo, p.layout_type);
          for (int k = 0; k < ksteps; ++k) {
            // advance 32 bytes (16 bf16) along K inside the swizzle atom: +2 in 16-byte units
            umma_bf16_warp(d_addr, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                      (it | k) ? 1u : 0u);
          }
          umma_commit_warp(&empty_bar[stage]);  // frees the smem slot when the MMAs retire
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_warp(&tmem_full_bar[as]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    // TMEM -> registers (+bias, ->bf16/fp32) -> swizzled shared-memory staging -> TMA store: every
    // global write is a full coalesced box; rows beyond To are clipped by the tensor map.
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;        // row of the 128-row tile
    uint8_t* stg_base = ring + p.stage_region;
    const uint32_t pitch = (uint32_t)(p.gw_y * p.es);          // 128 / 64 / 32 bytes
    const uint32_t xr = pitch == 128 ? (uint32_t)(row & 7) : (pitch == 64 ? (uint32_t)((row >> 1) & 3) : (uint32_t)((row >> 2) & 1));
    const bool issuer = (threadIdx.x == 64);                    // first epilogue thread
    const float ep_slope = (EP && p.ep_slope) ? __ldg(p.ep_slope) : 1.f;
    // batch statistics: thread et owns four adjacent columns (one 8-byte word of a staged row) of every staging round and
    // the row slice [part*st_rows, (part+1)*st_rows) of the 128-row tile; partial sums stay in registers across all tiles
    // of this persistent CTA (tiles_n == 1, <= 2 rounds) and are flushed once at the end.  (One column per thread and a
    // 128-row serial loop of 2-byte loads cost 1.4 us per tile - 0.69 ms instead of 0.29 ms on the 1x1 16->128 conv.)
    const int et = threadIdx.x - 64;
    const int st_nvec = p.ecols >> 2;
    const int st_parts = UM / st_nvec;
    const int st_rows = (UM + st_parts - 1) / st_parts;
    const int st_v = et % st_nvec, st_part = et / st_nvec;
    const bool st_on = st_part < st_parts;
    float st_s0[4] = {0.f, 0.f, 0.f, 0.f}, st_q0[4] = {0.f, 0.f, 0.f, 0.f};
    float st_s1[4] = {0.f, 0.f, 0.f, 0.f}, st_q1[4] = {0.f, 0.f, 0.f, 0.f};
    int local = 0;
    int sround = 0;                                             // staging rounds issued so far
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
      const int as = local & 1;
      const int n_tile = tile % p.tiles_n;
      int r = tile / p.tiles_n;
      const int f_blk = r % p.f_tiles;
      r /= p.f_tiles;
      const int t_blk = r % p.t_tiles;
      const int b = r / p.t_tiles;
      const int n0 = n_tile * p.block_n;
      mbar_wait(&tmem_full_bar[as], (local >> 1) & 1);
      fence_after();
      const int t0 = t_blk * p.t_tile, f0 = f_blk * p.fo_tile;
      const int rounds = p.block_n / p.ecols;
      for (int rd = 0; rd < rounds; ++rd, ++sround) {
        // the TMA store that last used this staging buffer must have finished reading it
        uint8_t* stg = stg_base + (size_t)((p.nstg == 2) ? (sround & 1) : 0) * p.staging_bytes;
        if (issuer) {
          if (p.nstg == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int cbeg = rd * p.ecols;
        for (int c = cbeg; c < cbeg + p.ecols; c += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.block_n + c), v);
          float o[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = __uint_as_float(v[e]);
          if (p.bias) {
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] += __ldg(p.bias + n0 + c + e);
          }
          if (EP && p.ep_scale) {
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = fmaf(o[e], __ldg(p.ep_scale + n0 + c + e), __ldg(p.ep_shift + n0 + c + e));
          }
          if (EP && p.ep_slope) {
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = o[e] > 0.f ? o[e] : o[e] * ep_slope;
          }
          const int cl = c - cbeg;
          const int sub = cl / p.gw_y, col = cl - sub * p.gw_y;
          uint8_t* rowp = stg + (size_t)sub * p.y_sub_bytes + (size_t)row * pitch;
          const uint32_t ch0 = (uint32_t)(col * p.es) >> 4;          // first 16-byte chunk of these 16 columns
          if (p.es == 2) {
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
              pk[e] = *reinterpret_cast<uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(rowp + (((ch0 + 0) ^ xr) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(rowp + (((ch0 + 1) ^ xr) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              *reinterpret_cast<float4*>(rowp + (((ch0 + e) ^ xr) << 4)) =
                  make_float4(o[4 * e], o[4 * e + 1], o[4 * e + 2], o[4 * e + 3]);
          }
        }
        if (rd == rounds - 1) {
          // accumulator drained: hand the TMEM buffer back to the MMA warp
          fence_before();
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[as])) : "memory");
        }
        // make the generic-proxy smem writes visible to the async proxy, then one thread stores
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          for (int sidx = 0; sidx < p.ecols / p.gw_y; ++sidx) {
            asm volatile(
                "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                    reinterpret_cast<uint64_t>(&tmY)),
                "r"(smem_u32(stg + (size_t)sidx * p.y_sub_bytes)), "r"(n0 + cbeg + sidx * p.gw_y), "r"(f0), "r"(t0), "r"(b)
                : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (EP && p.stats_sum) {
          // column sums of the staged (bf16-rounded) tile over its valid rows, read back from the swizzled
          // staging buffer (consecutive threads read consecutive columns of one row: conflict free); the
          // buffer is not rewritten before every epilogue thread has passed the next round's barriers
          int nvalid = (p.To - t0) * p.fo_tile;
          if (nvalid > UM) nvalid = UM;
          const int c4 = st_v * 4;
          const int sub = c4 / p.gw_y, cl = c4 - sub * p.gw_y;
          // 32-bit shared-space addresses (LDS.64, four rows in flight); the swizzle term of row rr is
          // ((chunk ^ ((rr >> xs) & xm)) << 4) with (xs, xm) = (0, 7) / (1, 3) / (2, 1) for 128 / 64 / 32-byte rows
          const uint32_t colb = smem_u32(stg) + (uint32_t)sub * p.y_sub_bytes + (uint32_t)((cl * 2) & 15);
          const uint32_t chk = (uint32_t)(cl * 2) >> 4;
          const uint32_t xs = pitch == 128 ? 0u : (pitch == 64 ? 1u : 2u), xm = 7u >> xs;
          int rend = (st_part + 1) * st_rows;
          if (rend > nvalid) rend = nvalid;
          if (!st_on) rend = 0;
          float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
          for (int r4 = st_part * st_rows; r4 < rend; r4 += 4) {
            uint32_t wx[4], wy[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint32_t rr = (uint32_t)(r4 + u);
              wx[u] = 0u;
              wy[u] = 0u;
              if ((int)rr < rend)
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                             : "=r"(wx[u]), "=r"(wy[u])
                             : "r"(colb + rr * pitch + ((chk ^ ((rr >> xs) & xm)) << 4)));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float v0 = __uint_as_float(wx[u] << 16), v1 = __uint_as_float(wx[u] & 0xffff0000u);
              const float v2 = __uint_as_float(wy[u] << 16), v3 = __uint_as_float(wy[u] & 0xffff0000u);
              s[0] += v0; q[0] = fmaf(v0, v0, q[0]);
              s[1] += v1; q[1] = fmaf(v1, v1, q[1]);
              s[2] += v2; q[2] = fmaf(v2, v2, q[2]);
              s[3] += v3; q[3] = fmaf(v3, v3, q[3]);
            }
          }
          if (rd == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { st_s0[j] += s[j]; st_q0[j] += q[j]; }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) { st_s1[j] += s[j]; st_q1[j] += q[j]; }
          }
        }
      }
    }
    if (EP && p.stats_sum) {
      // one atomic per column and CTA: same-address fp64 atomics retire at about 20 ns each, so per-thread flushes
      // (parts x CTAs per address) were a 25-100 us tail.  The partial sums meet in the (now idle) staging buffer.
      if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float* scr = reinterpret_cast<float*>(stg_base);           // [round][s|q][part][ecols]
      const int nrd = p.block_n > p.ecols ? 2 : 1;
      const int pstride = st_parts * p.ecols;
      if (st_on) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int o = st_part * p.ecols + st_v * 4 + j;
          scr[o] = st_s0[j];
          scr[pstride + o] = st_q0[j];
          if (nrd == 2) {
            scr[2 * pstride + o] = st_s1[j];
            scr[3 * pstride + o] = st_q1[j];
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int c = et; c < nrd * p.ecols; c += 128) {
        const int rd = c / p.ecols, col = c - rd * p.ecols;
        if (rd * p.ecols + col < p.N) {
          float s = 0.f, q = 0.f;
          for (int pt = 0; pt < st_parts; ++pt) {
            s += scr[(2 * rd) * pstride + pt * p.ecols + col];
            q += scr[(2 * rd + 1) * pstride + pt * p.ecols + col];
          }
          atomicAdd(p.stats_sum + rd * p.ecols + col, (double)s);
          atomicAdd(p.stats_sumsq + rd * p.ecols + col, (double)q);
        }
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
  }

  fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// dense source check + geometry; returns nullptr if supported, else a reason
const char* umma_unsupported_v1(const ClskdTapConv* d) {
  if (d->x_dtype != CLSKD_BF16) return "x must be bf16";
  const int Ctot = d->c0 + d->c1;
  if (Ctot % 16 || d->c0 % 16) return "channels must be multiples of 16";
  if (d->N % 16) return "N must be a multiple of 16";
  if (d->N > 256 && d->N % 128) return "N > 256 must be a multiple of 128";
  if (d->y_dtype != CLSKD_BF16 && d->N > 128 && d->N % 128) return "fp32 output: N > 128 must be a multiple of 128";
  if (d->sf != 1 && d->sf != 2) return "sf must be 1 or 2";
  if (!is_pow2(d->Fo) || (d->Fo > 128 && d->Fo % 128)) return "Fo must be a power of two";
  if (d->accumulate) return "accumulate unsupported";
  if ((d->ep_scale == nullptr) != (d->ep_shift == nullptr)) return "ep_scale and ep_shift come together";
  if ((d->stats_sum == nullptr) != (d->stats_sumsq == nullptr)) return "stats_sum and stats_sumsq come together";
  if (d->stats_sum) {
    if (d->y_dtype != CLSKD_BF16) return "fused statistics need a bf16 output";
    if (d->N != 16 && d->N != 32 && d->N != 64 && d->N != 128 && d->N != 256)
      return "fused statistics need N in {16,32,64,128,256}";
  }
  if (d->Fi % d->sf) return "Fi must be a multiple of sf";
  auto chk = [&](const void* x, int64_t sB, int64_t sT, int64_t sF) -> const char* {
    if ((uintptr_t)x % 16) return "x not 16-byte aligned";
    if ((sB * 2) % 16 || (sT * 2) % 16 || (sF * 2) % 16) return "x strides not 16-byte multiples";
    return nullptr;
  };
  if (const char* r = chk(d->x0, d->x0_sB, d->x0_sT, d->x0_sF)) return r;
  if (d->c1)
    if (const char* r = chk(d->x1, d->x1_sB, d->x1_sT, d->x1_sF)) return r;
  if ((uintptr_t)d->w % 16 || (uintptr_t)d->y % 16) return "w/y not 16-byte aligned";
  int ye = d->y_dtype == CLSKD_BF16 ? 2 : 4;
  if ((d->y_sB * ye) % 16 || (d->y_sT * ye) % 16 || (d->y_sF * ye) % 16)
    return "y strides not 16-byte multiples";
  if (!get_encode()) return "cuTensorMapEncodeTiled unavailable";
  return nullptr;
}

}  // namespace
}  // namespace clskd

using namespace clskd;


// round-1 kernel (one TMA box per tap, weights through the ring), kept as the A/B baseline of tools/kbench.py
extern "C" int clskd_tapconv_fwd_umma_v1(const ClskdTapConv* d, void* stream) {
  CLSKD_CHECK_ARG(d && d->x0 && d->w && d->y, "clskd_tapconv_fwd_umma_v1: null pointer");
  CLSKD_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= CLSKD_MAX_TAPS, "clskd_tapconv_fwd_umma_v1: ntaps");
  if (const char* why = umma_unsupported_v1(d)) {
    set_error("clskd_tapconv_fwd_umma_v1: unsupported: %s", why);
    return CLSKD_ERR_UNSUPPORTED;
  }
  const int64_t M = (int64_t)d->B * d->To * d->Fo;
  if (M == 0) return CLSKD_OK;
  EncodeTiledFn enc = get_encode();
  const int Ctot = d->c0 + d->c1;

  UmmaParamsV1 p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.To = d->To; p.Fo = d->Fo;
  p.fo_tile = d->Fo < UM ? d->Fo : UM;
  p.t_tile = UM / p.fo_tile;
  p.f_tiles = d->Fo / p.fo_tile;
  p.t_tiles = cdiv(d->To, p.t_tile);
  p.block_n = d->N <= 256 ? d->N : (d->N % 256 == 0 ? 256 : 128);
  p.es = d->y_dtype == CLSKD_BF16 ? 2 : 4;
  if (p.es == 4 && p.block_n > 128 && d->N % 128 == 0) p.block_n = 128;   // keep the fp32 staging tile <= 64 KB
  p.tiles_n = d->N / p.block_n;
  // largest K chunk that divides both sources
  int bk = 64;
  while (bk > 16 && (d->c0 % bk || (d->c1 % bk))) bk >>= 1;
  p.block_k = bk;
  p.chunks0 = d->c0 / bk;
  p.chunks_tot = Ctot / bk;
  p.ntaps = d->ntaps;
  for (int j = 0; j < d->ntaps; ++j) {
    int df = d->df[j];
    int fl = df >= 0 ? df / d->sf : -((-df + d->sf - 1) / d->sf);  // floor division
    p.tap_t[j] = d->dt[j];
    p.tap_f[j] = fl;
    p.tap_p[j] = df - fl * d->sf;
  }
  auto pad1k = [](uint32_t v) { return (v + 1023u) & ~1023u; };
  p.a_bytes = pad1k((uint32_t)UM * bk * 2);
  p.b_bytes = pad1k((uint32_t)p.block_n * bk * 2);
  p.tx_bytes = (uint32_t)UM * bk * 2 + (uint32_t)p.block_n * bk * 2;
  p.sbo = (uint32_t)(8 * bk * 2) >> 4;
  CUtensorMapSwizzle sw;
  if (bk == 64) { p.layout_type = 2; sw = CU_TENSOR_MAP_SWIZZLE_128B; }
  else if (bk == 32) { p.layout_type = 4; sw = CU_TENSOR_MAP_SWIZZLE_64B; }
  else { p.layout_type = 6; sw = CU_TENSOR_MAP_SWIZZLE_32B; }
  {
    const int max_gw = 128 / p.es;             // 64 bf16 or 32 fp32 columns per 128-byte swizzle row
    int gw = max_gw;
    while (gw > 16 && p.block_n % gw) gw >>= 1;   // >= 16 columns: one tcgen05.ld chunk never straddles sub-tiles
    p.gw_y = gw;
  }
  p.y_sub_bytes = (uint32_t)UM * p.gw_y * p.es;
  int cols = 32;
  while (cols < 2 * p.block_n) cols <<= 1;     // double-buffered accumulator
  p.tmem_cols = (uint32_t)cols;
  const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
  p.ecols = p.block_n > 128 ? 128 : p.block_n;     // block_n is 256 or <= 128 here... (256 = 2 rounds)
  if (p.block_n % p.ecols) p.ecols = p.block_n;
  const uint32_t staging_bytes = (uint32_t)UM * p.ecols * p.es;
  p.staging_bytes = staging_bytes;
  // two CTAs per SM when the accumulators (2 x 2 x block_n TMEM columns) and ~110 KB of smem each
  // allow it: their serial per-tile latencies (TMA -> MMA -> TMEM drain -> store) overlap
  const bool two_ctas = 2 * cols <= 512 && staging_bytes <= 32 * 1024;
  const uint32_t budget = two_ctas ? 108u * 1024u : 222u * 1024u;
  p.nstg = (2 * staging_bytes + 2 * stage_bytes <= budget) ? 2 : 1;
  int stages = (int)((budget - p.nstg * staging_bytes) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  const int num_k = p.ntaps * p.chunks_tot;
  // (the ring runs ahead across tiles of the persistent loop, so it is not limited by num_k)
  p.stages = stages;
  p.y = d->y; p.y_sB = d->y_sB; p.y_sT = d->y_sT; p.y_sF = d->y_sF; p.y_dtype = d->y_dtype;
  p.bias = d->bias; p.N = d->N;
  p.ep_scale = d->ep_scale; p.ep_shift = d->ep_shift; p.ep_slope = d->ep_slope;
  p.stats_sum = d->stats_sum; p.stats_sumsq = d->stats_sumsq;

  CUtensorMap tmA0, tmA1, tmB;
  int rc = encode_act(enc, &tmA0, d->x0, d->c0, d->sf, d->Fi, d->Ti, d->B, d->x0_sB, d->x0_sT,
                      d->x0_sF, bk, p.fo_tile, p.t_tile, sw);
  if (rc) { set_error("clskd_tapconv_fwd_umma_v1: cuTensorMapEncodeTiled(x0) failed: %d", rc); return CLSKD_ERR_CUDA; }
  if (d->c1) {
    rc = encode_act(enc, &tmA1, d->x1, d->c1, d->sf, d->Fi, d->Ti, d->B, d->x1_sB, d->x1_sT,
                    d->x1_sF, bk, p.fo_tile, p.t_tile, sw);
    if (rc) { set_error("clskd_tapconv_fwd_umma_v1: cuTensorMapEncodeTiled(x1) failed: %d", rc); return CLSKD_ERR_CUDA; }
  } else {
    tmA1 = tmA0;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)Ctot, (cuuint64_t)d->N, (cuuint64_t)d->ntaps};
    cuuint64_t strides[2] = {(cuuint64_t)Ctot * 2, (cuuint64_t)Ctot * 2 * d->N};
    cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)p.block_n, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d->w), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { set_error("clskd_tapconv_fwd_umma_v1: cuTensorMapEncodeTiled(w) failed: %d", (int)r); return CLSKD_ERR_CUDA; }
  }
  const int64_t tiles = (int64_t)d->B * p.t_tiles * p.f_tiles * p.tiles_n;
  CLSKD_CHECK_ARG(tiles <= 2147483647LL, "clskd_tapconv_fwd_umma_v1: too many tiles");
  p.num_tiles = (int)tiles;
  p.stage_region = (uint32_t)stages * stage_bytes;
  CUtensorMap tmY;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->N, (cuuint64_t)d->Fo, (cuuint64_t)d->To, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->y_sF * p.es, (cuuint64_t)d->y_sT * p.es, (cuuint64_t)d->y_sB * p.es};
    cuuint32_t box[4] = {(cuuint32_t)p.gw_y, (cuuint32_t)p.fo_tile, (cuuint32_t)p.t_tile, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmY, p.es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d->y,
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(p.gw_y * p.es),
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { set_error("clskd_tapconv_fwd_umma_v1: cuTensorMapEncodeTiled(y) failed: %d", (int)r); return CLSKD_ERR_CUDA; }
  }
  size_t smem = (size_t)stages * stage_bytes + (size_t)p.nstg * staging_bytes + 1024;
  const bool ep = d->ep_scale || d->ep_slope || d->stats_sum;
  static size_t smem_set[2] = {0, 0};
  if (smem > smem_set[ep ? 1 : 0]) {
    cudaError_t e = ep ? cudaFuncSetAttribute(tapconv_umma_v1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                       : cudaFuncSetAttribute(tapconv_umma_v1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("clskd_tapconv_fwd_umma_v1: smem attr: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
    smem_set[ep ? 1 : 0] = smem;
  }
  const int64_t max_ctas = (int64_t)sm_count() * (two_ctas ? 2 : 1);
  const unsigned grid = (unsigned)(tiles < max_ctas ? tiles : max_ctas);
  if (ep) tapconv_umma_v1_kernel<true><<<grid, kThreads, smem, (cudaStream_t)stream>>>(tmA0, tmA1, tmB, tmY, p);
  else tapconv_umma_v1_kernel<false><<<grid, kThreads, smem, (cudaStream_t)stream>>>(tmA0, tmA1, tmB, tmY, p);
  CLSKD_CHECK_LAUNCH("clskd_tapconv_fwd_umma_v1");
  return CLSKD_OK;
}
