// Tap-list implicit GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), bf16
// operands, fp32 accumulation in tensor memory.
//
//   Y[b,to,fo,n] = bias[n] + sum_j sum_c X[b, to+dt[j], fo*sf+df[j], c] * W[j][n][c]
//
// One CTA computes a 128 x BLOCK_N output tile: the 128 rows are a (t_tile x fo_tile) patch of one
// utterance, so that the A operand of a tap is a box of the channels-last activation tensor viewed as the
// 5-D tensor (C, f-parity, F/sf, T, B); frequency/time zero padding (and the ragged last time tile) come
// for free from TMA out-of-bounds zero fill.  The skip connection of the decoder is a second tensor map
// that supplies the upper K range (no materialised concat).  B is the packed block weight
// [tap][n][c] (K-major).  Both operands land in 128/64/32-byte swizzled K-major shared memory and are
// consumed by tcgen05.mma (M=128, N=BLOCK_N, K=16 per instruction) issued by one thread.
//
// Operand reuse (the kernel is bound by the L2 -> shared-memory fill rate, not by the tensor pipe):
//   * HALO PATCHES.  Taps are grouped; a group loads ONE activation patch per K chunk and every tap of the
//     group is issued from a shifted shared-memory descriptor instead of its own TMA box.  tcgen05 applies
//     the swizzle XOR to the absolute shared-memory address, so a descriptor start that is shifted by whole
//     rows (not a multiple of the 1024-byte pattern) reads exactly the shifted rows of a canonically
//     swizzled patch with matrix-base-offset 0 (measured: profiles/r02_hw_desc_shift.log).
//       - full mode (fo_tile = 128, stride 1): one (t_tile+tspan) x (128+fspan) patch serves all taps
//         (3x3: 390 rows instead of 9 x 128);
//       - time mode (any tile): taps that differ only in dt share a (t_tile+tspan) x fo_tile patch; the
//         shift is a multiple of fo_tile rows;
//       - legacy mode: one box per tap.
//   * RESIDENT WEIGHTS.  When the whole packed weight of the launch fits next to the pipeline it is loaded
//     once per persistent CTA and the ring carries activation patches only.
//
// Persistent: one CTA per SM (two when TMEM / shared memory allow) loops over its tiles.  Warp roles
// (192 threads): warp 0 = TMA producer (runs ahead across tiles through the shared-memory ring), warp 1 =
// TMEM allocator + MMA issuer, warps 2-5 = epilogue (tcgen05.ld -> bias / folded BatchNorm / PReLU ->
// bf16/fp32 -> swizzled staging -> TMA store, batch statistics from the staged tile).  The accumulator is
// double-buffered in TMEM (2 x BLOCK_N <= 512 columns), so the epilogue of tile i overlaps the main loop of
// tile i+1.
#include <mutex>
#include <unordered_map>
#include <string>
#include <stdlib.h>

#include "umma.cuh"

namespace clskd {
extern int g_wgrad_mode;               // tapconv_wgrad_umma.cu
extern int g_lstm_legacy;              // lstm.cu
namespace {
using namespace umma;

constexpr int UM = 128;       // UMMA M
constexpr int kThreads = 192;
constexpr int kMaxEpN = 256;  // epilogue constants staged in shared memory up to this N

// tuning overrides (clskd_set_tuning): 0 = automatic
int g_tune_mode = 0;       // 1 legacy (one box per tap), 2 time-grouped patches, 3 full halo patch where possible
int g_tune_resident = 0;   // 1 never keep the weights resident, 2 always when they fit
int g_tune_two_cta = 0;    // 1 force one CTA per SM, 2 force two when possible
int g_tune_v1 = 0;         // 1 route to the round-1 kernel
int g_tune_split = 0;      // 1 one TMA box per patch, 2 one box per time row of the patch (more loads in flight)
int g_tune_noauto = 0;     // 1 disable the per-shape autotuner (first-call timing of the candidate configurations)

// one launch configuration of the forward kernel (0 = automatic for every field)
struct FwdCfg {
  int v1;        // round-1 kernel
  int mode;      // 1 one box per tap, 2 time-grouped patches, 3 full halo patch where possible
  int resident;  // 1 never keep the weights resident
  int two_cta;   // 1 one CTA per SM
  int split;     // 1 one TMA box per patch, 2 one per time row
  int bk_cap;    // cap of the K chunk (32: smaller stages, two CTAs per SM fit more often)
};

struct UmmaParams {
  int B, To, Fo;
  int t_tile, fo_tile, f_tiles, t_tiles, tiles_n;
  int block_n, block_k;
  int chunks0, chunks_tot;  // K chunks of source 0 / total per tap
  int ntaps;
  // tap groups: taps [grp_beg[g], grp_beg[g+1]) (group order) share the patch whose box origin is shifted by
  // (grp_p, grp_f, grp_t) from the tile origin; tap_aoff = byte offset of the tap's 128-row A tile in the patch
  int ngroups, maxg;
  int grp_beg[CLSKD_MAX_TAPS + 1];
  int grp_p[CLSKD_MAX_TAPS], grp_f[CLSKD_MAX_TAPS], grp_t[CLSKD_MAX_TAPS];
  uint32_t tap_aoff[CLSKD_MAX_TAPS];
  int tap_w[CLSKD_MAX_TAPS];   // tap coordinate in the weight tensor map
  int stages;
  uint32_t a_bytes, b_bytes;   // patch / weight tile, padded to 1024
  uint32_t a_tx, b_tx;         // bytes actually delivered
  int a_nbox;                  // the patch is loaded as a_nbox boxes of a_box_t time rows each (a_box_bytes apart)
  int a_box_t;
  uint32_t a_box_bytes;
  uint32_t stage_bytes;        // a_bytes + maxg * b_bytes (resident: a_bytes)
  int resident;                // whole weight resident in shared memory
  uint32_t bres_off;           // its offset from the ring base
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
  uint32_t stg_off;            // offset of the staging buffers from the ring base
  uint32_t staging_bytes;      // one staging buffer: 128 rows x ecols x es
  int nstg;                    // 1 or 2 staging buffers (double-buffered TMA stores)
  int accum;                   // 1: TMA reduce-add instead of store (y += tile)
  int ecols;                   // columns staged per TMA-store round (<= 128): block_n / ecols rounds per tile
  // fused epilogue (see ClskdTapConv): folded eval BatchNorm, PReLU, batch statistics of the stored outputs
  const float* ep_scale;
  const float* ep_shift;
  const float* ep_slope;
  double* stats_sum;
  double* stats_sumsq;
};

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
      "%12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// EP = false: plain contraction (+bias); EP = true: the fused epilogue variants (kept out of the plain
// instantiation so that it stays at its lean register count)
template <bool EP>
__global__ void __launch_bounds__(kThreads, 2)
tapconv_umma_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmY,
                    const UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ __align__(8) uint64_t bres_bar;
  __shared__ uint32_t tmem_base_smem;

  // 1024-byte aligned operand ring
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                             ~(uintptr_t)1023);
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
    mbar_init(&bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  // persistent: this CTA owns tiles blockIdx.x, blockIdx.x + gridDim.x, ...  (n tile fastest, so
  // consecutive CTAs share the activation patch in L2)
  if (warp == 0) {
    // ===================== TMA producer (all lanes run the loops, one elected lane issues) =====================
    {
      if (p.resident) {
        mbar_expect_tx_warp(&bres_bar, (uint32_t)(p.ntaps * p.chunks_tot) * p.b_tx);
        uint8_t* dst = ring + p.bres_off;
        for (int j = 0; j < p.ntaps; ++j)
          for (int ch = 0; ch < p.chunks_tot; ++ch)
            tma_load_3d_warp(dst + (size_t)(j * p.chunks_tot + ch) * p.b_bytes, &tmB, &bres_bar, ch * p.block_k, 0, p.tap_w[j]);
      }
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
        for (int ch = 0; ch < p.chunks_tot; ++ch) {
          const bool src0 = ch < p.chunks0;
          const int cc = (src0 ? ch : ch - p.chunks0) * p.block_k;
          for (int g = 0; g < p.ngroups; ++g) {
            const int jb = p.grp_beg[g], je = p.grp_beg[g + 1];
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            mbar_expect_tx_warp(&full_bar[stage], p.a_tx + (p.resident ? 0u : (uint32_t)(je - jb) * p.b_tx));
            uint8_t* a_dst = ring + (size_t)stage * p.stage_bytes;
            for (int bx = 0; bx < p.a_nbox; ++bx)
              tma_load_5d_warp(a_dst + (size_t)bx * p.a_box_bytes, src0 ? &tmA0 : &tmA1, &full_bar[stage], cc, p.grp_p[g],
                          f0 + p.grp_f[g], t0 + p.grp_t[g] + bx * p.a_box_t, b);
            if (!p.resident) {
              uint8_t* b_dst = a_dst + p.a_bytes;
              for (int j = jb; j < je; ++j)
                tma_load_3d_warp(b_dst + (size_t)(j - jb) * p.b_bytes, &tmB, &full_bar[stage], ch * p.block_k, n0, p.tap_w[j]);
            }
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
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
      const uint32_t ring_addr = smem_u32(ring);
      const uint64_t desc_fixed = ((uint64_t)((p.sbo & 0x3FFFu) | (1u << 14) | (p.layout_type << 29)) << 32) | (1u << 16);
      if (p.resident) {
        mbar_wait(&bres_bar, 0);
        fence_after();
      }
      int local = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
        const int as = local & 1;
        mbar_wait(&tmem_empty_bar[as], ((local >> 1) & 1) ^ 1u);   // epilogue drained this accumulator
        fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t)(as * p.block_n);
        uint32_t first = 0;
        for (int ch = 0; ch < p.chunks_tot; ++ch) {
          for (int g = 0; g < p.ngroups; ++g) {
            mbar_wait(&full_bar[stage], phase);
            fence_after();
            const uint32_t a_addr = ring_addr + (uint32_t)stage * p.stage_bytes;
            // descriptors: hi word (SBO, version, layout) | lo word (start address >> 4, LBO = 1 (unused)); all offsets
            // are multiples of 16 bytes below 2^18, so the 14-bit address field adds up exactly.  The next tap's patch
            // offset is fetched BEFORE this tap's instructions are issued (the asm volatile MMAs are ordering points the
            // compiler does not move loads across: a parameter-table round trip per tap otherwise sits in front of
            // every tap's first instruction, and this warp's instruction stream paces narrow / short-K launches).
            const int jb = p.grp_beg[g], je = p.grp_beg[g + 1];
            const uint32_t a16 = a_addr >> 4;
            uint32_t b16 = p.resident ? (ring_addr + p.bres_off + (uint32_t)(jb * p.chunks_tot + ch) * p.b_bytes) >> 4
                                      : (a_addr + p.a_bytes) >> 4;
            const uint32_t bstep16 = (p.resident ? (uint32_t)p.chunks_tot * p.b_bytes : p.b_bytes) >> 4;
            uint32_t aoff_next = p.tap_aoff[jb] >> 4;
            for (int j = jb; j < je; ++j) {
              const uint32_t aoff = aoff_next;
              if (j + 1 < je) aoff_next = p.tap_aoff[j + 1] >> 4;
              // the tap's A tile: the patch rows shifted by tap_aoff (matrix base offset stays 0: the swizzle
              // XOR is a function of the absolute shared-memory address)
              const uint64_t adesc = desc_fixed | (uint64_t)(a16 + aoff);
              const uint64_t bdesc = desc_fixed | (uint64_t)b16;
              b16 += bstep16;
              for (int k = 0; k < ksteps; ++k) {
                // advance 32 bytes (16 bf16) along K inside the swizzle atom: +2 in 16-byte units
                umma_bf16_warp(d_addr, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, first);
                first = 1u;
              }
            }
            umma_commit_warp(&empty_bar[stage]);  // frees the smem slot when the MMAs retire
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
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
    uint8_t* stg_base = ring + p.stg_off;
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
        // two 16-column TMEM loads in flight per wait (the epilogue is a serial chain per tile: this halves the
        // exposed TMEM round trips)
        const bool pair = (p.ecols & 31) == 0;
        for (int c2 = cbeg; c2 < cbeg + p.ecols; c2 += pair ? 32 : 16) {
          uint32_t va[16], vb[16];
          const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.block_n + c2);
          tmem_ld16_async(tcol, va);
          if (pair) tmem_ld16_async(tcol + 16u, vb);
          tmem_ld_wait();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
          if (hh && !pair) break;
          const int c = c2 + 16 * hh;
          float o[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = __uint_as_float(hh ? vb[e] : va[e]);
          // columns >= N exist only as padding of an N that is not a multiple of 16 (zero weights; clipped by the
          // TMA store): their per-channel constants are not read
          const int nlive = p.N - (n0 + c);
          if (p.bias) {
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] += e < nlive ? __ldg(p.bias + n0 + c + e) : 0.f;
          }
          if (EP && p.ep_scale) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (e < nlive) o[e] = fmaf(o[e], __ldg(p.ep_scale + n0 + c + e), __ldg(p.ep_shift + n0 + c + e));
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
            if (p.accum)
              asm volatile(
                  "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                      reinterpret_cast<uint64_t>(&tmY)),
                  "r"(smem_u32(stg + (size_t)sidx * p.y_sub_bytes)), "r"(n0 + cbeg + sidx * p.gw_y), "r"(f0), "r"(t0), "r"(b)
                  : "memory");
            else
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
const char* umma_unsupported(const ClskdTapConv* d) {
  if (d->x_dtype != CLSKD_BF16) return "x must be bf16";
  // channel counts that are multiples of 8 but not of 16 (the reference's quarter-width student: 8-channel maps) run
  // padded: the TMA box is 16 channels wide over the 8 that exist (zero fill / clipped store) and the packed weight
  // carries zero rows / columns - see clskd_tapconv_umma_padded
  if (d->c0 % 8 || d->c1 % 8 || d->c0 < 8) return "channels must be multiples of 8";
  if (d->N % 8 || d->N < 8) return "N must be a multiple of 8";
  if (d->N > 256 && d->N % 128) return "N > 256 must be a multiple of 128";
  if (d->y_dtype != CLSKD_BF16 && d->N > 128 && d->N % 128) return "fp32 output: N > 128 must be a multiple of 128";
  if (d->sf != 1 && d->sf != 2) return "sf must be 1 or 2";
  if (!is_pow2(d->Fo) || (d->Fo > 128 && d->Fo % 128)) return "Fo must be a power of two";
  // y += result: the epilogue's TMA stores become TMA reduce-adds (performed in L2 in the output type); no fused
  // epilogue on such a launch
  if (d->accumulate && (d->stats_sum || d->ep_scale || d->ep_slope)) return "accumulate with a fused epilogue unsupported";
  if ((d->ep_scale == nullptr) != (d->ep_shift == nullptr)) return "ep_scale and ep_shift come together";
  if ((d->stats_sum == nullptr) != (d->stats_sumsq == nullptr)) return "stats_sum and stats_sumsq come together";
  if (d->stats_sum) {
    if (d->y_dtype != CLSKD_BF16) return "fused statistics need a bf16 output";
    if (d->N != 8 && d->N != 16 && d->N != 32 && d->N != 64 && d->N != 128 && d->N != 256)
      return "fused statistics need N in {8,16,32,64,128,256}";
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

// channels-last activation as the 5-D tensor (C, f-parity, Fi/sf, Ti, B) with an arbitrary (fp x tp) patch box
int encode_act_patch(EncodeTiledFn enc, CUtensorMap* tm, const void* x, int C, int sf, int Fi, int Ti, int B, int64_t sB,
                     int64_t sT, int64_t sF, int block_k, int fp, int tp, CUtensorMapSwizzle sw) {
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)sf, (cuuint64_t)(Fi / sf), (cuuint64_t)Ti, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)sF * 2, (cuuint64_t)sF * 2 * sf, (cuuint64_t)sT * 2, (cuuint64_t)sB * 2};
  cuuint32_t box[5] = {(cuuint32_t)block_k, 1, (cuuint32_t)fp, (cuuint32_t)tp, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return (int)enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace
}  // namespace clskd

using namespace clskd;

extern "C" int clskd_has_tcgen05(void) { return 1; }

extern "C" int clskd_set_tuning(int key, int value) {
  switch (key) {
    case 0: g_tune_mode = value; return CLSKD_OK;
    case 1: g_tune_resident = value; return CLSKD_OK;
    case 2: g_tune_two_cta = value; return CLSKD_OK;
    case 3: g_tune_v1 = value; return CLSKD_OK;
    case 4: g_tune_split = value; return CLSKD_OK;
    case 5: g_tune_noauto = value; return CLSKD_OK;
    case 6: g_wgrad_mode = value; return CLSKD_OK;
    case 7: g_lstm_legacy = value; return CLSKD_OK;
    default: set_error("clskd_set_tuning: unknown key %d", key); return CLSKD_ERR_ARG;
  }
}

extern "C" int clskd_tapconv_umma_supported(const ClskdTapConv* d) {
  if (!d) return 0;
  return umma_unsupported(d) == nullptr ? 1 : 0;
}

static inline int pad16(int v) { return (v + 15) & ~15; }

// padded channel extent of source 1: a multiple of 16, and - when source 1 is narrower than the K chunk that source 0
// alone would run with - that chunk (one zero-filled TMA box over the channels that exist), so that a narrow second
// source does not drag the whole contraction down to 16- or 32-channel chunks (ABF conv1 data gradient with the folded
// BatchNorm backward: 128 + 16 channels)
static inline int umma_c1p(int c0, int c1) {
  const int c0p = pad16(c0);
  int c1p = pad16(c1), bk0 = 64;
  while (bk0 > 16 && c0p % bk0) bk0 >>= 1;
  if (c1 && c1p < bk0) c1p = bk0;
  return c1p;
}

extern "C" int clskd_tapconv_umma_c1p(int c0, int c1) { return umma_c1p(c0, c1); }

static int launch_cfg(const ClskdTapConv* d, const FwdCfg& cfg, void* stream) {
  const bool padded = (d->c0 % 16) || (d->N % 16) || umma_c1p(d->c0, d->c1) != d->c1;
  if (cfg.v1 && !d->accumulate) {        // (the round-1 kernel has no reduce-add epilogue)
    if (padded) { set_error("clskd_tapconv_fwd_umma: the round-1 kernel needs multiples of 16 channels"); return CLSKD_ERR_UNSUPPORTED; }
    return clskd_tapconv_fwd_umma_v1(d, stream);
  }
  EncodeTiledFn enc = get_encode();
  // padded extents: what the kernel contracts / produces (the packed weight is [ntaps][Np][c0p + c1p])
  const int c0p = pad16(d->c0), c1p = umma_c1p(d->c0, d->c1), Np = pad16(d->N);
  const int Ctot = c0p + c1p;

  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.To = d->To; p.Fo = d->Fo;
  p.fo_tile = d->Fo < UM ? d->Fo : UM;
  p.t_tile = UM / p.fo_tile;
  p.f_tiles = d->Fo / p.fo_tile;
  p.t_tiles = cdiv(d->To, p.t_tile);
  p.block_n = Np <= 256 ? Np : (Np % 256 == 0 ? 256 : 128);
  p.es = d->y_dtype == CLSKD_BF16 ? 2 : 4;
  if (p.es == 4 && p.block_n > 128 && Np % 128 == 0) p.block_n = 128;   // keep the fp32 staging tile <= 64 KB
  p.tiles_n = Np / p.block_n;
  // largest K chunk that divides both sources
  int bk = 64;
  while (bk > 16 && (c0p % bk || (c1p % bk))) bk >>= 1;
  if (cfg.bk_cap >= 16 && bk > cfg.bk_cap) bk = cfg.bk_cap;
  p.block_k = bk;
  p.chunks0 = c0p / bk;
  p.chunks_tot = Ctot / bk;
  p.ntaps = d->ntaps;
  const uint32_t pitch_a = (uint32_t)bk * 2;                 // bytes per patch row (= swizzle span)

  // ---- tap geometry: (time offset, parity, floor(df / sf))
  int tt[CLSKD_MAX_TAPS], tp_[CLSKD_MAX_TAPS], tf[CLSKD_MAX_TAPS];
  int tmin = 1 << 30, tmax = -(1 << 30), fmin = 1 << 30, fmax = -(1 << 30);
  for (int j = 0; j < d->ntaps; ++j) {
    int df = d->df[j];
    int fl = df >= 0 ? df / d->sf : -((-df + d->sf - 1) / d->sf);  // floor division
    tt[j] = d->dt[j];
    tf[j] = fl;
    tp_[j] = df - fl * d->sf;
    tmin = tt[j] < tmin ? tt[j] : tmin; tmax = tt[j] > tmax ? tt[j] : tmax;
    fmin = fl < fmin ? fl : fmin; fmax = fl > fmax ? fl : fmax;
  }
  // ---- grouping mode
  int mode = 1;
  const bool full_ok = d->sf == 1 && p.fo_tile == UM && d->ntaps > 1 && (UM + (fmax - fmin)) <= 256 &&
                       (p.t_tile + (tmax - tmin)) <= 16;
  const bool time_ok = d->ntaps > 1 && (p.fo_tile % 8 == 0) && (p.t_tile + (tmax - tmin)) * p.fo_tile <= 1024;
  if (full_ok) mode = 3;
  else if (time_ok && p.t_tile > 1 && tmax > tmin) mode = 2;
  if (cfg.mode == 1) mode = 1;
  if (cfg.mode == 2) mode = time_ok ? 2 : 1;
  if (cfg.mode == 3) mode = full_ok ? 3 : (time_ok && p.t_tile > 1 && tmax > tmin ? 2 : 1);
  int box_f = p.fo_tile, box_t = p.t_tile;
  auto build_groups = [&](int md) {
    box_f = p.fo_tile;
    box_t = p.t_tile;
    if (md == 3) {
      box_f = (UM + (fmax - fmin) + 7) & ~7;     // whole 8-row swizzle atoms per time row: every per-row TMA box
      box_t = p.t_tile + (tmax - tmin);          // of the patch then starts on a pattern boundary
      p.ngroups = 1;
      p.grp_beg[0] = 0; p.grp_beg[1] = d->ntaps;
      p.grp_p[0] = 0; p.grp_f[0] = fmin; p.grp_t[0] = tmin;
      for (int j = 0; j < d->ntaps; ++j) {
        p.tap_w[j] = j;
        p.tap_aoff[j] = (uint32_t)((tt[j] - tmin) * box_f + (tf[j] - fmin)) * pitch_a;
      }
    } else if (md == 2) {
      box_t = p.t_tile + (tmax - tmin);
      bool used[CLSKD_MAX_TAPS] = {false};
      int n = 0;
      p.ngroups = 0;
      for (int j = 0; j < d->ntaps; ++j) {
        if (used[j]) continue;
        const int g = p.ngroups++;
        p.grp_beg[g] = n;
        p.grp_p[g] = tp_[j]; p.grp_f[g] = tf[j]; p.grp_t[g] = tmin;
        for (int i = j; i < d->ntaps; ++i)
          if (!used[i] && tp_[i] == tp_[j] && tf[i] == tf[j]) {
            used[i] = true;
            p.tap_w[n] = i;
            p.tap_aoff[n] = (uint32_t)((tt[i] - tmin) * box_f) * pitch_a;
            ++n;
          }
      }
      p.grp_beg[p.ngroups] = n;
    } else {
      p.ngroups = d->ntaps;
      for (int j = 0; j < d->ntaps; ++j) {
        p.grp_beg[j] = j;
        p.grp_p[j] = tp_[j]; p.grp_f[j] = tf[j]; p.grp_t[j] = tt[j];
        p.tap_w[j] = j;
        p.tap_aoff[j] = 0;
      }
      p.grp_beg[d->ntaps] = d->ntaps;
    }
    p.maxg = 0;
    for (int g = 0; g < p.ngroups; ++g) {
      const int n = p.grp_beg[g + 1] - p.grp_beg[g];
      p.maxg = n > p.maxg ? n : p.maxg;
    }
  };
  auto pad1k = [](uint32_t v) { return (v + 1023u) & ~1023u; };
  p.b_tx = (uint32_t)p.block_n * pitch_a;
  p.b_bytes = pad1k(p.b_tx);
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
  p.ecols = p.block_n > 128 ? 128 : p.block_n;     // block_n is 256 or <= 128 here (256 = 2 rounds)
  if (p.block_n % p.ecols) p.ecols = p.block_n;
  const uint32_t staging_bytes = (uint32_t)UM * p.ecols * p.es;
  p.staging_bytes = staging_bytes;
  const uint32_t w_bytes = (uint32_t)(d->ntaps * p.chunks_tot) * p.b_bytes;
  const uint32_t kSmemMax = 222u * 1024u;     // 227 KB per CTA minus the static shared memory (epilogue constants, barriers)
  bool resident = false, two_ctas = false;
  uint32_t fixed = 0;
  int stages = 0;
  for (;;) {
    build_groups(mode);
    p.a_tx = (uint32_t)box_f * box_t * pitch_a;
    p.a_bytes = pad1k(p.a_tx);
    // several TMA boxes per patch: the TMA unit keeps more row requests in flight across boxes than inside one
    int split = 1;       // measured: one box per patch is at least as fast as one per time row (tools/kbench.py)
    if (cfg.split) split = cfg.split;
    if (split == 2 && box_t > 1) { p.a_nbox = box_t; p.a_box_t = 1; }
    else { p.a_nbox = 1; p.a_box_t = box_t; }
    p.a_box_bytes = (uint32_t)box_f * p.a_box_t * pitch_a;
    // resident weights: the whole packed weight next to the pipeline (single n tile), leaving room for two
    // patch stages and one staging buffer
    resident = p.tiles_n == 1 && w_bytes <= 100u * 1024u && w_bytes + 2 * p.a_bytes + staging_bytes + 2048 <= kSmemMax;
    if (cfg.resident == 1) resident = false;
    p.resident = resident ? 1 : 0;
    p.stage_bytes = p.a_bytes + (resident ? 0u : (uint32_t)p.maxg * p.b_bytes);
    fixed = resident ? w_bytes : 0u;
    // two CTAs per SM when the accumulators (2 x 2 x block_n TMEM columns) and ~110 KB of smem each
    // allow it: their serial per-tile latencies (TMA -> MMA -> TMEM drain -> store) overlap
    two_ctas = 2 * cols <= 512 && staging_bytes <= 32 * 1024 &&
               fixed + 2 * p.stage_bytes + staging_bytes + 1024 <= 106u * 1024u;
    if (cfg.two_cta == 1) two_ctas = false;
    const uint32_t budget = two_ctas ? 106u * 1024u : kSmemMax;
    p.nstg = (fixed + 2 * staging_bytes + 2 * p.stage_bytes + 1024 <= budget) ? 2 : 1;
    stages = (int)((budget - 1024 - fixed - p.nstg * staging_bytes) / p.stage_bytes);
    if (stages > 8) stages = 8;
    if (stages >= 2) break;
    if (mode == 1) {
      set_error("clskd_tapconv_fwd_umma: tile does not fit shared memory (stage %u B, weights %u B)", p.stage_bytes, fixed);
      return CLSKD_ERR_UNSUPPORTED;
    }
    mode = 1;        // patches do not fit: one box per tap
  }
  p.stages = stages;
  p.bres_off = (uint32_t)stages * p.stage_bytes;
  p.stg_off = p.bres_off + fixed;
  p.y = d->y; p.y_sB = d->y_sB; p.y_sT = d->y_sT; p.y_sF = d->y_sF; p.y_dtype = d->y_dtype;
  p.bias = d->bias; p.N = d->N;
  p.ep_scale = d->ep_scale; p.ep_shift = d->ep_shift; p.ep_slope = d->ep_slope;
  p.stats_sum = d->stats_sum; p.stats_sumsq = d->stats_sumsq;
  p.accum = d->accumulate ? 1 : 0;

  CUtensorMap tmA0, tmA1, tmB;
  int rc = encode_act_patch(enc, &tmA0, d->x0, d->c0, d->sf, d->Fi, d->Ti, d->B, d->x0_sB, d->x0_sT,
                            d->x0_sF, bk, box_f, p.a_box_t, sw);
  if (rc) { set_error("clskd_tapconv_fwd_umma: cuTensorMapEncodeTiled(x0) failed: %d", rc); return CLSKD_ERR_CUDA; }
  if (d->c1) {
    rc = encode_act_patch(enc, &tmA1, d->x1, d->c1, d->sf, d->Fi, d->Ti, d->B, d->x1_sB, d->x1_sT,
                          d->x1_sF, bk, box_f, p.a_box_t, sw);
    if (rc) { set_error("clskd_tapconv_fwd_umma: cuTensorMapEncodeTiled(x1) failed: %d", rc); return CLSKD_ERR_CUDA; }
  } else {
    tmA1 = tmA0;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)Ctot, (cuuint64_t)Np, (cuuint64_t)d->ntaps};
    cuuint64_t strides[2] = {(cuuint64_t)Ctot * 2, (cuuint64_t)Ctot * 2 * Np};
    cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)p.block_n, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d->w), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { set_error("clskd_tapconv_fwd_umma: cuTensorMapEncodeTiled(w) failed: %d", (int)r); return CLSKD_ERR_CUDA; }
  }
  const int64_t tiles = (int64_t)d->B * p.t_tiles * p.f_tiles * p.tiles_n;
  CLSKD_CHECK_ARG(tiles <= 2147483647LL, "clskd_tapconv_fwd_umma: too many tiles");
  p.num_tiles = (int)tiles;
  CUtensorMap tmY;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->N, (cuuint64_t)d->Fo, (cuuint64_t)d->To, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->y_sF * p.es, (cuuint64_t)d->y_sT * p.es, (cuuint64_t)d->y_sB * p.es};
    cuuint32_t box[4] = {(cuuint32_t)p.gw_y, (cuuint32_t)p.fo_tile, (cuuint32_t)p.t_tile, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmY, p.es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d->y,
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(p.gw_y * p.es),
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { set_error("clskd_tapconv_fwd_umma: cuTensorMapEncodeTiled(y) failed: %d", (int)r); return CLSKD_ERR_CUDA; }
  }
  size_t smem = (size_t)p.stg_off + (size_t)p.nstg * staging_bytes + 1024;
  const bool ep = d->ep_scale || d->ep_slope || d->stats_sum;
  static size_t smem_set[2] = {0, 0};
  if (smem > smem_set[ep ? 1 : 0]) {
    cudaError_t e = ep ? cudaFuncSetAttribute(tapconv_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                       : cudaFuncSetAttribute(tapconv_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("clskd_tapconv_fwd_umma: smem attr: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
    smem_set[ep ? 1 : 0] = smem;
  }
  const int64_t max_ctas = (int64_t)sm_count() * (two_ctas ? 2 : 1);
  const unsigned grid = (unsigned)(tiles < max_ctas ? tiles : max_ctas);
  if (ep) tapconv_umma_kernel<true><<<grid, kThreads, smem, (cudaStream_t)stream>>>(tmA0, tmA1, tmB, tmY, p);
  else tapconv_umma_kernel<false><<<grid, kThreads, smem, (cudaStream_t)stream>>>(tmA0, tmA1, tmB, tmY, p);
  CLSKD_CHECK_LAUNCH("clskd_tapconv_fwd_umma");
  return CLSKD_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Per-shape autotuner.  Which configuration wins depends on what bounds the shape (L2 -> shared-memory fill
// rate, shared-memory operand reads of narrow-N MMAs, resident CTAs per SM, epilogue latency): the first call
// for a shape signature times a handful of candidates on the caller's own tensors and caches the winner
// (like cuDNN's find mode).  Candidate runs write the same outputs; their batch statistics go to a scratch
// buffer.  clskd_set_tuning keys 0-4 force a configuration instead, key 5 = 1 disables the tuner.
// ---------------------------------------------------------------------------------------------------------
namespace {
std::mutex g_tune_mu;
std::unordered_map<std::string, FwdCfg> g_tune_cache;
double* g_tune_scratch = nullptr;      // [2][256] fp64 statistics sink for candidate runs

std::string shape_key(const ClskdTapConv* d) {
  char buf[512];
  int n = snprintf(buf, sizeof(buf), "%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d%d%d%d|", d->B, d->To, d->Fo, d->Ti, d->Fi, d->sf,
                   d->ntaps, d->c0, d->c1, d->N, d->y_dtype, d->bias ? 1 : 0, d->ep_scale ? 1 : 0, d->ep_slope ? 1 : 0,
                   d->stats_sum ? 1 : 0);
  for (int j = 0; j < d->ntaps && n < (int)sizeof(buf) - 16; ++j) n += snprintf(buf + n, sizeof(buf) - n, "%d:%d;", d->dt[j], d->df[j]);
  return std::string(buf);
}

FwdCfg autotune(const ClskdTapConv* d, cudaStream_t st) {
  const FwdCfg cands[] = {
      {1, 0, 0, 0, 0, 0},    // round-1 kernel
      {0, 0, 0, 0, 0, 0},    // automatic grouping, resident weights when they fit
      {0, 0, 1, 0, 0, 0},    // ... weights through the ring
      {0, 1, 0, 0, 0, 0},    // one box per tap, resident weights
      {0, 0, 0, 0, 0, 32},   // K chunk 32 (smaller stages)
      {0, 0, 1, 0, 0, 32},
      {0, 2, 1, 0, 0, 0},    // time-grouped patches
      {0, 2, 1, 0, 0, 32},
  };
  if (!g_tune_scratch && cudaMalloc(&g_tune_scratch, sizeof(double) * 512) != cudaSuccess) {
    cudaGetLastError();
    return cands[0];
  }
  ClskdTapConv t = *d;
  if (t.stats_sum) {
    t.stats_sum = g_tune_scratch;
    t.stats_sumsq = g_tune_scratch + 256;
  }
  cudaDeviceSynchronize();           // nothing else may share the SMs while the candidates are timed
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int best = 0;
  float best_ms = 1e30f;
  for (int i = 0; i < (int)(sizeof(cands) / sizeof(cands[0])); ++i) {
    if (launch_cfg(&t, cands[i], st) != CLSKD_OK) { cudaGetLastError(); continue; }
    cudaEventRecord(e0, st);
    bool ok = true;
    for (int r = 0; r < 2 && ok; ++r) ok = launch_cfg(&t, cands[i], st) == CLSKD_OK;
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess || !ok) { cudaGetLastError(); continue; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best_ms) { best_ms = ms; best = i; }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return cands[best];
}
}  // namespace

extern "C" int clskd_tapconv_fwd_umma(const ClskdTapConv* d, void* stream) {
  CLSKD_CHECK_ARG(d && d->x0 && d->w && d->y, "clskd_tapconv_fwd_umma: null pointer");
  CLSKD_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= CLSKD_MAX_TAPS, "clskd_tapconv_fwd_umma: ntaps");
  if (const char* why = umma_unsupported(d)) {
    set_error("clskd_tapconv_fwd_umma: unsupported: %s", why);
    return CLSKD_ERR_UNSUPPORTED;
  }
  const int64_t M = (int64_t)d->B * d->To * d->Fo;
  if (M == 0) return CLSKD_OK;
  const bool forced = g_tune_v1 || g_tune_mode || g_tune_resident || g_tune_two_cta || g_tune_split;
  static const bool env_off = getenv("CLSKD_AUTOTUNE") && atoi(getenv("CLSKD_AUTOTUNE")) == 0;
  if (forced || g_tune_noauto || env_off || M < 65536) {
    // forced configuration, or a launch too small to be worth tuning (automatic configuration)
    const FwdCfg cfg = {g_tune_v1, g_tune_mode, g_tune_resident, g_tune_two_cta, g_tune_split, 0};
    return launch_cfg(d, cfg, stream);
  }
  FwdCfg cfg;
  {
    std::lock_guard<std::mutex> lk(g_tune_mu);
    const std::string key = shape_key(d);
    auto it = g_tune_cache.find(key);
    if (it == g_tune_cache.end()) {
      // an accumulating launch cannot be timed on the caller's tensors (every candidate would add to y): it uses the
      // configuration a plain launch of the same shape tuned earlier, else the automatic one
      if (d->accumulate) return launch_cfg(d, FwdCfg{0, 0, 0, 0, 0, 0}, stream);
      it = g_tune_cache.emplace(key, autotune(d, (cudaStream_t)stream)).first;
    }
    cfg = it->second;
  }
  return launch_cfg(d, cfg, stream);
}

// number of shapes tuned so far and how many chose the round-1 kernel (diagnostics for tools/kbench.py, bench.py)
extern "C" int clskd_tuning_stats(int* n_shapes, int* n_v1) {
  std::lock_guard<std::mutex> lk(g_tune_mu);
  int v = 0;
  for (auto& kv : g_tune_cache) v += kv.second.v1 ? 1 : 0;
  if (n_shapes) *n_shapes = (int)g_tune_cache.size();
  if (n_v1) *n_v1 = v;
  return CLSKD_OK;
}
