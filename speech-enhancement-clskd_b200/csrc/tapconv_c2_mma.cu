// First encoder layer of DCCRN on tensor cores (bf16 policy): a strided convolution over the TWO-channel fp32
// spectrogram (real | imaginary, DCCRN.py:178 / tools_for_model.py:237-246 ComplexConv2d with in_channels = 1 per part)
// onto N <= 64 channels, K = ntaps * 2 <= 32.  The CUDA-core kernel it replaces (tapconv_fwd_smallk_kernel) spent
// 900 instructions per output row on K * N scalar FMAs with the weights broadcast from shared memory
// (0.36 ms for 64 x 643 x 128 rows onto 32 channels, 5.7x its HBM floor).
//
// mma.sync m16n8k16: 16 consecutive output frequencies are the M rows, k = 2 * tap + channel, so a thread's A pair
// (k, k+1) is the (re, im) pair of ONE input position = one 8-byte load; the weights live in registers as B fragments
// for the whole kernel.  The fp32 inputs and the fp32 weights are both split into bf16 hi + lo
// (x w ~ xh wh + xl wh + xh wl, three MMAs, fp32 accumulation), so the layer keeps ~2^-16 relative accuracy - the
// spectrogram is the one fp32 activation of the bf16 policy and is not rounded to bf16 here either.
#include "common.cuh"

namespace clskd {
namespace {

__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// (v0, v1) fp32 -> packed bf16 pairs hi and lo with v ~ hi + lo (v0 in the low half: the lower k index)
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);                  // one packed conversion
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float r0 = v0 - __uint_as_float(hi << 16), r1 = v1 - __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int NT8>      // N / 8 column blocks
__global__ void __launch_bounds__(256, 2) tapconv_fwd_c2_mma_kernel(ClskdTapConv d) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int N = d.N;
  const float* w = reinterpret_cast<const float*>(d.w);            // [ntaps][2][N]
  const int ksteps = d.ntaps > 8 ? 2 : 1;
  // per-thread taps: k step s, half h -> tap 8s + 4h + t
  int tdt[2][2], tdf[2][2];
  bool tok[2][2];
  uint32_t bh[2][NT8][2], bl[2][NT8][2];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int tap = 8 * s + 4 * h + t;
      tok[s][h] = tap < d.ntaps;
      int vdt = 0, vdf = 0;
#pragma unroll
      for (int j = 0; j < CLSKD_MAX_TAPS; ++j)
        if (j == tap) { vdt = d.dt[j]; vdf = d.df[j]; }             // (constant-bank reads, no local-memory copy of d)
      tdt[s][h] = vdt;
      tdf[s][h] = vdf;
#pragma unroll
      for (int j = 0; j < NT8; ++j) {
        const int n = 8 * j + g;
        const float w0 = tok[s][h] ? w[(size_t)(tap * 2) * N + n] : 0.f;
        const float w1 = tok[s][h] ? w[(size_t)(tap * 2 + 1) * N + n] : 0.f;
        split2(w0, w1, bh[s][j][h], bl[s][j][h]);
      }
    }
  float bias2[NT8][2];
#pragma unroll
  for (int j = 0; j < NT8; ++j) {
    bias2[j][0] = d.bias ? d.bias[8 * j + 2 * t] : 0.f;
    bias2[j][1] = d.bias ? d.bias[8 * j + 2 * t + 1] : 0.f;
  }

  const float* x = reinterpret_cast<const float*>(d.x0);
  __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(d.y);
  const int tiles_f = d.Fo >> 4;
  // 32-bit tile arithmetic (the launcher checks the range): six 64-bit divisions per tile were most of the kernel
  const unsigned ntiles = (unsigned)d.B * (unsigned)d.To * (unsigned)tiles_f;
  const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
  // per-warp staging tile [16 rows][N] bf16 (+16 bytes per row: conflict-free) for row-contiguous 16-byte stores
  extern __shared__ __align__(16) uint8_t c2_sm[];
  constexpr int SROW = NT8 * 16 + 16;                      // bytes per staged row
  uint8_t* stg = c2_sm + (threadIdx.x >> 5) * (16 * SROW);

  // software pipeline: the 8-byte (re, im) loads of the NEXT tile are in flight while this tile's MMAs and stores run
  // (one tile per warp iteration is a serial load -> mma -> store chain otherwise: 2.2 us per tile, latency bound)
  float2 xv[2][2][2];
  // element offset of a tap relative to input position (ti = to, fi = f * sf): 32-bit (the launcher checks the range)
  int toff[2][2];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int h = 0; h < 2; ++h) toff[s][h] = tdt[s][h] * (int)d.x0_sT + tdf[s][h] * (int)d.x0_sF;
  const int fstep = d.sf * (int)d.x0_sF;                    // elements between consecutive output frequencies
  auto load_tile = [&](int b, int to, int f0) {
    const float* xr = x + (int64_t)b * d.x0_sB + (int64_t)to * d.x0_sT + (f0 + g) * fstep;
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ti = to + tdt[s][h];
        const bool okt = s < ksteps && tok[s][h] && ti >= 0 && ti < d.Ti;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int fi = (f0 + g + 8 * rr) * d.sf + tdf[s][h];
          xv[s][h][rr] = make_float2(0.f, 0.f);
          if (okt && fi >= 0 && fi < d.Fi) xv[s][h][rr] = __ldg(reinterpret_cast<const float2*>(xr + 8 * rr * fstep + toff[s][h]));
        }
      }
  };
  // tile -> (b, to, f tile) once, then incrementally (a division per tile and coordinate was most of the loop)
  unsigned tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int ft = (int)(tile % (unsigned)tiles_f);
  int to = (int)((tile / (unsigned)tiles_f) % (unsigned)d.To), b = (int)((tile / (unsigned)tiles_f) / (unsigned)d.To);
  const int d_ft = (int)(nwarps % (unsigned)tiles_f);
  const int d_to = (int)((nwarps / (unsigned)tiles_f) % (unsigned)d.To), d_b = (int)((nwarps / (unsigned)tiles_f) / (unsigned)d.To);
  auto advance = [&](int& ft_, int& to_, int& b_) {
    ft_ += d_ft;
    int c = ft_ >= tiles_f ? 1 : 0;
    ft_ -= c * tiles_f;
    to_ += d_to + c;
    c = to_ >= d.To ? 1 : 0;
    to_ -= c * d.To;
    b_ += d_b + c;
  };
  if (tile < ntiles) load_tile(b, to, ft << 4);
  for (; tile < ntiles; tile += nwarps) {
    const int f0 = ft << 4, to_c = to, b_c = b;
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) split2(xv[s][h][rr].x, xv[s][h][rr].y, ah[s][2 * h + rr], al[s][2 * h + rr]);
    advance(ft, to, b);
    if (tile + nwarps < ntiles) load_tile(b, to, ft << 4);
    float acc[NT8][4];
#pragma unroll
    for (int j = 0; j < NT8; ++j) {
      acc[j][0] = acc[j][2] = bias2[j][0];
      acc[j][1] = acc[j][3] = bias2[j][1];
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (s < ksteps) {
#pragma unroll
        for (int j = 0; j < NT8; ++j) {
          mma_bf16_16816(acc[j], ah[s], bh[s][j][0], bh[s][j][1]);
          mma_bf16_16816(acc[j], al[s], bh[s][j][0], bh[s][j][1]);
          mma_bf16_16816(acc[j], ah[s], bl[s][j][0], bl[s][j][1]);
        }
      }
    }
    // stage the [16][N] tile, then every lane stores 16 contiguous bytes of a row
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NT8; ++j) {
      const __nv_bfloat162 p0 = __floats2bfloat162_rn(acc[j][0], acc[j][1]);
      const __nv_bfloat162 p1 = __floats2bfloat162_rn(acc[j][2], acc[j][3]);
      *reinterpret_cast<__nv_bfloat162*>(stg + g * SROW + (8 * j + 2 * t) * 2) = p0;
      *reinterpret_cast<__nv_bfloat162*>(stg + (g + 8) * SROW + (8 * j + 2 * t) * 2) = p1;
    }
    __syncwarp();
    __nv_bfloat16* yt = y + (int64_t)b_c * d.y_sB + (int64_t)to_c * d.y_sT + (int64_t)f0 * d.y_sF;
#pragma unroll
    for (int i = lane; i < 16 * NT8; i += 32) {
      const int row = i / NT8, ch = i - row * NT8;
      const uint4 v = *reinterpret_cast<const uint4*>(stg + row * SROW + ch * 16);
      *reinterpret_cast<uint4*>(yt + (int64_t)row * d.y_sF + ch * 8) = v;
    }
  }
}

// Weight gradient of the same layer on mma.sync:  dW[k][n] = sum_rows x[row + tap(k)][c(k)] dY[row][n],  k = 2 tap + c.
// D[k (M, two 16-blocks for <= 16 taps)][n] += A[k][16 rows] B[16 rows][n]: the rows of a tile are 16 consecutive output
// frequencies of one (b, t) line; x is split into bf16 hi + lo (two MMAs), dY is bf16 already.  A warp keeps its partial dW
// in registers over all its tiles; CTAs combine in shared memory and add to the fp32 dW (zeroed by the caller).
// The CUDA-core kernel it replaces (tapconv_wgrad_c2_kernel) is bound by its shared-memory operand reads: 10 LDS.64 per
// 20 FMAs, 0.45 ms for 84 MB + 169 MB.
template <int NT8>
__global__ void __launch_bounds__(256, 2) tapconv_wgrad_c2_mma_kernel(ClskdTapConv d) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int N = d.N;
  __shared__ float red[32 * 64];
  for (int i = threadIdx.x; i < 32 * N; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int mtiles = d.ntaps > 8 ? 2 : 1;
  // per-thread contraction rows k = 16 mt + 8 h + g -> (tap, channel)
  int kdt[2][2], kdf[2][2], koff[2][2];
  bool kok[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = 16 * mt + 8 * h + g, tap = k >> 1, c = k & 1;
      kok[mt][h] = tap < d.ntaps;
      int vdt = 0, vdf = 0;
#pragma unroll
      for (int j = 0; j < CLSKD_MAX_TAPS; ++j)
        if (j == tap) { vdt = d.dt[j]; vdf = d.df[j]; }
      kdt[mt][h] = vdt;
      kdf[mt][h] = vdf;
      koff[mt][h] = vdt * (int)d.x0_sT + vdf * (int)d.x0_sF + c;
    }
  const int fstep = d.sf * (int)d.x0_sF;
  const float* x = reinterpret_cast<const float*>(d.x0);
  const __nv_bfloat16* dy = reinterpret_cast<const __nv_bfloat16*>(d.y);
  float acc[2][NT8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int j = 0; j < NT8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][j][e] = 0.f;

  const int tiles_f = d.Fo >> 4;
  const unsigned ntiles = (unsigned)d.B * (unsigned)d.To * (unsigned)tiles_f;
  const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
  unsigned tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int ft = (int)(tile % (unsigned)tiles_f);
  int to = (int)((tile / (unsigned)tiles_f) % (unsigned)d.To), b = (int)((tile / (unsigned)tiles_f) / (unsigned)d.To);
  const int d_ft = (int)(nwarps % (unsigned)tiles_f);
  const int d_to = (int)((nwarps / (unsigned)tiles_f) % (unsigned)d.To), d_b = (int)((nwarps / (unsigned)tiles_f) / (unsigned)d.To);
  for (; tile < ntiles; tile += nwarps) {
    const int f0 = ft << 4;
    // ---- B fragments: dY rows (2t, 2t+1) and (2t+8, 2t+9), column 8j + g
    const __nv_bfloat16* dyl = dy + (int64_t)b * d.y_sB + (int64_t)to * d.y_sT + (int64_t)f0 * d.y_sF + g;
    uint32_t bf[NT8][2];
#pragma unroll
    for (int j = 0; j < NT8; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const __nv_bfloat16* p0 = dyl + (int64_t)(2 * t + 8 * h) * d.y_sF + 8 * j;
        const uint32_t lo = (uint32_t)__bfloat16_as_ushort(p0[0]), hi = (uint32_t)__bfloat16_as_ushort(p0[d.y_sF]);
        bf[j][h] = lo | (hi << 16);
      }
    // ---- A fragments: x at (row + tap) for the thread's four k, rows (2t, 2t+1, 2t+8, 2t+9)
    const float* xr = x + (int64_t)b * d.x0_sB + (int64_t)to * d.x0_sT;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      if (mt < mtiles) {
        uint32_t ah[4], al[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ti = to + kdt[mt][h];
          const bool okt = kok[mt][h] && ti >= 0 && ti < d.Ti;
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            float v[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int fo = f0 + 2 * t + 8 * rr + q;
              const int fi = fo * d.sf + kdf[mt][h];
              v[q] = (okt && fi >= 0 && fi < d.Fi) ? __ldg(xr + fo * fstep + koff[mt][h]) : 0.f;
            }
            split2(v[0], v[1], ah[h + 2 * rr], al[h + 2 * rr]);
          }
        }
#pragma unroll
        for (int j = 0; j < NT8; ++j) {
          mma_bf16_16816(acc[mt][j], ah, bf[j][0], bf[j][1]);
          mma_bf16_16816(acc[mt][j], al, bf[j][0], bf[j][1]);
        }
      }
    }
    // next tile
    ft += d_ft;
    int c = ft >= tiles_f ? 1 : 0;
    ft -= c * tiles_f;
    to += d_to + c;
    c = to >= d.To ? 1 : 0;
    to -= c * d.To;
    b += d_b + c;
  }
  // ---- reduction: D rows k = 16 mt + g (+8), columns 8 j + 2 t (+1)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int j = 0; j < NT8; ++j) {
      const int k0 = 16 * mt + g, n0 = 8 * j + 2 * t;
      atomicAdd(&red[k0 * N + n0], acc[mt][j][0]);
      atomicAdd(&red[k0 * N + n0 + 1], acc[mt][j][1]);
      atomicAdd(&red[(k0 + 8) * N + n0], acc[mt][j][2]);
      atomicAdd(&red[(k0 + 8) * N + n0 + 1], acc[mt][j][3]);
    }
  __syncthreads();
  float* dw = reinterpret_cast<float*>(const_cast<void*>(d.w));
  for (int i = threadIdx.x; i < 2 * d.ntaps * N; i += blockDim.x) atomicAdd(dw + i, red[i]);
}

}  // namespace

namespace c2mma {

// true = handled (launch issued; the caller checks the launch status)
bool try_fwd(const ClskdTapConv* d, cudaStream_t st) {
  if (d->x_dtype != CLSKD_F32 || d->y_dtype != CLSKD_BF16 || d->c0 != 2 || d->c1 != 0 || d->accumulate) return false;
  if (d->N % 8 || d->N < 8 || d->N > 64 || d->ntaps > 16 || d->Fo % 16 || d->Fo <= 0) return false;
  if (d->x0_sF != 2 || d->x0_sT % 2 || d->x0_sB % 2 || (uintptr_t)d->x0 % 8) return false;          // 8-byte (re, im) loads
  if (d->y_sF % 8 || d->y_sT % 8 || d->y_sB % 8 || (uintptr_t)d->y % 16 || (uintptr_t)d->w % 4) return false;  // 16-byte stores
  const int64_t ntiles = (int64_t)d->B * d->To * (d->Fo / 16);
  if (ntiles < 1024 || ntiles > 2000000000LL) return false;      // tiny launches stay on the scalar kernel; 32-bit tile ids
  if ((int64_t)(d->Ti + 16) * d->x0_sT > 2000000000LL || d->x0_sT > 100000000LL) return false;      // 32-bit tap offsets
  int64_t blocks = (ntiles + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 2;          // resident CTAs: the weight fragments are built once per CTA
  if (blocks > cap) blocks = cap;
  const size_t sh = (size_t)8 * 16 * ((d->N / 8) * 16 + 16);
  switch (d->N / 8) {
    case 1: tapconv_fwd_c2_mma_kernel<1><<<(unsigned)blocks, 256, sh, st>>>(*d); break;
    case 2: tapconv_fwd_c2_mma_kernel<2><<<(unsigned)blocks, 256, sh, st>>>(*d); break;
    case 3: tapconv_fwd_c2_mma_kernel<3><<<(unsigned)blocks, 256, sh, st>>>(*d); break;
    case 4: tapconv_fwd_c2_mma_kernel<4><<<(unsigned)blocks, 256, sh, st>>>(*d); break;
    case 5: tapconv_fwd_c2_mma_kernel<5><<<(unsigned)blocks, 256, sh, st>>>(*d); break;
    case 6: tapconv_fwd_c2_mma_kernel<6><<<(unsigned)blocks, 256, sh, st>>>(*d); break;
    case 7: tapconv_fwd_c2_mma_kernel<7><<<(unsigned)blocks, 256, sh, st>>>(*d); break;
    default: tapconv_fwd_c2_mma_kernel<8><<<(unsigned)blocks, 256, sh, st>>>(*d); break;
  }
  return true;
}

// dW (fp32 [ntaps][2][N], zeroed by the caller or accumulated into) of the same layer; true = handled
bool try_wgrad(const ClskdTapConv* d, cudaStream_t st) {
  if (d->x_dtype != CLSKD_F32 || d->y_dtype != CLSKD_BF16 || d->c0 != 2 || d->c1 != 0) return false;
  if (d->N % 8 || d->N < 8 || d->N > 32 || d->ntaps > 16 || d->Fo % 16 || d->Fo <= 0) return false;
  if (d->x0_sF != 2 || (uintptr_t)d->x0 % 4 || (uintptr_t)d->y % 2 || (uintptr_t)d->w % 4) return false;
  const int64_t ntiles = (int64_t)d->B * d->To * (d->Fo / 16);
  if (ntiles < 1024 || ntiles > 2000000000LL) return false;
  if ((int64_t)(d->Ti + 16) * d->x0_sT > 2000000000LL) return false;      // 32-bit tap offsets
  int64_t blocks = (ntiles + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 2;
  if (blocks > cap) blocks = cap;
  switch (d->N / 8) {
    case 1: tapconv_wgrad_c2_mma_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(*d); break;
    case 2: tapconv_wgrad_c2_mma_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(*d); break;
    case 3: tapconv_wgrad_c2_mma_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(*d); break;
    default: tapconv_wgrad_c2_mma_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(*d); break;
  }
  return true;
}

}  // namespace c2mma
}  // namespace clskd
