// Pointwise (1x1, stride 1, dense operands) fast paths of the tap-list contraction for the two
// HBM-bound extremes that ABF's attention and 2-channel mask layers produce:
//   * N <= 2 outputs per row   (GEMV over <= 256 channels: attention logits, dgrad onto 2 channels)
//   * K <= 4 inputs per row    (outer product onto N channels: 1x1 conv of / dgrad onto a 2-channel map)
// and their weight gradients.  Rows are addressed flat (x + m*C), so there is no per-row index
// arithmetic; every global access is a 16-byte vector; several rows are in flight per lane group and
// the N<=2 forward reduces 4 rows x 2 outputs with a halving butterfly (9 shuffles for 8 sums).
#include "common.cuh"

namespace clskd {
namespace {

__device__ __forceinline__ void ld8(const float* p, float* o) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float* o) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    o[2 * i] = f.x;
    o[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void st8(float* p, const float* v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float* v) {
  uint4 u;
  __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2);
  u.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(p) = u;
}

constexpr int PT = 256;
constexpr int RQ = 4;   // rows in flight per lane group

// ------------------------------------------------------------------------------------------ N <= 2
// lane group of `up2` lanes (power of two >= U = Ctot/8 units) per row; unit u < U0 reads source 0
template <typename TX, typename TY>
__global__ void __launch_bounds__(PT) pw_fwd_n2_kernel(const TX* __restrict__ x0, const TX* __restrict__ x1, int c0,
                                                       int c1, const float* __restrict__ w, const float* __restrict__ bias,
                                                       int N, int64_t M, int up2, TY* __restrict__ y, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int ul = lane & (up2 - 1), sub = lane / up2, rpw = 32 / up2;
  const int U0 = c0 >> 3, U = (c0 + c1) >> 3;
  const bool act = ul < U;
  const bool s1 = ul >= U0;
  const TX* xb = s1 ? x1 + (ul - U0) * 8 : x0 + ul * 8;
  const int cs = s1 ? c1 : c0;                      // row pitch of this lane's source
  float wr[8][2];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = ul * 8 + e;                        // concatenated channel index
    wr[e][0] = act ? w[(int64_t)k * N] : 0.f;
    wr[e][1] = (act && N > 1) ? w[(int64_t)k * N + 1] : 0.f;
  }
  const float b0 = bias ? bias[0] : 0.f, b1 = (bias && N > 1) ? bias[1] : 0.f;
  const int64_t warp0 = ((int64_t)blockIdx.x * PT + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * PT) >> 5;
  const int64_t rows_it = (int64_t)rpw * RQ;
  for (int64_t m0 = warp0 * rows_it; m0 < M; m0 += nwarps * rows_it) {
    float xv[RQ][8];
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
      const int64_t m = m0 + (int64_t)q * rpw + sub;
      if (act && m < M) {
        ld8(xb + m * cs, xv[q]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) xv[q][e] = 0.f;
      }
    }
    float v[RQ * 2];
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        a0 = fmaf(xv[q][e], wr[e][0], a0);
        a1 = fmaf(xv[q][e], wr[e][1], a1);
      }
      v[2 * q] = a0;
      v[2 * q + 1] = a1;
    }
    // halving butterfly over the up2 lanes of a row group: after a halving step a lane keeps half of
    // its value slots (upper half when its offset bit is set); once one slot is left: plain sums
    int cnt = RQ * 2;
    int sel = 0;                                      // which of the original 8 slots slot 0 now holds
    for (int o = up2 >> 1; o > 0; o >>= 1) {
      if (cnt > 1) {
        const int h = cnt >> 1;
        const bool up = (ul & o) != 0;
#pragma unroll
        for (int i = 0; i < RQ; ++i) {
          if (i < h) {
            const float send = up ? v[i] : v[i + h];
            const float keep = up ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
        if (up) sel += h;
        cnt = h;
      } else {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
      }
    }
    // lanes whose plain-step bits are zero write their cnt slots
    int plain_mask = 0;
    {
      int c2 = RQ * 2;
      for (int o = up2 >> 1; o > 0; o >>= 1) {
        if (c2 > 1) c2 >>= 1;
        else plain_mask |= o;
      }
    }
    if ((ul & plain_mask) == 0) {
#pragma unroll
      for (int i = 0; i < RQ * 2; ++i) {
        if (i < cnt) {
          const int slot = sel + i;                   // original slot = q*2 + n
          const int q = slot >> 1, n = slot & 1;
          const int64_t m = m0 + (int64_t)q * rpw + sub;
          if (m < M && n < N) {
            float val = v[i] + (n ? b1 : b0);
            TY* yp = y + m * N + n;
            if (accumulate) val += ld_f(yp);
            st_f(yp, val);
          }
        }
      }
    }
  }
}

template <typename TX, typename TY>
__global__ void __launch_bounds__(PT) pw_wgrad_n2_kernel(const TX* __restrict__ x0, const TX* __restrict__ x1, int c0,
                                                         int c1, const TY* __restrict__ dy, int N, int64_t M, int up2,
                                                         float* __restrict__ dw) {
  extern __shared__ float red[];   // [Ctot][2]
  const int Ctot = c0 + c1;
  for (int i = threadIdx.x; i < Ctot * 2; i += PT) red[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int ul = lane & (up2 - 1), sub = lane / up2, rpw = 32 / up2;
  const int U0 = c0 >> 3, U = Ctot >> 3;
  const bool act = ul < U;
  const bool s1 = ul >= U0;
  const TX* xb = s1 ? x1 + (ul - U0) * 8 : x0 + ul * 8;
  const int cs = s1 ? c1 : c0;
  float acc[8][2];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e][0] = acc[e][1] = 0.f;
  const int64_t warp0 = ((int64_t)blockIdx.x * PT + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * PT) >> 5;
  const int64_t rows_it = (int64_t)rpw * RQ;
  for (int64_t m0 = warp0 * rows_it; m0 < M; m0 += nwarps * rows_it) {
    float xv[RQ][8], g0[RQ], g1[RQ];
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
      const int64_t m = m0 + (int64_t)q * rpw + sub;
      const bool live = act && m < M;
      if (live) {
        ld8(xb + m * cs, xv[q]);
        g0[q] = ld_f(dy + m * N);
        g1[q] = N > 1 ? ld_f(dy + m * N + 1) : 0.f;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) xv[q][e] = 0.f;
        g0[q] = g1[q] = 0.f;
      }
    }
#pragma unroll
    for (int q = 0; q < RQ; ++q)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        acc[e][0] = fmaf(xv[q][e], g0[q], acc[e][0]);
        acc[e][1] = fmaf(xv[q][e], g1[q], acc[e][1]);
      }
  }
  if (act) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(&red[(ul * 8 + e) * 2], acc[e][0]);
      atomicAdd(&red[(ul * 8 + e) * 2 + 1], acc[e][1]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Ctot * 2; i += PT) {
    const int n = i & 1, k = i >> 1;
    if (n < N) atomicAdd(dw + (int64_t)k * N + n, red[i]);
  }
}

// ------------------------------------------------------------------------------------------ K <= 4
template <typename TX, typename TY, int K>
__global__ void __launch_bounds__(PT) pw_fwd_k_kernel(const TX* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, int N, int64_t M,
                                                      TY* __restrict__ y) {
  const int tpr = N >> 3, rows_par = PT / tpr;
  const int ng = threadIdx.x % tpr, rs = threadIdx.x / tpr;
  if (rs >= rows_par) return;
  float wr[K][8], br[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    br[e] = bias ? bias[ng * 8 + e] : 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) wr[k][e] = w[(int64_t)k * N + ng * 8 + e];
  }
  constexpr int RB = 4;
  for (int64_t m0 = (int64_t)blockIdx.x * rows_par * RB + rs; m0 < M; m0 += (int64_t)gridDim.x * rows_par * RB) {
    float xv[RB][K];
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      const int64_t m = m0 + (int64_t)q * rows_par;
#pragma unroll
      for (int k = 0; k < K; ++k) xv[q][k] = m < M ? ld_f(x + m * K + k) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      const int64_t m = m0 + (int64_t)q * rows_par;
      if (m < M) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float a = br[e];
#pragma unroll
          for (int k = 0; k < K; ++k) a = fmaf(xv[q][k], wr[k][e], a);
          o[e] = a;
        }
        st8(y + m * N + ng * 8, o);
      }
    }
  }
}

template <typename TX, typename TY, int K>
__global__ void __launch_bounds__(PT) pw_wgrad_k_kernel(const TX* __restrict__ x, const TY* __restrict__ dy, int N,
                                                        int64_t M, float* __restrict__ dw) {
  extern __shared__ float red[];   // [K][N]
  for (int i = threadIdx.x; i < K * N; i += PT) red[i] = 0.f;
  __syncthreads();
  const int tpr = N >> 3, rows_par = PT / tpr;
  const int ng = threadIdx.x % tpr, rs = threadIdx.x / tpr;
  float acc[K][8];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
  constexpr int RB = 4;
  if (rs < rows_par) {
    for (int64_t m0 = (int64_t)blockIdx.x * rows_par * RB + rs; m0 < M; m0 += (int64_t)gridDim.x * rows_par * RB) {
      float g[RB][8], xv[RB][K];
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        const int64_t m = m0 + (int64_t)q * rows_par;
        if (m < M) {
          ld8(dy + m * N + ng * 8, g[q]);
#pragma unroll
          for (int k = 0; k < K; ++k) xv[q][k] = ld_f(x + m * K + k);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) g[q][e] = 0.f;
#pragma unroll
          for (int k = 0; k < K; ++k) xv[q][k] = 0.f;
        }
      }
#pragma unroll
      for (int q = 0; q < RB; ++q)
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[k][e] = fmaf(xv[q][k], g[q][e], acc[k][e]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(&red[k * N + ng * 8 + e], acc[k][e]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * N; i += PT) atomicAdd(dw + i, red[i]);
}

// ------------------------------------------------------------------------------------------ tap sum
struct TapList {
  int n;
  int dt[CLSKD_MAX_TAPS], df[CLSKD_MAX_TAPS];
};

template <typename TZ, typename TY>
__global__ void __launch_bounds__(PT) tapsum_fwd_kernel(const TZ* __restrict__ z, int B, int Ti, int Fi, int To,
                                                        int Fo, int sf, int Zc, TapList taps, int N,
                                                        const float* __restrict__ bias, TY* __restrict__ y) {
  const unsigned M = (unsigned)B * To * Fo;
  for (unsigned m = blockIdx.x * PT + threadIdx.x; m < M; m += gridDim.x * PT) {
    const unsigned r = m / (unsigned)Fo;
    const int f = (int)(m - r * (unsigned)Fo);
    const int b = (int)(r / (unsigned)To), t = (int)(r - (unsigned)b * (unsigned)To);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < taps.n; ++j) {
      const int tt = t + taps.dt[j];
      int ff = f + taps.df[j];
      if (tt < 0 || tt >= Ti || ff < 0) continue;
      if (sf > 1) {
        if (ff % sf) continue;
        ff /= sf;
      }
      if (ff >= Fi) continue;
      const TZ* zp = z + (((int64_t)b * Ti + tt) * Fi + ff) * Zc + j * N;
      for (int n = 0; n < N; ++n) acc[n] += ld_f(zp + n);
    }
    for (int n = 0; n < N; ++n) st_f(y + (int64_t)m * N + n, acc[n] + (bias ? bias[n] : 0.f));
  }
}


// ---- N == 2 (the mask-level maps: the only users on the training path) without per-element index arithmetic: one
// block row per (b, t) line (no divisions per thread), a (tap, 2-channel) pair is one 4-byte (bf16) / 8-byte (fp32)
// load, the taps are unrolled with all loads issued before the sum.  The generic kernels above spend ~15 instructions
// per 2-byte load (0.43 ms per launch at 10.5 M outputs against a 0.06 ms HBM floor).
template <typename T>
__device__ __forceinline__ float2 ld_pair(const T* p);
template <>
__device__ __forceinline__ float2 ld_pair<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <>
__device__ __forceinline__ float2 ld_pair<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ void st_pair(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void st_pair(__nv_bfloat16* p, float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<uint32_t*>(&h);
}

template <typename TZ, typename TY, int NT>
__global__ void __launch_bounds__(PT) tapsum_fwd_n2_kernel(const TZ* __restrict__ z, int Ti, int Fi, int To, int Fo,
                                                           int sf, int Zc, TapList taps, const float* __restrict__ bias,
                                                           TY* __restrict__ y) {
  const int row = blockIdx.x;                       // b * To + t
  const int b = row / To, t = row - b * To;
  const float b0 = bias ? bias[0] : 0.f, b1 = bias ? bias[1] : 0.f;
  const TZ* zb = z + (int64_t)b * Ti * Fi * Zc;
  for (int f = blockIdx.y * PT + threadIdx.x; f < Fo; f += gridDim.y * PT) {
    float2 v[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int tt = t + taps.dt[j];
      int ff = f + taps.df[j];
      bool ok = j < taps.n && tt >= 0 && tt < Ti && ff >= 0;
      if (sf == 2) {
        ok = ok && !(ff & 1);
        ff >>= 1;
      }
      ok = ok && ff < Fi;
      v[j] = ok ? ld_pair(zb + ((int64_t)tt * Fi + ff) * Zc + 2 * j) : make_float2(0.f, 0.f);
    }
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      a0 += v[j].x;
      a1 += v[j].y;
    }
    st_pair(y + ((int64_t)row * Fo + f) * 2, a0 + b0, a1 + b1);
  }
}

// one thread per (input position, 8-channel group = 4 taps) of dz
template <typename TD, typename TZ>
__global__ void __launch_bounds__(PT) tapsum_bwd_n2_kernel(const TD* __restrict__ dy, int Ti, int Fi, int To, int Fo,
                                                           int sf, int Zc, TapList taps, TZ* __restrict__ dz) {
  const int row = blockIdx.x;                       // b * Ti + t
  const int b = row / Ti, t = row - b * Ti;
  const int gpr = Zc >> 3;
  const TD* db = dy + (int64_t)b * To * Fo * 2;
  for (int i = blockIdx.y * PT + threadIdx.x; i < Fi * gpr; i += gridDim.y * PT) {
    const int f = i / gpr, g = i - f * gpr;
    float o[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = g * 4 + e;
      float2 v = make_float2(0.f, 0.f);
      if (j < taps.n) {
        const int tt = t - taps.dt[j], ff = f * sf - taps.df[j];
        if (tt >= 0 && tt < To && ff >= 0 && ff < Fo) v = ld_pair(db + ((int64_t)tt * Fo + ff) * 2);
      }
      o[2 * e] = v.x;
      o[2 * e + 1] = v.y;
    }
    st8(dz + ((int64_t)row * Fi + f) * Zc + g * 8, o);
  }
}

// ---- tiled tap-sum gradient (stride 1, Ti == To, Fi == Fo): one block owns a TT x TF patch of positions of one
// utterance; the patch of dy rows (N values each) is staged in shared memory, thread (t, f) assembles and stores its
// whole dz row (measured 0.61 -> 0.48 ms per launch; the same tiling of the FORWARD gather was slower than the direct
// kernel - its z rows are re-read from L1, not from HBM - and was dropped).
constexpr int TS_TT = 4, TS_TF = 64;        // 256 output positions per block

template <typename TD, typename TZ>
__global__ void __launch_bounds__(256) tapsum_bwd_tiled_kernel(const TD* __restrict__ dy, int B, int T, int F, int Zc,
                                                              TapList taps, int N, int tmin, int tspan, int fmin, int fspan,
                                                              TZ* __restrict__ dz) {
  extern __shared__ __align__(16) uint8_t ts_smem[];
  float* ps = reinterpret_cast<float*>(ts_smem);                     // [(TT + tspan) x (TF + fspan)][N] fp32
  const int PW = TS_TF + fspan, PH = TS_TT + tspan;
  const int f_tiles = (F + TS_TF - 1) / TS_TF, t_tiles = (T + TS_TT - 1) / TS_TT;
  for (int tile = blockIdx.x; tile < B * t_tiles * f_tiles; tile += gridDim.x) {
    const int ft = tile % f_tiles;
    int r = tile / f_tiles;
    const int tt = r % t_tiles, b = r / t_tiles;
    const int t0 = tt * TS_TT, f0 = ft * TS_TF;
    __syncthreads();
    // dz[(t,f)][j] = dy[(t - dt_j, f - df_j)]: the patch spans t0 - tmax .. t0 + TT - 1 - tmin (same extents, mirrored)
    for (int i = threadIdx.x; i < PH * PW * N; i += 256) {
      const int n = i % N;
      int rr = i / N;
      const int pf = rr % PW, pt = rr / PW;
      const int ti = t0 - (tmin + tspan) + pt, fi = f0 - (fmin + fspan) + pf;
      float v = 0.f;
      if (ti >= 0 && ti < T && fi >= 0 && fi < F) v = ld_f(dy + (((int64_t)b * T + ti) * F + fi) * N + n);
      ps[i] = v;
    }
    __syncthreads();
    const int lt = threadIdx.x / TS_TF, lf = threadIdx.x % TS_TF;
    const int t = t0 + lt, f = f0 + lf;
    if (t < T && f < F) {
      TZ* zo = dz + (((int64_t)b * T + t) * F + f) * Zc;
      for (int c0 = 0; c0 < Zc; c0 += 8) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = c0 + e;
          const int j = c / N, n = c - j * N;
          float v = 0.f;
          if (j < taps.n) {
            const int pr = (lt - taps.dt[j] + tmin + tspan) * PW + (lf - taps.df[j] + fmin + fspan);
            v = ps[pr * N + n];
          }
          o[e] = v;
        }
        st8(zo + c0, o);
      }
    }
  }
}

struct TapExt { int tmin, tspan, fmin, fspan; };
static TapExt tap_extents(const TapList& tl) {
  int t0 = 1 << 30, t1 = -(1 << 30), f0 = 1 << 30, f1 = -(1 << 30);
  for (int j = 0; j < tl.n; ++j) {
    t0 = tl.dt[j] < t0 ? tl.dt[j] : t0; t1 = tl.dt[j] > t1 ? tl.dt[j] : t1;
    f0 = tl.df[j] < f0 ? tl.df[j] : f0; f1 = tl.df[j] > f1 ? tl.df[j] : f1;
  }
  return TapExt{t0, t1 - t0, f0, f1 - f0};
}

// one thread per (input position, 8-channel group) of dz
template <typename TD, typename TZ>
__global__ void __launch_bounds__(PT) tapsum_bwd_kernel(const TD* __restrict__ dy, int B, int Ti, int Fi, int To,
                                                        int Fo, int sf, int Zc, TapList taps, int N,
                                                        TZ* __restrict__ dz) {
  const unsigned gpr = (unsigned)Zc >> 3;
  const unsigned total = (unsigned)B * Ti * Fi * gpr;
  for (unsigned i = blockIdx.x * PT + threadIdx.x; i < total; i += gridDim.x * PT) {
    const unsigned m = i / gpr;
    const int c0 = (int)(i - m * gpr) * 8;
    const unsigned r = m / (unsigned)Fi;
    const int f = (int)(m - r * (unsigned)Fi);
    const int b = (int)(r / (unsigned)Ti), t = (int)(r - (unsigned)b * (unsigned)Ti);
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = c0 + e;
      const int j = c / N, n = c - j * N;
      float v = 0.f;
      if (j < taps.n) {
        const int tt = t - taps.dt[j], ff = f * sf - taps.df[j];
        if (tt >= 0 && tt < To && ff >= 0 && ff < Fo) v = ld_f(dy + (((int64_t)b * To + tt) * Fo + ff) * N + n);
      }
      o[e] = v;
    }
    st8(dz + (int64_t)m * Zc + c0, o);
  }
}

inline bool al16(const void* p) { return ((uintptr_t)p % 16) == 0; }

// 1x1, stride 1, same extents, every operand dense over (B,T,F)
bool is_pointwise_dense(const ClskdTapConv* d) {
  if (d->ntaps != 1 || d->dt[0] != 0 || d->df[0] != 0 || d->sf != 1 || d->To != d->Ti || d->Fo != d->Fi) return false;
  auto dense = [&](int64_t sB, int64_t sT, int64_t sF, int c) {
    return sF == c && sT == (int64_t)d->Fi * c && sB == (int64_t)d->Ti * d->Fi * c;
  };
  if (!dense(d->x0_sB, d->x0_sT, d->x0_sF, d->c0)) return false;
  if (d->c1 && !dense(d->x1_sB, d->x1_sT, d->x1_sF, d->c1)) return false;
  if (!dense(d->y_sB, d->y_sT, d->y_sF, d->N)) return false;
  return true;
}

int pw_grid(int64_t iters) {
  int64_t blocks = (iters + PT - 1) / PT;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

namespace pw {

// returns true when the launch was handled here
bool try_fwd(const ClskdTapConv* d, cudaStream_t st) {
  if (!is_pointwise_dense(d)) return false;
  const int Ctot = d->c0 + d->c1;
  const int64_t M = (int64_t)d->B * d->To * d->Fo;
  if (M < 1) return false;
  if (d->N <= 2 && Ctot % 8 == 0 && d->c0 % 8 == 0 && Ctot <= 256 && al16(d->x0) && (!d->c1 || al16(d->x1))) {
    int up2 = 1;
    while (up2 < Ctot / 8) up2 <<= 1;
    const int64_t warps = (M + (32 / up2) * RQ - 1) / ((32 / up2) * RQ);
    const int grid = pw_grid(warps * 32 / 2);
#define L(TX, TY)                                                                                              \
  pw_fwd_n2_kernel<TX, TY><<<grid, PT, 0, st>>>((const TX*)d->x0, (const TX*)d->x1, d->c0, d->c1,             \
                                                (const float*)d->w, d->bias, d->N, M, up2, (TY*)d->y, d->accumulate)
    if (d->x_dtype == CLSKD_F32 && d->y_dtype == CLSKD_F32) L(float, float);
    else if (d->x_dtype == CLSKD_F32) L(float, __nv_bfloat16);
    else if (d->y_dtype == CLSKD_F32) L(__nv_bfloat16, float);
    else L(__nv_bfloat16, __nv_bfloat16);
#undef L
    return true;
  }
  if (!d->c1 && (Ctot == 2 || Ctot == 4) && d->N % 8 == 0 && d->N >= 8 && d->N <= 2048 && !d->accumulate &&
      al16(d->y)) {
    const int rows_par = PT / (d->N / 8);
    if (rows_par < 1) return false;
    const int grid = pw_grid((M + 3) / 4 * (d->N / 8));
#define L(TX, TY, K)                                                                                   \
  pw_fwd_k_kernel<TX, TY, K><<<grid, PT, 0, st>>>((const TX*)d->x0, (const float*)d->w, d->bias, d->N, M, (TY*)d->y)
#define LK(TX, TY) do { if (Ctot == 2) L(TX, TY, 2); else L(TX, TY, 4); } while (0)
    if (d->x_dtype == CLSKD_F32 && d->y_dtype == CLSKD_F32) LK(float, float);
    else if (d->x_dtype == CLSKD_F32) LK(float, __nv_bfloat16);
    else if (d->y_dtype == CLSKD_F32) LK(__nv_bfloat16, float);
    else LK(__nv_bfloat16, __nv_bfloat16);
#undef LK
#undef L
    return true;
  }
  return false;
}

// dW (fp32 [Ctot][N]) must already be zeroed (or hold the value to accumulate onto)
bool try_wgrad(const ClskdTapConv* d, cudaStream_t st) {
  if (!is_pointwise_dense(d)) return false;
  const int Ctot = d->c0 + d->c1;
  const int64_t M = (int64_t)d->B * d->To * d->Fo;
  if (M < 1024) return false;
  float* dw = reinterpret_cast<float*>(const_cast<void*>(d->w));
  if (d->N <= 2 && Ctot % 8 == 0 && d->c0 % 8 == 0 && Ctot <= 256 && al16(d->x0) && (!d->c1 || al16(d->x1))) {
    int up2 = 1;
    while (up2 < Ctot / 8) up2 <<= 1;
    const int grid = sm_count() * 4;
    const size_t sh = sizeof(float) * (size_t)Ctot * 2;
#define L(TX, TY)                                                                                     \
  pw_wgrad_n2_kernel<TX, TY><<<grid, PT, sh, st>>>((const TX*)d->x0, (const TX*)d->x1, d->c0, d->c1,  \
                                                   (const TY*)d->y, d->N, M, up2, dw)
    if (d->x_dtype == CLSKD_F32 && d->y_dtype == CLSKD_F32) L(float, float);
    else if (d->x_dtype == CLSKD_F32) L(float, __nv_bfloat16);
    else if (d->y_dtype == CLSKD_F32) L(__nv_bfloat16, float);
    else L(__nv_bfloat16, __nv_bfloat16);
#undef L
    return true;
  }
  if (!d->c1 && (Ctot == 2 || Ctot == 4) && d->N % 8 == 0 && d->N >= 8 && d->N <= 2048 && al16(d->y)) {
    const int grid = sm_count() * 4;
    const size_t sh = sizeof(float) * (size_t)Ctot * d->N;
#define L(TX, TY, K) pw_wgrad_k_kernel<TX, TY, K><<<grid, PT, sh, st>>>((const TX*)d->x0, (const TY*)d->y, d->N, M, dw)
#define LK(TX, TY) do { if (Ctot == 2) L(TX, TY, 2); else L(TX, TY, 4); } while (0)
    if (d->x_dtype == CLSKD_F32 && d->y_dtype == CLSKD_F32) LK(float, float);
    else if (d->x_dtype == CLSKD_F32) LK(float, __nv_bfloat16);
    else if (d->y_dtype == CLSKD_F32) LK(__nv_bfloat16, float);
    else LK(__nv_bfloat16, __nv_bfloat16);
#undef LK
#undef L
    return true;
  }
  return false;
}

}  // namespace pw
}  // namespace clskd

using namespace clskd;

extern "C" int clskd_tapsum_fwd(const void* z, int z_dtype, int B, int Ti, int Fi, int To, int Fo, int sf, int Zc,
                                int ntaps, const int32_t* dt_host, const int32_t* df_host, int N,
                                const float* bias, void* y, int y_dtype, void* stream) {
  CLSKD_CHECK_ARG(z && y && dt_host && df_host, "clskd_tapsum_fwd: null pointer");
  CLSKD_CHECK_ARG(ntaps >= 1 && ntaps <= CLSKD_MAX_TAPS && N >= 1 && N <= 4 && ntaps * N <= Zc && sf >= 1,
                  "clskd_tapsum_fwd: bad tap / channel counts");
  const int64_t M = (int64_t)B * To * Fo;
  CLSKD_CHECK_ARG(M < 4000000000LL && (int64_t)B * Ti * Fi * (Zc / 8 + 1) < 4000000000LL,
                  "clskd_tapsum_fwd: tensor too large");
  if (M == 0) return CLSKD_OK;
  TapList tl;
  tl.n = ntaps;
  for (int j = 0; j < ntaps; ++j) { tl.dt[j] = dt_host[j]; tl.df[j] = df_host[j]; }
  const int grid = pw_grid(M);
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 2 && (sf == 1 || sf == 2) && Zc % 2 == 0 && (uintptr_t)z % 8 == 0 && (uintptr_t)y % 8 == 0 &&
      (int64_t)B * To < 2147483647LL && ntaps <= 10) {
    dim3 g2(B * To, cdiv(Fo, PT));
#define LN(TZ, TY)                                                                                                   \
  do {                                                                                                               \
    if (ntaps <= 6) tapsum_fwd_n2_kernel<TZ, TY, 6><<<g2, PT, 0, st>>>((const TZ*)z, Ti, Fi, To, Fo, sf, Zc, tl, bias, (TY*)y); \
    else tapsum_fwd_n2_kernel<TZ, TY, 10><<<g2, PT, 0, st>>>((const TZ*)z, Ti, Fi, To, Fo, sf, Zc, tl, bias, (TY*)y);  \
  } while (0)
    if (z_dtype == CLSKD_F32 && y_dtype == CLSKD_F32) LN(float, float);
    else if (z_dtype == CLSKD_F32) LN(float, __nv_bfloat16);
    else if (y_dtype == CLSKD_F32) LN(__nv_bfloat16, float);
    else LN(__nv_bfloat16, __nv_bfloat16);
#undef LN
    CLSKD_CHECK_LAUNCH("clskd_tapsum_fwd(n2)");
    return CLSKD_OK;
  }
#define L(TZ, TY) tapsum_fwd_kernel<TZ, TY><<<grid, PT, 0, st>>>((const TZ*)z, B, Ti, Fi, To, Fo, sf, Zc, tl, N, bias, (TY*)y)
  if (z_dtype == CLSKD_F32 && y_dtype == CLSKD_F32) L(float, float);
  else if (z_dtype == CLSKD_F32) L(float, __nv_bfloat16);
  else if (y_dtype == CLSKD_F32) L(__nv_bfloat16, float);
  else L(__nv_bfloat16, __nv_bfloat16);
#undef L
  CLSKD_CHECK_LAUNCH("clskd_tapsum_fwd");
  return CLSKD_OK;
}

extern "C" int clskd_tapsum_bwd(const void* dy, int dy_dtype, int B, int Ti, int Fi, int To, int Fo, int sf, int Zc,
                                int ntaps, const int32_t* dt_host, const int32_t* df_host, int N, void* dz,
                                int dz_dtype, void* stream) {
  CLSKD_CHECK_ARG(dy && dz && dt_host && df_host, "clskd_tapsum_bwd: null pointer");
  CLSKD_CHECK_ARG(ntaps >= 1 && ntaps <= CLSKD_MAX_TAPS && N >= 1 && N <= 4 && ntaps * N <= Zc && Zc % 8 == 0 && sf >= 1,
                  "clskd_tapsum_bwd: bad tap / channel counts");
  CLSKD_CHECK_ARG(((uintptr_t)dz % 16) == 0, "clskd_tapsum_bwd: dz must be 16-byte aligned");
  const int64_t M = (int64_t)B * Ti * Fi;
  CLSKD_CHECK_ARG(M * (Zc / 8 + 1) < 4000000000LL, "clskd_tapsum_bwd: tensor too large");
  if (M == 0) return CLSKD_OK;
  TapList tl;
  tl.n = ntaps;
  for (int j = 0; j < ntaps; ++j) { tl.dt[j] = dt_host[j]; tl.df[j] = df_host[j]; }
  const int grid = pw_grid(M * (Zc / 8));
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 2 && (uintptr_t)dy % 8 == 0 && (int64_t)B * Ti < 2147483647LL) {
    dim3 g2(B * Ti, cdiv((int64_t)Fi * (Zc / 8), PT));
#define LN(TD, TZ) tapsum_bwd_n2_kernel<TD, TZ><<<g2, PT, 0, st>>>((const TD*)dy, Ti, Fi, To, Fo, sf, Zc, tl, (TZ*)dz)
    if (dy_dtype == CLSKD_F32 && dz_dtype == CLSKD_F32) LN(float, float);
    else if (dy_dtype == CLSKD_F32) LN(float, __nv_bfloat16);
    else if (dz_dtype == CLSKD_F32) LN(__nv_bfloat16, float);
    else LN(__nv_bfloat16, __nv_bfloat16);
#undef LN
    CLSKD_CHECK_LAUNCH("clskd_tapsum_bwd(n2)");
    return CLSKD_OK;
  }
  {
    const TapExt te = tap_extents(tl);
    const size_t sh = sizeof(float) * (size_t)(TS_TT + te.tspan) * (TS_TF + te.fspan) * N;
    if (sf == 1 && Ti == To && Fi == Fo && M >= 65536 && te.tspan <= 4 && te.fspan <= 8) {
      const int64_t tiles = (int64_t)B * cdiv(To, TS_TT) * cdiv(Fo, TS_TF);
      const int g2 = (int)(tiles < (int64_t)sm_count() * 8 ? tiles : (int64_t)sm_count() * 8);
#define LT(TD, TZ) tapsum_bwd_tiled_kernel<TD, TZ><<<g2, 256, sh, st>>>((const TD*)dy, B, To, Fo, Zc, tl, N, te.tmin, te.tspan, \
                                                                       te.fmin, te.fspan, (TZ*)dz)
      if (dy_dtype == CLSKD_F32 && dz_dtype == CLSKD_F32) LT(float, float);
      else if (dy_dtype == CLSKD_F32) LT(float, __nv_bfloat16);
      else if (dz_dtype == CLSKD_F32) LT(__nv_bfloat16, float);
      else LT(__nv_bfloat16, __nv_bfloat16);
#undef LT
      CLSKD_CHECK_LAUNCH("clskd_tapsum_bwd(tiled)");
      return CLSKD_OK;
    }
  }
#define L(TD, TZ) tapsum_bwd_kernel<TD, TZ><<<grid, PT, 0, st>>>((const TD*)dy, B, Ti, Fi, To, Fo, sf, Zc, tl, N, (TZ*)dz)
  if (dy_dtype == CLSKD_F32 && dz_dtype == CLSKD_F32) L(float, float);
  else if (dy_dtype == CLSKD_F32) L(float, __nv_bfloat16);
  else if (dz_dtype == CLSKD_F32) L(__nv_bfloat16, float);
  else L(__nv_bfloat16, __nv_bfloat16);
#undef L
  CLSKD_CHECK_LAUNCH("clskd_tapsum_bwd");
  return CLSKD_OK;
}
