// Tap-list implicit GEMM on CUDA cores (fp32 FMA, fp32 accumulate).
//
// This is the exact-fp32 policy of every convolution-like operator on the DCCRN path and the
// fallback for layers whose channel counts are below a UMMA tile (enc0: K=20, dec5: N=20, the
// ABF attention conv: N=2).  The bf16 tensor-core version of the same contraction lives in
// tapconv_umma.cu.  See clskd.h for the contraction definition.
#include "common.cuh"

namespace clskd {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct RowInfo {
  int b, t, f;
};

// ------------------------------------------------------------------------------------------
// forward: Y[M,N] = im2col(X)[M,Ktot] * W[Ktot,N]
// ------------------------------------------------------------------------------------------
template <typename TX, typename TY, bool VEC>
__global__ void __launch_bounds__(NT) tapconv_fwd_kernel(ClskdTapConv d) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ int rowB[BM], rowT[BM], rowF[BM];

  const int tid = threadIdx.x;
  const int64_t M = (int64_t)d.B * d.To * d.Fo;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int Ctot = d.c0 + d.c1;
  const int Ktot = d.ntaps * Ctot;
  const TX* x0 = reinterpret_cast<const TX*>(d.x0);
  const TX* x1 = reinterpret_cast<const TX*>(d.x1);
  const float* w = reinterpret_cast<const float*>(d.w);

  if (tid < BM) {
    int64_t m = m0 + tid;
    if (m < M) {
      int f = (int)(m % d.Fo);
      int64_t r = m / d.Fo;
      rowF[tid] = f;
      rowT[tid] = (int)(r % d.To);
      rowB[tid] = (int)(r / d.To);
    } else {
      rowB[tid] = -1;
      rowT[tid] = 0;
      rowF[tid] = 0;
    }
  }
  __syncthreads();

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int a_row = tid >> 2;        // 0..63
  const int a_k4 = (tid & 3) * 4;    // 0,4,8,12
  const int b_k = tid >> 4;          // 0..15
  const int b_n4 = (tid & 15) * 4;   // 0..60
  const int ty = tid >> 4, tx = tid & 15;

  const int rb = rowB[a_row], rt = rowT[a_row], rf = rowF[a_row];

  for (int kk = 0; kk < Ktot; kk += BK) {
    // ---- A tile (gather)
    float av[4] = {0.f, 0.f, 0.f, 0.f};
    if (rb >= 0) {
      if (VEC) {
        // Ctot % 16 == 0 and c0 % 16 == 0: the 16-wide chunk sits in one tap and one source
        int kg = kk + a_k4;
        int tap = kg / Ctot;
        int c = kg - tap * Ctot;
        int ti = rt + d.dt[tap], fi = rf * d.sf + d.df[tap];
        if (ti >= 0 && ti < d.Ti && fi >= 0 && fi < d.Fi) {
          float4 v;
          if (c < d.c0)
            v = ld4(x0 + (int64_t)rb * d.x0_sB + (int64_t)ti * d.x0_sT + (int64_t)fi * d.x0_sF + c);
          else
            v = ld4(x1 + (int64_t)rb * d.x1_sB + (int64_t)ti * d.x1_sT + (int64_t)fi * d.x1_sF +
                    (c - d.c0));
          av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int kg = kk + a_k4 + e;
          if (kg < Ktot) {
            int tap = kg / Ctot;
            int c = kg - tap * Ctot;
            int ti = rt + d.dt[tap], fi = rf * d.sf + d.df[tap];
            if (ti >= 0 && ti < d.Ti && fi >= 0 && fi < d.Fi) {
              if (c < d.c0)
                av[e] = ld_f(x0 + (int64_t)rb * d.x0_sB + (int64_t)ti * d.x0_sT +
                             (int64_t)fi * d.x0_sF + c);
              else
                av[e] = ld_f(x1 + (int64_t)rb * d.x1_sB + (int64_t)ti * d.x1_sT +
                             (int64_t)fi * d.x1_sF + (c - d.c0));
            }
          }
        }
      }
    }
    // ---- B tile
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    {
      int kg = kk + b_k;
      if (kg < Ktot) {
        int n = n0 + b_n4;
        const float* wp = w + (int64_t)kg * d.N + n;
        if (VEC && n + 3 < d.N) {
          float4 v = *reinterpret_cast<const float4*>(wp);
          bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (n + e < d.N) bv[e] = wp[e];
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) As[a_k4 + e][a_row] = av[e];
#pragma unroll
    for (int e = 0; e < 4; ++e) Bs[b_k][b_n4 + e] = bv[e];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float ar[4] = {a.x, a.y, a.z, a.w};
      float br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }

  // ---- epilogue
  TY* y = reinterpret_cast<TY*>(d.y);
  const int n = n0 + tx * 4;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (d.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < d.N) bias[j] = d.bias[n + j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = ty * 4 + i;
    int b = rowB[r];
    if (b < 0) continue;
    TY* yp = y + (int64_t)b * d.y_sB + (int64_t)rowT[r] * d.y_sT + (int64_t)rowF[r] * d.y_sF + n;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j < d.N) {
        float v = acc[i][j] + bias[j];
        if (d.accumulate) v += ld_f(yp + j);
        st_f(yp + j, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// wgrad: dW[Ktot,N] = im2col(X)^T [Ktot,M] * dY[M,N], split over M with fp32 atomics
// ------------------------------------------------------------------------------------------
constexpr int WK = 16;  // rows of M per step

template <typename TX, typename TY>
__global__ void __launch_bounds__(NT) tapconv_wgrad_kernel(ClskdTapConv d, int64_t rows_per_split) {
  __shared__ float As[WK][BM + 4];  // [m][kg]
  __shared__ float Bs[WK][BN + 4];  // [m][n]

  const int tid = threadIdx.x;
  const int64_t M = (int64_t)d.B * d.To * d.Fo;
  const int kg0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int Ctot = d.c0 + d.c1;
  const int Ktot = d.ntaps * Ctot;
  const TX* x0 = reinterpret_cast<const TX*>(d.x0);
  const TX* x1 = reinterpret_cast<const TX*>(d.x1);
  const TY* dy = reinterpret_cast<const TY*>(d.y);
  float* dw = reinterpret_cast<float*>(const_cast<void*>(d.w));

  const int64_t mbeg = (int64_t)blockIdx.z * rows_per_split;
  const int64_t mend = min(M, mbeg + rows_per_split);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: row r = tid/16 (0..15), 4 consecutive columns starting at (tid%16)*4
  const int l_r = tid >> 4;
  const int l_c4 = (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;

  // decode this thread's 4 kg columns once
  int tapv[4], cv[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int kg = kg0 + l_c4 + e;
    if (kg < Ktot) {
      tapv[e] = kg / Ctot;
      cv[e] = kg - tapv[e] * Ctot;
    } else {
      tapv[e] = -1;
      cv[e] = 0;
    }
  }

  for (int64_t mm = mbeg; mm < mend; mm += WK) {
    int64_t m = mm + l_r;
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < mend) {
      int f = (int)(m % d.Fo);
      int64_t r = m / d.Fo;
      int t = (int)(r % d.To);
      int b = (int)(r / d.To);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (tapv[e] >= 0) {
          int ti = t + d.dt[tapv[e]], fi = f * d.sf + d.df[tapv[e]];
          if (ti >= 0 && ti < d.Ti && fi >= 0 && fi < d.Fi) {
            int c = cv[e];
            if (c < d.c0)
              av[e] = ld_f(x0 + (int64_t)b * d.x0_sB + (int64_t)ti * d.x0_sT +
                           (int64_t)fi * d.x0_sF + c);
            else
              av[e] = ld_f(x1 + (int64_t)b * d.x1_sB + (int64_t)ti * d.x1_sT +
                           (int64_t)fi * d.x1_sF + (c - d.c0));
          }
        }
      }
      const TY* yp = dy + (int64_t)b * d.y_sB + (int64_t)t * d.y_sT + (int64_t)f * d.y_sF;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int n = n0 + l_c4 + e;
        if (n < d.N) bv[e] = ld_f(yp + n);
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      As[l_r][l_c4 + e] = av[e];
      Bs[l_r][l_c4 + e] = bv[e];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float ar[4] = {a.x, a.y, a.z, a.w};
      float br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int kg = kg0 + ty * 4 + i;
    if (kg >= Ktot) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < d.N) atomicAdd(dw + (int64_t)kg * d.N + n, acc[i][j]);
    }
  }
}


// ------------------------------------------------------------------------------------------
// N <= 2 outputs per row (the mask-producing last decoder layer, ABF's 2-logit attention conv and
// the data gradient of a 2-channel input): a GEMV per row, HBM-bound.  A warp reads a row's
// K = ntaps*Ctot inputs with 16-byte loads (lane -> fixed 8-channel unit, so the lane's weights
// live in registers), reduces with shuffles and writes the N outputs.  Rows whose K fits 16 lanes
// or fewer share a warp.
// ------------------------------------------------------------------------------------------
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

// (b, t, f) of output row m; 32-bit divisions when the row count allows (checked on the host)
__device__ __forceinline__ void row_decode(int64_t m, int To, int Fo, int& b, int& t, int& f) {
  if (m < 0x7fffffffLL) {
    const unsigned mm = (unsigned)m;
    const unsigned r = mm / (unsigned)Fo;
    f = (int)(mm - r * (unsigned)Fo);
    b = (int)(r / (unsigned)To);
    t = (int)(r - (unsigned)b * (unsigned)To);
  } else {
    f = (int)(m % Fo);
    const int64_t r = m / Fo;
    t = (int)(r % To);
    b = (int)(r / To);
  }
}

struct N2Geom {
  int U;      // 8-channel units per row = ntaps * Ctot / 8
  int cpt;    // units per tap = Ctot / 8
  int up2;    // lanes cooperating on one row (power of two, <= 32)
};

template <typename TX, int UPL>
struct N2Units {
  int tap[UPL];
  int dt[UPL], df[UPL];
  int64_t coff[UPL];   // channel offset inside the source
  bool src1[UPL], ok[UPL];
  __device__ __forceinline__ void init(const ClskdTapConv& d, const N2Geom& g, int ul) {
#pragma unroll
    for (int i = 0; i < UPL; ++i) {
      int u = ul + i * g.up2;
      ok[i] = u < g.U;
      int uu = ok[i] ? u : 0;
      tap[i] = uu / g.cpt;
      int c = (uu - tap[i] * g.cpt) * 8;
      src1[i] = c >= d.c0;
      coff[i] = src1[i] ? c - d.c0 : c;
      dt[i] = d.dt[tap[i]];
      df[i] = d.df[tap[i]];
    }
  }
  // loads the unit's 8 inputs for output row (b,t,f); returns false (zeros) when out of range
  __device__ __forceinline__ bool load(const ClskdTapConv& d, int i, int b, int t, int f, float* x) const {
    int ti = t + dt[i], fi = f * d.sf + df[i];
    if (!ok[i] || ti < 0 || ti >= d.Ti || fi < 0 || fi >= d.Fi) return false;
    if (src1[i])
      ld8(reinterpret_cast<const TX*>(d.x1) + (int64_t)b * d.x1_sB + (int64_t)ti * d.x1_sT + (int64_t)fi * d.x1_sF + coff[i], x);
    else
      ld8(reinterpret_cast<const TX*>(d.x0) + (int64_t)b * d.x0_sB + (int64_t)ti * d.x0_sT + (int64_t)fi * d.x0_sF + coff[i], x);
    return true;
  }
};

template <typename TX, typename TY, int UPL>
__global__ void __launch_bounds__(256) tapconv_fwd_n2_kernel(ClskdTapConv d, N2Geom g) {
  const int lane = threadIdx.x & 31;
  const int ul = lane & (g.up2 - 1), sub = lane / g.up2, rpw = 32 / g.up2;
  const int Ctot = d.c0 + d.c1;
  const float* w = reinterpret_cast<const float*>(d.w);
  N2Units<TX, UPL> un;
  un.init(d, g, ul);
  float wr[UPL][8][2];
#pragma unroll
  for (int i = 0; i < UPL; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        int64_t k = (int64_t)un.tap[i] * Ctot + (un.src1[i] ? d.c0 : 0) + un.coff[i] + e;
        wr[i][e][n] = (un.ok[i] && n < d.N) ? w[k * d.N + n] : 0.f;
      }
  const float b0 = d.bias ? d.bias[0] : 0.f;
  const float b1 = (d.bias && d.N > 1) ? d.bias[1] : 0.f;
  const int64_t M = (int64_t)d.B * d.To * d.Fo;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  TY* y = reinterpret_cast<TY*>(d.y);
  constexpr int RB = 1;   // rows per lane group and iteration (batching rows measured slower: the row index math dominates)
  for (int64_t m0 = warp0 * rpw * RB; m0 < M; m0 += nwarps * rpw * RB) {
    float x[RB][UPL][8];
    bool ok[RB][UPL];
    int rb[RB], rt[RB], rf[RB];
    bool live[RB];
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      const int64_t m = m0 + (int64_t)q * rpw + sub;
      live[q] = m < M;
      row_decode(live[q] ? m : 0, d.To, d.Fo, rb[q], rt[q], rf[q]);
#pragma unroll
      for (int i = 0; i < UPL; ++i) ok[q][i] = live[q] && un.load(d, i, rb[q], rt[q], rf[q], x[q][i]);
    }
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int i = 0; i < UPL; ++i) {
        if (ok[q][i]) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            a0 = fmaf(x[q][i][e], wr[i][e][0], a0);
            a1 = fmaf(x[q][i][e], wr[i][e][1], a1);
          }
        }
      }
      for (int o = g.up2 >> 1; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      }
      if (live[q] && ul == 0) {
        TY* yp = y + (int64_t)rb[q] * d.y_sB + (int64_t)rt[q] * d.y_sT + (int64_t)rf[q] * d.y_sF;
        float v0 = a0 + b0, v1 = a1 + b1;
        if (d.accumulate) {
          v0 += ld_f(yp);
          if (d.N > 1) v1 += ld_f(yp + 1);
        }
        st_f(yp, v0);
        if (d.N > 1) st_f(yp + 1, v1);
      }
    }
  }
}

// N <= 2 with many taps (ABF's 3x3 conv onto the 2-channel mask map: K = 9*128): lanes own the
// 8-channel units of ONE tap (cpt <= 128 -> at most 4 units per lane), the tap loop is in the
// kernel and the weights sit in shared memory transposed to [tap][e][unit] so that a warp's loads
// are conflict free.
template <typename TX, typename TY, int TB>
__global__ void __launch_bounds__(256) tapconv_fwd_n2_taps_kernel(ClskdTapConv d, int cpt, int up2) {
  extern __shared__ float2 w2[];   // [ntaps][8][cpt]
  const int Ctot = d.c0 + d.c1;
  const float* w = reinterpret_cast<const float*>(d.w);
  for (int i = threadIdx.x; i < d.ntaps * 8 * cpt; i += blockDim.x) {
    const int cu = i % cpt, e = (i / cpt) & 7, tap = i / (8 * cpt);
    const int64_t k = (int64_t)tap * Ctot + cu * 8 + e;
    w2[i] = make_float2(w[k * d.N], d.N > 1 ? w[k * d.N + 1] : 0.f);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int ul = lane & (up2 - 1), sub = lane / up2, rpw = 32 / up2;   // up2 lanes share a row
  const float b0 = d.bias ? d.bias[0] : 0.f;
  const float b1 = (d.bias && d.N > 1) ? d.bias[1] : 0.f;
  const int64_t M = (int64_t)d.B * d.To * d.Fo;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const TX* x0 = reinterpret_cast<const TX*>(d.x0);
  const TX* x1 = reinterpret_cast<const TX*>(d.x1);
  TY* y = reinterpret_cast<TY*>(d.y);
  for (int64_t m0 = warp0 * rpw; m0 < M; m0 += nwarps * rpw) {
    const int64_t m = m0 + sub;
    const bool live = m < M;
    int b, t, f;
    row_decode(live ? m : 0, d.To, d.Fo, b, t, f);
    float a0 = 0.f, a1 = 0.f;
    for (int cu = ul; cu < cpt; cu += up2) {          // the lane's units (same for every tap)
      const int c = cu * 8;
      const bool s1 = c >= d.c0;
      for (int tap0 = 0; tap0 < d.ntaps; tap0 += TB) {
        float xv[TB][8];
        bool ok[TB];
#pragma unroll
        for (int q = 0; q < TB; ++q) {                // TB taps' loads in flight together
          const int tap = tap0 + q;
          ok[q] = false;
          if (live && tap < d.ntaps) {
            const int ti = t + d.dt[tap], fi = f * d.sf + d.df[tap];
            if (ti >= 0 && ti < d.Ti && fi >= 0 && fi < d.Fi) {
              ok[q] = true;
              if (s1) ld8(x1 + (int64_t)b * d.x1_sB + (int64_t)ti * d.x1_sT + (int64_t)fi * d.x1_sF + (c - d.c0), xv[q]);
              else ld8(x0 + (int64_t)b * d.x0_sB + (int64_t)ti * d.x0_sT + (int64_t)fi * d.x0_sF + c, xv[q]);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < TB; ++q) {
          if (ok[q]) {
            const float2* wp = w2 + (size_t)(tap0 + q) * 8 * cpt + cu;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float2 ww = wp[e * cpt];
              a0 = fmaf(xv[q][e], ww.x, a0);
              a1 = fmaf(xv[q][e], ww.y, a1);
            }
          }
        }
      }
    }
    for (int o = up2 >> 1; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    if (live && ul == 0) {
      TY* yp = y + (int64_t)b * d.y_sB + (int64_t)t * d.y_sT + (int64_t)f * d.y_sF;
      float v0 = a0 + b0, v1 = a1 + b1;
      if (d.accumulate) {
        v0 += ld_f(yp);
        if (d.N > 1) v1 += ld_f(yp + 1);
      }
      st_f(yp, v0);
      if (d.N > 1) st_f(yp + 1, v1);
    }
  }
}

// dW[k][n] for N <= 2: every lane accumulates its own 8-channel units over the rows its warp
// visits (no cross-lane traffic in the loop), then shared-memory and global fp32 reductions.
template <typename TX, typename TY, int UPL>
__global__ void __launch_bounds__(256) tapconv_wgrad_n2_kernel(ClskdTapConv d, N2Geom g) {
  extern __shared__ float red[];   // [U*8*2]
  const int lane = threadIdx.x & 31;
  const int ul = lane & (g.up2 - 1), sub = lane / g.up2, rpw = 32 / g.up2;
  const int Ctot = d.c0 + d.c1;
  N2Units<TX, UPL> un;
  un.init(d, g, ul);
  float acc[UPL][8][2];
#pragma unroll
  for (int i = 0; i < UPL; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[i][e][0] = acc[i][e][1] = 0.f;
  for (int i = threadIdx.x; i < g.U * 16; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int64_t M = (int64_t)d.B * d.To * d.Fo;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const TY* dy = reinterpret_cast<const TY*>(d.y);
  constexpr int RB = 1;   // rows per lane group and iteration (batching rows measured slower: the row index math dominates)
  for (int64_t m0 = warp0 * rpw * RB; m0 < M; m0 += nwarps * rpw * RB) {
    float x[RB][UPL][8], g0[RB], g1[RB];
    bool ok[RB][UPL];
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      const int64_t m = m0 + (int64_t)q * rpw + sub;
      const bool live = m < M;
      int b, t, f;
      row_decode(live ? m : 0, d.To, d.Fo, b, t, f);
      const TY* yp = dy + (int64_t)b * d.y_sB + (int64_t)t * d.y_sT + (int64_t)f * d.y_sF;
      g0[q] = live ? ld_f(yp) : 0.f;
      g1[q] = (live && d.N > 1) ? ld_f(yp + 1) : 0.f;
#pragma unroll
      for (int i = 0; i < UPL; ++i) ok[q][i] = live && un.load(d, i, b, t, f, x[q][i]);
    }
#pragma unroll
    for (int q = 0; q < RB; ++q)
#pragma unroll
      for (int i = 0; i < UPL; ++i) {
        if (ok[q][i]) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            acc[i][e][0] = fmaf(x[q][i][e], g0[q], acc[i][e][0]);
            acc[i][e][1] = fmaf(x[q][i][e], g1[q], acc[i][e][1]);
          }
        }
      }
  }
#pragma unroll
  for (int i = 0; i < UPL; ++i) {
    if (!un.ok[i]) continue;
    int u = ul + i * g.up2;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(&red[(u * 8 + e) * 2], acc[i][e][0]);
      atomicAdd(&red[(u * 8 + e) * 2 + 1], acc[i][e][1]);
    }
  }
  __syncthreads();
  float* dw = reinterpret_cast<float*>(const_cast<void*>(d.w));
  for (int i = threadIdx.x; i < g.U * 16; i += blockDim.x) {
    int n = i & 1, ke = i >> 1;             // ke = u*8+e = tap*Ctot + c (units are laid out in k order)
    if (n < d.N) atomicAdd(dw + (int64_t)ke * d.N + n, red[i]);
  }
}

bool n2_ok(const ClskdTapConv* d, N2Geom* g) {
  const int Ctot = d->c0 + d->c1;
  if (d->N > 2 || Ctot % 8 || d->c0 % 8) return false;
  const int xe = d->x_dtype == CLSKD_F32 ? 4 : 2;
  auto al = [&](const void* p, int64_t a, int64_t b, int64_t c) {
    return ((uintptr_t)p % 16 == 0) && (a * xe) % 16 == 0 && (b * xe) % 16 == 0 && (c * xe) % 16 == 0;
  };
  if (!al(d->x0, d->x0_sB, d->x0_sT, d->x0_sF)) return false;
  if (d->c1 && !al(d->x1, d->x1_sB, d->x1_sT, d->x1_sF)) return false;
  g->cpt = Ctot / 8;
  g->U = d->ntaps * g->cpt;
  if (g->cpt > 128) return false;          // more than 1024 input channels: generic kernel
  int up2 = 1;
  while (up2 < g->U && up2 < 32) up2 <<= 1;
  g->up2 = up2;
  return true;
}


// ------------------------------------------------------------------------------------------
// K = ntaps*Ctot <= 32 inputs per row (the 2-channel spectrum entering the first encoder layer
// and ABF's 1x1 conv on the 2-channel mask map, the data gradients of the 2-logit attention conv
// and of the mask layer): an outer-product expansion, bound by the OUTPUT write.  One thread makes
// 8 consecutive outputs of one row (one 16/32-byte store); the row's few inputs are broadcast loads.
// ------------------------------------------------------------------------------------------
constexpr int SK_MAXK = 32;

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

// input element (b, ti, fi, c) of the (x0 | x1) channel concat, zero outside the map
template <typename TX>
__device__ __forceinline__ float smallk_at(const ClskdTapConv& d, int c, int b, int ti, int fi) {
  if (ti < 0 || ti >= d.Ti || fi < 0 || fi >= d.Fi) return 0.f;
  if (c < d.c0)
    return ld_f(reinterpret_cast<const TX*>(d.x0) + (int64_t)b * d.x0_sB + (int64_t)ti * d.x0_sT + (int64_t)fi * d.x0_sF + c);
  return ld_f(reinterpret_cast<const TX*>(d.x1) + (int64_t)b * d.x1_sB + (int64_t)ti * d.x1_sT + (int64_t)fi * d.x1_sF + (c - d.c0));
}

template <typename TX>
__device__ __forceinline__ float smallk_x(const ClskdTapConv& d, int k, int Ctot, int b, int t, int f) {
  const int tap = k / Ctot, c = k - tap * Ctot;
  const int ti = t + d.dt[tap], fi = f * d.sf + d.df[tap];
  if (ti < 0 || ti >= d.Ti || fi < 0 || fi >= d.Fi) return 0.f;
  if (c < d.c0)
    return ld_f(reinterpret_cast<const TX*>(d.x0) + (int64_t)b * d.x0_sB + (int64_t)ti * d.x0_sT + (int64_t)fi * d.x0_sF + c);
  return ld_f(reinterpret_cast<const TX*>(d.x1) + (int64_t)b * d.x1_sB + (int64_t)ti * d.x1_sT + (int64_t)fi * d.x1_sF + (c - d.c0));
}

// Y[m][n] for K = ntaps*C <= 32 (first encoder layer: 2 input channels, 5x2 taps; 1x1 convs on the
// 2-channel mask map).  A thread owns 8*NV consecutive outputs of one row: the row's K inputs are
// loaded once per thread (tap bounds and addresses hoisted out of the channel loop), the weights are
// broadcast from shared memory, so the kernel is bound by the output store, not by instruction issue.
template <typename TX, typename TY, typename I, int NV>
__global__ void __launch_bounds__(256) tapconv_fwd_smallk_kernel(ClskdTapConv d, int Ktot) {
  extern __shared__ float wsm[];   // [Ktot][N] + bias[N]
  const int N = d.N, Ctot = d.c0 + d.c1;
  const float* w = reinterpret_cast<const float*>(d.w);
  for (int i = threadIdx.x; i < Ktot * N; i += blockDim.x) wsm[i] = w[i];
  for (int i = threadIdx.x; i < N; i += blockDim.x) wsm[Ktot * N + i] = d.bias ? d.bias[i] : 0.f;
  __syncthreads();
  const I tpr = N / (8 * NV);                   // threads per row
  const I M = (I)d.B * d.To * d.Fo;
  const I total = M * tpr;                      // I = unsigned when it fits 32 bits (cheap divisions)
  TY* y = reinterpret_cast<TY*>(d.y);
  const TX* x0 = reinterpret_cast<const TX*>(d.x0);
  const TX* x1 = reinterpret_cast<const TX*>(d.x1);
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int n0 = (int)(i % tpr) * (8 * NV);
    const I m = i / tpr;
    const int f = (int)(m % (I)d.Fo);
    const I r = m / (I)d.Fo;
    const int t = (int)(r % (I)d.To), b = (int)(r / (I)d.To);
    float acc[NV][8];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[v][e] = wsm[Ktot * N + n0 + v * 8 + e];
    const float* wk = wsm + n0;
    for (int tap = 0; tap < d.ntaps; ++tap) {
      const int ti = t + d.dt[tap], fi = f * d.sf + d.df[tap];
      const bool ok = ti >= 0 && ti < d.Ti && fi >= 0 && fi < d.Fi;
      const TX* p0 = x0 + (int64_t)b * d.x0_sB + (int64_t)ti * d.x0_sT + (int64_t)fi * d.x0_sF;
      const TX* p1 = d.c1 ? x1 + (int64_t)b * d.x1_sB + (int64_t)ti * d.x1_sT + (int64_t)fi * d.x1_sF : p0;
      for (int c = 0; c < Ctot; ++c, wk += N) {
        const float xv = ok ? (c < d.c0 ? ld_f(p0 + c) : ld_f(p1 + (c - d.c0))) : 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const float4 w0 = *reinterpret_cast<const float4*>(wk + v * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(wk + v * 8 + 4);
          acc[v][0] = fmaf(xv, w0.x, acc[v][0]); acc[v][1] = fmaf(xv, w0.y, acc[v][1]);
          acc[v][2] = fmaf(xv, w0.z, acc[v][2]); acc[v][3] = fmaf(xv, w0.w, acc[v][3]);
          acc[v][4] = fmaf(xv, w1.x, acc[v][4]); acc[v][5] = fmaf(xv, w1.y, acc[v][5]);
          acc[v][6] = fmaf(xv, w1.z, acc[v][6]); acc[v][7] = fmaf(xv, w1.w, acc[v][7]);
        }
      }
    }
    TY* yo = y + (int64_t)b * d.y_sB + (int64_t)t * d.y_sT + (int64_t)f * d.y_sF + n0;
#pragma unroll
    for (int v = 0; v < NV; ++v) st8(yo + v * 8, acc[v]);
  }
}

// dW[k][n] for K <= 32: a thread owns 8 consecutive n and up to 8 k of the rows it visits (one
// 16-byte read of dY per row, inputs broadcast), then shared-memory and global fp32 reductions.
template <typename TX, typename TY>
__global__ void __launch_bounds__(256) tapconv_wgrad_smallk_kernel(ClskdTapConv d, int Ktot, int kchunks) {
  extern __shared__ float red[];   // [Ktot][N]
  const int N = d.N, Ctot = d.c0 + d.c1;
  for (int i = threadIdx.x; i < Ktot * N; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int tpr = N >> 3;
  const int per_row = tpr * kchunks;            // threads cooperating on one row
  const int rows_par = blockDim.x / per_row;    // rows in flight per block
  const int sub = threadIdx.x / per_row;
  const int rem = threadIdx.x - sub * per_row;
  const int kc = rem / tpr, n0 = (rem - kc * tpr) * 8;
  const int k0 = kc * 8, kcnt = min(8, Ktot - k0);
  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[a][e] = 0.f;
  const int64_t M = (int64_t)d.B * d.To * d.Fo;
  const TY* dy = reinterpret_cast<const TY*>(d.y);
  // (tap shift, channel) of the up-to-8 contraction indices this thread owns: decoded once
  int dta[8], dfa[8], ca[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int k = k0 + a;
    const int tap = a < kcnt ? k / Ctot : 0;
    dta[a] = d.dt[tap];
    dfa[a] = d.df[tap];
    ca[a] = a < kcnt ? k - tap * Ctot : -1;
  }
  if (sub < rows_par) {
    constexpr int RB = 2;
    for (int64_t m0 = (int64_t)blockIdx.x * rows_par * RB + sub; m0 < M; m0 += (int64_t)gridDim.x * rows_par * RB) {
      float g[RB][8], xv[RB][8];
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        const int64_t m = m0 + (int64_t)q * rows_par;
        const bool live = m < M;
        int b, t, f;
        row_decode(live ? m : 0, d.To, d.Fo, b, t, f);
        if (live) {
          ld8(dy + (int64_t)b * d.y_sB + (int64_t)t * d.y_sT + (int64_t)f * d.y_sF + n0, g[q]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) g[q][e] = 0.f;
        }
#pragma unroll
        for (int a = 0; a < 8; ++a)
          xv[q][a] = (live && ca[a] >= 0) ? smallk_at<TX>(d, ca[a], b, t + dta[a], f * d.sf + dfa[a]) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < RB; ++q)
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[a][e] = fmaf(xv[q][a], g[q][e], acc[a][e]);
    }
#pragma unroll
    for (int a = 0; a < 8; ++a)
      if (a < kcnt)
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(&red[(k0 + a) * N + n0 + e], acc[a][e]);
  }
  __syncthreads();
  float* dw = reinterpret_cast<float*>(const_cast<void*>(d.w));
  for (int i = threadIdx.x; i < Ktot * N; i += blockDim.x) atomicAdd(dw + i, red[i]);
}


// Weight gradient of a convolution with TWO input channels (first encoder layer: the interleaved spectrum, 5x2 taps,
// frequency stride 2).  One block walks (b, t) lines: the <= 2 input time rows a line's taps touch are staged in shared
// memory (zero padded in f, zero for t outside the map), thread (slot, n) owns output channel n of the output
// frequencies slot, slot + S, ...: one coalesced dY read and NT broadcast float2 reads per (fo, n) feed 2*NT register
// accumulators - no per-element bounds checks or address arithmetic (the generic small-K kernel spent its time there:
// 184 GB/s).  Cross-slot and cross-block reductions in shared / global fp32 atomics.
template <typename TY, int NT>
__global__ void __launch_bounds__(256) tapconv_wgrad_c2_kernel(ClskdTapConv d, int tmin, int tspan, int fmin, int fspan) {
  extern __shared__ float xs[];                  // [tspan][Fi + fspan][2] then red[NT*2][N]
  const int N = d.N;
  const int FW = d.Fi + fspan;                   // padded input line
  float* red = xs + (size_t)tspan * FW * 2;
  for (int i = threadIdx.x; i < NT * 2 * N; i += blockDim.x) red[i] = 0.f;
  const int slots = blockDim.x / N;
  const int n = threadIdx.x % N, slot = threadIdx.x / N;
  int off[NT];                                    // float2 index of tap j for output frequency 0
#pragma unroll
  for (int j = 0; j < NT; ++j) off[j] = (d.dt[j] - tmin) * FW + (d.df[j] - fmin);
  float2 acc[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j] = make_float2(0.f, 0.f);
  const float* x = reinterpret_cast<const float*>(d.x0);
  const TY* dy = reinterpret_cast<const TY*>(d.y);
  const int64_t lines = (int64_t)d.B * d.To;
  for (int64_t ln = blockIdx.x; ln < lines; ln += gridDim.x) {
    const int b = (int)(ln / d.To), t = (int)(ln - (int64_t)b * d.To);
    __syncthreads();
    for (int i = threadIdx.x; i < tspan * FW; i += blockDim.x) {
      const int tr = i / FW, fp = i - tr * FW;
      const int ti = t + tmin + tr, fi = fp + fmin;
      float2 v = make_float2(0.f, 0.f);
      if (ti >= 0 && ti < d.Ti && fi >= 0 && fi < d.Fi)
        v = *reinterpret_cast<const float2*>(x + (int64_t)b * d.x0_sB + (int64_t)ti * d.x0_sT + (int64_t)fi * d.x0_sF);
      reinterpret_cast<float2*>(xs)[i] = v;
    }
    __syncthreads();
    if (slot < slots) {
      const TY* dyl = dy + (int64_t)b * d.y_sB + (int64_t)t * d.y_sT + n;
      const float2* xl = reinterpret_cast<const float2*>(xs);
      for (int fo = slot; fo < d.Fo; fo += slots) {
        const float g = ld_f(dyl + (int64_t)fo * d.y_sF);
        const int base = fo * d.sf;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const float2 xv = xl[base + off[j]];
          acc[j].x = fmaf(xv.x, g, acc[j].x);
          acc[j].y = fmaf(xv.y, g, acc[j].y);
        }
      }
    }
  }
  if (slot < slots) {
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      atomicAdd(&red[(j * 2) * N + n], acc[j].x);
      atomicAdd(&red[(j * 2 + 1) * N + n], acc[j].y);
    }
  }
  __syncthreads();
  float* dw = reinterpret_cast<float*>(const_cast<void*>(d.w));
  for (int i = threadIdx.x; i < NT * 2 * N; i += blockDim.x) atomicAdd(dw + i, red[i]);
}

// fp32 2-channel source with channel pairs 8-byte aligned, <= 2..3 time rows per line, N a power of two <= 256
bool wgrad_c2_ok(const ClskdTapConv* d, int* tmin, int* tspan, int* fmin, int* fspan) {
  if (d->c0 != 2 || d->c1 != 0 || d->x_dtype != CLSKD_F32 || d->accumulate) return false;
  if (d->ntaps != 10 && d->ntaps != 9 && d->ntaps != 1) return false;
  if (d->N < 8 || d->N > 256 || (d->N & (d->N - 1))) return false;
  if (((uintptr_t)d->x0 % 8) || (d->x0_sB % 2) || (d->x0_sT % 2) || (d->x0_sF % 2)) return false;
  int t0 = 1 << 30, t1 = -(1 << 30), f0 = 1 << 30, f1 = -(1 << 30);
  for (int j = 0; j < d->ntaps; ++j) {
    t0 = d->dt[j] < t0 ? d->dt[j] : t0; t1 = d->dt[j] > t1 ? d->dt[j] : t1;
    f0 = d->df[j] < f0 ? d->df[j] : f0; f1 = d->df[j] > f1 ? d->df[j] : f1;
  }
  *tmin = t0; *tspan = t1 - t0 + 1; *fmin = f0;
  // the last output frequency reads input (Fo-1)*sf + f1: pad the staged line so that index stays inside it
  const int need = (d->Fo - 1) * d->sf + f1 - f0 + 1;
  *fspan = need > d->Fi ? need - d->Fi : 0;
  if (*fspan < f1 - f0) *fspan = f1 - f0;
  const size_t sh = sizeof(float) * ((size_t)(*tspan) * (d->Fi + *fspan) * 2 + (size_t)d->ntaps * 2 * d->N);
  return *tspan <= 3 && sh <= 96 * 1024;
}

bool smallk_ok(const ClskdTapConv* d, int* Ktot) {
  *Ktot = d->ntaps * (d->c0 + d->c1);
  if (*Ktot > SK_MAXK || d->N % 8 || d->N < 8 || d->N > 512 || d->accumulate) return false;
  if (((size_t)*Ktot * d->N + d->N) * sizeof(float) > 46 * 1024) return false;   // weights staged in static-limit smem
  const int ye = d->y_dtype == CLSKD_F32 ? 4 : 2;
  return ((uintptr_t)d->y % 16 == 0) && (d->y_sB * ye) % 16 == 0 && (d->y_sT * ye) % 16 == 0 &&
         (d->y_sF * ye) % 16 == 0;
}

int check_desc(const ClskdTapConv* d, const char* who) {
  CLSKD_CHECK_ARG(d != nullptr, "%s: null descriptor", who);
  CLSKD_CHECK_ARG(d->x0 && d->w && d->y, "%s: null tensor pointer", who);
  CLSKD_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= CLSKD_MAX_TAPS, "%s: ntaps=%d out of range", who,
                  d->ntaps);
  CLSKD_CHECK_ARG(d->c0 >= 1 && d->c1 >= 0 && (d->c1 == 0 || d->x1), "%s: bad channel split", who);
  CLSKD_CHECK_ARG(d->B >= 0 && d->To >= 0 && d->Fo >= 0 && d->N >= 1 && d->sf >= 1,
                  "%s: bad extents", who);
  CLSKD_CHECK_ARG((d->x_dtype == CLSKD_F32 || d->x_dtype == CLSKD_BF16) &&
                      (d->y_dtype == CLSKD_F32 || d->y_dtype == CLSKD_BF16),
                  "%s: bad dtype tag", who);
  return CLSKD_OK;
}

bool vec_ok(const ClskdTapConv* d) {
  int Ctot = d->c0 + d->c1;
  if (Ctot % 16 || d->c0 % 16 || d->N % 4) return false;
  int xe = d->x_dtype == CLSKD_F32 ? 4 : 2;
  // 4-element vector loads: 16 B (fp32) / 8 B (bf16) alignment of bases and strides
  auto al = [&](const void* p, int64_t a, int64_t b, int64_t c) {
    return ((uintptr_t)p % (4 * xe) == 0) && a % 4 == 0 && b % 4 == 0 && c % 4 == 0;
  };
  if (!al(d->x0, d->x0_sB, d->x0_sT, d->x0_sF)) return false;
  if (d->c1 && !al(d->x1, d->x1_sB, d->x1_sT, d->x1_sF)) return false;
  if ((uintptr_t)d->w % 16) return false;
  return true;
}

}  // namespace

}  // namespace clskd

using namespace clskd;

extern "C" int clskd_tapconv_fwd(const ClskdTapConv* d, void* stream) {
  int rc = check_desc(d, "clskd_tapconv_fwd");
  if (rc) return rc;
  CLSKD_CHECK_ARG(!(d->accumulate && d->y_dtype != CLSKD_F32),
                  "clskd_tapconv_fwd: accumulate needs fp32 output");
  if (d->ep_scale || d->ep_shift || d->ep_slope || d->stats_sum || d->stats_sumsq) {
    set_error("clskd_tapconv_fwd: the fused epilogue (ep_*, stats_*) exists on the tcgen05 kernel only");
    return CLSKD_ERR_UNSUPPORTED;
  }
  int64_t M = (int64_t)d->B * d->To * d->Fo;
  if (M == 0) return CLSKD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (pw::try_fwd(d, st)) {
    CLSKD_CHECK_LAUNCH("clskd_tapconv_fwd(pointwise)");
    return CLSKD_OK;
  }
  if (c2mma::try_fwd(d, st)) {
    CLSKD_CHECK_LAUNCH("clskd_tapconv_fwd(c2 mma)");
    return CLSKD_OK;
  }
  N2Geom g2;
  if (n2_ok(d, &g2) && g2.U > 128) {
    int tup2 = 1;                                     // lanes per row: power of two covering cpt, <= 32
    while (tup2 < g2.cpt && tup2 < 32) tup2 <<= 1;
    int64_t blocks = (M / (32 / tup2) + 7) / 8 / 4;
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    if (blocks < 1) blocks = 1;
    const size_t sh = sizeof(float2) * (size_t)d->ntaps * 8 * g2.cpt;   // <= 16 taps * 8 * 128 * 8 B = 128 KB
    if (sh <= 46 * 1024) {
      if (d->x_dtype == CLSKD_F32 && d->y_dtype == CLSKD_F32)
        tapconv_fwd_n2_taps_kernel<float, float, 3><<<(unsigned)blocks, 256, sh, st>>>(*d, g2.cpt, tup2);
      else if (d->x_dtype == CLSKD_F32)
        tapconv_fwd_n2_taps_kernel<float, __nv_bfloat16, 3><<<(unsigned)blocks, 256, sh, st>>>(*d, g2.cpt, tup2);
      else if (d->y_dtype == CLSKD_F32)
        tapconv_fwd_n2_taps_kernel<__nv_bfloat16, float, 3><<<(unsigned)blocks, 256, sh, st>>>(*d, g2.cpt, tup2);
      else
        tapconv_fwd_n2_taps_kernel<__nv_bfloat16, __nv_bfloat16, 3><<<(unsigned)blocks, 256, sh, st>>>(*d, g2.cpt, tup2);
      CLSKD_CHECK_LAUNCH("clskd_tapconv_fwd(n2 taps)");
      return CLSKD_OK;
    }
  } else if (n2_ok(d, &g2)) {
    const int upl = cdiv(g2.U, g2.up2);   // 1, 2, 3 or 4 units per lane
    const int rbat = 1;
    const int64_t rows_per_warp_iter = (int64_t)(32 / g2.up2) * rbat;
    const int64_t warps_needed = (M + rows_per_warp_iter - 1) / rows_per_warp_iter;
    int64_t blocks = (warps_needed + 7) / 8 / 4;
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    if (blocks < 1) blocks = 1;
#define LAUNCH_N2(TX, TY)                                                                       \
  do {                                                                                          \
    if (upl <= 1) tapconv_fwd_n2_kernel<TX, TY, 1><<<(unsigned)blocks, 256, 0, st>>>(*d, g2);   \
    else if (upl == 2) tapconv_fwd_n2_kernel<TX, TY, 2><<<(unsigned)blocks, 256, 0, st>>>(*d, g2); \
    else tapconv_fwd_n2_kernel<TX, TY, 4><<<(unsigned)blocks, 256, 0, st>>>(*d, g2);            \
  } while (0)
    if (d->x_dtype == CLSKD_F32 && d->y_dtype == CLSKD_F32) LAUNCH_N2(float, float);
    else if (d->x_dtype == CLSKD_F32) LAUNCH_N2(float, __nv_bfloat16);
    else if (d->y_dtype == CLSKD_F32) LAUNCH_N2(__nv_bfloat16, float);
    else LAUNCH_N2(__nv_bfloat16, __nv_bfloat16);
#undef LAUNCH_N2
    CLSKD_CHECK_LAUNCH("clskd_tapconv_fwd(n2)");
    return CLSKD_OK;
  }
  int ktot_sk;
  if (smallk_ok(d, &ktot_sk)) {
    const int nv = d->N % 32 == 0 ? 4 : (d->N % 16 == 0 ? 2 : 1);
    const int64_t total = M * (d->N / (8 * nv));
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
    const size_t sh = sizeof(float) * ((size_t)ktot_sk * d->N + d->N);
    const bool small = total + (int64_t)blocks * 256 < 4000000000LL && M * (d->N / 8) < 4000000000LL;
#define LAUNCH_SK2(TX, TY, NV)                                                                               \
  do {                                                                                                       \
    if (small) tapconv_fwd_smallk_kernel<TX, TY, unsigned, NV><<<(unsigned)blocks, 256, sh, st>>>(*d, ktot_sk); \
    else tapconv_fwd_smallk_kernel<TX, TY, int64_t, NV><<<(unsigned)blocks, 256, sh, st>>>(*d, ktot_sk);     \
  } while (0)
#define LAUNCH_SK(TX, TY)                  \
  do {                                     \
    if (nv == 4) LAUNCH_SK2(TX, TY, 4);    \
    else if (nv == 2) LAUNCH_SK2(TX, TY, 2); \
    else LAUNCH_SK2(TX, TY, 1);            \
  } while (0)
    if (d->x_dtype == CLSKD_F32 && d->y_dtype == CLSKD_F32) LAUNCH_SK(float, float);
    else if (d->x_dtype == CLSKD_F32) LAUNCH_SK(float, __nv_bfloat16);
    else if (d->y_dtype == CLSKD_F32) LAUNCH_SK(__nv_bfloat16, float);
    else LAUNCH_SK(__nv_bfloat16, __nv_bfloat16);
#undef LAUNCH_SK2
#undef LAUNCH_SK
    CLSKD_CHECK_LAUNCH("clskd_tapconv_fwd(smallk)");
    return CLSKD_OK;
  }
  dim3 grid(cdiv(M, BM), cdiv(d->N, BN));
  bool vec = vec_ok(d);
#define LAUNCH(TX, TY)                                                        \
  do {                                                                        \
    if (vec) tapconv_fwd_kernel<TX, TY, true><<<grid, NT, 0, st>>>(*d);       \
    else tapconv_fwd_kernel<TX, TY, false><<<grid, NT, 0, st>>>(*d);          \
  } while (0)
  if (d->x_dtype == CLSKD_F32 && d->y_dtype == CLSKD_F32) LAUNCH(float, float);
  else if (d->x_dtype == CLSKD_F32) LAUNCH(float, __nv_bfloat16);
  else if (d->y_dtype == CLSKD_F32) LAUNCH(__nv_bfloat16, float);
  else LAUNCH(__nv_bfloat16, __nv_bfloat16);
#undef LAUNCH
  CLSKD_CHECK_LAUNCH("clskd_tapconv_fwd");
  return CLSKD_OK;
}

extern "C" int clskd_tapconv_wgrad(const ClskdTapConv* d, void* stream) {
  int rc = check_desc(d, "clskd_tapconv_wgrad");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  int Ktot = d->ntaps * (d->c0 + d->c1);
  int64_t M = (int64_t)d->B * d->To * d->Fo;
  if (!d->accumulate) {
    cudaError_t e = cudaMemsetAsync(const_cast<void*>(d->w), 0, sizeof(float) * (size_t)Ktot * d->N, st);
    if (e != cudaSuccess) {
      set_error("clskd_tapconv_wgrad: memset failed: %s", cudaGetErrorString(e));
      return CLSKD_ERR_CUDA;
    }
  }
  if (M == 0) return CLSKD_OK;
  if (pw::try_wgrad(d, st)) {
    CLSKD_CHECK_LAUNCH("clskd_tapconv_wgrad(pointwise)");
    return CLSKD_OK;
  }
  N2Geom g2;
  if (n2_ok(d, &g2) && M >= 1024 && g2.U > 128) {
    // many taps: one pass over the rows per group of taps whose units fit the register budget
    const int tpg = 128 / g2.cpt;           // taps per group (cpt <= 128 -> >= 1)
    const int Ctot_ = d->c0 + d->c1;
    for (int t0 = 0; t0 < d->ntaps; t0 += tpg) {
      ClskdTapConv sub = *d;
      sub.ntaps = d->ntaps - t0 < tpg ? d->ntaps - t0 : tpg;
      for (int j = 0; j < sub.ntaps; ++j) {
        sub.dt[j] = d->dt[t0 + j];
        sub.df[j] = d->df[t0 + j];
      }
      sub.w = reinterpret_cast<const float*>(d->w) + (size_t)t0 * Ctot_ * d->N;
      sub.accumulate = 1;                   // dW was zeroed above (or the caller accumulates)
      int rc2 = clskd_tapconv_wgrad(&sub, stream);
      if (rc2) return rc2;
    }
    return CLSKD_OK;
  }
  if (n2_ok(d, &g2) && M >= 1024) {
    const int upl = cdiv(g2.U, g2.up2);
    int blocks = sm_count() * 4;
    const size_t sh = sizeof(float) * (size_t)g2.U * 16;
#define LAUNCH_W2(TX, TY)                                                                         \
  do {                                                                                            \
    if (upl <= 1) tapconv_wgrad_n2_kernel<TX, TY, 1><<<blocks, 256, sh, st>>>(*d, g2);            \
    else if (upl == 2) tapconv_wgrad_n2_kernel<TX, TY, 2><<<blocks, 256, sh, st>>>(*d, g2);       \
    else tapconv_wgrad_n2_kernel<TX, TY, 4><<<blocks, 256, sh, st>>>(*d, g2);                     \
  } while (0)
    if (d->x_dtype == CLSKD_F32 && d->y_dtype == CLSKD_F32) LAUNCH_W2(float, float);
    else if (d->x_dtype == CLSKD_F32) LAUNCH_W2(float, __nv_bfloat16);
    else if (d->y_dtype == CLSKD_F32) LAUNCH_W2(__nv_bfloat16, float);
    else LAUNCH_W2(__nv_bfloat16, __nv_bfloat16);
#undef LAUNCH_W2
    CLSKD_CHECK_LAUNCH("clskd_tapconv_wgrad(n2)");
    return CLSKD_OK;
  }
  if (c2mma::try_wgrad(d, st)) {
    CLSKD_CHECK_LAUNCH("clskd_tapconv_wgrad(c2 mma)");
    return CLSKD_OK;
  }
  {
    int tmin, tspan, fmin, fspan;
    if (M >= 4096 && wgrad_c2_ok(d, &tmin, &tspan, &fmin, &fspan)) {
      const size_t sh = sizeof(float) * ((size_t)tspan * (d->Fi + fspan) * 2 + (size_t)d->ntaps * 2 * d->N);
      int64_t lines = (int64_t)d->B * d->To;
      int blocks = sm_count() * 4;
      if (blocks > lines) blocks = (int)lines;
#define LAUNCH_C2(TY, NT)                                                                                       \
  do {                                                                                                          \
    cudaFuncSetAttribute(tapconv_wgrad_c2_kernel<TY, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); \
    tapconv_wgrad_c2_kernel<TY, NT><<<blocks, 256, sh, st>>>(*d, tmin, tspan, fmin, fspan);                     \
  } while (0)
      if (d->y_dtype == CLSKD_F32) {
        if (d->ntaps == 10) LAUNCH_C2(float, 10); else if (d->ntaps == 9) LAUNCH_C2(float, 9); else LAUNCH_C2(float, 1);
      } else {
        if (d->ntaps == 10) LAUNCH_C2(__nv_bfloat16, 10); else if (d->ntaps == 9) LAUNCH_C2(__nv_bfloat16, 9); else LAUNCH_C2(__nv_bfloat16, 1);
      }
#undef LAUNCH_C2
      CLSKD_CHECK_LAUNCH("clskd_tapconv_wgrad(c2)");
      return CLSKD_OK;
    }
  }
  int ktot_sk;
  if (smallk_ok(d, &ktot_sk) && M >= 1024 && d->N <= 256) {
    const int kchunks = cdiv(ktot_sk, 8);
    const int per_row = (d->N / 8) * kchunks;
    if (per_row <= 256) {
      int blocks = sm_count() * 4;
      const size_t sh = sizeof(float) * (size_t)ktot_sk * d->N;
      if (d->x_dtype == CLSKD_F32 && d->y_dtype == CLSKD_F32)
        tapconv_wgrad_smallk_kernel<float, float><<<blocks, 256, sh, st>>>(*d, ktot_sk, kchunks);
      else if (d->x_dtype == CLSKD_F32)
        tapconv_wgrad_smallk_kernel<float, __nv_bfloat16><<<blocks, 256, sh, st>>>(*d, ktot_sk, kchunks);
      else if (d->y_dtype == CLSKD_F32)
        tapconv_wgrad_smallk_kernel<__nv_bfloat16, float><<<blocks, 256, sh, st>>>(*d, ktot_sk, kchunks);
      else
        tapconv_wgrad_smallk_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, sh, st>>>(*d, ktot_sk, kchunks);
      CLSKD_CHECK_LAUNCH("clskd_tapconv_wgrad(smallk)");
      return CLSKD_OK;
    }
  }
  int tiles = cdiv(Ktot, BM) * cdiv(d->N, BN);
  int target = sm_count() * 6;
  int splits = target / tiles;
  if (splits < 1) splits = 1;
  int64_t max_splits = (M + 255) / 256;
  if (splits > max_splits) splits = (int)max_splits;
  if (splits > 65535) splits = 65535;
  int64_t rows = (M + splits - 1) / splits;
  rows = (rows + WK - 1) / WK * WK;
  splits = (int)((M + rows - 1) / rows);
  dim3 grid(cdiv(Ktot, BM), cdiv(d->N, BN), splits);
  if (d->x_dtype == CLSKD_F32 && d->y_dtype == CLSKD_F32)
    tapconv_wgrad_kernel<float, float><<<grid, NT, 0, st>>>(*d, rows);
  else if (d->x_dtype == CLSKD_F32)
    tapconv_wgrad_kernel<float, __nv_bfloat16><<<grid, NT, 0, st>>>(*d, rows);
  else if (d->y_dtype == CLSKD_F32)
    tapconv_wgrad_kernel<__nv_bfloat16, float><<<grid, NT, 0, st>>>(*d, rows);
  else
    tapconv_wgrad_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, NT, 0, st>>>(*d, rows);
  CLSKD_CHECK_LAUNCH("clskd_tapconv_wgrad");
  return CLSKD_OK;
}
