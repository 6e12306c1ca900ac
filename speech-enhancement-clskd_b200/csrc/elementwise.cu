// Memory-bound kernels of the DCCRN distillation path: layout/packing helpers, BatchNorm(+PReLU)
// forward/backward, complex BatchNorm, the polar mask, iSTFT overlap-add, ABF helpers, Adam.
// All are HBM-bound: coalesced, 128-bit vectorised where the shape allows, grids sized in
// multiples of the SM count (grid-stride loops).
#include <stdarg.h>

#include "common.cuh"

namespace clskd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

static inline int ew_grid(int64_t work_items, int threads) {
  int64_t blocks = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

namespace {

// ------------------------------------------------------------------------------- layout helpers
template <typename TS, typename TD>
__global__ void strided_copy4d_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int64_t n0,
                                      int64_t n1, int64_t n2, int64_t n3, int64_t s0, int64_t s1,
                                      int64_t s2, int64_t s3, int64_t d0, int64_t d1, int64_t d2,
                                      int64_t d3) {
  int64_t total = n0 * n1 * n2 * n3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t i3 = i % n3, r = i / n3;
    int64_t i2 = r % n2;
    r /= n2;
    int64_t i1 = r % n1, i0 = r / n1;
    st_f(dst + i0 * d0 + i1 * d1 + i2 * d2 + i3 * d3,
         ld_f(src + i0 * s0 + i1 * s1 + i2 * s2 + i3 * s3));
  }
}

// innermost dimension contiguous on both sides: a thread moves V consecutive elements with 8/16-byte
// accesses and decodes its index once per vector (32-bit arithmetic whenever the count allows)
template <typename TS, typename TD, int V, typename I>
__global__ void __launch_bounds__(256)
strided_copy4d_vec_kernel(const TS* __restrict__ src, TD* __restrict__ dst, I n1, I n2, I n3v, I total,
                          int64_t s0, int64_t s1, int64_t s2, int64_t d0, int64_t d1, int64_t d2) {
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const I i3 = i % n3v;
    I r = i / n3v;
    const I i2 = r % n2;
    r /= n2;
    const I i1 = r % n1, i0 = r / n1;
    const TS* sp = src + (int64_t)i0 * s0 + (int64_t)i1 * s1 + (int64_t)i2 * s2 + (int64_t)i3 * V;
    TD* dp = dst + (int64_t)i0 * d0 + (int64_t)i1 * d1 + (int64_t)i2 * d2 + (int64_t)i3 * V;
    if (V >= 4) {
#pragma unroll
      for (int j = 0; j < V / 4; ++j) st4(dp + 4 * j, ld4(sp + 4 * j));
    } else {
      const float a = ld_f(sp), b = ld_f(sp + 1);
      st_f(dp, a);
      st_f(dp + 1, b);
    }
  }
}

// The (channel, frequency) transposes either side of the LSTM (DCCRN.py:178-199: [B,T,F,C] maps <-> [T,B,C*F] features)
// have ONE dimension of extent 4 (the frequency bins left after the encoder).  The generic kernel above moves them one
// element at a time with three 64-bit divisions each (0.65 TB/s); here a thread owns a 4 x 4 block - four 4-element vector
// accesses on the strided side, 16 contiguous elements on the dense side - with one 32-bit index decode per block.
//   DIR 0: shape [n0,n1,C,4], src (.., .., 1, Ls) -> dst (.., .., 4, 1)     src rows [f][c] (pitch Ls), dst dense [c][f]
//   DIR 1: shape [n0,n1,4,C], src (.., .., 1, 4)  -> dst (.., .., Ld, 1)    src dense [c][f], dst rows [f][c] (pitch Ld)
__device__ __forceinline__ void st16(float* p, const float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(p)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void st16(__nv_bfloat16* p, const float* v) {
  uint32_t w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    w[j] = *reinterpret_cast<const uint32_t*>(&h);
  }
  reinterpret_cast<uint4*>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
  reinterpret_cast<uint4*>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
__device__ __forceinline__ void ld16(const float* p, float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 a = reinterpret_cast<const float4*>(p)[j];
    v[4 * j] = a.x; v[4 * j + 1] = a.y; v[4 * j + 2] = a.z; v[4 * j + 3] = a.w;
  }
}
__device__ __forceinline__ void ld16(const __nv_bfloat16* p, float* v) {
  const uint4 a = reinterpret_cast<const uint4*>(p)[0], b = reinterpret_cast<const uint4*>(p)[1];
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[2 * j] = __uint_as_float(w[j] << 16);
    v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}

template <typename TS, typename TD, int DIR>
__global__ void __launch_bounds__(256)
transpose4_kernel(const TS* __restrict__ src, TD* __restrict__ dst, unsigned n1, unsigned cgroups, unsigned total,
                  int64_t s0, int64_t s1, int64_t d0, int64_t d1, int64_t pitch) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned cg = i % cgroups, u = i / cgroups;
    const unsigned i1 = u % n1, i0 = u / n1;
    const TS* sp = src + (int64_t)i0 * s0 + (int64_t)i1 * s1;
    TD* dp = dst + (int64_t)i0 * d0 + (int64_t)i1 * d1;
    float v[16];
    if (DIR == 0) {
      float4 r[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) r[f] = ld4(sp + (int64_t)f * pitch + cg * 4);
#pragma unroll
      for (int f = 0; f < 4; ++f) { v[f] = r[f].x; v[4 + f] = r[f].y; v[8 + f] = r[f].z; v[12 + f] = r[f].w; }
      st16(dp + cg * 16, v);
    } else {
      ld16(sp + cg * 16, v);
#pragma unroll
      for (int f = 0; f < 4; ++f) st4(dp + (int64_t)f * pitch + cg * 4, make_float4(v[f], v[4 + f], v[8 + f], v[12 + f]));
    }
  }
}

template <typename TD>
__global__ void pack_gather_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                   const int32_t* __restrict__ table, int64_t n,
                                   TD* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      int32_t e = table[2 * i + k];
      if (e >= 0) {
        const float* s = (e & 1) ? b : a;
        float t = s[e >> 2];
        v += (e & 2) ? -t : t;
      }
    }
    st_f(out + i, v);
  }
}

// every registered packed weight in ONE launch: desc[e] = {a, b, table, out, start, n*2 + (out is bf16)} (6 x int64);
// `start` is the entry's offset in the concatenated index space (ascending), `total` its end
__global__ void multi_pack_gather_kernel(const int64_t* __restrict__ desc, int n_entries, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_entries - 1;
    while (lo < hi) {                       // last entry with start <= i
      const int mid = (lo + hi + 1) >> 1;
      if (desc[6 * mid + 4] <= i) lo = mid; else hi = mid - 1;
    }
    const int64_t* de = desc + 6 * lo;
    const float* a = reinterpret_cast<const float*>(de[0]);
    const float* b = reinterpret_cast<const float*>(de[1]);
    const int32_t* table = reinterpret_cast<const int32_t*>(de[2]);
    const int64_t j = i - de[4];
    if (j >= (de[5] >> 1)) continue;
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int32_t e = table[2 * j + k];
      if (e >= 0) {
        const float t = ((e & 1) ? b : a)[e >> 2];
        v += (e & 2) ? -t : t;
      }
    }
    if (de[5] & 1) reinterpret_cast<__nv_bfloat16*>(de[3])[j] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(de[3])[j] = v;
  }
}

__global__ void unpack_gather2_kernel(const float* __restrict__ src,
                                      const int32_t* __restrict__ table2, int64_t n,
                                      float* __restrict__ dst, int accumulate) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float v = accumulate ? dst[i] : 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      int32_t e = table2[2 * i + k];
      if (e >= 0) {
        float s = src[e >> 1];
        v += (e & 1) ? -s : s;
      }
    }
    dst[i] = v;
  }
}

__device__ __forceinline__ int pad_map(int j, int L, int mode) {
  // j = index into the unpadded signal (may be out of range)
  if (j >= 0 && j < L) return j;
  if (mode == 0) return -1;
  // reflect (no edge repeat), valid for |overshoot| < L
  if (j < 0) j = -j;
  if (j >= L) j = 2 * (L - 1) - j;
  return (j >= 0 && j < L) ? j : -1;
}

template <typename TS>
__global__ void pad1d_kernel(const TS* __restrict__ src, int64_t src_sB, int B, int L, int left,
                             int right, int mode, float* __restrict__ dst) {
  int64_t Lp = (int64_t)L + left + right;
  int64_t total = Lp * B;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(i / Lp);
    int p = (int)(i - (int64_t)b * Lp);
    int j = pad_map(p - left, L, mode);
    dst[i] = j >= 0 ? ld_f(src + (int64_t)b * src_sB + j) : 0.f;
  }
}

__global__ void pad1d_bwd_kernel(const float* __restrict__ ddst, int B, int L, int left, int right,
                                 int mode, float* __restrict__ dsrc, int accumulate) {
  int64_t Lp = (int64_t)L + left + right;
  int64_t total = (int64_t)L * B;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(i / L);
    int j = (int)(i - (int64_t)b * L);
    const float* row = ddst + (int64_t)b * Lp;
    float v = row[j + left];
    if (mode == 1) {
      // left reflection: padded index p = left - j (j in 1..left); right: p = left + 2(L-1) - j
      if (j >= 1 && j <= left) v += row[left - j];
      int p = left + 2 * (L - 1) - j;
      if (j <= L - 2 && p >= left + L && p < Lp) v += row[p];
    }
    dsrc[i] = accumulate ? dsrc[i] + v : v;
  }
}

// ------------------------------------------------------------------------------- BatchNorm
// Column statistics of a dense [M, C] matrix.  Threads of a block are laid out so that a thread
// always sees the same channel; per-thread fp32 partials over a bounded number of rows, then a
// shared-memory fold and one fp64 atomic per (block, channel).
constexpr int CS_ROWS_PER_THREAD = 64;

template <typename T, int MODE>  // MODE 0: (x, x^2); MODE 1: bn backward sums
__global__ void colstats_kernel(const T* __restrict__ x, const T* __restrict__ dy, int64_t M, int C,
                                int rpb, const float* __restrict__ mean,
                                const float* __restrict__ invstd, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ slope,
                                double* __restrict__ out0, double* __restrict__ out1,
                                double* __restrict__ out2) {
  extern __shared__ float sh[];  // [2 or 3][blockDim.x]
  const int tid = threadIdx.x;
  const int active = rpb * C;
  const int c = tid % C;
  const int r0 = tid / C;
  const int64_t rows_per_block = (int64_t)rpb * CS_ROWS_PER_THREAD;
  const int64_t mbeg = (int64_t)blockIdx.x * rows_per_block;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  if (tid < active) {
    float mu = 0.f, is = 1.f, g = 1.f, bt = 0.f, sl = 1.f;
    if (MODE == 1) {
      mu = mean[c];
      is = invstd[c];
      g = gamma ? gamma[c] : 1.f;
      bt = beta ? beta[c] : 0.f;
      sl = slope ? slope[0] : 1.f;
    }
#pragma unroll 4
    for (int it = 0; it < CS_ROWS_PER_THREAD; ++it) {
      int64_t m = mbeg + r0 + (int64_t)it * rpb;
      if (m >= M) break;
      float v = ld_f(x + m * C + c);
      if (MODE == 0) {
        a0 += v;
        a1 += v * v;
      } else {
        float xh = (v - mu) * is;
        float u = xh * g + bt;
        float d = ld_f(dy + m * C + c);
        float dz = u > 0.f ? d : d * sl;
        a0 += dz;
        a1 += dz * xh;
        a2 += u > 0.f ? 0.f : d * u;
      }
    }
  }
  sh[tid] = a0;
  sh[blockDim.x + tid] = a1;
  if (MODE == 1) sh[2 * blockDim.x + tid] = a2;
  __syncthreads();
  if (tid < C) {
    double s0 = 0., s1 = 0., s2 = 0.;
    for (int r = 0; r < rpb; ++r) {
      s0 += sh[r * C + tid];
      s1 += sh[blockDim.x + r * C + tid];
      if (MODE == 1) s2 += sh[2 * blockDim.x + r * C + tid];
    }
    atomicAdd(out0 + tid, s0);
    atomicAdd(out1 + tid, s1);
    if (MODE == 1) {
      // dslope is a single scalar: fold the C partials first
      sh[tid] = (float)s2;
    }
  }
  if (MODE == 1) {
    __syncthreads();
    if (tid == 0 && out2) {
      double s = 0.;
      for (int i = 0; i < C; ++i) s += sh[i];
      atomicAdd(out2, s);
    }
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq,
                                   int64_t M, int C, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ rmean, float* __restrict__ rvar) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mu = sum[c] / (double)M;
  double var = sumsq[c] / (double)M - mu * mu;
  if (var < 0.) var = 0.;
  mean[c] = (float)mu;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mu;
  if (rvar) {
    double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
  }
}

__global__ void bn_fold_kernel(const float* __restrict__ rmean, const float* __restrict__ rvar,
                               const float* __restrict__ gamma, const float* __restrict__ beta, int C, float eps,
                               float* __restrict__ scale, float* __restrict__ shift) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = (gamma ? gamma[c] : 1.f) * (1.f / sqrtf(rvar[c] + eps));
  scale[c] = sc;
  shift[c] = (beta ? beta[c] : 0.f) - rmean[c] * sc;
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ rmean, const float* __restrict__ rvar,
                                     int C, float eps, float* __restrict__ mean,
                                     float* __restrict__ invstd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean[c] = rmean[c];
  invstd[c] = 1.f / sqrtf(rvar[c] + eps);
}

template <typename TX, typename TY>
__global__ void bn_act_fwd_kernel(const TX* __restrict__ x, int64_t n4, int C,
                                  const float* __restrict__ mean, const float* __restrict__ invstd,
                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                  const float* __restrict__ slope, TY* __restrict__ y) {
  const float sl = slope ? slope[0] : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)((i * 4) % C);
    float4 v = ld4(x + i * 4);
    float in[4] = {v.x, v.y, v.z, v.w}, o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float sc = invstd[c + e] * (gamma ? gamma[c + e] : 1.f);
      float u = (in[e] - mean[c + e]) * sc + (beta ? beta[c + e] : 0.f);
      o[e] = u > 0.f ? u : u * sl;
    }
    st4(y + i * 4, make_float4(o[0], o[1], o[2], o[3]));
  }
}

template <typename TX, typename TY>
__global__ void bn_act_fwd_scalar_kernel(const TX* __restrict__ x, int64_t n, int C,
                                         const float* __restrict__ mean,
                                         const float* __restrict__ invstd,
                                         const float* __restrict__ gamma,
                                         const float* __restrict__ beta,
                                         const float* __restrict__ slope, TY* __restrict__ y) {
  const float sl = slope ? slope[0] : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    float sc = invstd[c] * (gamma ? gamma[c] : 1.f);
    float u = (ld_f(x + i) - mean[c]) * sc + (beta ? beta[c] : 0.f);
    st_f(y + i, u > 0.f ? u : u * sl);
  }
}

template <typename TX, typename TD, typename TO>
__global__ void bn_act_bwd_apply_kernel(const TX* __restrict__ x, const TD* __restrict__ dy,
                                        int64_t n, int64_t M, int C, const float* __restrict__ mean,
                                        const float* __restrict__ invstd,
                                        const float* __restrict__ gamma,
                                        const float* __restrict__ beta,
                                        const float* __restrict__ slope,
                                        const double* __restrict__ sum_dz,
                                        const double* __restrict__ sum_dz_xhat, int training,
                                        TO* __restrict__ dx) {
  const float sl = slope ? slope[0] : 1.f;
  const float invM = 1.f / (float)M;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    float g = gamma ? gamma[c] : 1.f;
    float is = invstd[c];
    float xh = (ld_f(x + i) - mean[c]) * is;
    float u = xh * g + (beta ? beta[c] : 0.f);
    float d = ld_f(dy + i);
    float dz = u > 0.f ? d : d * sl;
    float r = dz;
    if (training) r = dz - (float)sum_dz[c] * invM - xh * (float)sum_dz_xhat[c] * invM;
    st_f(dx + i, g * is * r);
  }
}

__global__ void bn_param_grads_kernel(const double* __restrict__ sum_dz,
                                      const double* __restrict__ sum_dz_xhat,
                                      const double* __restrict__ dslope, int C,
                                      float* __restrict__ dgamma, float* __restrict__ dbeta,
                                      float* __restrict__ dslope_out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    if (dgamma) dgamma[c] = (float)sum_dz_xhat[c];
    if (dbeta) dbeta[c] = (float)sum_dz[c];
  }
  if (c == 0 && dslope_out && dslope) dslope_out[0] = (float)dslope[0];
}

// ------------------------------------------------------------------------------- complex BN
template <typename T>
__global__ void cbn_moments_kernel(const T* __restrict__ x, int64_t M, int Cc, int rpb,
                                   double* __restrict__ s) {
  extern __shared__ float sh[];  // [5][blockDim.x]
  const int tid = threadIdx.x;
  const int active = rpb * Cc;
  const int c = tid % Cc, r0 = tid / Cc;
  const int64_t mbeg = (int64_t)blockIdx.x * rpb * CS_ROWS_PER_THREAD;
  float a[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (tid < active) {
    for (int it = 0; it < CS_ROWS_PER_THREAD; ++it) {
      int64_t m = mbeg + r0 + (int64_t)it * rpb;
      if (m >= M) break;
      float xr = ld_f(x + m * 2 * Cc + c), xi = ld_f(x + m * 2 * Cc + Cc + c);
      a[0] += xr; a[1] += xi; a[2] += xr * xr; a[3] += xr * xi; a[4] += xi * xi;
    }
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) sh[k * blockDim.x + tid] = a[k];
  __syncthreads();
  if (tid < Cc) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      double v = 0.;
      for (int r = 0; r < rpb; ++r) v += sh[k * blockDim.x + r * Cc + tid];
      atomicAdd(s + (int64_t)k * Cc + tid, v);
    }
  }
}

__global__ void cbn_finalize_kernel(const double* __restrict__ s, int64_t M, int Cc, float eps,
                                    float momentum, int training, const float* __restrict__ Wrr,
                                    const float* __restrict__ Wri, const float* __restrict__ Wii,
                                    float* RMr, float* RMi, float* RVrr, float* RVri, float* RVii,
                                    float* __restrict__ coef) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cc) return;
  float Mr, Mi, Vrr, Vri, Vii;
  if (training) {
    double mr = s[c] / M, mi = s[Cc + c] / M;
    Mr = (float)mr;
    Mi = (float)mi;
    Vrr = (float)(s[2 * Cc + c] / M - mr * mr);
    Vri = (float)(s[3 * Cc + c] / M - mr * mi);
    Vii = (float)(s[4 * Cc + c] / M - mi * mi);
    if (RMr) {  // lerp_(new, w): old + w*(new-old)   (tools_for_model.py:433-434,455-457)
      RMr[c] += momentum * (Mr - RMr[c]);
      RMi[c] += momentum * (Mi - RMi[c]);
      RVrr[c] += momentum * (Vrr - RVrr[c]);
      RVri[c] += momentum * (Vri - RVri[c]);
      RVii[c] += momentum * (Vii - RVii[c]);
    }
  } else {
    Mr = RMr[c]; Mi = RMi[c]; Vrr = RVrr[c]; Vri = RVri[c]; Vii = RVii[c];
  }
  Vrr += eps;
  Vii += eps;
  float tau = Vrr + Vii;
  float delta = Vrr * Vii - Vri * Vri;
  float sq = sqrtf(delta);
  float t = sqrtf(tau + 2.f * sq);
  float rst = 1.f / (sq * t);
  float Urr = (sq + Vii) * rst, Uii = (sq + Vrr) * rst, Uri = -Vri * rst;
  float Zrr = Urr, Zri = Uri, Zir = Uri, Zii = Uii;
  if (Wrr) {
    float wrr = Wrr[c], wri = Wri[c], wii = Wii[c];
    Zrr = wrr * Urr + wri * Uri;
    Zri = wrr * Uri + wri * Uii;
    Zir = wri * Urr + wii * Uri;
    Zii = wri * Uri + wii * Uii;
  }
  coef[c] = Mr;
  coef[Cc + c] = Mi;
  coef[2 * Cc + c] = Zrr;
  coef[3 * Cc + c] = Zri;
  coef[4 * Cc + c] = Zir;
  coef[5 * Cc + c] = Zii;
}

template <typename TX, typename TY>
__global__ void cbn_apply_kernel(const TX* __restrict__ x, int64_t M, int Cc,
                                 const float* __restrict__ coef, const float* __restrict__ Br,
                                 const float* __restrict__ Bi, TY* __restrict__ y) {
  int64_t total = M * Cc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % Cc);
    int64_t m = i / Cc;
    float xr = ld_f(x + m * 2 * Cc + c) - coef[c];
    float xi = ld_f(x + m * 2 * Cc + Cc + c) - coef[Cc + c];
    float yr = coef[2 * Cc + c] * xr + coef[3 * Cc + c] * xi + (Br ? Br[c] : 0.f);
    float yi = coef[4 * Cc + c] * xr + coef[5 * Cc + c] * xi + (Bi ? Bi[c] : 0.f);
    st_f(y + m * 2 * Cc + c, yr);
    st_f(y + m * 2 * Cc + Cc + c, yi);
  }
}

// ---- complex BN backward (autograd of tools_for_model.py:398-508, batch statistics included)
// pass 1: s6[0..5][Cc] = sum dyr, dyi, dyr*xr, dyr*xi, dyi*xr, dyi*xi
template <typename T>
__global__ void cbn_bwd_moments_kernel(const T* __restrict__ x, const T* __restrict__ dy, int64_t M, int Cc,
                                       int rpb, double* __restrict__ s) {
  extern __shared__ float sh[];  // [6][blockDim.x]
  const int tid = threadIdx.x;
  const int active = rpb * Cc;
  const int c = tid % Cc, r0 = tid / Cc;
  const int64_t mbeg = (int64_t)blockIdx.x * rpb * CS_ROWS_PER_THREAD;
  float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (tid < active) {
    for (int it = 0; it < CS_ROWS_PER_THREAD; ++it) {
      int64_t m = mbeg + r0 + (int64_t)it * rpb;
      if (m >= M) break;
      const float xr = ld_f(x + m * 2 * Cc + c), xi = ld_f(x + m * 2 * Cc + Cc + c);
      const float gr = ld_f(dy + m * 2 * Cc + c), gi = ld_f(dy + m * 2 * Cc + Cc + c);
      a[0] += gr; a[1] += gi; a[2] += gr * xr; a[3] += gr * xi; a[4] += gi * xr; a[5] += gi * xi;
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) sh[k * blockDim.x + tid] = a[k];
  __syncthreads();
  if (tid < Cc) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double v = 0.;
      for (int r = 0; r < rpb; ++r) v += sh[k * blockDim.x + r * Cc + tid];
      atomicAdd(s + (int64_t)k * Cc + tid, v);
    }
  }
}

// per-channel closed-form backward of the 2x2 whitening (fp64): parameter gradients and the
// coefficients of  dx = Z^T (dy - mean dy) + G (x - mean x)   (G = 0 with running statistics)
__global__ void cbn_bwd_finalize_kernel(const double* __restrict__ s, const double* __restrict__ s6, int64_t M,
                                        int Cc, float eps, int training, const float* __restrict__ Wrr,
                                        const float* __restrict__ Wri, const float* __restrict__ Wii,
                                        const float* RMr, const float* RMi, const float* RVrr, const float* RVri,
                                        const float* RVii, float* __restrict__ coefb, float* dWrr, float* dWri,
                                        float* dWii, float* dBr, float* dBi) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cc) return;
  double mr, mi, a, b, d;
  if (training) {
    mr = s[c] / M; mi = s[Cc + c] / M;
    a = s[2 * Cc + c] / M - mr * mr;
    b = s[3 * Cc + c] / M - mr * mi;
    d = s[4 * Cc + c] / M - mi * mi;
  } else {
    mr = RMr[c]; mi = RMi[c]; a = RVrr[c]; b = RVri[c]; d = RVii[c];
  }
  a += (double)eps;
  d += (double)eps;
  const double sq = sqrt(a * d - b * b), t = sqrt(a + d + 2. * sq), r = 1. / (sq * t);
  const double Urr = (sq + d) * r, Uii = (sq + a) * r, Uri = -b * r;
  const double wrr = Wrr ? Wrr[c] : 1., wri = Wri ? Wri[c] : 0., wii = Wii ? Wii[c] : 1.;
  const double Zrr = wrr * Urr + wri * Uri, Zri = wrr * Uri + wri * Uii;
  const double Zir = wri * Urr + wii * Uri, Zii = wri * Uri + wii * Uii;
  const double gr = s6[c], gi = s6[Cc + c];
  // A[i][j] = sum dy_i * (x_j - mean_j)
  const double Arr = s6[2 * Cc + c] - mr * gr, Ari = s6[3 * Cc + c] - mi * gr;
  const double Air = s6[4 * Cc + c] - mr * gi, Aii = s6[5 * Cc + c] - mi * gi;
  // Z = W U  ->  dW = A U^T (U symmetric), dU = W^T A
  if (dWrr) dWrr[c] = (float)(Arr * Urr + Ari * Uri);
  if (dWri) dWri[c] = (float)((Arr * Uri + Ari * Uii) + (Air * Urr + Aii * Uri));
  if (dWii) dWii[c] = (float)(Air * Uri + Aii * Uii);
  if (dBr) dBr[c] = (float)gr;
  if (dBi) dBi[c] = (float)gi;
  double Grr = 0., Gri = 0., Gii = 0.;
  if (training) {
    const double dUrr = wrr * Arr + wri * Air;
    const double dUri = (wrr * Ari + wri * Aii) + (wri * Arr + wii * Air);
    const double dUii = wri * Ari + wii * Aii;
    // derivatives of U(a, b, d)
    const double ds[3] = {d / (2. * sq), -b / sq, a / (2. * sq)};
    double dv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double dt = ((k == 1 ? 0. : 1.) + 2. * ds[k]) / (2. * t);
      const double dr = -r * (ds[k] / sq + dt / t);
      const double dUrr_k = (ds[k] + (k == 2 ? 1. : 0.)) * r + (sq + d) * dr;
      const double dUii_k = (ds[k] + (k == 0 ? 1. : 0.)) * r + (sq + a) * dr;
      const double dUri_k = -(k == 1 ? 1. : 0.) * r - b * dr;
      dv[k] = dUrr * dUrr_k + dUii * dUii_k + dUri * dUri_k;
    }
    Grr = 2. * dv[0] / M; Gri = dv[1] / M; Gii = 2. * dv[2] / M;
  }
  const double mgr = training ? gr / M : 0., mgi = training ? gi / M : 0.;
  coefb[c] = (float)mr;
  coefb[Cc + c] = (float)mi;
  coefb[2 * Cc + c] = (float)Zrr;   // dx_r = Zrr dyr + Zir dyi + Grr xcr + Gri xci - cr
  coefb[3 * Cc + c] = (float)Zir;
  coefb[4 * Cc + c] = (float)Zri;   // dx_i = Zri dyr + Zii dyi + Gri xcr + Gii xci - ci
  coefb[5 * Cc + c] = (float)Zii;
  coefb[6 * Cc + c] = (float)Grr;
  coefb[7 * Cc + c] = (float)Gri;
  coefb[8 * Cc + c] = (float)Gii;
  coefb[9 * Cc + c] = (float)(Zrr * mgr + Zir * mgi);
  coefb[10 * Cc + c] = (float)(Zri * mgr + Zii * mgi);
}

template <typename T>
__global__ void cbn_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy, int64_t M, int Cc,
                                     const float* __restrict__ cb, T* __restrict__ dx) {
  const int64_t total = M * Cc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cc);
    const int64_t m = i / Cc;
    const float xr = ld_f(x + m * 2 * Cc + c) - cb[c], xi = ld_f(x + m * 2 * Cc + Cc + c) - cb[Cc + c];
    const float gr = ld_f(dy + m * 2 * Cc + c), gi = ld_f(dy + m * 2 * Cc + Cc + c);
    const float dr = cb[2 * Cc + c] * gr + cb[3 * Cc + c] * gi + cb[6 * Cc + c] * xr + cb[7 * Cc + c] * xi - cb[9 * Cc + c];
    const float di = cb[4 * Cc + c] * gr + cb[5 * Cc + c] * gi + cb[7 * Cc + c] * xr + cb[8 * Cc + c] * xi - cb[10 * Cc + c];
    st_f(dx + m * 2 * Cc + c, dr);
    st_f(dx + m * 2 * Cc + Cc + c, di);
  }
}

// ------------------------------------------------------------------------------- mask
// One thread per (b,t,bin).  Trig-free polar mask: cos/sin of atan2 are the normalised
// components, so est = tanh|m| * |s|_eps * (unit(s) * unit(m)) as a complex product.
template <typename TM>
__global__ void mask_fwd_kernel(const float* __restrict__ spec, const TM* __restrict__ mask,
                                int64_t m_sB, int64_t m_sT, int B, int T, int nb, int mode,
                                float* __restrict__ out, float* __restrict__ mask_padded) {
  int64_t total = (int64_t)B * T * nb;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int f = (int)(i % nb);
    int64_t bt = i / nb;
    int t = (int)(bt % T), b = (int)(bt / T);
    float2 s = *reinterpret_cast<const float2*>(spec + i * 2);
    float mr = 0.f, mi = 0.f;
    if (f > 0) {
      const TM* mp = mask + (int64_t)b * m_sB + (int64_t)t * m_sT + (int64_t)(f - 1) * 2;
      mr = ld_f(mp);
      mi = ld_f(mp + 1);
    }
    float orr, oi;
    if (mode == 0) {
      float mags = sqrtf(s.x * s.x + s.y * s.y + 1e-8f);
      float hs = sqrtf(s.x * s.x + s.y * s.y);
      float cp = 1.f, sp = 0.f;
      if (hs > 0.f) {
        cp = s.x / hs;
        sp = s.y / hs;
      }
      float mm = sqrtf(mr * mr + mi * mi);
      float cq = 1.f, sq = 0.f;
      if (mm > 0.f) {
        cq = mr / mm;
        sq = mi / mm;
      }
      float a = tanhf(mm) * mags;
      orr = a * (cp * cq - sp * sq);
      oi = a * (sp * cq + cp * sq);
    } else if (mode == 1) {
      orr = s.x * mr - s.y * mi;
      oi = s.x * mi + s.y * mr;
    } else {
      orr = s.x * mr;
      oi = s.y * mi;
    }
    *reinterpret_cast<float2*>(out + i * 2) = make_float2(orr, oi);
    if (mask_padded) *reinterpret_cast<float2*>(mask_padded + i * 2) = make_float2(mr, mi);
  }
}

template <typename TM, typename TD>
__global__ void mask_bwd_kernel(const float* __restrict__ spec, const TM* __restrict__ mask,
                                int64_t m_sB, int64_t m_sT, int B, int T, int nb, int mode,
                                const float* __restrict__ dout, TD* __restrict__ dmask,
                                int64_t dm_sB, int64_t dm_sT) {
  int nbm = nb - 1;
  int64_t total = (int64_t)B * T * nbm;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int fm = (int)(i % nbm);
    int64_t bt = i / nbm;
    int t = (int)(bt % T), b = (int)(bt / T);
    int64_t si = (bt * nb + fm + 1) * 2;
    float2 s = *reinterpret_cast<const float2*>(spec + si);
    float2 g = *reinterpret_cast<const float2*>(dout + si);
    const TM* mp = mask + (int64_t)b * m_sB + (int64_t)t * m_sT + (int64_t)fm * 2;
    float mr = ld_f(mp), mi = ld_f(mp + 1);
    float dmr = 0.f, dmi = 0.f;
    if (mode == 0) {
      float mags = sqrtf(s.x * s.x + s.y * s.y + 1e-8f);
      float hs = sqrtf(s.x * s.x + s.y * s.y);
      float cp = 1.f, sp = 0.f;
      if (hs > 0.f) {
        cp = s.x / hs;
        sp = s.y / hs;
      }
      float mm = sqrtf(mr * mr + mi * mi);
      if (mm > 0.f) {
        float cq = mr / mm, sq = mi / mm;
        float th = tanhf(mm);
        float A = th * mags, dA = (1.f - th * th) * mags;
        float ct = cp * cq - sp * sq, stt = sp * cq + cp * sq;
        float Aom = A / mm;
        // d out_r / d(mr,mi), d out_i / d(mr,mi)
        float orr_mr = dA * cq * ct + Aom * stt * sq;
        float orr_mi = dA * sq * ct - Aom * stt * cq;
        float oi_mr = dA * cq * stt - Aom * ct * sq;
        float oi_mi = dA * sq * stt + Aom * ct * cq;
        dmr = g.x * orr_mr + g.y * oi_mr;
        dmi = g.x * orr_mi + g.y * oi_mi;
      }
    } else if (mode == 1) {
      dmr = g.x * s.x + g.y * s.y;
      dmi = -g.x * s.y + g.y * s.x;
    } else {
      dmr = g.x * s.x;
      dmi = g.y * s.y;
    }
    TD* dp = dmask + (int64_t)b * dm_sB + (int64_t)t * dm_sT + (int64_t)fm * 2;
    st_f(dp, dmr);
    st_f(dp + 1, dmi);
  }
}

// ------------------------------------------------------------------------------- overlap-add
__global__ void ola_fwd_kernel(const float* __restrict__ frames, const float* __restrict__ window,
                               int B, int T, int win, int hop, int trim, int do_clamp,
                               float* __restrict__ wav) {
  const int L = (T - 1) * hop + win - 2 * trim;
  int64_t total = (int64_t)B * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(i / L);
    int p = (int)(i - (int64_t)b * L) + trim;
    int thi = p / hop;
    if (thi > T - 1) thi = T - 1;
    int tlo = (p - win + hop) / hop;  // ceil((p-win+1)/hop) for p-win+1 >= 0
    if (p - win + 1 <= 0) tlo = 0;
    float acc = 0.f, coff = 0.f;
    for (int t = tlo; t <= thi; ++t) {
      int k = p - t * hop;
      if (window) {
        float w = window[k];
        coff += w * w;
      }
      acc += frames[((int64_t)b * T + t) * win + k];
    }
    float v = window ? acc / (coff + 1e-8f) : acc;
    if (do_clamp) v = fminf(1.f, fmaxf(-1.f, v));
    wav[i] = v;
  }
}

__global__ void ola_bwd_kernel(const float* __restrict__ dwav, const float* __restrict__ wav,
                               const float* __restrict__ window, int B, int T, int win, int hop,
                               int trim, int do_clamp, float* __restrict__ dframes) {
  const int L = (T - 1) * hop + win - 2 * trim;
  int64_t total = (int64_t)B * T * win;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int k = (int)(i % win);
    int64_t bt = i / win;
    int t = (int)(bt % T), b = (int)(bt / T);
    int p = t * hop + k;
    int j = p - trim;
    float v = 0.f;
    if (j >= 0 && j < L) {
      bool pass = true;
      if (do_clamp) {
        float wv = wav[(int64_t)b * L + j];
        pass = wv > -1.f && wv < 1.f;
      }
      if (pass) {
        v = dwav[(int64_t)b * L + j];
        if (window) {
          int thi = p / hop;
          if (thi > T - 1) thi = T - 1;
          int tlo = (p - win + hop) / hop;
          if (p - win + 1 <= 0) tlo = 0;
          float coff = 0.f;
          for (int tt = tlo; tt <= thi; ++tt) {
            float w = window[p - tt * hop];
            coff += w * w;
          }
          v /= (coff + 1e-8f);
        }
      }
    }
    dframes[i] = v;
  }
}

// ------------------------------------------------------------------------------- ABF helpers
template <typename T>
__global__ void resize_f_fwd_kernel(const T* __restrict__ x, int64_t BT, int Fi, int Fo, int C,
                                    T* __restrict__ y) {
  int64_t total = BT * Fo * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t r = i / C;
    int fo = (int)(r % Fo);
    int64_t bt = r / Fo;
    int fi = (int)(((int64_t)fo * Fi) / Fo);  // nearest: floor(fo * Fi / Fo)
    y[i] = x[(bt * Fi + fi) * C + c];
  }
}

template <typename T>
__global__ void resize_f_bwd_kernel(const T* __restrict__ dy, int64_t BT, int Fi, int Fo, int C,
                                    T* __restrict__ dx) {
  int64_t total = BT * Fi * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t r = i / C;
    int fi = (int)(r % Fi);
    int64_t bt = r / Fi;
    // outputs fo with floor(fo*Fi/Fo) == fi  <=>  fo in [ceil(fi*Fo/Fi), ceil((fi+1)*Fo/Fi) )
    int lo = (int)(((int64_t)fi * Fo + Fi - 1) / Fi);
    int hi = (int)(((int64_t)(fi + 1) * Fo + Fi - 1) / Fi);
    float acc = 0.f;
    for (int fo = lo; fo < hi && fo < Fo; ++fo) acc += ld_f(dy + (bt * Fo + fo) * C + c);
    st_f(dx + i, acc);
  }
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }

template <typename T>
__global__ void att_blend_fwd_kernel(const T* __restrict__ x, const T* __restrict__ y,
                                     const float* __restrict__ z, int64_t M, int C,
                                     T* __restrict__ out) {
  int64_t total = M * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t m = i / C;
    float z0 = sigmoidf_(z[2 * m]), z1 = sigmoidf_(z[2 * m + 1]);
    st_f(out + i, ld_f(x + i) * z0 + ld_f(y + i) * z1);
  }
}

// one warp per row: dx, dy elementwise; dz via warp reduction over channels
template <typename T>
__global__ void att_blend_bwd_kernel(const T* __restrict__ x, const T* __restrict__ y,
                                     const float* __restrict__ z, const T* __restrict__ dout,
                                     int64_t M, int C, T* __restrict__ dx, T* __restrict__ dy,
                                     float* __restrict__ dz) {
  int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t m = warp; m < M; m += nwarps) {
    float z0 = sigmoidf_(z[2 * m]), z1 = sigmoidf_(z[2 * m + 1]);
    float s0 = 0.f, s1 = 0.f;
    for (int c = lane; c < C; c += 32) {
      int64_t i = m * C + c;
      float g = ld_f(dout + i);
      float xv = ld_f(x + i), yv = ld_f(y + i);
      st_f(dx + i, g * z0);
      st_f(dy + i, g * z1);
      s0 += g * xv;
      s1 += g * yv;
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    if (lane == 0) {
      dz[2 * m] = s0 * z0 * (1.f - z0);
      dz[2 * m + 1] = s1 * z1 * (1.f - z1);
    }
  }
}

// ------------------------------------------------------------------------------- hcl pooling
// out[b,c,i,j] = mean over h in [floor(i*H/l), ceil((i+1)*H/l)), w likewise; H=F, W=T.
// one block per (b,i,j); threads over channels, coalesced along C.
template <typename T>
__global__ void adaptive_pool_fwd_kernel(const T* __restrict__ x, int B, int Tn, int F, int C,
                                         int l, float* __restrict__ out) {
  int cell = blockIdx.x;
  int j = cell % l, i = (cell / l) % l, b = cell / (l * l);
  int h0 = (i * F) / l, h1 = ((i + 1) * F + l - 1) / l;
  int w0 = (j * Tn) / l, w1 = ((j + 1) * Tn + l - 1) / l;
  float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int w = w0; w < w1; ++w)
      for (int h = h0; h < h1; ++h) acc += ld_f(x + (((int64_t)b * Tn + w) * F + h) * C + c);
    out[(((int64_t)b * C + c) * l + i) * l + j] = acc * inv;
  }
}

template <typename T>
__global__ void adaptive_pool_bwd_kernel(const float* __restrict__ dout, int B, int Tn, int F,
                                         int C, int l, T* __restrict__ dx, int accumulate) {
  int64_t total = (int64_t)B * Tn * F * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(idx % C);
    int64_t r = idx / C;
    int h = (int)(r % F);
    r /= F;
    int w = (int)(r % Tn);
    int b = (int)(r / Tn);
    float acc = 0.f;
    // cells whose window contains (h,w); windows of adjacent cells may overlap by one
    for (int i = 0; i < l; ++i) {
      int h0 = (i * F) / l, h1 = ((i + 1) * F + l - 1) / l;
      if (h < h0 || h >= h1) continue;
      for (int j = 0; j < l; ++j) {
        int w0 = (j * Tn) / l, w1 = ((j + 1) * Tn + l - 1) / l;
        if (w < w0 || w >= w1) continue;
        acc += dout[(((int64_t)b * C + c) * l + i) * l + j] / (float)((h1 - h0) * (w1 - w0));
      }
    }
    float v = accumulate ? ld_f(dx + idx) + acc : acc;
    st_f(dx + idx, v);
  }
}

template <typename TA, typename TB>
__global__ void sqdiff_sum_kernel(const TA* __restrict__ a, const TB* __restrict__ b, int64_t n,
                                  double* __restrict__ out) {
  __shared__ double sh[32];
  double acc = 0.;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float d = ld_f(a + i) - ld_f(b + i);
    acc += (double)(d * d);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

template <typename TA, typename TB, typename TD>
__global__ void sqdiff_bwd_kernel(const TA* __restrict__ a, const TB* __restrict__ b, int64_t n,
                                  const float* __restrict__ gout, float scale, TD* __restrict__ da,
                                  int accumulate) {
  float g = gout[0] * scale * 2.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float v = g * (ld_f(a + i) - ld_f(b + i));
    if (accumulate) v += ld_f(da + i);
    st_f(da + i, v);
  }
}

// ------------------------------------------------------------------------------- optimizer etc
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2,
                            float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float pi = p[i];
    if (wd != 0.f) gi += wd * pi;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    // torch.optim.Adam: p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}


// out = in0 + in1 (+ in2 + in3): gradient accumulation of a tensor with several consumers (what autograd's own
// at::add would do), fp32 accumulation, 16-byte vectors with a scalar tail
template <typename T, int K>
__global__ void sum_n_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c,
                             const T* __restrict__ d, int64_t n, T* __restrict__ out) {
  constexpr int V = 16 / sizeof(T);
  const int64_t nv = n / V;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    float acc[V];
    const T* src[4] = {a, b, c, d};
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const uint4 u = *reinterpret_cast<const uint4*>(src[k] + i * V);
      const T* h = reinterpret_cast<const T*>(&u);
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] += ld_f(h + e);
    }
    uint4 o;
    T* ho = reinterpret_cast<T*>(&o);
#pragma unroll
    for (int e = 0; e < V; ++e) st_f(ho + e, acc[e]);
    *reinterpret_cast<uint4*>(out + i * V) = o;
  }
  for (int64_t i = nv * V + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = ld_f(a + i) + ld_f(b + i);
    if (K > 2) v += ld_f(c + i);
    if (K > 3) v += ld_f(d + i);
    st_f(out + i, v);
  }
}

__global__ void fill_kernel(float* p, int64_t n, float v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = v;
}
__global__ void axpy_kernel(float* y, const float* x, int64_t n, float a) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    y[i] += a * x[i];
}
__global__ void axpby_kernel(const float* __restrict__ x, const float* __restrict__ y, float a,
                             float b, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = a * x[i] + (y ? b * y[i] : 0.f);
}
// flat[offsets[i] .. offsets[i+1]) = tensor i (zeros when its pointer is null)
__global__ void multi_pack_kernel(const unsigned long long* __restrict__ ptrs,
                                  const int64_t* __restrict__ offsets, int n,
                                  float* __restrict__ flat) {
  int64_t total = offsets[n];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {  // last tensor whose offset <= i
      int mid = (lo + hi + 1) >> 1;
      if (offsets[mid] <= i) lo = mid; else hi = mid - 1;
    }
    const float* src = reinterpret_cast<const float*>(ptrs[lo]);
    flat[i] = src ? src[i - offsets[lo]] : 0.f;
  }
}
__global__ void f64_to_f32_kernel(const double* in, int n, double scale, float* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)(in[i] * scale);
}

struct CsGeom {
  int threads, rpb, grid;
};
static int colstats_geom(int64_t M, int C, CsGeom* g) {
  if (C < 1 || C > 1024) return -1;
  if (C <= 256) {
    g->rpb = 256 / C;
    g->threads = 256;
  } else {
    g->rpb = 1;
    g->threads = (C + 31) / 32 * 32;
  }
  int64_t rows_per_block = (int64_t)g->rpb * CS_ROWS_PER_THREAD;
  g->grid = (int)((M + rows_per_block - 1) / rows_per_block);
  if (g->grid < 1) g->grid = 1;
  return 0;
}

}  // namespace
}  // namespace clskd

using namespace clskd;
#define ST ((cudaStream_t)stream)

extern "C" const char* clskd_last_error(void) { return g_err; }
extern "C" int clskd_abi_version(void) { return 1; }

extern "C" int clskd_strided_copy4d(const void* src, int src_dtype, const int64_t* ss, void* dst,
                                    int dst_dtype, const int64_t* ds, const int64_t* shape,
                                    void* stream) {
  CLSKD_CHECK_ARG(src && dst && ss && ds && shape, "clskd_strided_copy4d: null pointer");
  int64_t total = shape[0] * shape[1] * shape[2] * shape[3];
  if (total == 0) return CLSKD_OK;
  if (ss[3] == 1 && ds[3] == 1 && shape[3] % 2 == 0) {
    // vector width: largest of 8/4/2 that divides the row and keeps every access naturally aligned
    const int se = src_dtype == CLSKD_F32 ? 4 : 2, de = dst_dtype == CLSKD_F32 ? 4 : 2;
    auto fits = [&](int v) {
      if (shape[3] % v) return false;
      const int av = v > 4 ? 4 : v;            // elements per single access (ld4/st4 or scalar)
      const int64_t sa = (int64_t)(av == 1 ? 1 : av) * se, da = (int64_t)(av == 1 ? 1 : av) * de;
      if (v == 2) return true;                 // scalar accesses: always aligned
      for (int k = 0; k < 3; ++k)
        if ((ss[k] * se) % sa || (ds[k] * de) % da) return false;
      return ((uintptr_t)src % sa) == 0 && ((uintptr_t)dst % da) == 0;
    };
    const int v = fits(8) ? 8 : (fits(4) ? 4 : 2);
    const int64_t n3v = shape[3] / v;
    const int64_t tv = shape[0] * shape[1] * shape[2] * n3v;
    const int gridv = ew_grid(tv, 256);
    const bool small = tv + (int64_t)gridv * 256 < 4000000000LL;
#define LV3(TS, TD, V, I)                                                                                   \
  strided_copy4d_vec_kernel<TS, TD, V, I><<<gridv, 256, 0, ST>>>((const TS*)src, (TD*)dst, (I)shape[1],       \
                                                                  (I)shape[2], (I)n3v, (I)tv, ss[0], ss[1],  \
                                                                  ss[2], ds[0], ds[1], ds[2])
#define LV2(TS, TD, V) do { if (small) LV3(TS, TD, V, unsigned); else LV3(TS, TD, V, int64_t); } while (0)
#define LV(TS, TD) do { if (v == 8) LV2(TS, TD, 8); else if (v == 4) LV2(TS, TD, 4); else LV2(TS, TD, 2); } while (0)
    if (src_dtype == CLSKD_F32 && dst_dtype == CLSKD_F32) LV(float, float);
    else if (src_dtype == CLSKD_F32) LV(float, __nv_bfloat16);
    else if (dst_dtype == CLSKD_F32) LV(__nv_bfloat16, float);
    else LV(__nv_bfloat16, __nv_bfloat16);
#undef LV
#undef LV2
#undef LV3
    CLSKD_CHECK_LAUNCH("clskd_strided_copy4d(vec)");
    return CLSKD_OK;
  }
  {
    // 4 x 4 block transposes (see transpose4_kernel): dir 0 = [.., .., C, 4] gathered from rows of pitch ss[3], dir 1 =
    // [.., .., 4, C] scattered to rows of pitch ds[2]
    const int se = src_dtype == CLSKD_F32 ? 4 : 2, de = dst_dtype == CLSKD_F32 ? 4 : 2;
    int dir = -1;
    int64_t C = 0, pitch = 0;
    if (shape[3] == 4 && ss[2] == 1 && ds[3] == 1 && ds[2] == 4 && shape[2] % 4 == 0 && ss[3] >= shape[2]) {
      dir = 0; C = shape[2]; pitch = ss[3];
    } else if (shape[2] == 4 && ss[2] == 1 && ss[3] == 4 && ds[3] == 1 && shape[3] % 4 == 0 && ds[2] >= shape[3]) {
      dir = 1; C = shape[3]; pitch = ds[2];
    }
    const int64_t items = dir >= 0 ? shape[0] * shape[1] * (C / 4) : 0;
    const bool al = dir >= 0 && ss[0] % 4 == 0 && ss[1] % 4 == 0 && ds[0] % 4 == 0 && ds[1] % 4 == 0 && pitch % 4 == 0 &&
                    ((uintptr_t)src % (4 * se)) == 0 && ((uintptr_t)dst % (4 * de)) == 0 &&
                    (dir == 0 ? ((uintptr_t)dst % 16) == 0 : ((uintptr_t)src % 16) == 0) &&
                    // the 16-element side is accessed with 16-byte vectors: its unit strides must keep that alignment
                    (dir == 0 ? (ds[0] * de) % 16 == 0 && (ds[1] * de) % 16 == 0 : (ss[0] * se) % 16 == 0 && (ss[1] * se) % 16 == 0);
    if (al && items > 0 && items < 2000000000LL && shape[1] < 2000000000LL) {
      const int gridt = ew_grid(items, 256);
#define LT2(TS, TD, DIR)                                                                                            \
  transpose4_kernel<TS, TD, DIR><<<gridt, 256, 0, ST>>>((const TS*)src, (TD*)dst, (unsigned)shape[1], (unsigned)(C / 4), \
                                                        (unsigned)items, ss[0], ss[1], ds[0], ds[1], pitch)
#define LT(TS, TD) do { if (dir == 0) LT2(TS, TD, 0); else LT2(TS, TD, 1); } while (0)
      if (src_dtype == CLSKD_F32 && dst_dtype == CLSKD_F32) LT(float, float);
      else if (src_dtype == CLSKD_F32) LT(float, __nv_bfloat16);
      else if (dst_dtype == CLSKD_F32) LT(__nv_bfloat16, float);
      else LT(__nv_bfloat16, __nv_bfloat16);
#undef LT
#undef LT2
      CLSKD_CHECK_LAUNCH("clskd_strided_copy4d(transpose4)");
      return CLSKD_OK;
    }
  }
  int grid = ew_grid(total, 256);
#define L(TS, TD)                                                                               \
  strided_copy4d_kernel<TS, TD><<<grid, 256, 0, ST>>>((const TS*)src, (TD*)dst, shape[0], shape[1], \
                                                      shape[2], shape[3], ss[0], ss[1], ss[2], ss[3], \
                                                      ds[0], ds[1], ds[2], ds[3])
  if (src_dtype == CLSKD_F32 && dst_dtype == CLSKD_F32) L(float, float);
  else if (src_dtype == CLSKD_F32) L(float, __nv_bfloat16);
  else if (dst_dtype == CLSKD_F32) L(__nv_bfloat16, float);
  else L(__nv_bfloat16, __nv_bfloat16);
#undef L
  CLSKD_CHECK_LAUNCH("clskd_strided_copy4d");
  return CLSKD_OK;
}

extern "C" int clskd_pack_gather(const float* a, const float* b, const int32_t* table, int64_t n,
                                 void* out, int out_dtype, void* stream) {
  CLSKD_CHECK_ARG(a && table && out, "clskd_pack_gather: null pointer");
  if (n == 0) return CLSKD_OK;
  int grid = ew_grid(n, 256);
  if (out_dtype == CLSKD_F32)
    pack_gather_kernel<float><<<grid, 256, 0, ST>>>(a, b ? b : a, table, n, (float*)out);
  else
    pack_gather_kernel<__nv_bfloat16><<<grid, 256, 0, ST>>>(a, b ? b : a, table, n,
                                                           (__nv_bfloat16*)out);
  CLSKD_CHECK_LAUNCH("clskd_pack_gather");
  return CLSKD_OK;
}

extern "C" int clskd_multi_pack_gather(const int64_t* desc, int n_entries, int64_t total, void* stream) {
  CLSKD_CHECK_ARG(desc || n_entries == 0, "clskd_multi_pack_gather: null pointer");
  if (n_entries <= 0 || total <= 0) return CLSKD_OK;
  multi_pack_gather_kernel<<<ew_grid(total, 256), 256, 0, ST>>>(desc, n_entries, total);
  CLSKD_CHECK_LAUNCH("clskd_multi_pack_gather");
  return CLSKD_OK;
}

extern "C" int clskd_unpack_gather2(const float* src, const int32_t* table2, int64_t n, float* dst,
                                    int accumulate, void* stream) {
  CLSKD_CHECK_ARG(src && table2 && dst, "clskd_unpack_gather2: null pointer");
  if (n == 0) return CLSKD_OK;
  unpack_gather2_kernel<<<ew_grid(n, 256), 256, 0, ST>>>(src, table2, n, dst, accumulate);
  CLSKD_CHECK_LAUNCH("clskd_unpack_gather2");
  return CLSKD_OK;
}

extern "C" int clskd_pad1d(const void* src, int src_dtype, int64_t src_sB, int B, int L, int left,
                           int right, int mode, float* dst, void* stream) {
  CLSKD_CHECK_ARG(src && dst, "clskd_pad1d: null pointer");
  CLSKD_CHECK_ARG(mode == 0 || (left < L && right < L), "clskd_pad1d: reflect pad >= length");
  int64_t total = ((int64_t)L + left + right) * B;
  if (total == 0) return CLSKD_OK;
  if (src_dtype == CLSKD_F32)
    pad1d_kernel<float><<<ew_grid(total, 256), 256, 0, ST>>>((const float*)src, src_sB, B, L, left,
                                                           right, mode, dst);
  else
    pad1d_kernel<__nv_bfloat16><<<ew_grid(total, 256), 256, 0, ST>>>(
        (const __nv_bfloat16*)src, src_sB, B, L, left, right, mode, dst);
  CLSKD_CHECK_LAUNCH("clskd_pad1d");
  return CLSKD_OK;
}

extern "C" int clskd_pad1d_bwd(const float* ddst, int B, int L, int left, int right, int mode,
                               float* dsrc, int accumulate, void* stream) {
  CLSKD_CHECK_ARG(ddst && dsrc, "clskd_pad1d_bwd: null pointer");
  int64_t total = (int64_t)L * B;
  if (total == 0) return CLSKD_OK;
  pad1d_bwd_kernel<<<ew_grid(total, 256), 256, 0, ST>>>(ddst, B, L, left, right, mode, dsrc,
                                                        accumulate);
  CLSKD_CHECK_LAUNCH("clskd_pad1d_bwd");
  return CLSKD_OK;
}

extern "C" int clskd_colstats(const void* x, int dtype, int64_t M, int C, double* sum,
                              double* sumsq, void* stream) {
  CLSKD_CHECK_ARG(x && sum && sumsq, "clskd_colstats: null pointer");
  CsGeom g;
  CLSKD_CHECK_ARG(colstats_geom(M, C, &g) == 0, "clskd_colstats: C=%d unsupported (1..1024)", C);
  {
    cudaError_t e__ = cudaSuccess;
    e__ = zero_spans(ST, sum, sizeof(double) * C, sumsq, sizeof(double) * C);
    if (e__ != cudaSuccess) { set_error("memset failed: %s", cudaGetErrorString(e__)); return CLSKD_ERR_CUDA; }
  }
  if (M == 0) return CLSKD_OK;
  if (vec::colstats(x, nullptr, dtype, 0, M, C, nullptr, nullptr, nullptr, nullptr, nullptr, sum, sumsq, nullptr, ST)) {
    CLSKD_CHECK_LAUNCH("clskd_colstats(vec)");
    return CLSKD_OK;
  }
  size_t sh = 2 * g.threads * sizeof(float);
  CLSKD_DISPATCH_DTYPE(dtype, T,
                       (colstats_kernel<T, 0><<<g.grid, g.threads, sh, ST>>>(
                           (const T*)x, nullptr, M, C, g.rpb, nullptr, nullptr, nullptr, nullptr,
                           nullptr, sum, sumsq, nullptr)));
  CLSKD_CHECK_LAUNCH("clskd_colstats");
  return CLSKD_OK;
}

extern "C" int clskd_bn_finalize(const double* sum, const double* sumsq, int64_t M, int C,
                                 float eps, float momentum, float* mean, float* invstd,
                                 float* running_mean, float* running_var, void* stream) {
  CLSKD_CHECK_ARG(sum && sumsq && mean && invstd && M > 0, "clskd_bn_finalize: bad arguments");
  bn_finalize_kernel<<<cdiv(C, 128), 128, 0, ST>>>(sum, sumsq, M, C, eps, momentum, mean, invstd,
                                                   running_mean, running_var);
  CLSKD_CHECK_LAUNCH("clskd_bn_finalize");
  return CLSKD_OK;
}

extern "C" int clskd_bn_eval_stats(const float* running_mean, const float* running_var, int C,
                                   float eps, float* mean, float* invstd, void* stream) {
  CLSKD_CHECK_ARG(running_mean && running_var && mean && invstd, "clskd_bn_eval_stats: null pointer");
  bn_eval_stats_kernel<<<cdiv(C, 128), 128, 0, ST>>>(running_mean, running_var, C, eps, mean,
                                                     invstd);
  CLSKD_CHECK_LAUNCH("clskd_bn_eval_stats");
  return CLSKD_OK;
}

extern "C" int clskd_bn_fold(const float* running_mean, const float* running_var, const float* gamma,
                             const float* beta, int C, float eps, float* scale, float* shift, void* stream) {
  CLSKD_CHECK_ARG(running_mean && running_var && scale && shift, "clskd_bn_fold: null pointer");
  bn_fold_kernel<<<cdiv(C, 128), 128, 0, ST>>>(running_mean, running_var, gamma, beta, C, eps, scale, shift);
  CLSKD_CHECK_LAUNCH("clskd_bn_fold");
  return CLSKD_OK;
}

extern "C" int clskd_bn_act_fwd(const void* x, int x_dtype, int64_t M, int C, const float* mean,
                                const float* invstd, const float* gamma, const float* beta,
                                const float* slope, void* y, int y_dtype, void* stream) {
  CLSKD_CHECK_ARG(x && y && mean && invstd, "clskd_bn_act_fwd: null pointer");
  int64_t n = M * C;
  if (n == 0) return CLSKD_OK;
  if (vec::bn_act_fwd(x, x_dtype, M, C, mean, invstd, gamma, beta, slope, y, y_dtype, ST)) {
    CLSKD_CHECK_LAUNCH("clskd_bn_act_fwd(vec)");
    return CLSKD_OK;
  }
  bool vec = (C % 4 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0);
#define L(TX, TY)                                                                                 \
  do {                                                                                            \
    if (vec)                                                                                      \
      bn_act_fwd_kernel<TX, TY><<<ew_grid(n / 4, 256), 256, 0, ST>>>(                             \
          (const TX*)x, n / 4, C, mean, invstd, gamma, beta, slope, (TY*)y);                      \
    else                                                                                          \
      bn_act_fwd_scalar_kernel<TX, TY><<<ew_grid(n, 256), 256, 0, ST>>>(                          \
          (const TX*)x, n, C, mean, invstd, gamma, beta, slope, (TY*)y);                          \
  } while (0)
  if (x_dtype == CLSKD_F32 && y_dtype == CLSKD_F32) L(float, float);
  else if (x_dtype == CLSKD_F32) L(float, __nv_bfloat16);
  else if (y_dtype == CLSKD_F32) L(__nv_bfloat16, float);
  else L(__nv_bfloat16, __nv_bfloat16);
#undef L
  CLSKD_CHECK_LAUNCH("clskd_bn_act_fwd");
  return CLSKD_OK;
}

extern "C" int clskd_bn_act_bwd_stats(const void* x, int x_dtype, const void* dy, int dy_dtype,
                                      int64_t M, int C, const float* mean, const float* invstd,
                                      const float* gamma, const float* beta, const float* slope,
                                      double* sum_dz, double* sum_dz_xhat, double* dslope,
                                      void* stream) {
  CLSKD_CHECK_ARG(x && dy && mean && invstd && sum_dz && sum_dz_xhat,
                  "clskd_bn_act_bwd_stats: null pointer");
  CLSKD_CHECK_ARG(x_dtype == dy_dtype, "clskd_bn_act_bwd_stats: x/dy dtype must match");
  CsGeom g;
  CLSKD_CHECK_ARG(colstats_geom(M, C, &g) == 0, "clskd_bn_act_bwd_stats: C=%d unsupported", C);
  {
    cudaError_t e__ = cudaSuccess;
    e__ = zero_spans(ST, sum_dz, sizeof(double) * C, sum_dz_xhat, sizeof(double) * C, dslope, sizeof(double));
    if (e__ != cudaSuccess) { set_error("memset failed: %s", cudaGetErrorString(e__)); return CLSKD_ERR_CUDA; }
  }
  if (M == 0) return CLSKD_OK;
  if (vec::colstats(x, dy, x_dtype, 1, M, C, mean, invstd, gamma, beta, slope, sum_dz, sum_dz_xhat, dslope, ST)) {
    CLSKD_CHECK_LAUNCH("clskd_bn_act_bwd_stats(vec)");
    return CLSKD_OK;
  }
  size_t sh = 3 * g.threads * sizeof(float);
  CLSKD_DISPATCH_DTYPE(x_dtype, T,
                       (colstats_kernel<T, 1><<<g.grid, g.threads, sh, ST>>>(
                           (const T*)x, (const T*)dy, M, C, g.rpb, mean, invstd, gamma, beta, slope,
                           sum_dz, sum_dz_xhat, dslope)));
  CLSKD_CHECK_LAUNCH("clskd_bn_act_bwd_stats");
  return CLSKD_OK;
}

extern "C" int clskd_bn_act_bwd_apply(const void* x, int x_dtype, const void* dy, int dy_dtype,
                                      int64_t M, int C, const float* mean, const float* invstd,
                                      const float* gamma, const float* beta, const float* slope,
                                      const double* sum_dz, const double* sum_dz_xhat,
                                      const double* dslope, int training, void* dx, int dx_dtype,
                                      float* dgamma, float* dbeta, float* dslope_out,
                                      void* stream) {
  CLSKD_CHECK_ARG(x && dy && dx && mean && invstd && sum_dz && sum_dz_xhat,
                  "clskd_bn_act_bwd_apply: null pointer");
  CLSKD_CHECK_ARG(x_dtype == dy_dtype && x_dtype == dx_dtype,
                  "clskd_bn_act_bwd_apply: dtypes must match");
  int64_t n = M * C;
  if (n > 0 && vec::bn_act_bwd_apply(x, dy, x_dtype, M, C, mean, invstd, gamma, beta, slope, sum_dz, sum_dz_xhat,
                                     training, dx, ST)) {
    CLSKD_CHECK_LAUNCH("clskd_bn_act_bwd_apply(vec)");
  } else if (n > 0) {
    CLSKD_DISPATCH_DTYPE(x_dtype, T,
                         (bn_act_bwd_apply_kernel<T, T, T><<<ew_grid(n, 256), 256, 0, ST>>>(
                             (const T*)x, (const T*)dy, n, M, C, mean, invstd, gamma, beta, slope,
                             sum_dz, sum_dz_xhat, training, (T*)dx)));
    CLSKD_CHECK_LAUNCH("clskd_bn_act_bwd_apply");
  }
  if (dgamma || dbeta || dslope_out) {
    bn_param_grads_kernel<<<cdiv(C, 128), 128, 0, ST>>>(sum_dz, sum_dz_xhat, dslope, C, dgamma,
                                                        dbeta, dslope_out);
    CLSKD_CHECK_LAUNCH("clskd_bn_act_bwd_apply(param grads)");
  }
  return CLSKD_OK;
}

extern "C" int clskd_cbn_moments(const void* x, int dtype, int64_t M, int Cc, double* s,
                                 void* stream) {
  CLSKD_CHECK_ARG(x && s, "clskd_cbn_moments: null pointer");
  CsGeom g;
  CLSKD_CHECK_ARG(colstats_geom(M, Cc, &g) == 0, "clskd_cbn_moments: Cc=%d unsupported", Cc);
  {
    cudaError_t e__ = cudaSuccess;
    e__ = cudaMemsetAsync(s, 0, sizeof(double) * 5 * Cc, ST);
    if (e__ != cudaSuccess) { set_error("memset failed: %s", cudaGetErrorString(e__)); return CLSKD_ERR_CUDA; }
  }
  if (M == 0) return CLSKD_OK;
  size_t sh = 5 * g.threads * sizeof(float);
  CLSKD_DISPATCH_DTYPE(dtype, T,
                       (cbn_moments_kernel<T><<<g.grid, g.threads, sh, ST>>>((const T*)x, M, Cc,
                                                                              g.rpb, s)));
  CLSKD_CHECK_LAUNCH("clskd_cbn_moments");
  return CLSKD_OK;
}

extern "C" int clskd_cbn_finalize(const double* s, int64_t M, int Cc, float eps, float momentum,
                                  int training, const float* Wrr, const float* Wri,
                                  const float* Wii, float* RMr, float* RMi, float* RVrr,
                                  float* RVri, float* RVii, float* coef, void* stream) {
  CLSKD_CHECK_ARG(coef && (training ? (s != nullptr && M > 0) : (RMr && RMi && RVrr && RVri && RVii)),
                  "clskd_cbn_finalize: bad arguments");
  cbn_finalize_kernel<<<cdiv(Cc, 128), 128, 0, ST>>>(s, M, Cc, eps, momentum, training, Wrr, Wri,
                                                     Wii, RMr, RMi, RVrr, RVri, RVii, coef);
  CLSKD_CHECK_LAUNCH("clskd_cbn_finalize");
  return CLSKD_OK;
}

extern "C" int clskd_cbn_apply(const void* x, int x_dtype, int64_t M, int Cc, const float* coef,
                               const float* Br, const float* Bi, void* y, int y_dtype,
                               void* stream) {
  CLSKD_CHECK_ARG(x && y && coef, "clskd_cbn_apply: null pointer");
  CLSKD_CHECK_ARG(x_dtype == y_dtype, "clskd_cbn_apply: dtypes must match");
  if (M * Cc == 0) return CLSKD_OK;
  CLSKD_DISPATCH_DTYPE(x_dtype, T,
                       (cbn_apply_kernel<T, T><<<ew_grid(M * Cc, 256), 256, 0, ST>>>(
                           (const T*)x, M, Cc, coef, Br, Bi, (T*)y)));
  CLSKD_CHECK_LAUNCH("clskd_cbn_apply");
  return CLSKD_OK;
}

extern "C" int clskd_cbn_bwd_moments(const void* x, const void* dy, int dtype, int64_t M, int Cc, double* s6,
                                     void* stream) {
  CLSKD_CHECK_ARG(x && dy && s6, "clskd_cbn_bwd_moments: null pointer");
  CsGeom g;
  CLSKD_CHECK_ARG(colstats_geom(M, Cc, &g) == 0, "clskd_cbn_bwd_moments: Cc=%d unsupported", Cc);
  cudaError_t e__ = cudaMemsetAsync(s6, 0, sizeof(double) * 6 * Cc, ST);
  if (e__ != cudaSuccess) { set_error("memset failed: %s", cudaGetErrorString(e__)); return CLSKD_ERR_CUDA; }
  if (M == 0) return CLSKD_OK;
  size_t sh = 6 * g.threads * sizeof(float);
  CLSKD_DISPATCH_DTYPE(dtype, T,
                       (cbn_bwd_moments_kernel<T><<<g.grid, g.threads, sh, ST>>>((const T*)x, (const T*)dy, M, Cc,
                                                                                  g.rpb, s6)));
  CLSKD_CHECK_LAUNCH("clskd_cbn_bwd_moments");
  return CLSKD_OK;
}

extern "C" int clskd_cbn_bwd_finalize(const double* s, const double* s6, int64_t M, int Cc, float eps, int training,
                                      const float* Wrr, const float* Wri, const float* Wii, const float* RMr,
                                      const float* RMi, const float* RVrr, const float* RVri, const float* RVii,
                                      float* coefb, float* dWrr, float* dWri, float* dWii, float* dBr, float* dBi,
                                      void* stream) {
  CLSKD_CHECK_ARG(coefb && s6 && (training ? (s != nullptr && M > 0) : (RMr && RMi && RVrr && RVri && RVii)),
                  "clskd_cbn_bwd_finalize: bad arguments");
  cbn_bwd_finalize_kernel<<<cdiv(Cc, 128), 128, 0, ST>>>(s, s6, M, Cc, eps, training, Wrr, Wri, Wii, RMr, RMi, RVrr,
                                                         RVri, RVii, coefb, dWrr, dWri, dWii, dBr, dBi);
  CLSKD_CHECK_LAUNCH("clskd_cbn_bwd_finalize");
  return CLSKD_OK;
}

extern "C" int clskd_cbn_bwd_apply(const void* x, const void* dy, int dtype, int64_t M, int Cc, const float* coefb,
                                   void* dx, void* stream) {
  CLSKD_CHECK_ARG(x && dy && coefb && dx, "clskd_cbn_bwd_apply: null pointer");
  if (M * Cc == 0) return CLSKD_OK;
  CLSKD_DISPATCH_DTYPE(dtype, T,
                       (cbn_bwd_apply_kernel<T><<<ew_grid(M * Cc, 256), 256, 0, ST>>>((const T*)x, (const T*)dy, M, Cc,
                                                                                       coefb, (T*)dx)));
  CLSKD_CHECK_LAUNCH("clskd_cbn_bwd_apply");
  return CLSKD_OK;
}

extern "C" int clskd_mask_fwd(const float* spec, const void* mask, int mask_dtype, int64_t m_sB,
                              int64_t m_sT, int B, int T, int nbins, int mode, float* out_spec,
                              float* mask_padded, void* stream) {
  CLSKD_CHECK_ARG(spec && mask && out_spec, "clskd_mask_fwd: null pointer");
  CLSKD_CHECK_ARG(mode >= 0 && mode <= 2, "clskd_mask_fwd: mode %d (0='E',1='C',2='R')", mode);
  int64_t total = (int64_t)B * T * nbins;
  if (total == 0) return CLSKD_OK;
  CLSKD_DISPATCH_DTYPE(mask_dtype, TM,
                       (mask_fwd_kernel<TM><<<ew_grid(total, 256), 256, 0, ST>>>(
                           spec, (const TM*)mask, m_sB, m_sT, B, T, nbins, mode, out_spec,
                           mask_padded)));
  CLSKD_CHECK_LAUNCH("clskd_mask_fwd");
  return CLSKD_OK;
}

extern "C" int clskd_mask_bwd(const float* spec, const void* mask, int mask_dtype, int64_t m_sB,
                              int64_t m_sT, int B, int T, int nbins, int mode,
                              const float* dout_spec, void* dmask, int dmask_dtype, int64_t dm_sB,
                              int64_t dm_sT, void* stream) {
  CLSKD_CHECK_ARG(spec && mask && dout_spec && dmask, "clskd_mask_bwd: null pointer");
  CLSKD_CHECK_ARG(mask_dtype == dmask_dtype, "clskd_mask_bwd: mask/dmask dtype must match");
  int64_t total = (int64_t)B * T * (nbins - 1);
  if (total == 0) return CLSKD_OK;
  CLSKD_DISPATCH_DTYPE(mask_dtype, TM,
                       (mask_bwd_kernel<TM, TM><<<ew_grid(total, 256), 256, 0, ST>>>(
                           spec, (const TM*)mask, m_sB, m_sT, B, T, nbins, mode, dout_spec,
                           (TM*)dmask, dm_sB, dm_sT)));
  CLSKD_CHECK_LAUNCH("clskd_mask_bwd");
  return CLSKD_OK;
}

extern "C" int clskd_ola_fwd(const float* frames, const float* window, int B, int T, int win,
                             int hop, int trim, int do_clamp, float* wav, void* stream) {
  CLSKD_CHECK_ARG(win >= hop && hop > 0 && T >= 1 && trim >= 0, "clskd_ola_fwd: bad geometry");
  int L = (T - 1) * hop + win - 2 * trim;
  CLSKD_CHECK_ARG(L >= 0, "clskd_ola_fwd: too few frames");
  int64_t total = (int64_t)B * L;
  if (total == 0) return CLSKD_OK;  // an input shorter than one hop yields an empty waveform
  CLSKD_CHECK_ARG(frames && wav, "clskd_ola_fwd: null pointer");
  ola_fwd_kernel<<<ew_grid(total, 256), 256, 0, ST>>>(frames, window, B, T, win, hop, trim, do_clamp,
                                                      wav);
  CLSKD_CHECK_LAUNCH("clskd_ola_fwd");
  return CLSKD_OK;
}

extern "C" int clskd_ola_bwd(const float* dwav, const float* wav, const float* window, int B, int T,
                             int win, int hop, int trim, int do_clamp, float* dframes,
                             void* stream) {
  CLSKD_CHECK_ARG(dwav && dframes && (wav || !do_clamp), "clskd_ola_bwd: null pointer");
  int64_t total = (int64_t)B * T * win;
  if (total == 0) return CLSKD_OK;
  ola_bwd_kernel<<<ew_grid(total, 256), 256, 0, ST>>>(dwav, wav, window, B, T, win, hop, trim,
                                                      do_clamp, dframes);
  CLSKD_CHECK_LAUNCH("clskd_ola_bwd");
  return CLSKD_OK;
}

extern "C" int clskd_resize_f_fwd(const void* x, int dtype, int64_t BT, int Fi, int Fo, int C,
                                  void* y, void* stream) {
  CLSKD_CHECK_ARG(x && y && Fi > 0 && Fo > 0, "clskd_resize_f_fwd: bad arguments");
  int64_t total = BT * Fo * C;
  if (total == 0) return CLSKD_OK;
  if (vec::resize_f(x, dtype, BT, Fi, Fo, C, y, false, ST)) {
    CLSKD_CHECK_LAUNCH("clskd_resize_f_fwd(vec)");
    return CLSKD_OK;
  }
  CLSKD_DISPATCH_DTYPE(dtype, T,
                       (resize_f_fwd_kernel<T><<<ew_grid(total, 256), 256, 0, ST>>>(
                           (const T*)x, BT, Fi, Fo, C, (T*)y)));
  CLSKD_CHECK_LAUNCH("clskd_resize_f_fwd");
  return CLSKD_OK;
}

extern "C" int clskd_resize_f_bwd(const void* dy, int dtype, int64_t BT, int Fi, int Fo, int C,
                                  void* dx, void* stream) {
  CLSKD_CHECK_ARG(dy && dx && Fi > 0 && Fo > 0, "clskd_resize_f_bwd: bad arguments");
  int64_t total = BT * Fi * C;
  if (total == 0) return CLSKD_OK;
  if (vec::resize_f(dy, dtype, BT, Fi, Fo, C, dx, true, ST)) {
    CLSKD_CHECK_LAUNCH("clskd_resize_f_bwd(vec)");
    return CLSKD_OK;
  }
  CLSKD_DISPATCH_DTYPE(dtype, T,
                       (resize_f_bwd_kernel<T><<<ew_grid(total, 256), 256, 0, ST>>>(
                           (const T*)dy, BT, Fi, Fo, C, (T*)dx)));
  CLSKD_CHECK_LAUNCH("clskd_resize_f_bwd");
  return CLSKD_OK;
}

extern "C" int clskd_att_blend_fwd(const void* x, const void* y, int dtype, const float* z,
                                   int64_t M, int C, void* out, void* stream) {
  CLSKD_CHECK_ARG(x && y && z && out, "clskd_att_blend_fwd: null pointer");
  if (M * C == 0) return CLSKD_OK;
  if (vec::att_blend_fwd(x, y, dtype, z, M, C, out, ST)) {
    CLSKD_CHECK_LAUNCH("clskd_att_blend_fwd(vec)");
    return CLSKD_OK;
  }
  CLSKD_DISPATCH_DTYPE(dtype, T,
                       (att_blend_fwd_kernel<T><<<ew_grid(M * C, 256), 256, 0, ST>>>(
                           (const T*)x, (const T*)y, z, M, C, (T*)out)));
  CLSKD_CHECK_LAUNCH("clskd_att_blend_fwd");
  return CLSKD_OK;
}

extern "C" int clskd_att_blend_bwd(const void* x, const void* y, int dtype, const float* z,
                                   const void* dout, int64_t M, int C, void* dx, void* dy,
                                   float* dz, void* stream) {
  CLSKD_CHECK_ARG(x && y && z && dout && dx && dy && dz, "clskd_att_blend_bwd: null pointer");
  if (M * C == 0) return CLSKD_OK;
  if (vec::att_blend_bwd(x, y, dtype, z, dout, M, C, dx, dy, dz, ST)) {
    CLSKD_CHECK_LAUNCH("clskd_att_blend_bwd(vec)");
    return CLSKD_OK;
  }
  CLSKD_DISPATCH_DTYPE(dtype, T,
                       (att_blend_bwd_kernel<T><<<ew_grid(M * 32, 256), 256, 0, ST>>>(
                           (const T*)x, (const T*)y, z, (const T*)dout, M, C, (T*)dx, (T*)dy, dz)));
  CLSKD_CHECK_LAUNCH("clskd_att_blend_bwd");
  return CLSKD_OK;
}

extern "C" int clskd_adaptive_pool_fwd(const void* x, int dtype, int B, int T, int F, int C, int l,
                                       float* out, void* stream) {
  CLSKD_CHECK_ARG(x && out && l >= 1, "clskd_adaptive_pool_fwd: bad arguments");
  if ((int64_t)B * T * F * C == 0) return CLSKD_OK;
  CLSKD_DISPATCH_DTYPE(dtype, TT,
                       (adaptive_pool_fwd_kernel<TT><<<B * l * l, 128, 0, ST>>>((const TT*)x, B, T,
                                                                                 F, C, l, out)));
  CLSKD_CHECK_LAUNCH("clskd_adaptive_pool_fwd");
  return CLSKD_OK;
}

extern "C" int clskd_adaptive_pool_bwd(const float* dout, int B, int T, int F, int C, int l,
                                       void* dx, int dtype, int accumulate, void* stream) {
  CLSKD_CHECK_ARG(dout && dx && l >= 1, "clskd_adaptive_pool_bwd: bad arguments");
  int64_t total = (int64_t)B * T * F * C;
  if (total == 0) return CLSKD_OK;
  CLSKD_DISPATCH_DTYPE(dtype, TT,
                       (adaptive_pool_bwd_kernel<TT><<<ew_grid(total, 256), 256, 0, ST>>>(
                           dout, B, T, F, C, l, (TT*)dx, accumulate)));
  CLSKD_CHECK_LAUNCH("clskd_adaptive_pool_bwd");
  return CLSKD_OK;
}

extern "C" int clskd_sqdiff_sum(const void* a, int a_dtype, const void* b, int b_dtype, int64_t n,
                                double* out, void* stream) {
  CLSKD_CHECK_ARG(a && b && out, "clskd_sqdiff_sum: null pointer");
  if (n == 0) return CLSKD_OK;
  int grid = ew_grid(n, 256);
  if (grid > sm_count() * 4) grid = sm_count() * 4;
#define L(TA, TB) sqdiff_sum_kernel<TA, TB><<<grid, 256, 0, ST>>>((const TA*)a, (const TB*)b, n, out)
  if (a_dtype == CLSKD_F32 && b_dtype == CLSKD_F32) L(float, float);
  else if (a_dtype == CLSKD_F32) L(float, __nv_bfloat16);
  else if (b_dtype == CLSKD_F32) L(__nv_bfloat16, float);
  else L(__nv_bfloat16, __nv_bfloat16);
#undef L
  CLSKD_CHECK_LAUNCH("clskd_sqdiff_sum");
  return CLSKD_OK;
}

extern "C" int clskd_sqdiff_bwd(const void* a, int a_dtype, const void* b, int b_dtype, int64_t n,
                                const float* gout, float scale, void* da, int da_dtype,
                                int accumulate, void* stream) {
  CLSKD_CHECK_ARG(a && b && gout && da, "clskd_sqdiff_bwd: null pointer");
  CLSKD_CHECK_ARG(a_dtype == da_dtype, "clskd_sqdiff_bwd: a/da dtype must match");
  if (n == 0) return CLSKD_OK;
  int grid = ew_grid(n, 256);
#define L(TA, TB)                                                                              \
  sqdiff_bwd_kernel<TA, TB, TA><<<grid, 256, 0, ST>>>((const TA*)a, (const TB*)b, n, gout, scale, \
                                                      (TA*)da, accumulate)
  if (a_dtype == CLSKD_F32 && b_dtype == CLSKD_F32) L(float, float);
  else if (a_dtype == CLSKD_F32) L(float, __nv_bfloat16);
  else if (b_dtype == CLSKD_F32) L(__nv_bfloat16, float);
  else L(__nv_bfloat16, __nv_bfloat16);
#undef L
  CLSKD_CHECK_LAUNCH("clskd_sqdiff_bwd");
  return CLSKD_OK;
}

extern "C" int clskd_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr,
                               double beta1, double beta2, double eps, double weight_decay, int step,
                               double grad_scale, void* stream) {
  CLSKD_CHECK_ARG(p && g && m && v && step >= 1, "clskd_adam_step: bad arguments");
  if (n == 0) return CLSKD_OK;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  adam_kernel<<<ew_grid(n, 256), 256, 0, ST>>>(p, g, m, v, n, (float)lr, (float)beta1, (float)beta2,
                                               (float)eps, (float)weight_decay, (float)bc1,
                                               (float)sqrt(bc2), (float)grad_scale);
  CLSKD_CHECK_LAUNCH("clskd_adam_step");
  return CLSKD_OK;
}

extern "C" int clskd_fill_f32(float* p, int64_t n, float v, void* stream) {
  CLSKD_CHECK_ARG(p || n == 0, "clskd_fill_f32: null pointer");
  if (n == 0) return CLSKD_OK;
  fill_kernel<<<ew_grid(n, 256), 256, 0, ST>>>(p, n, v);
  CLSKD_CHECK_LAUNCH("clskd_fill_f32");
  return CLSKD_OK;
}

extern "C" int clskd_axpy_f32(float* y, const float* x, int64_t n, float a, void* stream) {
  CLSKD_CHECK_ARG((y && x) || n == 0, "clskd_axpy_f32: null pointer");
  if (n == 0) return CLSKD_OK;
  axpy_kernel<<<ew_grid(n, 256), 256, 0, ST>>>(y, x, n, a);
  CLSKD_CHECK_LAUNCH("clskd_axpy_f32");
  return CLSKD_OK;
}

extern "C" int clskd_axpby_f32(const float* x, const float* y, float a, float b, float* out,
                               int64_t n, void* stream) {
  CLSKD_CHECK_ARG((x && out) || n == 0, "clskd_axpby_f32: null pointer");
  if (n == 0) return CLSKD_OK;
  axpby_kernel<<<ew_grid(n, 256), 256, 0, ST>>>(x, y, a, b, out, n);
  CLSKD_CHECK_LAUNCH("clskd_axpby_f32");
  return CLSKD_OK;
}

extern "C" int clskd_multi_pack_f32(const void* ptrs, const int64_t* offsets, int n, int64_t total,
                                    float* flat, void* stream) {
  CLSKD_CHECK_ARG(ptrs && offsets && flat && n >= 1, "clskd_multi_pack_f32: bad arguments");
  if (total == 0) return CLSKD_OK;
  multi_pack_kernel<<<ew_grid(total, 256), 256, 0, ST>>>((const unsigned long long*)ptrs, offsets, n,
                                                        flat);
  CLSKD_CHECK_LAUNCH("clskd_multi_pack_f32");
  return CLSKD_OK;
}

extern "C" int clskd_f64_to_f32(const double* in, int n, double scale, float* out, void* stream) {
  CLSKD_CHECK_ARG(in && out, "clskd_f64_to_f32: null pointer");
  if (n == 0) return CLSKD_OK;
  f64_to_f32_kernel<<<cdiv(n, 128), 128, 0, ST>>>(in, n, scale, out);
  CLSKD_CHECK_LAUNCH("clskd_f64_to_f32");
  return CLSKD_OK;
}

extern "C" int clskd_sum_n(const void* in0, const void* in1, const void* in2, const void* in3, int k, int dtype,
                           int64_t n, void* out, void* stream) {
  CLSKD_CHECK_ARG(in0 && in1 && out && k >= 2 && k <= 4 && (k < 3 || in2) && (k < 4 || in3), "clskd_sum_n: bad arguments");
  CLSKD_CHECK_ARG(!(((uintptr_t)in0 | (uintptr_t)in1 | (uintptr_t)in2 | (uintptr_t)in3 | (uintptr_t)out) & 15),
                  "clskd_sum_n: tensors must be 16-byte aligned");
  if (n == 0) return CLSKD_OK;
  const int grid = ew_grid(n / 4 + 1, 256);
#define L(T, K) sum_n_kernel<T, K><<<grid, 256, 0, ST>>>((const T*)in0, (const T*)in1, (const T*)in2, (const T*)in3, n, (T*)out)
  if (dtype == CLSKD_F32) { if (k == 2) L(float, 2); else if (k == 3) L(float, 3); else L(float, 4); }
  else { if (k == 2) L(__nv_bfloat16, 2); else if (k == 3) L(__nv_bfloat16, 3); else L(__nv_bfloat16, 4); }
#undef L
  CLSKD_CHECK_LAUNCH("clskd_sum_n");
  return CLSKD_OK;
}
