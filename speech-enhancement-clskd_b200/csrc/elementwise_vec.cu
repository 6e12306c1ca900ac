// 128-bit vectorised versions of the HBM-bound channels-last kernels (BatchNorm statistics /
// normalise+PReLU / backward, ABF nearest-resize and attention blend).  A thread always owns the
// same 8-channel group, so per-channel constants live in registers and every global access is one
// 16-byte (bf16) or two 16-byte (fp32) transactions; a row of C channels is covered by C/8
// adjacent threads (fully coalesced).  Used when C % 8 == 0 and the tensors are 16-byte aligned;
// the scalar kernels in elementwise.cu remain the general fallback.
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

constexpr int VT = 256;   // threads per block

// ------------------------------------------------------------------------------- column statistics
template <typename T, int MODE>
__global__ void __launch_bounds__(VT, MODE == 1 ? 3 : 1) colstats_vec_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                          int64_t M, int C, const float* __restrict__ mean,
                                                          const float* __restrict__ invstd,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta,
                                                          const float* __restrict__ slope, double* __restrict__ out0,
                                                          double* __restrict__ out1, double* __restrict__ out2) {
  extern __shared__ float red[];   // [2][C] (+1)
  const int tpr = C >> 3, rows_par = VT / tpr;
  const int cg = threadIdx.x % tpr, rs = threadIdx.x / tpr;
  for (int i = threadIdx.x; i < 2 * C + 1; i += VT) red[i] = 0.f;
  __syncthreads();
  float a0[8], a1[8], a2 = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) a0[e] = a1[e] = 0.f;
  float nmi[8], is[8], g[8], bt[8];          // nmi = -mean * invstd: xhat = x * invstd + nmi
  float sl = 1.f;
  float2 a2v = make_float2(0.f, 0.f);
  if (MODE == 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = cg * 8 + e;
      is[e] = invstd[c];
      nmi[e] = -mean[c] * is[e];
      g[e] = gamma ? gamma[c] : 1.f;
      bt[e] = beta ? beta[c] : 0.f;
    }
    sl = slope ? slope[0] : 1.f;
  }
  if (rs < rows_par) {
    // four rows in flight per thread (raw 16-byte loads first): the grid is two CTAs per SM, so that the per-column
    // fp64 atomics at the end - about 20 ns each on one address - stay a few microseconds
    constexpr int UR = MODE == 0 ? 8 : 4;       // one input tensor in MODE 0: twice the rows for the same bytes in flight
    constexpr int NV = sizeof(T) == 2 ? 1 : 2;                 // 16-byte vectors per 8 channels
    const int64_t step = (int64_t)gridDim.x * rows_par;
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
    for (int64_t m0 = (int64_t)blockIdx.x * rows_par + rs; m0 < M; m0 += UR * step) {
      uint4 rx[UR][NV], rd[UR][NV];
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        const int64_t m = m0 + u * step;
        const bool ok = m < M;
        const uint4* px = reinterpret_cast<const uint4*>(x + (ok ? m : 0) * C + cg * 8);
#pragma unroll
        for (int q = 0; q < NV; ++q) rx[u][q] = ok ? px[q] : z4;
        if (MODE == 1) {
          const uint4* pd = reinterpret_cast<const uint4*>(dy + (ok ? m : 0) * C + cg * 8);
#pragma unroll
          for (int q = 0; q < NV; ++q) rd[u][q] = ok ? pd[q] : z4;
        }
      }
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        float v[8];
        ld8(reinterpret_cast<const T*>(&rx[u][0]), v);
        if (MODE == 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            a0[e] += v[e];
            a1[e] = fmaf(v[e], v[e], a1[e]);
          }
        } else {
          float d[8];
          ld8(reinterpret_cast<const T*>(&rd[u][0]), d);
          // packed fp32 pairs (FFMA2 / FADD2 / FMUL2: one issue slot per two channels); the PReLU branch is the
          // 0/1 step st = (u > 0):  dz = d (sl + (1 - sl) st),  dslope += d (u - u st)
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const float2 v2 = make_float2(v[e], v[e + 1]), d2 = make_float2(d[e], d[e + 1]);
            const float2 xh = __ffma2_rn(v2, make_float2(is[e], is[e + 1]), make_float2(nmi[e], nmi[e + 1]));
            const float2 uu = __ffma2_rn(xh, make_float2(g[e], g[e + 1]), make_float2(bt[e], bt[e + 1]));
            const float2 st = make_float2(uu.x > 0.f ? 1.f : 0.f, uu.y > 0.f ? 1.f : 0.f);
            const float2 dz = __fmul2_rn(d2, __ffma2_rn(st, make_float2(1.f - sl, 1.f - sl), make_float2(sl, sl)));
            const float2 s0 = __fadd2_rn(make_float2(a0[e], a0[e + 1]), dz);
            const float2 s1 = __ffma2_rn(dz, xh, make_float2(a1[e], a1[e + 1]));
            a0[e] = s0.x; a0[e + 1] = s0.y;
            a1[e] = s1.x; a1[e + 1] = s1.y;
            const float2 ng = __ffma2_rn(make_float2(-uu.x, -uu.y), st, uu);      // min(u, 0)
            a2v = __ffma2_rn(d2, ng, a2v);
          }
        }
      }
    }
    // Per-CTA reduction.  Float atomics on shared memory are compare-and-swap loops: with C = 16 the 128 row slots of a
    // CTA hit each column 128 ways and the loops retried 26 times on average (ncu: 1 M shared-atomic instructions for
    // 38 K additions, a fixed ~25 us per launch, 43 launches per step).  When the lanes of a row are an aligned
    // power-of-two group, the row slots of a warp are first summed with shuffles and one lane group per warp adds.
    float a2t = a2 + a2v.x + a2v.y;
    const bool pow2 = (tpr & (tpr - 1)) == 0 && tpr <= 32;      // (then every thread of the CTA is inside this branch)
    if (pow2) {
      for (int o = tpr; o < 32; o <<= 1) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          a0[e] += __shfl_xor_sync(0xffffffffu, a0[e], o);
          a1[e] += __shfl_xor_sync(0xffffffffu, a1[e], o);
        }
        if (MODE == 1) a2t += __shfl_xor_sync(0xffffffffu, a2t, o);
      }
      if (MODE == 1)
        for (int o = 1; o < tpr; o <<= 1) a2t += __shfl_xor_sync(0xffffffffu, a2t, o);
    }
    if (!pow2 || (threadIdx.x & 31) < tpr) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        atomicAdd(&red[cg * 8 + e], a0[e]);
        atomicAdd(&red[C + cg * 8 + e], a1[e]);
      }
    }
    if (MODE == 1 && (!pow2 || (threadIdx.x & 31) == 0)) atomicAdd(&red[2 * C], a2t);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += VT) {
    atomicAdd(out0 + c, (double)red[c]);
    atomicAdd(out1 + c, (double)red[C + c]);
  }
  if (MODE == 1 && out2 && threadIdx.x == 0) atomicAdd(out2, (double)red[2 * C]);
}

// ------------------------------------------------------------------------------- normalise + PReLU
template <typename TX, typename TY>
__global__ void __launch_bounds__(VT) bn_act_fwd_vec_kernel(const TX* __restrict__ x, int64_t M, int C,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            const float* __restrict__ slope, TY* __restrict__ y) {
  const int tpr = C >> 3, rows_par = VT / tpr;
  const int cg = threadIdx.x % tpr, rs = threadIdx.x / tpr;
  if (rs >= rows_par) return;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = cg * 8 + e;
    sc[e] = invstd[c] * (gamma ? gamma[c] : 1.f);
    sh[e] = (beta ? beta[c] : 0.f) - mean[c] * sc[e];
  }
  const float sl = slope ? slope[0] : 1.f;
  // (one row per iteration: four rows in flight measured 1.95 -> 2.75 ms per step here, while the same change took the
  // backward apply pass below from 3.22 to 2.76 ms)
  for (int64_t m = (int64_t)blockIdx.x * rows_par + rs; m < M; m += (int64_t)gridDim.x * rows_par) {
    float v[8];
    ld8(x + m * C + cg * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float u = fmaf(v[e], sc[e], sh[e]);
      v[e] = u > 0.f ? u : u * sl;
    }
    st8(y + m * C + cg * 8, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(VT) bn_act_bwd_apply_vec_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                  int64_t M, int C, const float* __restrict__ mean,
                                                                  const float* __restrict__ invstd,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta,
                                                                  const float* __restrict__ slope,
                                                                  const double* __restrict__ sum_dz,
                                                                  const double* __restrict__ sum_dz_xhat, int training,
                                                                  T* __restrict__ dx) {
  const int tpr = C >> 3, rows_par = VT / tpr;
  const int cg = threadIdx.x % tpr, rs = threadIdx.x / tpr;
  if (rs >= rows_par) return;
  float mu[8], is[8], g[8], bt[8], k1[8], k2[8];
  const float invM = 1.f / (float)M;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = cg * 8 + e;
    mu[e] = mean[c];
    is[e] = invstd[c];
    g[e] = gamma ? gamma[c] : 1.f;
    bt[e] = beta ? beta[c] : 0.f;
    k1[e] = training ? (float)sum_dz[c] * invM : 0.f;
    k2[e] = training ? (float)sum_dz_xhat[c] * invM : 0.f;
  }
  const float sl = slope ? slope[0] : 1.f;
  // UR rows in flight per thread (raw 16-byte loads first)
  constexpr int NV = sizeof(T) == 2 ? 1 : 2;
  constexpr int UR = sizeof(T) == 2 ? 4 : 2;
  const int64_t step = (int64_t)gridDim.x * rows_par;
  for (int64_t m0 = (int64_t)blockIdx.x * rows_par + rs; m0 < M; m0 += UR * step) {
    uint4 rx[UR][NV], rd[UR][NV];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const int64_t m = m0 + u * step;
      if (m < M) {
        const uint4* px = reinterpret_cast<const uint4*>(x + m * C + cg * 8);
        const uint4* pd = reinterpret_cast<const uint4*>(dy + m * C + cg * 8);
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          rx[u][q] = px[q];
          rd[u][q] = pd[q];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const int64_t m = m0 + u * step;
      if (m < M) {
        float v[8], d[8];
        ld8(reinterpret_cast<const T*>(&rx[u][0]), v);
        ld8(reinterpret_cast<const T*>(&rd[u][0]), d);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = (v[e] - mu[e]) * is[e];
          const float uu = fmaf(xh, g[e], bt[e]);
          const float dz = uu > 0.f ? d[e] : d[e] * sl;
          v[e] = g[e] * is[e] * (dz - k1[e] - xh * k2[e]);
        }
        st8(dx + m * C + cg * 8, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------- ABF helpers
template <typename T, typename I>
__global__ void __launch_bounds__(VT) resize_f_fwd_vec_kernel(const T* __restrict__ x, int64_t BT, int Fi, int Fo,
                                                              int C, T* __restrict__ y) {
  const I tpr = C >> 3;
  const I total = (I)BT * Fo * tpr;
  for (I i = (I)blockIdx.x * VT + threadIdx.x; i < total; i += (I)gridDim.x * VT) {
    const int cg = (int)(i % tpr);
    const I r = i / tpr;
    const int fo = (int)(r % (I)Fo);
    const int64_t bt = r / (I)Fo;
    const int fi = (int)(((int64_t)fo * Fi) / Fo);
    *reinterpret_cast<uint4*>(y + i * 8) = *reinterpret_cast<const uint4*>(x + ((bt * Fi + fi) * C + cg * 8));
    if (sizeof(T) == 4)
      *(reinterpret_cast<uint4*>(y + i * 8) + 1) = *(reinterpret_cast<const uint4*>(x + ((bt * Fi + fi) * C + cg * 8)) + 1);
  }
}

template <typename T, typename I>
__global__ void __launch_bounds__(VT) resize_f_bwd_vec_kernel(const T* __restrict__ dy, int64_t BT, int Fi, int Fo,
                                                              int C, T* __restrict__ dx) {
  const I tpr = C >> 3;
  const I total = (I)BT * Fi * tpr;
  for (I i = (I)blockIdx.x * VT + threadIdx.x; i < total; i += (I)gridDim.x * VT) {
    const int cg = (int)(i % tpr);
    const I r = i / tpr;
    const int fi = (int)(r % (I)Fi);
    const int64_t bt = r / (I)Fi;
    const int lo = (int)(((int64_t)fi * Fo + Fi - 1) / Fi);
    int hi = (int)(((int64_t)(fi + 1) * Fo + Fi - 1) / Fi);
    if (hi > Fo) hi = Fo;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int fo = lo; fo < hi; ++fo) {
      float v[8];
      ld8(dy + ((bt * Fo + fo) * C + cg * 8), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
    st8(dx + i * 8, acc);
  }
}

__device__ __forceinline__ float sigm(float v) { return 1.f / (1.f + __expf(-v)); }

template <typename T, typename I>
__global__ void __launch_bounds__(VT) att_blend_fwd_vec_kernel(const T* __restrict__ x, const T* __restrict__ y,
                                                               const float* __restrict__ z, int64_t M, int C,
                                                               T* __restrict__ out) {
  const I tpr = C >> 3;
  const I total = (I)M * tpr;
  for (I i = (I)blockIdx.x * VT + threadIdx.x; i < total; i += (I)gridDim.x * VT) {
    const int64_t m = i / tpr;
    const float2 zz = *reinterpret_cast<const float2*>(z + 2 * m);
    const float z0 = sigm(zz.x), z1 = sigm(zz.y);
    float a[8], b[8];
    ld8(x + i * 8, a);
    ld8(y + i * 8, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = a[e] * z0 + b[e] * z1;
    st8(out + i * 8, a);
  }
}

// tpr (a power of two <= 32) adjacent lanes share a row: dx, dy 8-wide; dz by a sub-warp shuffle sum
template <typename T>
__global__ void __launch_bounds__(VT) att_blend_bwd_vec_kernel(const T* __restrict__ x, const T* __restrict__ y,
                                                               const float* __restrict__ z, const T* __restrict__ dout,
                                                               int64_t M, int C, T* __restrict__ dx, T* __restrict__ dy,
                                                               float* __restrict__ dz) {
  const int tpr = C >> 3;
  const int64_t total = M * tpr;
  const int64_t stride = (int64_t)gridDim.x * VT;
  const int64_t rounds = (total + stride - 1) / stride;
  for (int64_t it = 0; it < rounds; ++it) {
    const int64_t i = it * stride + (int64_t)blockIdx.x * VT + threadIdx.x;
    const bool live = i < total;
    const int64_t ii = live ? i : 0;
    const int64_t m = ii < 0x7fffffffLL ? (int64_t)((unsigned)ii / (unsigned)tpr) : ii / tpr;
    const float2 zz = *reinterpret_cast<const float2*>(z + 2 * m);
    const float z0 = sigm(zz.x), z1 = sigm(zz.y);
    float g[8], a[8], b[8];
    ld8(dout + ii * 8, g);
    ld8(x + ii * 8, a);
    ld8(y + ii * 8, b);
    float s0 = 0.f, s1 = 0.f, o0[8], o1[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s0 = fmaf(g[e], a[e], s0);
      s1 = fmaf(g[e], b[e], s1);
      o0[e] = g[e] * z0;
      o1[e] = g[e] * z1;
    }
    if (live) {
      st8(dx + ii * 8, o0);
      st8(dy + ii * 8, o1);
    } else {
      s0 = s1 = 0.f;
    }
    for (int o = tpr >> 1; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (live && (ii % tpr) == 0) {
      dz[2 * m] = s0 * z0 * (1.f - z0);
      dz[2 * m + 1] = s1 * z1 * (1.f - z1);
    }
  }
}

inline bool al16(const void* p) { return ((uintptr_t)p % 16) == 0; }
inline bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

inline int row_grid(int64_t M, int rows_par) {
  int64_t blocks = (M + (int64_t)rows_par * 8 - 1) / ((int64_t)rows_par * 8);
  int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}
inline int flat_grid(int64_t items) {
  int64_t blocks = (items + VT - 1) / VT;
  int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

namespace vec {

bool colstats(const void* x, const void* dy, int dtype, int mode, int64_t M, int C, const float* mean,
              const float* invstd, const float* gamma, const float* beta, const float* slope, double* o0,
              double* o1, double* o2, cudaStream_t st) {
  if (C % 8 || C > 2048 || !al16(x) || (dy && !al16(dy)) || M < 1) return false;
  const int rows_par = VT / (C / 8);
  int grid = row_grid(M, rows_par * 4);
  // one wave: the statistics-only pass (43 registers) holds four CTAs per SM - 1.04 -> 0.80 ms per step over two; the
  // backward statistics pass (100 registers) two.  (The per-column fp64 atomics of a few hundred CTAs at the end cost
  // 2-4 us: tools/hwtests/atomic_tail_test.cu.)
  const int per_sm = mode == 0 ? 4 : 3;
  if (grid > per_sm * sm_count()) grid = per_sm * sm_count();
  const size_t sh = sizeof(float) * (2 * (size_t)C + 1);
  if (mode == 0) {
    CLSKD_DISPATCH_DTYPE(dtype, T, (colstats_vec_kernel<T, 0><<<grid, VT, sh, st>>>(
                                       (const T*)x, nullptr, M, C, nullptr, nullptr, nullptr, nullptr, nullptr, o0, o1, nullptr)));
  } else {
    CLSKD_DISPATCH_DTYPE(dtype, T, (colstats_vec_kernel<T, 1><<<grid, VT, sh, st>>>(
                                       (const T*)x, (const T*)dy, M, C, mean, invstd, gamma, beta, slope, o0, o1, o2)));
  }
  return true;
}

bool bn_act_fwd(const void* x, int x_dtype, int64_t M, int C, const float* mean, const float* invstd,
                const float* gamma, const float* beta, const float* slope, void* y, int y_dtype, cudaStream_t st) {
  if (C % 8 || C > 2048 || !al16(x) || !al16(y) || M < 1) return false;
  const int grid = row_grid(M, VT / (C / 8));
#define L(TX, TY) bn_act_fwd_vec_kernel<TX, TY><<<grid, VT, 0, st>>>((const TX*)x, M, C, mean, invstd, gamma, beta, slope, (TY*)y)
  if (x_dtype == CLSKD_F32 && y_dtype == CLSKD_F32) L(float, float);
  else if (x_dtype == CLSKD_F32) L(float, __nv_bfloat16);
  else if (y_dtype == CLSKD_F32) L(__nv_bfloat16, float);
  else L(__nv_bfloat16, __nv_bfloat16);
#undef L
  return true;
}

bool bn_act_bwd_apply(const void* x, const void* dy, int dtype, int64_t M, int C, const float* mean,
                      const float* invstd, const float* gamma, const float* beta, const float* slope,
                      const double* sum_dz, const double* sum_dz_xhat, int training, void* dx, cudaStream_t st) {
  if (C % 8 || C > 2048 || !al16(x) || !al16(dy) || !al16(dx) || M < 1) return false;
  const int grid = row_grid(M, VT / (C / 8));
  CLSKD_DISPATCH_DTYPE(dtype, T, (bn_act_bwd_apply_vec_kernel<T><<<grid, VT, 0, st>>>(
                                     (const T*)x, (const T*)dy, M, C, mean, invstd, gamma, beta, slope, sum_dz,
                                     sum_dz_xhat, training, (T*)dx)));
  return true;
}

bool resize_f(const void* src, int dtype, int64_t BT, int Fi, int Fo, int C, void* dst, bool backward, cudaStream_t st) {
  if (C % 8 || !al16(src) || !al16(dst)) return false;
  const int64_t items = BT * (backward ? Fi : Fo) * (C / 8);
  const int grid = flat_grid(items);
  const bool small = items + (int64_t)grid * VT < 4000000000LL;
  if (backward) {
    if (small) { CLSKD_DISPATCH_DTYPE(dtype, T, (resize_f_bwd_vec_kernel<T, unsigned><<<grid, VT, 0, st>>>((const T*)src, BT, Fi, Fo, C, (T*)dst))); }
    else { CLSKD_DISPATCH_DTYPE(dtype, T, (resize_f_bwd_vec_kernel<T, int64_t><<<grid, VT, 0, st>>>((const T*)src, BT, Fi, Fo, C, (T*)dst))); }
  } else {
    if (small) { CLSKD_DISPATCH_DTYPE(dtype, T, (resize_f_fwd_vec_kernel<T, unsigned><<<grid, VT, 0, st>>>((const T*)src, BT, Fi, Fo, C, (T*)dst))); }
    else { CLSKD_DISPATCH_DTYPE(dtype, T, (resize_f_fwd_vec_kernel<T, int64_t><<<grid, VT, 0, st>>>((const T*)src, BT, Fi, Fo, C, (T*)dst))); }
  }
  return true;
}

bool att_blend_fwd(const void* x, const void* y, int dtype, const float* z, int64_t M, int C, void* out, cudaStream_t st) {
  if (C % 8 || !al16(x) || !al16(y) || !al16(out) || ((uintptr_t)z % 8)) return false;
  const int grid = flat_grid(M * (C / 8));
  if (M * (C / 8) + (int64_t)grid * VT < 4000000000LL) {
    CLSKD_DISPATCH_DTYPE(dtype, T, (att_blend_fwd_vec_kernel<T, unsigned><<<grid, VT, 0, st>>>((const T*)x, (const T*)y, z, M, C, (T*)out)));
  } else {
    CLSKD_DISPATCH_DTYPE(dtype, T, (att_blend_fwd_vec_kernel<T, int64_t><<<grid, VT, 0, st>>>((const T*)x, (const T*)y, z, M, C, (T*)out)));
  }
  return true;
}

bool att_blend_bwd(const void* x, const void* y, int dtype, const float* z, const void* dout, int64_t M, int C,
                   void* dx, void* dy, float* dz, cudaStream_t st) {
  const int tpr = C / 8;
  if (C % 8 || !pow2(tpr) || tpr > 32 || !al16(x) || !al16(y) || !al16(dout) || !al16(dx) || !al16(dy) ||
      ((uintptr_t)z % 8))
    return false;
  const int grid = flat_grid(M * tpr);
  CLSKD_DISPATCH_DTYPE(dtype, T, (att_blend_bwd_vec_kernel<T><<<grid, VT, 0, st>>>(
                                     (const T*)x, (const T*)y, z, (const T*)dout, M, C, (T*)dx, (T*)dy, dz)));
  return true;
}

}  // namespace vec
}  // namespace clskd
