// Fused middle stage of the attention-based fusion block (ABF, framework.py:206-220):
//
//   xp     = BatchNorm(z1)                         (z1 = 1x1 conv of the student map, framework.py:209)
//   yv     = nearest-resize(y_prev) along F        (framework.py:213-215, factor 1 or 2)
//   logit  = W_att [xp ; yv] + b_att               (1x1 conv onto 2 logits, framework.py:217-218)
//   xb     = xp * sigmoid(logit0) + yv * sigmoid(logit1)        (framework.py:219)
//
// Everything is row-local, so one kernel does it in a single pass (read z1, read y_prev, write xb
// and the 2 logits per row) instead of the 9 tensor passes of BN-apply + resize + attention conv +
// blend.  The backward is two passes (BatchNorm needs batch sums of the gradient first) that
// recompute xp / yv / sigmoid from z1, y_prev and the saved logits instead of reading saved copies,
// and produce dz1, dy_prev (resize adjoint fused), dW_att, db_att, dgamma, dbeta.
//
// Thread mapping: C/8 adjacent lanes own one row (each lane a fixed 8-channel group, so BN and
// attention constants live in registers); rows are addressed flat; all accesses are 16-byte vectors.
#include "common.cuh"

namespace clskd {
namespace {

__device__ __forceinline__ void ld8(const float* p, float* o) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
// bf16 pair -> two floats: one shift, one mask (the bf16x2 -> float2 intrinsic compiles to PRMT + shift per high half;
// these kernels are instruction-issue bound)
__device__ __forceinline__ void unpack8(const uint4& u, float* o) {
  o[0] = __uint_as_float(u.x << 16); o[1] = __uint_as_float(u.x & 0xffff0000u);
  o[2] = __uint_as_float(u.y << 16); o[3] = __uint_as_float(u.y & 0xffff0000u);
  o[4] = __uint_as_float(u.z << 16); o[5] = __uint_as_float(u.z & 0xffff0000u);
  o[6] = __uint_as_float(u.w << 16); o[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float* o) {
  unpack8(*reinterpret_cast<const uint4*>(p), o);
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
// 1 / (1 + 2^(-v log2 e)) on MUFU ex2 / rcp (approx, flush-to-zero: 4 instructions; the IEEE division and the denormal
// range check of __expf were 12).  v -> -inf: ex2 = +inf, rcp = 0; v -> +inf: ex2 = 0, rcp(1) = 1
__device__ __forceinline__ float sigm(float v) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}

constexpr int AT = 256;

// ---- cp.async software pipeline: every thread owns private 16-byte slots in shared memory
// ([stage][slot][thread], so a warp's read-back is conflict free).  Loads for iteration i+D-1 are
// in flight while iteration i computes: the bytes in flight per SM no longer depend on registers
// or occupancy (these kernels are HBM-latency bound otherwise).
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(void* sdst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp8(void* sdst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp4(void* sdst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 8 consecutive channels of type T = VB<T> bytes = NCP 16-byte slots
template <typename T> struct Vec8 { static constexpr int NCP = sizeof(T) / 2; };
template <typename T>
__device__ __forceinline__ void cp_vec8(uint8_t* slot0, int slot_stride, const T* g) {
#pragma unroll
  for (int i = 0; i < Vec8<T>::NCP; ++i) cp16(slot0 + i * slot_stride, reinterpret_cast<const uint8_t*>(g) + 16 * i);
}
__device__ __forceinline__ void ld_vec8(const uint8_t* slot0, int slot_stride, const __nv_bfloat16*, float* o) {
  unpack8(*reinterpret_cast<const uint4*>(slot0), o);
}
__device__ __forceinline__ void ld_vec8(const uint8_t* slot0, int slot_stride, const float*, float* o) {
  const float4 a = *reinterpret_cast<const float4*>(slot0);
  const float4 b = *reinterpret_cast<const float4*>(slot0 + slot_stride);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
// XS variants: the 1x1 conv in front of the block has only TWO input channels (the mask-level map of
// the decoder side), so z1 = W1 x is recomputed per row from the 2-channel input instead of being
// stored and re-read (4 or 8 bytes per row instead of 2C or 4C), and the backward produces dx and dW1
// directly instead of writing dz1 for a separate data/weight-gradient pass.
// n consecutive 2-channel rows of type T -> copy into a pipeline slot / decode from it
template <typename T>
__device__ __forceinline__ void cp_xs(uint8_t* slot, const T* g, int nrows) {
  // bf16: 4 bytes per row, fp32: 8 bytes per row
  if (sizeof(T) == 2) { if (nrows == 2) cp8(slot, g); else cp4(slot, g); }
  else { cp8(slot, g); if (nrows == 2) cp8(slot + 8, g + 2); }
}
__device__ __forceinline__ void ld_xs(const uint8_t* slot, const __nv_bfloat16*, int row, float& x0, float& x1) {
  const uint32_t u = *reinterpret_cast<const uint32_t*>(slot + 4 * row);
  x0 = __uint_as_float(u << 16);
  x1 = __uint_as_float(u & 0xffff0000u);
}
__device__ __forceinline__ void ld_xs(const uint8_t* slot, const float*, int row, float& x0, float& x1) {
  const float2 v = *reinterpret_cast<const float2*>(slot + 8 * row);
  x0 = v.x;
  x1 = v.y;
}
constexpr int PD = 3;   // pipeline depth

struct AbfGeom {
  int64_t M;      // rows B*T*F
  int F, Fy, C;   // F of this level, F of y_prev (F or F/2), channels
  int tpr;        // lanes per row = C/8 (power of two <= 32)
  int cshift;     // log2(C)  (C = 8 * tpr is a power of two)
  int yshift;     // 0: y_prev has F rows, 1: F/2 rows
};

// row m = bt*F + f -> flat row of y_prev = bt*Fy + (f >> yshift).  F is even, so with Fy = F/2 this
// is simply m >> 1: no division on the per-row path (the kernels are instruction-issue bound).
__device__ __forceinline__ int64_t yrow(const AbfGeom& g, int64_t m) { return m >> g.yshift; }

// per-channel constants staged in shared memory as [NCONST][C] floats (a lane reads its 8 channels
// of one constant with two 16-byte loads; lanes of a row are contiguous -> conflict free).  Keeping
// them out of registers is what lets two or three CTAs stay resident per SM.
enum { K_SC = 0, K_SH, K_WX0, K_WY0, K_WX1, K_WY1, K_MU, K_IS, K_GI, K_K1, K_K2, K_W10, K_W11, NCONST };

// channel c = 8*cg + e of a constant vector lives at [e >= 4][cg][e & 3]
__device__ __forceinline__ int cpos(int c, int C) { return ((c >> 2) & 1) * (C >> 1) + (c >> 3) * 4 + (c & 3); }

__device__ __forceinline__ void stage_consts(float* cs, int C, const float* mean, const float* invstd,
                                             const float* gamma, const float* beta, const float* watt,
                                             const double* sums, double invM, int training,
                                             const float* w1 = nullptr) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = gamma ? gamma[c] : 1.f;
    const float sc = invstd[c] * g;
    cs[K_SC * C + cpos(c, C)] = sc;
    cs[K_SH * C + cpos(c, C)] = (beta ? beta[c] : 0.f) - mean[c] * sc;
    cs[K_WX0 * C + cpos(c, C)] = watt[c];
    cs[K_WY0 * C + cpos(c, C)] = watt[C + c];
    cs[K_WX1 * C + cpos(c, C)] = watt[2 * C + c];
    cs[K_WY1 * C + cpos(c, C)] = watt[3 * C + c];
    cs[K_MU * C + cpos(c, C)] = mean[c];
    cs[K_IS * C + cpos(c, C)] = invstd[c];
    cs[K_GI * C + cpos(c, C)] = g * invstd[c];
    cs[K_K1 * C + cpos(c, C)] = (sums && training) ? (float)(sums[c] * invM) : 0.f;
    cs[K_K2 * C + cpos(c, C)] = (sums && training) ? (float)(sums[C + c] * invM) : 0.f;
    cs[K_W10 * C + cpos(c, C)] = w1 ? w1[2 * c] : 0.f;          // conv1 weight [C][2] (XS variants)
    cs[K_W11 * C + cpos(c, C)] = w1 ? w1[2 * c + 1] : 0.f;
  }
  __syncthreads();
}
__device__ __forceinline__ void ldc(const float* cs, int which, int C, int cg, float* o) {
  // planar halves (cpos): the 16 lanes of a row read 16 consecutive float4 per load - conflict free.  (With the 8
  // channels of a lane contiguous the two float4 loads strode 32 bytes: 2-way bank conflicts on every constant load,
  // 36 % of the kernels' shared-memory wavefronts in ncu.)
  const float4 a = *reinterpret_cast<const float4*>(cs + which * C + cg * 4);
  const float4 b = *reinterpret_cast<const float4*>(cs + which * C + (C >> 1) + cg * 4);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

template <typename T, bool XS>
__global__ void __launch_bounds__(AT, 2) abf_mid_fwd_kernel(const T* __restrict__ z1, const T* __restrict__ y, AbfGeom g,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ watt, const float* __restrict__ batt,
                                                            const float* __restrict__ w1,
                                                            T* __restrict__ xb, float* __restrict__ logits) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* cs = reinterpret_cast<float*>(smem_raw);
  constexpr int RQ = 2;
  constexpr int NCP = Vec8<T>::NCP;
  constexpr int NSLOT = 2 * RQ * NCP;                    // z1[q], y[q]
  constexpr int SLOT_STRIDE = AT * 16;                   // bytes between consecutive slots of a thread
  uint8_t* pipe = smem_raw + ((sizeof(float) * NCONST * g.C + 15) & ~(size_t)15) + threadIdx.x * 16;
  stage_consts(cs, g.C, mean, invstd, gamma, beta, watt, nullptr, 0., 0, XS ? w1 : nullptr);
  const int lane = threadIdx.x & 31;
  const int cg = lane & (g.tpr - 1), sub = lane / g.tpr, rpw = 32 / g.tpr;
  const float b0 = batt ? batt[0] : 0.f, b1 = batt ? batt[1] : 0.f;
  const int64_t warp0 = ((int64_t)blockIdx.x * AT + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * AT) >> 5;
  const int64_t stride = nwarps * rpw * RQ;
  const int cs_ = g.cshift;
  const int coff = cg * 8;

  // the issue side runs PD-1 iterations ahead of the compute side; both walk rows incrementally
  int64_t mi = warp0 * rpw * RQ + sub;      // first row of this lane group in the iteration being issued
  int si = 0;                               // its pipeline stage
  // running source pointers (rows mi, mi + rpw; stride and rpw * RQ are even, so the y_prev row m >> yshift advances by
  // a constant too): 64-bit address arithmetic per load was a large part of the loop's integer instructions
  const T* zp_i[RQ];
  const T* yp_i[RQ];
#pragma unroll
  for (int q = 0; q < RQ; ++q) {
    const int64_t m = mi + q * rpw;
    zp_i[q] = XS ? z1 + 2 * m : z1 + (m << cs_) + coff;
    yp_i[q] = y + (yrow(g, m) << cs_) + coff;
  }
  const int64_t inc_z = XS ? 2 * stride : stride << cs_, inc_y = yrow(g, stride) << cs_;
  auto issue = [&]() {
    uint8_t* st = pipe + (size_t)si * NSLOT * SLOT_STRIDE;
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
      if (mi + q * rpw < g.M) {
        if (XS) cp_xs<T>(st + (q * NCP) * SLOT_STRIDE, zp_i[q], 1);
        else cp_vec8<T>(st + (q * NCP) * SLOT_STRIDE, SLOT_STRIDE, zp_i[q]);
        cp_vec8<T>(st + ((RQ + q) * NCP) * SLOT_STRIDE, SLOT_STRIDE, yp_i[q]);
      }
      zp_i[q] += inc_z;
      yp_i[q] += inc_y;
    }
    cp_commit();
    mi += stride;
    si = si + 1 == PD ? 0 : si + 1;
  };
#pragma unroll
  for (int i = 0; i < PD - 1; ++i) issue();

  int sc_ = 0;
  for (int64_t mb = warp0 * rpw * RQ; mb < g.M; mb += stride) {     // mb is warp-uniform
    issue();
    cp_wait<PD - 1>();
    const uint8_t* st = pipe + (size_t)sc_ * NSLOT * SLOT_STRIDE;
    sc_ = sc_ + 1 == PD ? 0 : sc_ + 1;
    const int64_t m0 = mb + sub;
    float xv[RQ][8], yv[RQ][8];
    bool live[RQ];
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
      const int64_t m = m0 + q * rpw;
      live[q] = m < g.M;
      // a row beyond M computes on whatever its slots hold: nothing is accumulated across rows here, the shuffles stay
      // inside the row's own lane group, and the stores below are predicated
      {
        if (XS) {
          float x0, x1, wa[8], wb[8];
          ld_xs(st + (q * NCP) * SLOT_STRIDE, (const T*)nullptr, 0, x0, x1);
          ldc(cs, K_W10, g.C, cg, wa);
          ldc(cs, K_W11, g.C, cg, wb);
#pragma unroll
          for (int e = 0; e < 8; ++e) xv[q][e] = fmaf(wa[e], x0, wb[e] * x1);      // z1 = W1 x
        } else {
          ld_vec8(st + (q * NCP) * SLOT_STRIDE, SLOT_STRIDE, (const T*)nullptr, xv[q]);
        }
        ld_vec8(st + ((RQ + q) * NCP) * SLOT_STRIDE, SLOT_STRIDE, (const T*)nullptr, yv[q]);
      }
    }
    float l0[RQ], l1[RQ];
    {
      float sc[8], sh[8];
      ldc(cs, K_SC, g.C, cg, sc);
      ldc(cs, K_SH, g.C, cg, sh);
#pragma unroll
      for (int q = 0; q < RQ; ++q)
#pragma unroll
        for (int e = 0; e < 8; ++e) xv[q][e] = fmaf(xv[q][e], sc[e], sh[e]);      // xp
    }
    {
      float wx[8], wy[8];
      ldc(cs, K_WX0, g.C, cg, wx);
      ldc(cs, K_WY0, g.C, cg, wy);
#pragma unroll
      for (int q = 0; q < RQ; ++q) {
        float a = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) a = fmaf(xv[q][e], wx[e], fmaf(yv[q][e], wy[e], a));
        l0[q] = a;
      }
      ldc(cs, K_WX1, g.C, cg, wx);
      ldc(cs, K_WY1, g.C, cg, wy);
#pragma unroll
      for (int q = 0; q < RQ; ++q) {
        float a = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) a = fmaf(xv[q][e], wx[e], fmaf(yv[q][e], wy[e], a));
        l1[q] = a;
      }
    }
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
      _Pragma("unroll") for (int o = 16; o > 0; o >>= 1) if (o < g.tpr) {
        l0[q] += __shfl_xor_sync(0xffffffffu, l0[q], o);
        l1[q] += __shfl_xor_sync(0xffffffffu, l1[q], o);
      }
      l0[q] += b0;
      l1[q] += b1;
      const float s0 = sigm(l0[q]), s1 = sigm(l1[q]);
      if (live[q]) {
        const int64_t m = m0 + q * rpw;
        float o8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o8[e] = xv[q][e] * s0 + yv[q][e] * s1;
        st8(xb + (m << cs_) + coff, o8);
        if (cg == 0) *reinterpret_cast<float2*>(logits + 2 * m) = make_float2(l0[q], l1[q]);
      }
    }
  }
}

// MODE 0: statistics pass (sums for BatchNorm backward, dW_att, db_att)
// MODE 1: apply pass (dz1, dy_prev)
// MODE 2: ONE pass (clskd_abf_mid_bwd_fold): the statistics of MODE 0 and, to `dz1`, the gradient dxp of the BatchNorm
//         OUTPUT (row-local), plus dy_prev.  The BatchNorm-backward affine dz1 = gi (dxp - k1 - k2 xhat) is then folded
//         into the weights of conv1's data / weight gradients (abf_fold_* kernels below): dz1 is never materialised and
//         z1 / gout / y_prev are read once instead of twice.
// A lane group always processes the PAIR of rows (f = 2j, 2j+1) that share one y_prev row when
// Fy = F/2 (so dy_prev is written once, without atomics); with Fy = F the pair is two plain rows.
// NT threads per CTA: 256 (two CTAs per SM, 128 registers) for the two-pass kernels; the one-pass kernel (MODE 2) keeps the
// accumulators of MODE 0 AND the output rows of MODE 1 live - at 128 registers it spilled inside the row loop and its
// global stores evicted the spill lines from L1 (1.83 ms at F = 128 where MODE 0 + MODE 1 took 2.17): it runs as three
// 128-thread CTAs per SM (168 registers)
template <typename T, int MODE, bool XS, int NT = AT>
__global__ void __launch_bounds__(NT, NT == AT ? 2 : 3) abf_mid_bwd_kernel(const T* __restrict__ gout, const T* __restrict__ z1,
                                                            const T* __restrict__ y, AbfGeom g,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ watt, const float* __restrict__ logits,
                                                            double* __restrict__ sums, double* __restrict__ dwatt,
                                                            double* __restrict__ dbatt, int training, T* __restrict__ dz1,
                                                            T* __restrict__ dy, const float* __restrict__ w1,
                                                            double* __restrict__ dw1) {
  extern __shared__ __align__(16) uint8_t smem_raw[];   // constants, (MODE 0) reduction slots, cp.async pipeline
  float* cs = reinterpret_cast<float*>(smem_raw);
  const int C = g.C;
  float* red = cs + NCONST * C;
  constexpr int NCP = Vec8<T>::NCP;
  constexpr int NSLOT = 6 * NCP + 1;                     // g[2], z1[2], y[2] vectors + one slot for both logit pairs
  constexpr int SLOT_STRIDE = NT * 16;
  uint8_t* pipe = smem_raw + ((sizeof(float) * ((NCONST + 6) * C + 2) + 15) & ~(size_t)15) + threadIdx.x * 16;
  if (MODE != 1 || XS)
    for (int i = threadIdx.x; i < 6 * C + 2; i += NT) red[i] = 0.f;
  // the pipeline slots start as zeros: a lane group beyond the last row pair reads finite (stale or zero) data and only
  // has to zero its gout rows to drop out of every sum (ordered before the first cp.async by stage_consts' barrier)
  for (int i = 0; i < PD * NSLOT; ++i) *reinterpret_cast<uint4*>(pipe + (size_t)i * SLOT_STRIDE) = make_uint4(0u, 0u, 0u, 0u);
  stage_consts(cs, C, mean, invstd, gamma, beta, watt, MODE == 1 ? sums : nullptr, 1.0 / (double)g.M, training,
               XS ? w1 : nullptr);
  // XS apply pass: dW1[c][k] partial sums of this thread's 8 channels
  float a_w10[8], a_w11[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a_w10[e] = a_w11[e] = 0.f;
  const int lane = threadIdx.x & 31;
  const int cg = lane & (g.tpr - 1), sub = lane / g.tpr, rpw = 32 / g.tpr;
  // MODE 0 accumulators
  float a_s0[8], a_s1[8], a_wx0[8], a_wx1[8], a_wy0[8], a_wy1[8], a_b0 = 0.f, a_b1 = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) a_s0[e] = a_s1[e] = a_wx0[e] = a_wx1[e] = a_wy0[e] = a_wy1[e] = 0.f;

  const int64_t pairs = g.M >> 1;           // M is even (F is even)
  const int64_t warp0 = ((int64_t)blockIdx.x * NT + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * NT) >> 5;
  const int64_t stride = nwarps * rpw;
  const int cs_ = g.cshift;
  const int coff = cg * 8;
  const bool y_full = g.yshift == 0;

  int64_t pi = warp0 * rpw + sub;           // pair issued next by this lane group
  int si = 0;
  // running source pointers of the pair being issued (64-bit address arithmetic per load was a quarter of the loop's
  // integer instructions); row pair p = rows 2p, 2p+1; y_prev row of row m = m >> yshift
  const int64_t one_row = (int64_t)1 << cs_;
  const int64_t inc_g = (2 * stride) << cs_, inc_y = yrow(g, 2 * stride) << cs_;
  const T* gp_i = gout + ((2 * pi) << cs_) + coff;
  const T* zp_i = XS ? z1 + 4 * pi : z1 + ((2 * pi) << cs_) + coff;       // XS: two rows of the 2-channel input
  const T* yp_i = y + (yrow(g, 2 * pi) << cs_) + coff;
  const float* lp_i = logits + 4 * pi;                                    // logits of rows 2p and 2p+1 (16 bytes)
  auto issue = [&]() {
    if (pi < pairs) {
      uint8_t* st = pipe + (size_t)si * NSLOT * SLOT_STRIDE;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        cp_vec8<T>(st + (q * NCP) * SLOT_STRIDE, SLOT_STRIDE, gp_i + q * one_row);
        if (!XS) cp_vec8<T>(st + ((2 + q) * NCP) * SLOT_STRIDE, SLOT_STRIDE, zp_i + q * one_row);
      }
      if (XS) cp_xs<T>(st + (2 * NCP) * SLOT_STRIDE, zp_i, 2);
      cp_vec8<T>(st + (4 * NCP) * SLOT_STRIDE, SLOT_STRIDE, yp_i);
      if (y_full) cp_vec8<T>(st + (5 * NCP) * SLOT_STRIDE, SLOT_STRIDE, yp_i + one_row);
      cp16(st + (6 * NCP) * SLOT_STRIDE, lp_i);
    }
    cp_commit();
    pi += stride;
    gp_i += inc_g;
    zp_i += XS ? 4 * stride : inc_g;
    yp_i += inc_y;
    lp_i += 4 * stride;
    si = si + 1 == PD ? 0 : si + 1;
  };
#pragma unroll
  for (int i = 0; i < PD - 1; ++i) issue();

  int sc_ = 0;
  for (int64_t pb = warp0 * rpw; pb < pairs; pb += stride) {       // pb is warp-uniform
    issue();
    cp_wait<PD - 1>();
    const uint8_t* st = pipe + (size_t)sc_ * NSLOT * SLOT_STRIDE;
    sc_ = sc_ + 1 == PD ? 0 : sc_ + 1;
    const int64_t p = pb + sub;
    const bool live = p < pairs;
    const int64_t m0 = live ? 2 * p : 0;
    const int64_t yr0 = yrow(g, m0);
    float gv[2][8], xv[2][8], yv[2][8];
    float2 lg[2];
    float xs0[2] = {0.f, 0.f}, xs1[2] = {0.f, 0.f};
    if (live || !XS) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        ld_vec8(st + (q * NCP) * SLOT_STRIDE, SLOT_STRIDE, (const T*)nullptr, gv[q]);
        if (!XS) ld_vec8(st + ((2 + q) * NCP) * SLOT_STRIDE, SLOT_STRIDE, (const T*)nullptr, xv[q]);
      }
      if (XS) {
        float wa[8], wb[8];
        ldc(cs, K_W10, C, cg, wa);
        ldc(cs, K_W11, C, cg, wb);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          ld_xs(st + (2 * NCP) * SLOT_STRIDE, (const T*)nullptr, q, xs0[q], xs1[q]);
#pragma unroll
          for (int e = 0; e < 8; ++e) xv[q][e] = fmaf(wa[e], xs0[q], wb[e] * xs1[q]);      // z1 = W1 x
        }
      }
      ld_vec8(st + (4 * NCP) * SLOT_STRIDE, SLOT_STRIDE, (const T*)nullptr, yv[0]);
      if (y_full) ld_vec8(st + (5 * NCP) * SLOT_STRIDE, SLOT_STRIDE, (const T*)nullptr, yv[1]);
      else {
#pragma unroll
        for (int e = 0; e < 8; ++e) yv[1][e] = yv[0][e];
      }
      const float4 l4 = *reinterpret_cast<const float4*>(st + (6 * NCP) * SLOT_STRIDE);
      lg[0] = make_float2(l4.x, l4.y);
      lg[1] = make_float2(l4.z, l4.w);
      if (!XS) {
        // a group beyond the last pair (tail of the grid-stride loop): every sum below is linear in gout
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int e = 0; e < 8; ++e) gv[q][e] = live ? gv[q][e] : 0.f;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int e = 0; e < 8; ++e) gv[q][e] = xv[q][e] = yv[q][e] = 0.f;
        lg[q] = make_float2(0.f, 0.f);
      }
    }
    // xp, and (apply pass) xhat kept in place of z1.  The statistics pass keeps the raw z1: it accumulates
    // sum dxp z1 and converts to sum dxp xhat = invstd (sum dxp z1 - mean sum dxp) once per CTA - two constant
    // vectors and two instructions per element less in the pass that is instruction-issue bound
    float xp[2][8];
    {
      float sc[8], sh[8];
      ldc(cs, K_SC, C, cg, sc);
      ldc(cs, K_SH, C, cg, sh);
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int e = 0; e < 8; ++e) xp[q][e] = fmaf(xv[q][e], sc[e], sh[e]);
      if (MODE == 1) {
        float mu[8], is[8];
        ldc(cs, K_MU, C, cg, mu);
        ldc(cs, K_IS, C, cg, is);
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int e = 0; e < 8; ++e) xv[q][e] = (xv[q][e] - mu[e]) * is[e];              // xhat
      }
    }
    float s0[2], s1[2], dl0[2], dl1[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      s0[q] = sigm(lg[q].x);
      s1[q] = sigm(lg[q].y);
      float t0 = 0.f, t1 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        t0 = fmaf(gv[q][e], xp[q][e], t0);
        t1 = fmaf(gv[q][e], yv[q][e], t1);
      }
      _Pragma("unroll") for (int o = 16; o > 0; o >>= 1) if (o < g.tpr) {
        t0 += __shfl_xor_sync(0xffffffffu, t0, o);
        t1 += __shfl_xor_sync(0xffffffffu, t1, o);
      }
      dl0[q] = t0 * s0[q] * (1.f - s0[q]);
      dl1[q] = t1 * s1[q] * (1.f - s1[q]);
    }
    if (MODE == 0 || MODE == 2) {
      float wx0[8], wx1[8];
      ldc(cs, K_WX0, C, cg, wx0);
      ldc(cs, K_WX1, C, cg, wx1);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float dx8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float dxp = gv[q][e] * s0[q] + wx0[e] * dl0[q] + wx1[e] * dl1[q];
          dx8[e] = dxp;
          a_s0[e] += dxp;
          a_s1[e] = fmaf(dxp, xv[q][e], a_s1[e]);
          a_wx0[e] = fmaf(dl0[q], xp[q][e], a_wx0[e]);
          a_wx1[e] = fmaf(dl1[q], xp[q][e], a_wx1[e]);
          a_wy0[e] = fmaf(dl0[q], yv[q][e], a_wy0[e]);
          a_wy1[e] = fmaf(dl1[q], yv[q][e], a_wy1[e]);
        }
        if (cg == 0) {
          a_b0 += dl0[q];
          a_b1 += dl1[q];
        }
        if (MODE == 2 && live) st8(dz1 + ((m0 + q) << cs_) + coff, dx8);       // dxp (gradient of the BatchNorm output)
      }
      if (MODE == 2) {
        float wy0[8], wy1[8];
        ldc(cs, K_WY0, C, cg, wy0);
        ldc(cs, K_WY1, C, cg, wy1);
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int e = 0; e < 8; ++e) yv[q][e] = gv[q][e] * s1[q] + wy0[e] * dl0[q] + wy1[e] * dl1[q];   // dyv
        if (live) {
          T* yo = dy + (yr0 << cs_) + coff;
          if (y_full) {
            st8(yo, yv[0]);
            st8(yo + ((int64_t)1 << cs_), yv[1]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) yv[0][e] += yv[1][e];
            st8(yo, yv[0]);
          }
        }
      }
    } else {
      {
        float wx0[8], wx1[8], gi[8], k1[8], k2[8];
        ldc(cs, K_WX0, C, cg, wx0);
        ldc(cs, K_WX1, C, cg, wx1);
        ldc(cs, K_GI, C, cg, gi);
        ldc(cs, K_K1, C, cg, k1);
        ldc(cs, K_K2, C, cg, k2);
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float dxp = gv[q][e] * s0[q] + wx0[e] * dl0[q] + wx1[e] * dl1[q];
            xp[q][e] = gi[e] * (dxp - k1[e] - xv[q][e] * k2[e]);      // dz1
          }
      }
      {
        float wy0[8], wy1[8];
        ldc(cs, K_WY0, C, cg, wy0);
        ldc(cs, K_WY1, C, cg, wy1);
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int e = 0; e < 8; ++e) yv[q][e] = gv[q][e] * s1[q] + wy0[e] * dl0[q] + wy1[e] * dl1[q];   // dyv
      }
      if (XS) {
        // conv1 backward in place: dx[k] = sum_c W1[c][k] dz1[c] (row-wise), dW1[c][k] += dz1[c] x[k]
        float wa[8], wb[8];
        ldc(cs, K_W10, C, cg, wa);
        ldc(cs, K_W11, C, cg, wb);
        float p[2][2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float p0 = 0.f, p1 = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            p0 = fmaf(wa[e], xp[q][e], p0);
            p1 = fmaf(wb[e], xp[q][e], p1);
            a_w10[e] = fmaf(xp[q][e], xs0[q], a_w10[e]);
            a_w11[e] = fmaf(xp[q][e], xs1[q], a_w11[e]);
          }
          _Pragma("unroll") for (int o = 16; o > 0; o >>= 1) if (o < g.tpr) {
            p0 += __shfl_xor_sync(0xffffffffu, p0, o);
            p1 += __shfl_xor_sync(0xffffffffu, p1, o);
          }
          p[q][0] = p0;
          p[q][1] = p1;
        }
        if (live && cg == 0) {
          T* xo = dz1 + 2 * m0;                       // dx rows m0, m0+1 of the 2-channel input
          st_f(xo, p[0][0]); st_f(xo + 1, p[0][1]); st_f(xo + 2, p[1][0]); st_f(xo + 3, p[1][1]);
        }
      }
      if (live) {
        T* zo = dz1 + (m0 << cs_) + coff;
        T* yo = dy + (yr0 << cs_) + coff;
        if (!XS) {
          st8(zo, xp[0]);
          st8(zo + ((int64_t)1 << cs_), xp[1]);
        }
        if (y_full) {
          st8(yo, yv[0]);
          st8(yo + ((int64_t)1 << cs_), yv[1]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) yv[0][e] += yv[1][e];
          st8(yo, yv[0]);
        }
      }
    }
  }
  if (MODE == 1 && XS) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = cg * 8 + e;
      atomicAdd(&red[c], a_w10[e]);
      atomicAdd(&red[C + c], a_w11[e]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += NT) {
      atomicAdd(dw1 + 2 * i, (double)red[i]);
      atomicAdd(dw1 + 2 * i + 1, (double)red[C + i]);
    }
  }
  if (MODE == 0 || MODE == 2) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = cg * 8 + e;
      atomicAdd(&red[c], a_s0[e]);
      atomicAdd(&red[C + c], a_s1[e]);
      atomicAdd(&red[2 * C + c], a_wx0[e]);
      atomicAdd(&red[3 * C + c], a_wy0[e]);
      atomicAdd(&red[4 * C + c], a_wx1[e]);
      atomicAdd(&red[5 * C + c], a_wy1[e]);
    }
    if (cg == 0) {
      atomicAdd(&red[6 * C], a_b0);
      atomicAdd(&red[6 * C + 1], a_b1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += NT) {
      atomicAdd(sums + i, (double)red[i]);
      // red[C + i] is this CTA's sum dxp z1 (raw z1): -> sum dxp xhat
      atomicAdd(sums + C + i, (double)invstd[i] * ((double)red[C + i] - (double)mean[i] * (double)red[i]));
      // dwatt layout = W_att layout [2][2C]: row k, columns [x channels | y channels]
      atomicAdd(dwatt + i, (double)red[2 * C + i]);
      atomicAdd(dwatt + C + i, (double)red[3 * C + i]);
      atomicAdd(dwatt + 2 * C + i, (double)red[4 * C + i]);
      atomicAdd(dwatt + 3 * C + i, (double)red[5 * C + i]);
    }
    if (threadIdx.x == 0) {
      atomicAdd(dbatt, (double)red[6 * C]);
      atomicAdd(dbatt + 1, (double)red[6 * C + 1]);
    }
  }
}


// =====================================================================================================
// XS2: the 2-channel ("mask level") ABF middle stage with the rank-2 structure of z1 = W1 x folded through
// every reduction (framework.py:209-219 with a 2-channel conv1 input).  With
//     xp_c   = P_c x0 + Q_c x1 + sh_c              (P = sc w10, Q = sc w11, sc = invstd gamma, sh = beta - mean sc)
//     xhat_c = al_c x0 + be_c x1 + ga_c            (al = w10 invstd, be = w11 invstd, ga = -mean invstd)
// the x half of the attention logits is a pair of scalars per row, the BatchNorm-backward batch sums
//     sum_rows dxp_c,  sum_rows dxp_c x0,  sum_rows dxp_c x1        (dxp = g s0 + wx0 dl0 + wx1 dl1)
// need only  G0_c = sum g_c s0,  G1_c = sum g_c s0 x0,  G2_c = sum g_c s0 x1  per channel plus six scalar sums, dW_att's
// x half and dW1 follow from those sums and the 2x2 moments of x, and
//     dx_k = sum_c w1k_c dz1_c = s0 <g, U_k> + dl0 c_ka + dl1 c_kb - c_kc - x0 c_kd - x1 c_ke      (U_k = w1k gamma invstd)
// needs only the two per-row dot products <g, U_k>.  So the backward is ONE pass over g / y_prev (it writes dy_prev and 16
// bytes per row: s0<g,U0>, s0<g,U1>, dl0, dl1), a one-block finalisation, and a pass over 24 bytes per row that emits dx -
// instead of two full passes that recompute z1, xp and xhat per element.
// Thread mapping as above: C/8 adjacent lanes own a PAIR of rows (f = 2j, 2j+1: they share the y_prev row when Fy = F/2).
// =====================================================================================================
enum { X_P = 0, X_Q, X_SH, X_WY0, X_WY1, X_U0, X_U1, X_NCONST };
// scalar slots behind the per-channel constants
enum { XS_A0 = 0, XS_B0, XS_K0, XS_A1, XS_B1, XS_K1, XS_NSCAL };

__device__ __forceinline__ void xs2_stage_consts(float* cs, int C, const float* mean, const float* invstd, const float* gamma,
                                                 const float* beta, const float* watt, const float* batt, const float* w1) {
  // per-channel vectors
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = gamma ? gamma[c] : 1.f;
    const float is = invstd[c], sc = is * g, sh = (beta ? beta[c] : 0.f) - mean[c] * sc;
    const float a = w1[2 * c], b = w1[2 * c + 1];
    cs[X_P * C + cpos(c, C)] = sc * a;
    cs[X_Q * C + cpos(c, C)] = sc * b;
    cs[X_SH * C + cpos(c, C)] = sh;
    cs[X_WY0 * C + cpos(c, C)] = watt[C + c];
    cs[X_WY1 * C + cpos(c, C)] = watt[3 * C + c];
    cs[X_U0 * C + cpos(c, C)] = a * g * is;
    cs[X_U1 * C + cpos(c, C)] = b * g * is;
  }
  __syncthreads();
  // scalars of the x half of the logits: l_k = x0 A_k + x1 B_k + K_k + <yv, wy_k>  (K_k includes the bias)
  float* sc_ = cs + X_NCONST * C;
  if (threadIdx.x < 32) {
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = threadIdx.x; c < C; c += 32) {
      const float p = cs[X_P * C + cpos(c, C)], q = cs[X_Q * C + cpos(c, C)], sh = cs[X_SH * C + cpos(c, C)];
      const float wx0 = watt[c], wx1 = watt[2 * C + c];
      v[0] = fmaf(p, wx0, v[0]); v[1] = fmaf(q, wx0, v[1]); v[2] = fmaf(sh, wx0, v[2]);
      v[3] = fmaf(p, wx1, v[3]); v[4] = fmaf(q, wx1, v[4]); v[5] = fmaf(sh, wx1, v[5]);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) v[i] = warp_sum(v[i]);
    if (threadIdx.x == 0) {
      sc_[XS_A0] = v[0]; sc_[XS_B0] = v[1]; sc_[XS_K0] = v[2] + (batt ? batt[0] : 0.f);
      sc_[XS_A1] = v[3]; sc_[XS_B1] = v[4]; sc_[XS_K1] = v[5] + (batt ? batt[1] : 0.f);
    }
  }
  __syncthreads();
}

// n partial sums per lane -> totals over the tpr lanes of a row group (butterfly); tpr is a power of two
template <int N>
__device__ __forceinline__ void group_sum(float* v, int tpr) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    if (o < tpr) {
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    }
}

template <typename T>
__global__ void __launch_bounds__(AT, 2) abf_xs2_fwd_kernel(const T* __restrict__ x, const T* __restrict__ y, AbfGeom g,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ watt, const float* __restrict__ batt,
                                                            const float* __restrict__ w1, T* __restrict__ xb,
                                                            float* __restrict__ logits) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* cs = reinterpret_cast<float*>(smem_raw);
  const int C = g.C;
  constexpr int NCP = Vec8<T>::NCP;
  constexpr int NSLOT = 2 * NCP + 1;                      // y rows (1 or 2) + one slot for the two 2-channel rows
  constexpr int SLOT_STRIDE = AT * 16;
  uint8_t* pipe = smem_raw + ((sizeof(float) * (X_NCONST * C + XS_NSCAL) + 15) & ~(size_t)15) + threadIdx.x * 16;
  xs2_stage_consts(cs, C, mean, invstd, gamma, beta, watt, batt, w1);
  const float* sc_ = cs + X_NCONST * C;
  const float A0 = sc_[XS_A0], B0 = sc_[XS_B0], K0 = sc_[XS_K0], A1 = sc_[XS_A1], B1 = sc_[XS_B1], K1 = sc_[XS_K1];
  const int lane = threadIdx.x & 31;
  const int cg = lane & (g.tpr - 1), sub = lane / g.tpr, rpw = 32 / g.tpr;
  const int64_t pairs = g.M >> 1;
  const int64_t warp0 = ((int64_t)blockIdx.x * AT + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * AT) >> 5;
  const int64_t stride = nwarps * rpw;
  const int cs_ = g.cshift, coff = cg * 8;
  const bool y_full = g.yshift == 0;

  int64_t pi = warp0 * rpw + sub;
  int si = 0;
  auto issue = [&]() {
    if (pi < pairs) {
      uint8_t* st = pipe + (size_t)si * NSLOT * SLOT_STRIDE;
      const int64_t m0 = 2 * pi;
      const T* yp = y + (yrow(g, m0) << cs_) + coff;
      cp_vec8<T>(st, SLOT_STRIDE, yp);
      if (y_full) cp_vec8<T>(st + NCP * SLOT_STRIDE, SLOT_STRIDE, yp + ((int64_t)1 << cs_));
      cp_xs<T>(st + (2 * NCP) * SLOT_STRIDE, x + 2 * m0, 2);
    }
    cp_commit();
    pi += stride;
    si = si + 1 == PD ? 0 : si + 1;
  };
#pragma unroll
  for (int i = 0; i < PD - 1; ++i) issue();

  int sc2 = 0;
  for (int64_t pb = warp0 * rpw; pb < pairs; pb += stride) {     // warp-uniform
    issue();
    cp_wait<PD - 1>();
    const uint8_t* st = pipe + (size_t)sc2 * NSLOT * SLOT_STRIDE;
    sc2 = sc2 + 1 == PD ? 0 : sc2 + 1;
    const int64_t p = pb + sub;
    const bool live = p < pairs;
    const int64_t m0 = live ? 2 * p : 0;
    float yv[2][8], x0[2] = {0.f, 0.f}, x1[2] = {0.f, 0.f};
    if (live) {
      ld_vec8(st, SLOT_STRIDE, (const T*)nullptr, yv[0]);
      if (y_full) ld_vec8(st + NCP * SLOT_STRIDE, SLOT_STRIDE, (const T*)nullptr, yv[1]);
      ld_xs(st + (2 * NCP) * SLOT_STRIDE, (const T*)nullptr, 0, x0[0], x1[0]);
      ld_xs(st + (2 * NCP) * SLOT_STRIDE, (const T*)nullptr, 1, x0[1], x1[1]);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) yv[0][e] = yv[1][e] = 0.f;
    }
    // y half of the logits (once per y row)
    float ly[4] = {0.f, 0.f, 0.f, 0.f};
    {
      float w0[8], w1v[8];
      ldc(cs, X_WY0, C, cg, w0);
      ldc(cs, X_WY1, C, cg, w1v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        ly[0] = fmaf(yv[0][e], w0[e], ly[0]);
        ly[1] = fmaf(yv[0][e], w1v[e], ly[1]);
      }
      if (y_full) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          ly[2] = fmaf(yv[1][e], w0[e], ly[2]);
          ly[3] = fmaf(yv[1][e], w1v[e], ly[3]);
        }
      }
    }
    if (y_full) {
      group_sum<4>(ly, g.tpr);
    } else {
      group_sum<2>(ly, g.tpr);
      ly[2] = ly[0];
      ly[3] = ly[1];
#pragma unroll
      for (int e = 0; e < 8; ++e) yv[1][e] = yv[0][e];
    }
    float P[8], Q[8], SH[8];
    ldc(cs, X_P, C, cg, P);
    ldc(cs, X_Q, C, cg, Q);
    ldc(cs, X_SH, C, cg, SH);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float l0 = fmaf(x0[q], A0, fmaf(x1[q], B0, K0)) + ly[2 * q];
      const float l1 = fmaf(x0[q], A1, fmaf(x1[q], B1, K1)) + ly[2 * q + 1];
      const float s0 = sigm(l0), s1 = sigm(l1);
      if (live) {
        const float a = x0[q] * s0, b = x1[q] * s0;
        float o8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o8[e] = fmaf(P[e], a, fmaf(Q[e], b, fmaf(SH[e], s0, yv[q][e] * s1)));
        st8(xb + ((m0 + q) << cs_) + coff, o8);
        if (cg == 0) *reinterpret_cast<float2*>(logits + 2 * (m0 + q)) = make_float2(l0, l1);
      }
    }
  }
}

// per-CTA scalar sums of the backward pass
enum { R_D0 = 0, R_D1, R_D0X0, R_D0X1, R_D1X0, R_D1X1, R_X0, R_X1, R_X00, R_X01, R_X11, R_NSCAL };

// acc layout (fp64, zeroed by the launcher): G0[C] G1[C] G2[C] WY0[C] WY1[C] scal[R_NSCAL]
// 128-thread CTAs, three per SM (168 registers): at 128 registers the kernel spilled inside the row loop and its
// dy_prev / workspace stores evicted the spill lines from L1 (as in the one-pass abf_mid_bwd_kernel)
constexpr int XT = 128;
template <typename T>
__global__ void __launch_bounds__(XT, 3) abf_xs2_bwd_kernel(const T* __restrict__ gout, const T* __restrict__ x,
                                                            const T* __restrict__ y, AbfGeom g,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ watt, const float* __restrict__ w1,
                                                            const float* __restrict__ logits, double* __restrict__ acc,
                                                            float4* __restrict__ rows, T* __restrict__ dy) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* cs = reinterpret_cast<float*>(smem_raw);
  const int C = g.C;
  float* red = cs + X_NCONST * C + XS_NSCAL;              // [5*C + R_NSCAL]
  constexpr int NCP = Vec8<T>::NCP;
  constexpr int NSLOT = 4 * NCP + 2;                      // g[2], y[1..2] vectors, x pair, logits pair
  constexpr int SLOT_STRIDE = XT * 16;
  uint8_t* pipe = smem_raw + ((sizeof(float) * (X_NCONST * C + XS_NSCAL + 5 * C + R_NSCAL) + 15) & ~(size_t)15) + threadIdx.x * 16;
  for (int i = threadIdx.x; i < 5 * C + R_NSCAL; i += XT) red[i] = 0.f;
  xs2_stage_consts(cs, C, mean, invstd, gamma, beta, watt, nullptr, w1);
  const int lane = threadIdx.x & 31;
  const int cg = lane & (g.tpr - 1), sub = lane / g.tpr, rpw = 32 / g.tpr;
  const int64_t pairs = g.M >> 1;
  const int64_t warp0 = ((int64_t)blockIdx.x * XT + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * XT) >> 5;
  const int64_t stride = nwarps * rpw;
  const int cs_ = g.cshift, coff = cg * 8;
  const bool y_full = g.yshift == 0;
  float aG0[8], aG1[8], aG2[8], aW0[8], aW1[8], aS[R_NSCAL];
#pragma unroll
  for (int e = 0; e < 8; ++e) aG0[e] = aG1[e] = aG2[e] = aW0[e] = aW1[e] = 0.f;
#pragma unroll
  for (int e = 0; e < R_NSCAL; ++e) aS[e] = 0.f;

  int64_t pi = warp0 * rpw + sub;
  int si = 0;
  // running source pointers of the row pair being issued (rows 2p, 2p+1)
  const int64_t one_row = (int64_t)1 << cs_;
  const int64_t inc_g = (2 * stride) << cs_, inc_y = yrow(g, 2 * stride) << cs_;
  const T* gp_i = gout + ((2 * pi) << cs_) + coff;
  const T* yp_i = y + (yrow(g, 2 * pi) << cs_) + coff;
  const T* xp_i = x + 4 * pi;
  const float* lp_i = logits + 4 * pi;
  auto issue = [&]() {
    if (pi < pairs) {
      uint8_t* st = pipe + (size_t)si * NSLOT * SLOT_STRIDE;
      cp_vec8<T>(st, SLOT_STRIDE, gp_i);
      cp_vec8<T>(st + NCP * SLOT_STRIDE, SLOT_STRIDE, gp_i + one_row);
      cp_vec8<T>(st + (2 * NCP) * SLOT_STRIDE, SLOT_STRIDE, yp_i);
      if (y_full) cp_vec8<T>(st + (3 * NCP) * SLOT_STRIDE, SLOT_STRIDE, yp_i + one_row);
      cp_xs<T>(st + (4 * NCP) * SLOT_STRIDE, xp_i, 2);
      cp16(st + (4 * NCP + 1) * SLOT_STRIDE, lp_i);
    }
    cp_commit();
    pi += stride;
    gp_i += inc_g;
    yp_i += inc_y;
    xp_i += 4 * stride;
    lp_i += 4 * stride;
    si = si + 1 == PD ? 0 : si + 1;
  };
#pragma unroll
  for (int i = 0; i < PD - 1; ++i) issue();

  int sc2 = 0;
  for (int64_t pb = warp0 * rpw; pb < pairs; pb += stride) {
    issue();
    cp_wait<PD - 1>();
    const uint8_t* st = pipe + (size_t)sc2 * NSLOT * SLOT_STRIDE;
    sc2 = sc2 + 1 == PD ? 0 : sc2 + 1;
    const int64_t p = pb + sub;
    const bool live = p < pairs;
    const int64_t m0 = live ? 2 * p : 0;
    float gv[2][8], yv[2][8], x0[2] = {0.f, 0.f}, x1[2] = {0.f, 0.f};
    float4 l4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      ld_vec8(st, SLOT_STRIDE, (const T*)nullptr, gv[0]);
      ld_vec8(st + NCP * SLOT_STRIDE, SLOT_STRIDE, (const T*)nullptr, gv[1]);
      ld_vec8(st + (2 * NCP) * SLOT_STRIDE, SLOT_STRIDE, (const T*)nullptr, yv[0]);
      if (y_full) ld_vec8(st + (3 * NCP) * SLOT_STRIDE, SLOT_STRIDE, (const T*)nullptr, yv[1]);
      ld_xs(st + (4 * NCP) * SLOT_STRIDE, (const T*)nullptr, 0, x0[0], x1[0]);
      ld_xs(st + (4 * NCP) * SLOT_STRIDE, (const T*)nullptr, 1, x0[1], x1[1]);
      l4 = *reinterpret_cast<const float4*>(st + (4 * NCP + 1) * SLOT_STRIDE);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) gv[0][e] = gv[1][e] = yv[0][e] = yv[1][e] = 0.f;
    }
    if (!y_full) {
#pragma unroll
      for (int e = 0; e < 8; ++e) yv[1][e] = yv[0][e];
    }
    // six dot products per row: <g,P>, <g,Q>, <g,sh>, <g,yv>, <g,U0>, <g,U1>
    float d[12];
    {
      float P[8], Q[8], SH[8], U0[8], U1[8];
      ldc(cs, X_P, C, cg, P);
      ldc(cs, X_Q, C, cg, Q);
      ldc(cs, X_SH, C, cg, SH);
      ldc(cs, X_U0, C, cg, U0);
      ldc(cs, X_U1, C, cg, U1);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float a = 0.f, b = 0.f, c = 0.f, dd = 0.f, u0 = 0.f, u1 = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float ge = gv[q][e];
          a = fmaf(ge, P[e], a);
          b = fmaf(ge, Q[e], b);
          c = fmaf(ge, SH[e], c);
          dd = fmaf(ge, yv[q][e], dd);
          u0 = fmaf(ge, U0[e], u0);
          u1 = fmaf(ge, U1[e], u1);
        }
        d[6 * q] = a; d[6 * q + 1] = b; d[6 * q + 2] = c; d[6 * q + 3] = dd; d[6 * q + 4] = u0; d[6 * q + 5] = u1;
      }
    }
    group_sum<12>(d, g.tpr);
    float s0[2], s1[2], dl0[2], dl1[2];
    const float lg[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      s0[q] = sigm(lg[2 * q]);
      s1[q] = sigm(lg[2 * q + 1]);
      const float t0 = fmaf(x0[q], d[6 * q], fmaf(x1[q], d[6 * q + 1], d[6 * q + 2]));
      const float t1 = d[6 * q + 3];
      dl0[q] = live ? t0 * s0[q] * (1.f - s0[q]) : 0.f;
      dl1[q] = live ? t1 * s1[q] * (1.f - s1[q]) : 0.f;
    }
    if (live && cg == 0) {
      rows[m0] = make_float4(s0[0] * d[4], s0[0] * d[5], dl0[0], dl1[0]);
      rows[m0 + 1] = make_float4(s0[1] * d[10], s0[1] * d[11], dl0[1], dl1[1]);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        aS[R_D0] += dl0[q]; aS[R_D1] += dl1[q];
        aS[R_D0X0] = fmaf(dl0[q], x0[q], aS[R_D0X0]); aS[R_D0X1] = fmaf(dl0[q], x1[q], aS[R_D0X1]);
        aS[R_D1X0] = fmaf(dl1[q], x0[q], aS[R_D1X0]); aS[R_D1X1] = fmaf(dl1[q], x1[q], aS[R_D1X1]);
        aS[R_X0] += x0[q]; aS[R_X1] += x1[q];
        aS[R_X00] = fmaf(x0[q], x0[q], aS[R_X00]); aS[R_X01] = fmaf(x0[q], x1[q], aS[R_X01]);
        aS[R_X11] = fmaf(x1[q], x1[q], aS[R_X11]);
      }
    }
    // per-channel sums and dy_prev
    {
      float wy0[8], wy1[8];
      ldc(cs, X_WY0, C, cg, wy0);
      ldc(cs, X_WY1, C, cg, wy1);
      float dyv[2][8];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float gs = live ? s0[q] : 0.f;
        const float gx0 = gs * x0[q], gx1 = gs * x1[q];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float ge = gv[q][e];
          aG0[e] = fmaf(ge, gs, aG0[e]);
          aG1[e] = fmaf(ge, gx0, aG1[e]);
          aG2[e] = fmaf(ge, gx1, aG2[e]);
          aW0[e] = fmaf(dl0[q], yv[q][e], aW0[e]);
          aW1[e] = fmaf(dl1[q], yv[q][e], aW1[e]);
          dyv[q][e] = fmaf(ge, s1[q], fmaf(wy0[e], dl0[q], wy1[e] * dl1[q]));
        }
      }
      if (live) {
        T* yo = dy + (yrow(g, m0) << cs_) + coff;
        if (y_full) {
          st8(yo, dyv[0]);
          st8(yo + ((int64_t)1 << cs_), dyv[1]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) dyv[0][e] += dyv[1][e];
          st8(yo, dyv[0]);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = cg * 8 + e;
    atomicAdd(&red[c], aG0[e]);
    atomicAdd(&red[C + c], aG1[e]);
    atomicAdd(&red[2 * C + c], aG2[e]);
    atomicAdd(&red[3 * C + c], aW0[e]);
    atomicAdd(&red[4 * C + c], aW1[e]);
  }
  if (cg == 0) {
#pragma unroll
    for (int e = 0; e < R_NSCAL; ++e) atomicAdd(&red[5 * C + e], aS[e]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 5 * C + R_NSCAL; i += XT) atomicAdd(acc + i, (double)red[i]);
}

// one block: batch sums -> dgamma / dbeta (sums), dW_att, db_att, dW1 and the 10 constants of the dx pass
__global__ void abf_xs2_finalize_kernel(const double* __restrict__ acc, int C, double invM, int training,
                                        const float* __restrict__ mean, const float* __restrict__ invstd,
                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                        const float* __restrict__ watt, const float* __restrict__ w1,
                                        double* __restrict__ sums, double* __restrict__ dwatt, double* __restrict__ dbatt,
                                        double* __restrict__ dw1, float* __restrict__ dxc) {
  __shared__ double sh[10][32];
  const double* S = acc + 5 * C;
  double part[10];
  for (int i = 0; i < 10; ++i) part[i] = 0.;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double g = gamma ? gamma[c] : 1.f, is = invstd[c], mu = mean[c];
    const double a = w1[2 * c], b = w1[2 * c + 1];
    const double sc = is * g, shv = (beta ? (double)beta[c] : 0.) - mu * sc;
    const double P = sc * a, Q = sc * b, al = a * is, be = b * is, ga = -mu * is, gi = g * is;
    const double wx0 = watt[c], wx1 = watt[2 * C + c];
    const double S1 = acc[c] + wx0 * S[R_D0] + wx1 * S[R_D1];
    const double SX0 = acc[C + c] + wx0 * S[R_D0X0] + wx1 * S[R_D1X0];
    const double SX1 = acc[2 * C + c] + wx0 * S[R_D0X1] + wx1 * S[R_D1X1];
    const double S2 = al * SX0 + be * SX1 + ga * S1;
    sums[c] = S1;                 // dbeta
    sums[C + c] = S2;             // dgamma
    dwatt[c] = P * S[R_D0X0] + Q * S[R_D0X1] + shv * S[R_D0];
    dwatt[C + c] = acc[3 * C + c];
    dwatt[2 * C + c] = P * S[R_D1X0] + Q * S[R_D1X1] + shv * S[R_D1];
    dwatt[3 * C + c] = acc[4 * C + c];
    const double k1 = training ? S1 * invM : 0., k2 = training ? S2 * invM : 0.;
    dw1[2 * c] = gi * (SX0 - k1 * S[R_X0] - k2 * (al * S[R_X00] + be * S[R_X01] + ga * S[R_X0]));
    dw1[2 * c + 1] = gi * (SX1 - k1 * S[R_X1] - k2 * (al * S[R_X01] + be * S[R_X11] + ga * S[R_X1]));
    const double U0 = a * gi, U1 = b * gi;
    part[0] += U0 * wx0; part[1] += U0 * wx1; part[2] += U0 * (k1 + ga * k2); part[3] += U0 * al * k2; part[4] += U0 * be * k2;
    part[5] += U1 * wx0; part[6] += U1 * wx1; part[7] += U1 * (k1 + ga * k2); part[8] += U1 * al * k2; part[9] += U1 * be * k2;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = 0; i < 10; ++i) {
    part[i] = warp_sum(part[i]);
    if (lane == 0) sh[i][w] = part[i];
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    double t = 0.;
    for (int k = 0; k < (int)((blockDim.x + 31) >> 5); ++k) t += sh[threadIdx.x][k];
    dxc[threadIdx.x] = (float)t;
  }
  if (threadIdx.x == 0) {
    dbatt[0] = S[R_D0];
    dbatt[1] = S[R_D1];
  }
}

// dx_k = e_k + dl0 c_ka + dl1 c_kb - c_kc - x0 c_kd - x1 c_ke   (24 bytes per row in, 4 or 8 out)
template <typename T>
__global__ void abf_xs2_dx_kernel(const float4* __restrict__ rows, const T* __restrict__ x, int64_t M,
                                  const float* __restrict__ dxc, T* __restrict__ dx) {
  float c[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) c[i] = __ldg(dxc + i);
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    const float4 r = rows[m];
    const float x0 = ld_f(x + 2 * m), x1 = ld_f(x + 2 * m + 1);
    st_f(dx + 2 * m, r.x + r.z * c[0] + r.w * c[1] - c[2] - x0 * c[3] - x1 * c[4]);
    st_f(dx + 2 * m + 1, r.y + r.z * c[5] + r.w * c[6] - c[7] - x0 * c[8] - x1 * c[9]);
  }
}

// =====================================================================================================
// BatchNorm backward of the ABF's conv1 (framework.py:209) folded into conv1's own gradients.  With z1 = W1 x (1x1 conv,
// no bias), xhat = (z1 - mu) is, gi = gamma is, k1 = mean dxp, k2 = mean dxp xhat (batch statistics; 0 in eval mode):
//     dz1 = gi (dxp - k1 - k2 xhat)
//     dx  = dz1 W1   = dxp (diag(gi) W1)  -  x (W1^T diag(gi k2 is) W1)  +  sum_c (gi (k2 is mu - k1))_c W1[c,:]
//     dW1 = dz1^T x  = diag(gi) ( dxp^T x  -  k1 (sum x)^T  -  diag(k2 is) ( W1 (x^T x) - mu (sum x)^T ) )
// so the data gradient is ONE two-source 1x1 contraction [dxp | x] with the weights below (+ a bias), and the weight
// gradient needs dxp^T x (the usual weight-gradient launch on dxp), the Cin x Cin Gram matrix of x and its column sums.
// =====================================================================================================
// weff: bf16 [Cin][C + c1p] (row n = output channel of the data gradient; columns: dxp channels, then x channels,
// zero padded to c1p);  bias: fp32 [Cin]
__global__ void abf_fold_dgrad_kernel(const float* __restrict__ w1, const float* __restrict__ gamma,
                                      const float* __restrict__ mean, const float* __restrict__ invstd,
                                      const double* __restrict__ sums, double invM, int training, int C, int Cin, int c1p,
                                      __nv_bfloat16* __restrict__ weff, float* __restrict__ bias) {
  __shared__ float s_gi[256], s_q[256], s_r[256], s_wn[256];
  __shared__ float s_red[8];
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float gi = (gamma ? gamma[c] : 1.f) * invstd[c];
    const float k1 = training ? (float)(sums[c] * invM) : 0.f;
    const float k2 = training ? (float)(sums[C + c] * invM) : 0.f;
    s_gi[c] = gi;
    s_q[c] = gi * k2 * invstd[c];
    s_r[c] = gi * (k2 * invstd[c] * mean[c] - k1);
    s_wn[c] = w1[(size_t)c * Cin + n];
  }
  __syncthreads();
  const int Kt = C + c1p;
  for (int k = threadIdx.x; k < Kt; k += blockDim.x) {
    float v = 0.f;
    if (k < C) v = s_gi[k] * s_wn[k];
    else if (k - C < Cin) {
      const int j = k - C;
      float a = 0.f;
      for (int c = 0; c < C; ++c) a = fmaf(s_wn[c] * s_q[c], w1[(size_t)c * Cin + j], a);
      v = -a;
    }
    weff[(size_t)n * Kt + k] = __float2bfloat16(v);
  }
  float b = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) b = fmaf(s_r[c], s_wn[c], b);
  for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = b;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];
    bias[n] = t;
  }
}

// Batch statistics of z1 = W1 x from the moments of x (clskd_colgram): sum_c = W1[c] . sx,  sumsq_c = W1[c]^T G W1[c].
// W1 is rounded to bf16 first - the tcgen05 conv contracts the bf16 copy of the weight, and the statistics must belong to
// the z1 it produces.  A block stages G as fp32 in shared memory once and serves 8 channels (one warp each): the Cin^2
// inner products run as fp32 FMAs (the fp64 pipe of this GPU made the all-double version 28 us per launch), the final
// sums over j in fp64.
constexpr int FOLD_CPB = 8;      // channels per block
__global__ void __launch_bounds__(256) abf_fold_stats_kernel(const double* __restrict__ G, const double* __restrict__ sx,
                                                             const float* __restrict__ w1, int C, int Cin,
                                                             double* __restrict__ sum, double* __restrict__ sumsq) {
  extern __shared__ float fs_sm[];                 // Gf[Cin*Cin], wq[FOLD_CPB][Cin]
  float* Gf = fs_sm;
  float* wq = fs_sm + (size_t)Cin * Cin;
  for (int i = threadIdx.x; i < Cin * Cin; i += blockDim.x) Gf[i] = (float)G[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * FOLD_CPB + warp;
  if (c < C)
    for (int j = lane; j < Cin; j += 32) wq[warp * Cin + j] = __bfloat162float(__float2bfloat16_rn(w1[(size_t)c * Cin + j]));
  __syncthreads();
  if (c >= C) return;
  const float* w = wq + warp * Cin;
  double s = 0.0, q = 0.0;
  for (int j = lane; j < Cin; j += 32) {
    float t = 0.f;
    for (int k = 0; k < Cin; ++k) t = fmaf(Gf[k * Cin + j], w[k], t);        // G is symmetric: column j, conflict free
    s += (double)w[j] * sx[j];
    q += (double)w[j] * (double)t;
  }
  s = warp_sum(s);
  q = warp_sum(q);
  if (lane == 0) {
    sum[c] = s;
    sumsq[c] = q;
  }
}

// P: fp32 [Cin][C] = x^T dxp;  G: fp64 [Cin][Cin] = x^T x;  sx: fp64 [Cin] column sums of x;  dw1: fp32 [C][Cin].
// A block stages G as fp32 in shared memory and serves FOLD_CPB channels; W1 (x^T x) in fp32 FMAs, the rest in fp64.
__global__ void __launch_bounds__(256) abf_fold_dw1_kernel(const float* __restrict__ P, const double* __restrict__ G,
                                                           const double* __restrict__ sx, const float* __restrict__ w1,
                                                           const float* __restrict__ gamma, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, const double* __restrict__ sums,
                                                           double invM, int training, int C, int Cin,
                                                           float* __restrict__ dw1) {
  extern __shared__ float fs_sm[];                 // Gf[Cin*Cin] (training only)
  float* Gf = fs_sm;
  if (training) {
    for (int i = threadIdx.x; i < Cin * Cin; i += blockDim.x) Gf[i] = (float)G[i];
    __syncthreads();
  }
  const int c_beg = blockIdx.x * FOLD_CPB;
  const int n = min(FOLD_CPB, C - c_beg) * Cin;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = c_beg + i / Cin, k = i % Cin;
    const double gi = (double)(gamma ? gamma[c] : 1.f) * (double)invstd[c];
    double v = (double)P[(size_t)k * C + c];
    if (training) {
      const double k1 = sums[c] * invM, k2 = sums[C + c] * invM;
      const float* w = w1 + (size_t)c * Cin;
      float t = 0.f;
      for (int j = 0; j < Cin; ++j) t = fmaf(w[j], Gf[j * Cin + k], t);
      v -= k1 * sx[k] + k2 * (double)invstd[c] * ((double)t - (double)mean[c] * sx[k]);
    }
    dw1[(size_t)c * Cin + k] = (float)(gi * v);
  }
}

const char* abf_unsupported(int B, int T, int F, int Fy, int C, const void* a, const void* b, const void* c) {
  if (C % 8 || C > 256 || C < 8) return "C must be a multiple of 8 in [8,256]";
  const int tpr = C / 8;
  if (tpr & (tpr - 1)) return "C/8 must be a power of two";
  if (!(Fy == F || 2 * Fy == F)) return "y_prev must have F or F/2 frequency rows";
  if (F % 2) return "F must be even";
  if (((uintptr_t)a % 16) || ((uintptr_t)b % 16) || (c && ((uintptr_t)c % 16))) return "tensors must be 16-byte aligned";
  return nullptr;
}

int ilog2(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }

int ew_grid_abf(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

int abf_grid(int64_t warp_iters) {
  int64_t blocks = (warp_iters + 7) / 8;
  // the kernels hold two CTAs per SM (launch bounds): a grid of exactly that is one wave - the constants are staged
  // once per SM slot and the per-column fp64 atomics of the statistics pass (about 20 ns each on one address) come from
  // 296 CTAs instead of 1184
  const int64_t cap = (int64_t)sm_count() * 2;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace
}  // namespace clskd

using namespace clskd;

extern "C" int clskd_abf_mid_supported(int B, int T, int F, int Fy, int C) {
  return abf_unsupported(B, T, F, Fy, C, nullptr, nullptr, nullptr) == nullptr ? 1 : 0;
}

// shared launchers: xs == false: `z1` is the stored conv1 output [M, C]; xs == true: `z1` is the 2-channel
// conv1 INPUT [M, 2] and w1 its weight [C][2]
static int abf_fwd_launch(const char* who, bool xs, const void* z1, const void* y, int dtype, int B, int T, int F,
                          int Fy, int C, const float* mean, const float* invstd, const float* gamma,
                          const float* beta, const float* watt, const float* batt, const float* w1, void* xb,
                          float* logits, void* stream) {
  CLSKD_CHECK_ARG(z1 && y && mean && invstd && watt && xb && logits && (!xs || w1), "%s: null pointer", who);
  if (const char* why = abf_unsupported(B, T, F, Fy, C, xs ? y : z1, y, xb)) {
    set_error("%s: unsupported: %s", who, why);
    return CLSKD_ERR_UNSUPPORTED;
  }
  CLSKD_CHECK_ARG(!xs || ((uintptr_t)z1 % 16) == 0, "%s: x must be 16-byte aligned", who);
  AbfGeom g;
  g.M = (int64_t)B * T * F; g.F = F; g.Fy = Fy; g.C = C; g.tpr = C / 8;
  g.cshift = ilog2(C); g.yshift = Fy == F ? 0 : 1;
  if (g.M == 0) return CLSKD_OK;
  const int rows_per_warp_iter = (32 / g.tpr) * 2;
  const int grid = abf_grid((g.M + rows_per_warp_iter - 1) / rows_per_warp_iter / 4);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t es = dtype == CLSKD_F32 ? 4 : 2;
  const size_t cbytes = (sizeof(float) * NCONST * (size_t)C + 15) & ~(size_t)15;
  const size_t sh = cbytes + (size_t)PD * (4 * (es / 2)) * AT * 16;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(abf_mid_fwd_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(abf_mid_fwd_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(abf_mid_fwd_kernel<__nv_bfloat16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(abf_mid_fwd_kernel<__nv_bfloat16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  if (xs)
    CLSKD_DISPATCH_DTYPE(dtype, TT, (abf_mid_fwd_kernel<TT, true><<<grid, AT, sh, st>>>(
                                        (const TT*)z1, (const TT*)y, g, mean, invstd, gamma, beta, watt, batt, w1,
                                        (TT*)xb, logits)));
  else
    CLSKD_DISPATCH_DTYPE(dtype, TT, (abf_mid_fwd_kernel<TT, false><<<grid, AT, sh, st>>>(
                                        (const TT*)z1, (const TT*)y, g, mean, invstd, gamma, beta, watt, batt, nullptr,
                                        (TT*)xb, logits)));
  CLSKD_CHECK_LAUNCH(who);
  return CLSKD_OK;
}

static int abf_bwd_launch(const char* who, bool xs, const void* gout, const void* z1, const void* y, int dtype, int B,
                          int T, int F, int Fy, int C, const float* mean, const float* invstd, const float* gamma,
                          const float* beta, const float* watt, const float* logits, int training, double* sums,
                          double* dwatt, double* dbatt, void* dz1, void* dy, const float* w1, double* dw1,
                          void* stream, bool fold = false) {
  CLSKD_CHECK_ARG(gout && z1 && y && mean && invstd && watt && logits && sums && dwatt && dbatt && dz1 && dy &&
                      (!xs || (w1 && dw1)), "%s: null pointer", who);
  if (const char* why = abf_unsupported(B, T, F, Fy, C, gout, xs ? gout : z1, xs ? nullptr : dz1)) {
    set_error("%s: unsupported: %s", who, why);
    return CLSKD_ERR_UNSUPPORTED;
  }
  CLSKD_CHECK_ARG(((uintptr_t)y % 16) == 0 && ((uintptr_t)dy % 16) == 0, "%s: y / dy must be 16-byte aligned", who);
  CLSKD_CHECK_ARG(!xs || (((uintptr_t)z1 % 16) == 0 && ((uintptr_t)dz1 % 16) == 0), "%s: x / dx must be 16-byte aligned", who);
  AbfGeom g;
  g.M = (int64_t)B * T * F; g.F = F; g.Fy = Fy; g.C = C; g.tpr = C / 8;
  g.cshift = ilog2(C); g.yshift = Fy == F ? 0 : 1;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = zero_spans(st, sums, sizeof(double) * 2 * C, dwatt, sizeof(double) * 4 * C, dbatt, sizeof(double) * 2,
                             xs ? dw1 : nullptr, sizeof(double) * 2 * C);
  if (e != cudaSuccess) { set_error("%s: memset: %s", who, cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
  if (g.M == 0) return CLSKD_OK;
  const int64_t pairs = g.M / 2;
  const int pairs_per_warp = 32 / g.tpr;
  const int grid = abf_grid((pairs + pairs_per_warp - 1) / pairs_per_warp / 4);
  const size_t es = dtype == CLSKD_F32 ? 4 : 2;
  const size_t cbytes = (sizeof(float) * ((NCONST + 6) * (size_t)C + 2) + 15) & ~(size_t)15;
  const size_t sh = cbytes + (size_t)PD * (6 * (es / 2) + 1) * AT * 16;
  static bool attr = false;
  if (!attr) {
#define ABF_ATTR(TT, MD, XX) cudaFuncSetAttribute(abf_mid_bwd_kernel<TT, MD, XX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
    ABF_ATTR(float, 0, false); ABF_ATTR(float, 1, false); ABF_ATTR(__nv_bfloat16, 0, false); ABF_ATTR(__nv_bfloat16, 1, false);
    ABF_ATTR(float, 0, true); ABF_ATTR(float, 1, true); ABF_ATTR(__nv_bfloat16, 0, true); ABF_ATTR(__nv_bfloat16, 1, true);
#undef ABF_ATTR
    cudaFuncSetAttribute(abf_mid_bwd_kernel<float, 2, false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(abf_mid_bwd_kernel<__nv_bfloat16, 2, false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
#define ABF_BWD(MD, XX, DZ, DY)                                                                                     \
  CLSKD_DISPATCH_DTYPE(dtype, TT, (abf_mid_bwd_kernel<TT, MD, XX><<<grid, AT, sh, st>>>(                            \
                                      (const TT*)gout, (const TT*)z1, (const TT*)y, g, mean, invstd, gamma, beta,   \
                                      watt, logits, sums, dwatt, dbatt, training, (TT*)(DZ), (TT*)(DY), w1, dw1)))
  if (fold) {
    // one pass: statistics + dxp + dy_prev; 128-thread CTAs, three per SM (one wave)
    constexpr int FT = 128;
    int64_t blocks = ((pairs + pairs_per_warp - 1) / pairs_per_warp / 4 + FT / 32 - 1) / (FT / 32);
    const int64_t cap = (int64_t)sm_count() * 3;
    const int fgrid = (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
    const size_t fsh = cbytes + (size_t)PD * (6 * (es / 2) + 1) * FT * 16;
    CLSKD_DISPATCH_DTYPE(dtype, TT, (abf_mid_bwd_kernel<TT, 2, false, FT><<<fgrid, FT, fsh, st>>>(
                                        (const TT*)gout, (const TT*)z1, (const TT*)y, g, mean, invstd, gamma, beta, watt,
                                        logits, sums, dwatt, dbatt, training, (TT*)dz1, (TT*)dy, w1, dw1)));
    CLSKD_CHECK_LAUNCH(who);
    return CLSKD_OK;
  }
  if (xs) ABF_BWD(0, true, nullptr, nullptr); else ABF_BWD(0, false, nullptr, nullptr);
  CLSKD_CHECK_LAUNCH(who);
  if (xs) ABF_BWD(1, true, dz1, dy); else ABF_BWD(1, false, dz1, dy);
#undef ABF_BWD
  CLSKD_CHECK_LAUNCH(who);
  return CLSKD_OK;
}

extern "C" int clskd_abf_mid_fwd(const void* z1, const void* y, int dtype, int B, int T, int F, int Fy, int C,
                                 const float* mean, const float* invstd, const float* gamma, const float* beta,
                                 const float* watt, const float* batt, void* xb, float* logits, void* stream) {
  return abf_fwd_launch("clskd_abf_mid_fwd", false, z1, y, dtype, B, T, F, Fy, C, mean, invstd, gamma, beta, watt, batt,
                        nullptr, xb, logits, stream);
}

extern "C" int clskd_abf_mid_bwd(const void* gout, const void* z1, const void* y, int dtype, int B, int T, int F, int Fy,
                                 int C, const float* mean, const float* invstd, const float* gamma, const float* beta,
                                 const float* watt, const float* logits, int training, double* sums, double* dwatt,
                                 double* dbatt, void* dz1, void* dy, void* stream) {
  return abf_bwd_launch("clskd_abf_mid_bwd", false, gout, z1, y, dtype, B, T, F, Fy, C, mean, invstd, gamma, beta, watt,
                        logits, training, sums, dwatt, dbatt, dz1, dy, nullptr, nullptr, stream);
}

extern "C" int clskd_abf_mid_bwd_fold(const void* gout, const void* z1, const void* y, int dtype, int B, int T, int F,
                                      int Fy, int C, const float* mean, const float* invstd, const float* gamma,
                                      const float* beta, const float* watt, const float* logits, double* sums,
                                      double* dwatt, double* dbatt, void* dxp, void* dy, void* stream) {
  return abf_bwd_launch("clskd_abf_mid_bwd_fold", false, gout, z1, y, dtype, B, T, F, Fy, C, mean, invstd, gamma, beta,
                        watt, logits, 1, sums, dwatt, dbatt, dxp, dy, nullptr, nullptr, stream, true);
}

extern "C" int clskd_abf_fold_dgrad(const float* w1, const float* gamma, const float* mean, const float* invstd,
                                    const double* sums, int64_t M, int training, int C, int Cin, int c1p, void* weff,
                                    float* bias, void* stream) {
  CLSKD_CHECK_ARG(w1 && mean && invstd && sums && weff && bias, "clskd_abf_fold_dgrad: null pointer");
  CLSKD_CHECK_ARG(C >= 1 && C <= 256 && Cin >= 1 && Cin <= 256 && c1p >= Cin && M >= 1, "clskd_abf_fold_dgrad: extents");
  abf_fold_dgrad_kernel<<<Cin, 256, 0, (cudaStream_t)stream>>>(w1, gamma, mean, invstd, sums, 1.0 / (double)M, training, C,
                                                               Cin, c1p, (__nv_bfloat16*)weff, bias);
  CLSKD_CHECK_LAUNCH("clskd_abf_fold_dgrad");
  return CLSKD_OK;
}

extern "C" int clskd_abf_fold_stats(const double* G, const double* sx, const float* w1, int C, int Cin, double* sum,
                                    double* sumsq, void* stream) {
  CLSKD_CHECK_ARG(G && sx && w1 && sum && sumsq, "clskd_abf_fold_stats: null pointer");
  CLSKD_CHECK_ARG(C >= 1 && Cin >= 1 && Cin <= 1024, "clskd_abf_fold_stats: extents");
  CLSKD_CHECK_ARG(Cin <= 128, "clskd_abf_fold_stats: Cin <= 128");
  const size_t sh = sizeof(float) * ((size_t)Cin * Cin + (size_t)FOLD_CPB * Cin);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(abf_fold_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    attr = true;
  }
  abf_fold_stats_kernel<<<(C + FOLD_CPB - 1) / FOLD_CPB, 256, sh, (cudaStream_t)stream>>>(G, sx, w1, C, Cin, sum, sumsq);
  CLSKD_CHECK_LAUNCH("clskd_abf_fold_stats");
  return CLSKD_OK;
}

extern "C" int clskd_abf_fold_dw1(const float* P, const double* G, const double* sx, const float* w1, const float* gamma,
                                  const float* mean, const float* invstd, const double* sums, int64_t M, int training,
                                  int C, int Cin, float* dw1, void* stream) {
  CLSKD_CHECK_ARG(P && w1 && mean && invstd && sums && dw1 && (!training || (G && sx)), "clskd_abf_fold_dw1: null pointer");
  CLSKD_CHECK_ARG(C >= 1 && C <= 256 && Cin >= 1 && Cin <= 256 && M >= 1, "clskd_abf_fold_dw1: extents");
  CLSKD_CHECK_ARG(!training || Cin <= 128, "clskd_abf_fold_dw1: Cin <= 128");
  const size_t sh = training ? sizeof(float) * (size_t)Cin * Cin : 0;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(abf_fold_dw1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    attr = true;
  }
  abf_fold_dw1_kernel<<<(C + FOLD_CPB - 1) / FOLD_CPB, 256, sh, (cudaStream_t)stream>>>(P, G, sx, w1, gamma, mean, invstd, sums,
                                                                                      1.0 / (double)M, training, C, Cin, dw1);
  CLSKD_CHECK_LAUNCH("clskd_abf_fold_dw1");
  return CLSKD_OK;
}

extern "C" int clskd_abf_mid_xs_fwd(const void* x, const float* w1, const void* y, int dtype, int B, int T, int F, int Fy,
                                    int C, const float* mean, const float* invstd, const float* gamma,
                                    const float* beta, const float* watt, const float* batt, void* xb, float* logits,
                                    void* stream) {
  return abf_fwd_launch("clskd_abf_mid_xs_fwd", true, x, y, dtype, B, T, F, Fy, C, mean, invstd, gamma, beta, watt, batt,
                        w1, xb, logits, stream);
}

extern "C" int clskd_abf_mid_xs_bwd(const void* gout, const void* x, const float* w1, const void* y, int dtype, int B,
                                    int T, int F, int Fy, int C, const float* mean, const float* invstd,
                                    const float* gamma, const float* beta, const float* watt, const float* logits,
                                    int training, double* sums, double* dwatt, double* dbatt, double* dw1, void* dx,
                                    void* dy, void* stream) {
  return abf_bwd_launch("clskd_abf_mid_xs_bwd", true, gout, x, y, dtype, B, T, F, Fy, C, mean, invstd, gamma, beta, watt,
                        logits, training, sums, dwatt, dbatt, dx, dy, w1, dw1, stream);
}


// ---- XS2 launchers (rank-2 folded kernels above)
extern "C" int clskd_abf_xs2_fwd(const void* x, const float* w1, const void* y, int dtype, int B, int T, int F, int Fy,
                                 int C, const float* mean, const float* invstd, const float* gamma, const float* beta,
                                 const float* watt, const float* batt, void* xb, float* logits, void* stream) {
  const char* who = "clskd_abf_xs2_fwd";
  CLSKD_CHECK_ARG(x && w1 && y && mean && invstd && watt && xb && logits, "%s: null pointer", who);
  if (const char* why = abf_unsupported(B, T, F, Fy, C, y, y, xb)) {
    set_error("%s: unsupported: %s", who, why);
    return CLSKD_ERR_UNSUPPORTED;
  }
  CLSKD_CHECK_ARG(((uintptr_t)x % 16) == 0 && C >= 16, "%s: x must be 16-byte aligned and C >= 16", who);
  AbfGeom g;
  g.M = (int64_t)B * T * F; g.F = F; g.Fy = Fy; g.C = C; g.tpr = C / 8;
  g.cshift = ilog2(C); g.yshift = Fy == F ? 0 : 1;
  if (g.M == 0) return CLSKD_OK;
  const int64_t pairs = g.M / 2;
  const int ppw = 32 / g.tpr;
  const int grid = abf_grid((pairs + ppw - 1) / ppw / 4);
  const size_t es = dtype == CLSKD_F32 ? 4 : 2;
  const size_t cbytes = (sizeof(float) * (X_NCONST * (size_t)C + XS_NSCAL) + 15) & ~(size_t)15;
  const size_t sh = cbytes + (size_t)PD * (2 * (es / 2) + 1) * AT * 16;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(abf_xs2_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(abf_xs2_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  CLSKD_DISPATCH_DTYPE(dtype, TT, (abf_xs2_fwd_kernel<TT><<<grid, AT, sh, (cudaStream_t)stream>>>(
                                      (const TT*)x, (const TT*)y, g, mean, invstd, gamma, beta, watt, batt, w1, (TT*)xb, logits)));
  CLSKD_CHECK_LAUNCH(who);
  return CLSKD_OK;
}

// workspace of clskd_abf_xs2_bwd: 16 bytes per row + the fp64 sums + the dx constants
static int64_t xs2_ws_bytes(int B, int T, int F, int C) {
  const int64_t M = (int64_t)B * T * F;
  return M * 16 + (int64_t)sizeof(double) * (5 * (int64_t)C + R_NSCAL) + 64 + 256;
}
extern "C" int clskd_abf_xs2_bwd_workspace(int B, int T, int F, int C, int64_t* bytes) {
  CLSKD_CHECK_ARG(bytes, "clskd_abf_xs2_bwd_workspace: null pointer");
  *bytes = xs2_ws_bytes(B, T, F, C);
  return CLSKD_OK;
}

extern "C" int clskd_abf_xs2_bwd(const void* gout, const void* x, const float* w1, const void* y, int dtype, int B, int T,
                                 int F, int Fy, int C, const float* mean, const float* invstd, const float* gamma,
                                 const float* beta, const float* watt, const float* logits, int training, double* sums,
                                 double* dwatt, double* dbatt, double* dw1, void* dx, void* dy, void* workspace,
                                 int64_t ws_bytes, void* stream) {
  const char* who = "clskd_abf_xs2_bwd";
  CLSKD_CHECK_ARG(gout && x && w1 && y && mean && invstd && watt && logits && sums && dwatt && dbatt && dw1 && dx && dy &&
                      workspace, "%s: null pointer", who);
  if (const char* why = abf_unsupported(B, T, F, Fy, C, gout, y, dy)) {
    set_error("%s: unsupported: %s", who, why);
    return CLSKD_ERR_UNSUPPORTED;
  }
  CLSKD_CHECK_ARG(((uintptr_t)x % 16) == 0 && ((uintptr_t)workspace % 16) == 0 && C >= 16 && C <= 1024,
                  "%s: x / workspace must be 16-byte aligned, 16 <= C <= 1024", who);
  CLSKD_CHECK_ARG(ws_bytes >= xs2_ws_bytes(B, T, F, C), "%s: workspace too small", who);
  AbfGeom g;
  g.M = (int64_t)B * T * F; g.F = F; g.Fy = Fy; g.C = C; g.tpr = C / 8;
  g.cshift = ilog2(C); g.yshift = Fy == F ? 0 : 1;
  cudaStream_t st = (cudaStream_t)stream;
  float4* rows = reinterpret_cast<float4*>(workspace);
  double* acc = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(workspace) + g.M * 16);
  float* dxc = reinterpret_cast<float*>(acc + 5 * (size_t)C + R_NSCAL);
  cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(double) * (5 * (size_t)C + R_NSCAL), st);
  if (e != cudaSuccess) { set_error("%s: memset: %s", who, cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
  const size_t es = dtype == CLSKD_F32 ? 4 : 2;
  if (g.M > 0) {
    const int64_t pairs = g.M / 2;
    const int ppw = 32 / g.tpr;
    int64_t blocks = ((pairs + ppw - 1) / ppw / 4 + XT / 32 - 1) / (XT / 32);
    const int64_t cap = (int64_t)sm_count() * 3;
    const int grid = (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
    const size_t cbytes = (sizeof(float) * (X_NCONST * (size_t)C + XS_NSCAL + 5 * (size_t)C + R_NSCAL) + 15) & ~(size_t)15;
    const size_t sh = cbytes + (size_t)PD * (4 * (es / 2) + 2) * XT * 16;
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(abf_xs2_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(abf_xs2_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr = true;
    }
    CLSKD_DISPATCH_DTYPE(dtype, TT, (abf_xs2_bwd_kernel<TT><<<grid, XT, sh, st>>>(
                                        (const TT*)gout, (const TT*)x, (const TT*)y, g, mean, invstd, gamma, beta, watt, w1,
                                        logits, acc, rows, (TT*)dy)));
    CLSKD_CHECK_LAUNCH(who);
  }
  abf_xs2_finalize_kernel<<<1, 256, 0, st>>>(acc, C, g.M > 0 ? 1.0 / (double)g.M : 0.0, training, mean, invstd, gamma, beta,
                                            watt, w1, sums, dwatt, dbatt, dw1, dxc);
  CLSKD_CHECK_LAUNCH(who);
  if (g.M > 0) {
    const int grid = ew_grid_abf(g.M);
    CLSKD_DISPATCH_DTYPE(dtype, TT, (abf_xs2_dx_kernel<TT><<<grid, 256, 0, st>>>(rows, (const TT*)x, g.M, dxc, (TT*)dx)));
    CLSKD_CHECK_LAUNCH(who);
  }
  return CLSKD_OK;
}

// column sums of z = W1 x for a 2-channel x from the moments of x: sum_c = w0 S0 + w1 S1,
// sumsq_c = w0^2 S00 + 2 w0 w1 S01 + w1^2 S11   (s5 = clskd_cbn_moments of x viewed as one complex channel)
__global__ void rank2_colstats_kernel(const double* __restrict__ s5, const float* __restrict__ w1, int C,
                                      double* __restrict__ sum, double* __restrict__ sumsq) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double a = w1[2 * c], b = w1[2 * c + 1];
  sum[c] = a * s5[0] + b * s5[1];
  sumsq[c] = a * a * s5[2] + 2. * a * b * s5[3] + b * b * s5[4];
}

extern "C" int clskd_rank2_colstats(const double* s5, const float* w1, int C, double* sum, double* sumsq,
                                    void* stream) {
  CLSKD_CHECK_ARG(s5 && w1 && sum && sumsq && C >= 1, "clskd_rank2_colstats: bad arguments");
  rank2_colstats_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(s5, w1, C, sum, sumsq);
  CLSKD_CHECK_LAUNCH("clskd_rank2_colstats");
  return CLSKD_OK;
}
