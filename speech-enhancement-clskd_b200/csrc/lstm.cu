// LSTM recurrence of the complex LSTM (NavieComplexLSTM, tools_for_model.py:138-178).
//
// The input projections (x W_ih^T + b_ih + b_hh) are batched GEMMs done by tapconv; what is left
// is strictly sequential: per time step a [RB x H] x [H x 4H] product plus the gate math.  That is
// latency bound, so one persistent CTA owns RB=4 batch rows of one weight set for the whole
// sequence: W_hh stays resident in shared memory (fp32 up to H=64, bf16 up to H=128; larger falls
// back to L2-resident global reads), h lives in shared memory, c in a register, and the next step's
// pre-activations are prefetched while the current step computes.  The four LSTM passes of a
// complex layer (2 weight sets x {real,imag} rows) run concurrently as independent CTAs.
#include "common.cuh"

namespace clskd {
int g_lstm_legacy = 0;     // clskd_set_tuning key 7: 1 = CUDA-core recurrence kernels only (A/B timing, tests)
namespace {

constexpr int RB = 4;

__device__ __forceinline__ float sigm(float v) { return 1.f / (1.f + expf(-v)); }

// One thread owns hidden unit j of RB = 4 batch rows: all four gates of its cells, so the gate
// math needs no exchange and the only per-step traffic through shared memory is the broadcast read
// of h (one float4 = the 4 rows of h_k per k) and the 4-gate weight vector W[k][j][0..3] (8 bytes
// in bf16, 16 bytes in fp32): 2 shared loads per 16 FMAs.  h is double-buffered -> one barrier per
// time step.  WMODE 0: weights fp32 in smem; 1: bf16 in smem; 2: fp32 from global / L2.
template <int WMODE, int RBF>
__global__ void lstm_fwd_kernel(const float* __restrict__ pre, const float* __restrict__ whh_t,
                                int T, int R, int Bp, int H, int64_t pre_pstride,
                                int64_t pre_tstride, int64_t pre_ld, int64_t pre_set_stride,
                                int64_t whh_set_stride, float* __restrict__ h_out,
                                float* __restrict__ gates_out, float* __restrict__ c_out,
                                const float* h0, const float* c0, float* hN, float* cN) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int G = 4 * H;
  const int j = threadIdx.x;  // hidden unit
  const int set = blockIdx.y;
  const int r0 = blockIdx.x * RBF;
  const int64_t out_set_stride = (int64_t)T * R * H;
  pre += (int64_t)set * pre_set_stride;
  whh_t += (int64_t)set * whh_set_stride;
  h_out += (int64_t)set * out_set_stride;
  if (gates_out) gates_out += (int64_t)set * out_set_stride * 4;
  if (c_out) c_out += (int64_t)set * out_set_stride;

  float* h_s = reinterpret_cast<float*>(smem_raw);                   // [2][H][RBF] : (rows of h_k)
  float4* w_f = reinterpret_cast<float4*>(h_s + 2 * H * 4);           // [H][H] float4 (WMODE 0)
  uint2* w_b = reinterpret_cast<uint2*>(h_s + 2 * H * 4);             // [H][H] 4 x bf16 (WMODE 1)

  // W_hh^T is [k][g*H + j]; stage it as [k][j][g]
  if (WMODE == 0) {
    for (int i = j; i < H * H; i += blockDim.x) {
      const int k = i / H, jj = i - k * H;
      const float* wp = whh_t + (int64_t)k * G + jj;
      w_f[i] = make_float4(wp[0], wp[H], wp[2 * H], wp[3 * H]);
    }
  } else if (WMODE == 1) {
    for (int i = j; i < H * H; i += blockDim.x) {
      const int k = i / H, jj = i - k * H;
      const float* wp = whh_t + (int64_t)k * G + jj;
      __nv_bfloat162 a = __floats2bfloat162_rn(wp[0], wp[H]), b = __floats2bfloat162_rn(wp[2 * H], wp[3 * H]);
      w_b[i] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    }
  }
  // initial state [nsets][R][H] (streaming inference carries it across time chunks); zero when NULL
  const int64_t st_set = (int64_t)set * R * H;
  float c_state[RBF];
#pragma unroll
  for (int r = 0; r < RBF; ++r) {
    const bool ok = r0 + r < R;
    h_s[j * RBF + r] = (h0 && ok) ? h0[st_set + (int64_t)(r0 + r) * H + j] : 0.f;
    c_state[r] = (c0 && ok) ? c0[st_set + (int64_t)(r0 + r) * H + j] : 0.f;
  }

  // pre-activation addresses of (row r, gate g, unit j) without the time term
  int64_t prow[RBF];
  int64_t orow[RBF];
#pragma unroll
  for (int r = 0; r < RBF; ++r) {
    const int rr = min(r0 + r, R - 1);
    prow[r] = (int64_t)(rr / Bp) * pre_pstride + (int64_t)(rr % Bp) * pre_ld + j;
    orow[r] = (int64_t)(rr / Bp) * T * Bp + (rr % Bp);                // output row ((part*T + t)*Bp + b) at t = 0
  }
  float pcur[RBF][4], pnext[RBF][4];
#pragma unroll
  for (int r = 0; r < RBF; ++r)
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      pcur[r][g] = (r0 + r < R && T > 0) ? pre[prow[r] + g * H] : 0.f;
      pnext[r][g] = 0.f;
    }
  __syncthreads();

  for (int t = 0; t < T; ++t) {
    if (t + 1 < T) {
#pragma unroll
      for (int r = 0; r < RBF; ++r)
        if (r0 + r < R) {
#pragma unroll
          for (int g = 0; g < 4; ++g) pnext[r][g] = pre[prow[r] + (int64_t)(t + 1) * pre_tstride + g * H];
        }
    }
    float acc[RBF][4];
#pragma unroll
    for (int r = 0; r < RBF; ++r)
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[r][g] = pcur[r][g];
    const float* hb = h_s + (t & 1) * H * RBF;
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
      float w[4];
      if (WMODE == 0) {
        const float4 wv = w_f[k * H + j];
        w[0] = wv.x; w[1] = wv.y; w[2] = wv.z; w[3] = wv.w;
      } else if (WMODE == 1) {
        const uint2 wv = w_b[k * H + j];
        w[0] = __uint_as_float(wv.x << 16);
        w[1] = __uint_as_float(wv.x & 0xffff0000u);
        w[2] = __uint_as_float(wv.y << 16);
        w[3] = __uint_as_float(wv.y & 0xffff0000u);
      } else {
        const float* wp = whh_t + (int64_t)k * G + j;
        w[0] = __ldg(wp); w[1] = __ldg(wp + H); w[2] = __ldg(wp + 2 * H); w[3] = __ldg(wp + 3 * H);
      }
      float hr[RBF];
      if (RBF == 4) {
        const float4 hv = *reinterpret_cast<const float4*>(hb + k * 4);
        hr[0] = hv.x; hr[1] = hv.y; hr[2 % RBF] = hv.z; hr[3 % RBF] = hv.w;
      } else {
        const float2 hv = *reinterpret_cast<const float2*>(hb + k * 2);
        hr[0] = hv.x; hr[1] = hv.y;
      }
#pragma unroll
      for (int r = 0; r < RBF; ++r)
#pragma unroll
        for (int g = 0; g < 4; ++g) acc[r][g] = fmaf(w[g], hr[r], acc[r][g]);
    }
    float hn[RBF];
#pragma unroll
    for (int r = 0; r < RBF; ++r) {
      const float gi = sigm(acc[r][0]), gf = sigm(acc[r][1]), gg = tanhf(acc[r][2]), go = sigm(acc[r][3]);
      c_state[r] = gf * c_state[r] + gi * gg;
      hn[r] = go * tanhf(c_state[r]);
      if (r0 + r < R) {
        const int64_t row = orow[r] + (int64_t)t * Bp;
        h_out[row * H + j] = hn[r];
        if (gates_out) {
          float* gp = gates_out + row * G;
          gp[j] = gi;
          gp[H + j] = gf;
          gp[2 * H + j] = gg;
          gp[3 * H + j] = go;
        }
        if (c_out) c_out[row * H + j] = c_state[r];
      }
    }
#pragma unroll
    for (int r = 0; r < RBF; ++r) h_s[((t + 1) & 1) * H * RBF + j * RBF + r] = hn[r];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RBF; ++r)
#pragma unroll
      for (int g = 0; g < 4; ++g) pcur[r][g] = pnext[r][g];
  }
#pragma unroll
  for (int r = 0; r < RBF; ++r)
    if (r0 + r < R) {
      if (hN) hN[st_set + (int64_t)(r0 + r) * H + j] = h_s[(T & 1) * H * RBF + j * RBF + r];
      if (cN) cN[st_set + (int64_t)(r0 + r) * H + j] = c_state[r];
    }
}

// Same recurrence with TWO threads per hidden unit (threads 2j, 2j+1): each contracts half of the k
// range for all RBF rows, the pair swaps partial sums with one shuffle per value, and each thread then
// finalises (gate math, state, stores) RBF/2 of the rows.  Halves the serial FMA chain per time step
// and doubles the warps per scheduler that hide the shared-memory latency - the step time is what
// bounds this kernel.  Weights in shared memory only (WMODE 0 fp32, 1 bf16).
template <int WMODE, int RBF>
__global__ void lstm_fwd_ks2_kernel(const float* __restrict__ pre, const float* __restrict__ whh_t,
                                    int T, int R, int Bp, int H, int64_t pre_pstride,
                                    int64_t pre_tstride, int64_t pre_ld, int64_t pre_set_stride,
                                    int64_t whh_set_stride, float* __restrict__ h_out,
                                    float* __restrict__ gates_out, float* __restrict__ c_out,
                                    const float* h0, const float* c0, float* hN, float* cN) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int RH = RBF / 2;
  const int G = 4 * H;
  const int j = threadIdx.x >> 1, half = threadIdx.x & 1;
  const int set = blockIdx.y;
  const int r0 = blockIdx.x * RBF + half * RH;      // first row this thread finalises
  const int64_t out_set_stride = (int64_t)T * R * H;
  pre += (int64_t)set * pre_set_stride;
  whh_t += (int64_t)set * whh_set_stride;
  h_out += (int64_t)set * out_set_stride;
  if (gates_out) gates_out += (int64_t)set * out_set_stride * 4;
  if (c_out) c_out += (int64_t)set * out_set_stride;

  float* h_s = reinterpret_cast<float*>(smem_raw);                   // [2][H][RBF]
  float4* w_f = reinterpret_cast<float4*>(h_s + 2 * H * 4);
  uint2* w_b = reinterpret_cast<uint2*>(h_s + 2 * H * 4);
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) {
    const int k = i / H, jj = i - k * H;
    const float* wp = whh_t + (int64_t)k * G + jj;
    if (WMODE == 0) {
      w_f[i] = make_float4(wp[0], wp[H], wp[2 * H], wp[3 * H]);
    } else {
      __nv_bfloat162 a = __floats2bfloat162_rn(wp[0], wp[H]), b = __floats2bfloat162_rn(wp[2 * H], wp[3 * H]);
      w_b[i] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    }
  }
  const int64_t st_set = (int64_t)set * R * H;
  float c_state[RH];
  int64_t prow[RH], orow[RH];
#pragma unroll
  for (int r = 0; r < RH; ++r) {
    const bool ok = r0 + r < R;
    h_s[j * RBF + half * RH + r] = (h0 && ok) ? h0[st_set + (int64_t)(r0 + r) * H + j] : 0.f;
    c_state[r] = (c0 && ok) ? c0[st_set + (int64_t)(r0 + r) * H + j] : 0.f;
    const int rr = min(r0 + r, R - 1);
    prow[r] = (int64_t)(rr / Bp) * pre_pstride + (int64_t)(rr % Bp) * pre_ld + j;
    orow[r] = (int64_t)(rr / Bp) * T * Bp + (rr % Bp);
  }
  float pcur[RH][4], pnext[RH][4];
#pragma unroll
  for (int r = 0; r < RH; ++r)
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      pcur[r][g] = (r0 + r < R) ? pre[prow[r] + g * H] : 0.f;
      pnext[r][g] = 0.f;
    }
  const int kbeg = half * (H >> 1), kend = kbeg + (H >> 1);
  __syncthreads();

  for (int t = 0; t < T; ++t) {
    if (t + 1 < T) {
#pragma unroll
      for (int r = 0; r < RH; ++r)
        if (r0 + r < R) {
#pragma unroll
          for (int g = 0; g < 4; ++g) pnext[r][g] = pre[prow[r] + (int64_t)(t + 1) * pre_tstride + g * H];
        }
    }
    float acc[RBF][4];
#pragma unroll
    for (int r = 0; r < RBF; ++r)
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[r][g] = 0.f;
    const float* hb = h_s + (t & 1) * H * RBF;
#pragma unroll 8
    for (int k = kbeg; k < kend; ++k) {
      float w[4];
      if (WMODE == 0) {
        const float4 wv = w_f[k * H + j];
        w[0] = wv.x; w[1] = wv.y; w[2] = wv.z; w[3] = wv.w;
      } else {
        const uint2 wv = w_b[k * H + j];
        w[0] = __uint_as_float(wv.x << 16);
        w[1] = __uint_as_float(wv.x & 0xffff0000u);
        w[2] = __uint_as_float(wv.y << 16);
        w[3] = __uint_as_float(wv.y & 0xffff0000u);
      }
      float hr[RBF];
      if (RBF == 4) {
        const float4 hv = *reinterpret_cast<const float4*>(hb + k * 4);
        hr[0] = hv.x; hr[1] = hv.y; hr[2 % RBF] = hv.z; hr[3 % RBF] = hv.w;
      } else {
        const float2 hv = *reinterpret_cast<const float2*>(hb + k * 2);
        hr[0] = hv.x; hr[1] = hv.y;
      }
#pragma unroll
      for (int r = 0; r < RBF; ++r)
#pragma unroll
        for (int g = 0; g < 4; ++g) acc[r][g] = fmaf(w[g], hr[r], acc[r][g]);
    }
    // the pair swaps the partial sums of the rows the OTHER thread finalises
    float tot[RH][4];
#pragma unroll
    for (int r = 0; r < RH; ++r)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float mine = half ? acc[RH + r][g] : acc[r][g];
        const float send = half ? acc[r][g] : acc[RH + r][g];
        tot[r][g] = pcur[r][g] + mine + __shfl_xor_sync(0xffffffffu, send, 1);
      }
    float hn[RH];
#pragma unroll
    for (int r = 0; r < RH; ++r) {
      const float gi = sigm(tot[r][0]), gf = sigm(tot[r][1]), gg = tanhf(tot[r][2]), go = sigm(tot[r][3]);
      c_state[r] = gf * c_state[r] + gi * gg;
      hn[r] = go * tanhf(c_state[r]);
      if (r0 + r < R) {
        const int64_t row = orow[r] + (int64_t)t * Bp;
        h_out[row * H + j] = hn[r];
        if (gates_out) {
          float* gp = gates_out + row * G;
          gp[j] = gi;
          gp[H + j] = gf;
          gp[2 * H + j] = gg;
          gp[3 * H + j] = go;
        }
        if (c_out) c_out[row * H + j] = c_state[r];
      }
    }
#pragma unroll
    for (int r = 0; r < RH; ++r) h_s[((t + 1) & 1) * H * RBF + j * RBF + half * RH + r] = hn[r];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RH; ++r)
#pragma unroll
      for (int g = 0; g < 4; ++g) pcur[r][g] = pnext[r][g];
  }
#pragma unroll
  for (int r = 0; r < RH; ++r)
    if (r0 + r < R) {
      if (hN) hN[st_set + (int64_t)(r0 + r) * H + j] = h_s[(T & 1) * H * RBF + j * RBF + half * RH + r];
      if (cN) cN[st_set + (int64_t)(r0 + r) * H + j] = c_state[r];
    }
}

// ------------------------------------------------------------------------------------------------------------
// Tensor-core recurrence (bf16 policy, H = 32 / 64 / 128).  gates^T [4H x rows] = W_hh [4H x H] . h^T [H x rows]
// per time step as mma.sync m16n8k16: eight batch rows are the N of the MMA, the gate rows the M.  W_hh does not
// change over the sequence, so every warp keeps its A fragments (bf16) in REGISTERS for all T steps - the step
// reads nothing but h from shared memory - and h is contracted as a bf16 hi + lo pair (16 mantissa bits), so the
// only rounding next to the CUDA-core kernel above is the bf16 W_hh that kernel (WMODE 1) uses as well.
// M tiles are laid out so that the four gates of a cell land in ONE thread: tile 0 of a unit octet holds gate i
// (rows 0-7) and gate f (rows 8-15) of units u0..u0+7, tile 1 gates g and o; thread (lane) then owns cells
// (unit u0 + lane/4, rows 2*(lane%4) + {0,1}) and the gate math needs no exchange.  One barrier per step.
// (tcgen05 would need a TMEM round trip per step; at M = 4H <= 512, N = 8 the step is latency, not throughput.)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int kLstmDepth = 4;                    // time steps of operands in flight (cp.async)
constexpr int kLstmStages = kLstmDepth + 1;
// MUFU.RCP / MUFU.EX2 (about 1 ulp each; __frcp_rn is a correctly rounded multi-instruction sequence that measured
// 0.37 us per cell - three quarters of the step)
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sigm_fast(float v) { return rcp_approx(1.f + __expf(-v)); }
__device__ __forceinline__ float tanh_fast(float v) { return 1.f - 2.f * rcp_approx(1.f + __expf(2.f * v)); }

template <int H, int NT>
__global__ void __launch_bounds__((H / 8 < 8 ? H / 8 : 8) * 32, 1)
lstm_fwd_mma_kernel(const float* __restrict__ pre, const float* __restrict__ whh_t, int T, int R, int Bp,
                    int64_t pre_pstride, int64_t pre_tstride, int64_t pre_ld, int64_t pre_set_stride,
                    int64_t whh_set_stride, float* __restrict__ h_out, float* __restrict__ gates_out,
                    float* __restrict__ c_out, const float* h0, const float* c0, float* hN, float* cN) {
  constexpr int NW = H / 8 < 8 ? H / 8 : 8;      // warps
  constexpr int OPW = (H / 8) / NW;              // unit octets per warp
  constexpr int KT = H / 16;                     // k tiles
  constexpr int G = 4 * H;
  constexpr int HP = H + 8;                      // padded row of the h buffers (bank-conflict free fragments)
  constexpr int NR = 8 * NT;                     // batch rows per CTA
  constexpr int PP = G + 4;                      // padded row of the pre-activation ring (floats)
  constexpr int NTHR = NW * 32;
  constexpr int CH = NR * H / NTHR;              // 16-byte chunks of one step's pre-activations per thread
  __shared__ __align__(16) __nv_bfloat16 hs[2][2][NR][HP];   // [buffer][hi|lo][row][k]
  extern __shared__ __align__(16) float pre_ring[];          // [kLstmStages][NR][PP]: steps t .. t+kLstmDepth
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lq = lane >> 2, lr = lane & 3;
  const int set = blockIdx.y;
  const int r0 = blockIdx.x * NR;
  const int64_t out_set_stride = (int64_t)T * R * H;
  pre += (int64_t)set * pre_set_stride;
  whh_t += (int64_t)set * whh_set_stride;
  h_out += (int64_t)set * out_set_stride;
  if (gates_out) gates_out += (int64_t)set * out_set_stride * 4;
  if (c_out) c_out += (int64_t)set * out_set_stride;
  const int64_t st_set = (int64_t)set * R * H;

  // A fragments: W_hh[gate*H + u][k] = whh_t[k*G + gate*H + u]
  uint32_t afr[OPW][2][KT][4];
#pragma unroll
  for (int o = 0; o < OPW; ++o) {
    const int u = (warp * OPW + o) * 8 + lq;
#pragma unroll
    for (int tt = 0; tt < 2; ++tt)
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        const int k = kt * 16 + 2 * lr;
        const float* wlo = whh_t + (2 * tt) * H + u;       // gate 2tt   (tile rows 0-7)
        const float* whi = whh_t + (2 * tt + 1) * H + u;   // gate 2tt+1 (tile rows 8-15)
        afr[o][tt][kt][0] = pack_bf16(wlo[(int64_t)k * G], wlo[(int64_t)(k + 1) * G]);
        afr[o][tt][kt][1] = pack_bf16(whi[(int64_t)k * G], whi[(int64_t)(k + 1) * G]);
        afr[o][tt][kt][2] = pack_bf16(wlo[(int64_t)(k + 8) * G], wlo[(int64_t)(k + 9) * G]);
        afr[o][tt][kt][3] = pack_bf16(whi[(int64_t)(k + 8) * G], whi[(int64_t)(k + 9) * G]);
      }
  }
  // initial state
  for (int i = threadIdx.x; i < NR * H; i += blockDim.x) {
    const int n = i / H, k = i - n * H;
    const float v = (h0 && r0 + n < R) ? h0[st_set + (int64_t)(r0 + n) * H + k] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    hs[0][0][n][k] = hi;
    hs[0][1][n][k] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
  // this thread's cells: (octet o, n tile nt, e): unit (warp*OPW+o)*8 + lq, row r0 + nt*8 + 2*lr + e
  float c_state[OPW][NT][2];
  int64_t orow[NT][2];
  bool rok[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int row = r0 + nt * 8 + 2 * lr + e;
      rok[nt][e] = row < R;
      const int rr = min(row, R - 1);
      orow[nt][e] = (int64_t)(rr / Bp) * T * Bp + (rr % Bp);
#pragma unroll
      for (int o = 0; o < OPW; ++o) {
        const int u = (warp * OPW + o) * 8 + lq;
        c_state[o][nt][e] = (c0 && rok[nt][e]) ? c0[st_set + (int64_t)rr * H + u] : 0.f;
      }
    }
  // The step is shorter than a DRAM round trip, so the pre-activations of the next kLstmDepth steps are kept in
  // flight with cp.async (rows >= R read row R-1; their cells are never stored).
  int64_t csrc[CH];
  uint32_t cdst[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int ci = threadIdx.x + i * NTHR;
    const int n = ci / H, c4 = ci - n * H;
    const int rr = min(r0 + n, R - 1);
    csrc[i] = (int64_t)(rr / Bp) * pre_pstride + (int64_t)(rr % Bp) * pre_ld + 4 * c4;
    cdst[i] = smem_u32(pre_ring) + (uint32_t)(n * PP + 4 * c4) * 4u;
  }
  auto prefetch = [&](int tt) {
    if (tt < T) {
      const uint32_t so = (uint32_t)(tt % kLstmStages) * (uint32_t)(NR * PP * 4);
#pragma unroll
      for (int i = 0; i < CH; ++i)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(cdst[i] + so),
                     "l"(pre + csrc[i] + (int64_t)tt * pre_tstride)
                     : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int d = 0; d < kLstmDepth; ++d) prefetch(d);

  for (int t = 0; t < T; ++t) {
    const int cur = t & 1, nxt = cur ^ 1;
    // step t's group has landed (kLstmDepth - 1 younger groups may be pending); the barrier also publishes the h
    // written by the previous step and frees the ring stage of step t - 1 for step t + kLstmDepth
    asm volatile("cp.async.wait_group %0;" ::"n"(kLstmDepth - 1) : "memory");
    __syncthreads();
    prefetch(t + kLstmDepth);
    const float* ps = pre_ring + (size_t)(t % kLstmStages) * (NR * PP);
    // accumulators start from the pre-activations: tile 0 = {i(e0), i(e1), f(e0), f(e1)}, tile 1 = {g, g, o, o}
    float acc[OPW][NT][2][4];
#pragma unroll
    for (int o = 0; o < OPW; ++o)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          const float* pr = ps + (nt * 8 + 2 * lr) * PP + (warp * OPW + o) * 8 + lq;
          acc[o][nt][tt][0] = pr[(2 * tt) * H];
          acc[o][nt][tt][1] = pr[PP + (2 * tt) * H];
          acc[o][nt][tt][2] = pr[(2 * tt + 1) * H];
          acc[o][nt][tt][3] = pr[PP + (2 * tt + 1) * H];
        }
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int n = nt * 8 + lq, k = kt * 16 + 2 * lr;
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(&hs[cur][0][n][k]);
        const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(&hs[cur][0][n][k + 8]);
        const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(&hs[cur][1][n][k]);
        const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(&hs[cur][1][n][k + 8]);
#if !defined(LSTM_PROBE) || LSTM_PROBE != 3
#pragma unroll
        for (int o = 0; o < OPW; ++o)
#pragma unroll
          for (int tt = 0; tt < 2; ++tt) mma_bf16_16816(acc[o][nt][tt], afr[o][tt][kt], bh0, bh1);
#pragma unroll
        for (int o = 0; o < OPW; ++o)
#pragma unroll
          for (int tt = 0; tt < 2; ++tt) mma_bf16_16816(acc[o][nt][tt], afr[o][tt][kt], bl0, bl1);
#else
        acc[0][nt][0][0] += __uint_as_float(bh0 ^ bh1 ^ bl0 ^ bl1 ^ afr[0][0][kt][0]) * 1e-30f;
#endif
      }
    }
#pragma unroll
    for (int o = 0; o < OPW; ++o) {
      const int u = (warp * OPW + o) * 8 + lq;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
#if defined(LSTM_PROBE) && LSTM_PROBE == 2
          const float gi = acc[o][nt][0][e], gf = acc[o][nt][0][2 + e], gg = acc[o][nt][1][e], go = acc[o][nt][1][2 + e];
          const float cs = 0.5f * c_state[o][nt][e] + 0.01f * gi * gg + 0.01f * gf;
          c_state[o][nt][e] = cs;
          const float hn = 0.01f * go + cs;
#else
          const float gi = sigm_fast(acc[o][nt][0][e]), gf = sigm_fast(acc[o][nt][0][2 + e]);
          const float gg = tanh_fast(acc[o][nt][1][e]), go = sigm_fast(acc[o][nt][1][2 + e]);
          const float cs = gf * c_state[o][nt][e] + gi * gg;
          c_state[o][nt][e] = cs;
          const float hn = go * tanh_fast(cs);
#endif
          const __nv_bfloat16 hi = __float2bfloat16_rn(hn);
          const int n = nt * 8 + 2 * lr + e;
          hs[nxt][0][n][u] = hi;
          hs[nxt][1][n][u] = __float2bfloat16_rn(hn - __bfloat162float(hi));
#if defined(LSTM_PROBE) && LSTM_PROBE == 1
          if (rok[nt][e] && t == T - 1) {
#else
          if (rok[nt][e]) {
#endif
            const int64_t row = orow[nt][e] + (int64_t)t * Bp;
            h_out[row * H + u] = hn;
            if (gates_out) {
              float* gp = gates_out + row * G + u;
              gp[0] = gi;
              gp[H] = gf;
              gp[2 * H] = gg;
              gp[3 * H] = go;
            }
            if (c_out) c_out[row * H + u] = cs;
            if (t == T - 1) {
              const int64_t srow = st_set + (int64_t)(r0 + n) * H + u;
              if (hN) hN[srow] = hn;
              if (cN) cN[srow] = cs;
            }
          }
        }
    }
  }
}

// BPTT.  Thread tid = (q, k): phase A treats it as cell (row q, unit k); phase B as the partial
// dot product over gate quarter q for hidden unit k.  W_hh [4H][H] in smem (WMODE 0) or global (2).
template <int WMODE>
__global__ void lstm_bwd_kernel(const float* __restrict__ dh_out, const float* __restrict__ whh,
                                const float* __restrict__ gates, const float* __restrict__ c_all,
                                int T, int R, int Bp, int H, int64_t whh_set_stride,
                                int64_t pre_pstride, int64_t pre_tstride, int64_t pre_ld,
                                int64_t pre_set_stride, float* __restrict__ dpre) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int G = 4 * H;
  const int tid = threadIdx.x;
  const int set = blockIdx.y;
  const int r0 = blockIdx.x * RB;
  const int64_t set_stride_h = (int64_t)T * R * H;
  dh_out += (int64_t)set * set_stride_h;
  c_all += (int64_t)set * set_stride_h;
  gates += (int64_t)set * set_stride_h * 4;
  dpre += (int64_t)set * pre_set_stride;
  whh += (int64_t)set * whh_set_stride;

  float* dp_s = reinterpret_cast<float*>(smem_raw);  // [RB][4H]
  float* part_s = dp_s + RB * G;                     // [4][RB][H]
  float* w_s = part_s + 4 * RB * H;                  // [4H][H] (WMODE 0)
  if (WMODE == 0)
    for (int i = tid; i < G * H; i += blockDim.x) w_s[i] = whh[i];
  for (int i = tid; i < 4 * RB * H; i += blockDim.x) part_s[i] = 0.f;
  __syncthreads();

  const int q = tid / H, k = tid - q * H;  // q in 0..3 (== RB rows)
  const bool row_ok = (r0 + q) < R;
  const int rq = min(r0 + q, R - 1);
  const int64_t orow_base = (int64_t)(rq / Bp) * T * Bp + (rq % Bp);
  const int64_t prow_base = (int64_t)(rq / Bp) * pre_pstride + (int64_t)(rq % Bp) * pre_ld;
  float dc_next = 0.f;

  for (int t = T - 1; t >= 0; --t) {
    // ---- phase A: cell (row q, unit k)
    float di = 0.f, df = 0.f, dg = 0.f, dout = 0.f;
    if (row_ok) {
      int64_t row = orow_base + (int64_t)t * Bp;
      float dh = dh_out[row * H + k] + part_s[(0 * RB + q) * H + k] + part_s[(1 * RB + q) * H + k] +
                 part_s[(2 * RB + q) * H + k] + part_s[(3 * RB + q) * H + k];
      const float* gp = gates + row * G;
      float gi = gp[k], gf = gp[H + k], gg = gp[2 * H + k], go = gp[3 * H + k];
      float ct = c_all[row * H + k];
      float cprev = t > 0 ? c_all[(row - Bp) * H + k] : 0.f;
      float tc = tanhf(ct);
      dout = dh * tc * go * (1.f - go);
      float dc = dh * go * (1.f - tc * tc) + dc_next;
      di = dc * gg * gi * (1.f - gi);
      df = dc * cprev * gf * (1.f - gf);
      dg = dc * gi * (1.f - gg * gg);
      dc_next = dc * gf;
      float* dp = dpre + prow_base + (int64_t)t * pre_tstride;
      dp[k] = di;
      dp[H + k] = df;
      dp[2 * H + k] = dg;
      dp[3 * H + k] = dout;
    }
    __syncthreads();  // everyone has consumed part_s
    dp_s[q * G + k] = di;
    dp_s[q * G + H + k] = df;
    dp_s[q * G + 2 * H + k] = dg;
    dp_s[q * G + 3 * H + k] = dout;
    __syncthreads();
    // ---- phase B: partial dh_rec[r][k] over gate quarter q
    float acc[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) acc[r] = 0.f;
    const int gbeg = q * H;
    for (int g = 0; g < H; g += 4) {
      float w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (WMODE == 0) w[e] = w_s[(gbeg + g + e) * H + k];
        else w[e] = __ldg(whh + (int64_t)(gbeg + g + e) * H + k);
      }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        float4 dv = *reinterpret_cast<const float4*>(dp_s + r * G + gbeg + g);
        acc[r] = fmaf(w[0], dv.x, acc[r]);
        acc[r] = fmaf(w[1], dv.y, acc[r]);
        acc[r] = fmaf(w[2], dv.z, acc[r]);
        acc[r] = fmaf(w[3], dv.w, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) part_s[(q * RB + r) * H + k] = acc[r];
    __syncthreads();
  }
}

// BPTT on the same tensor-core scheme: dh_rec^T [H x rows] = W_hh^T [H x 4H] . dgates^T [4H x rows].  Warp w owns
// hidden units 16w..16w+15 (one M tile, its W_hh^T fragments in registers for the whole sequence), the eight batch
// rows are N, and the step's gate gradients - written by the threads that own the cells - are the B operand as a bf16
// hi + lo pair in shared memory (double buffered: one barrier per step).  Thread (lane) owns cells (unit 16w + lane/4
// [+8], rows 2*(lane%4) + {0,1}), exactly the elements of its accumulator fragment.
template <int H, int NT>
__global__ void __launch_bounds__((H / 16) * 32, 1)
lstm_bwd_mma_kernel(const float* __restrict__ dh_out, const float* __restrict__ whh, const float* __restrict__ gates,
                    const float* __restrict__ c_all, int T, int R, int Bp, int64_t whh_set_stride,
                    int64_t pre_pstride, int64_t pre_tstride, int64_t pre_ld, int64_t pre_set_stride,
                    float* __restrict__ dpre) {
  constexpr int G = 4 * H;
  constexpr int KT = G / 16;
  constexpr int GP = G + 8;
  constexpr int NR = 8 * NT;
  constexpr int NTHR = (H / 16) * 32;
  constexpr int PG = G + 4, PH = H + 4;                       // padded rows of the operand ring (floats)
  constexpr int STG = NR * (PG + 2 * PH);                     // one stage: gates [NR][PG], dh [NR][PH], c [NR][PH]
  constexpr int CH = NR * (H + H / 2) / NTHR;                 // 16-byte chunks of one step's operands per thread
  __shared__ __align__(16) __nv_bfloat16 dg[2][2][NR][GP];   // [buffer][hi|lo][row][gate row]
  extern __shared__ __align__(16) float op_ring[];           // [kLstmStages][STG]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lq = lane >> 2, lr = lane & 3;
  const int set = blockIdx.y;
  const int r0 = blockIdx.x * NR;
  const int64_t set_stride_h = (int64_t)T * R * H;
  dh_out += (int64_t)set * set_stride_h;
  c_all += (int64_t)set * set_stride_h;
  gates += (int64_t)set * set_stride_h * 4;
  dpre += (int64_t)set * pre_set_stride;
  whh += (int64_t)set * whh_set_stride;

  // A fragments: A[k][m] = W_hh[m][k] = whh[m*H + k], rows k = 16*warp + lq (+8), columns m
  uint32_t afr[KT][4];
  {
    const int k = warp * 16 + lq;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      const int m = kt * 16 + 2 * lr;
      afr[kt][0] = pack_bf16(whh[(int64_t)m * H + k], whh[(int64_t)(m + 1) * H + k]);
      afr[kt][1] = pack_bf16(whh[(int64_t)m * H + k + 8], whh[(int64_t)(m + 1) * H + k + 8]);
      afr[kt][2] = pack_bf16(whh[(int64_t)(m + 8) * H + k], whh[(int64_t)(m + 9) * H + k]);
      afr[kt][3] = pack_bf16(whh[(int64_t)(m + 8) * H + k + 8], whh[(int64_t)(m + 9) * H + k + 8]);
    }
  }
  for (int i = threadIdx.x; i < 2 * NR * GP; i += blockDim.x) {
    (&dg[0][0][0][0])[i] = __float2bfloat16_rn(0.f);          // dh_rec of the last step is zero
  }
  // cells (j, nt, e): unit 16*warp + lq + 8j, row r0 + 8nt + 2lr + e  <->  accumulator element [nt][2j + e]
  int64_t prow[NT][2], orow[NT][2];
  bool rok[NT][2];
  float dc_next[2][NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int row = r0 + nt * 8 + 2 * lr + e;
      rok[nt][e] = row < R;
      const int rr = min(row, R - 1);
      prow[nt][e] = (int64_t)(rr / Bp) * pre_pstride + (int64_t)(rr % Bp) * pre_ld;
      orow[nt][e] = (int64_t)(rr / Bp) * T * Bp + (rr % Bp);
      dc_next[0][nt][e] = 0.f;
      dc_next[1][nt][e] = 0.f;
    }
  // operands of the next kLstmDepth steps in flight (cp.async): the step is shorter than a DRAM round trip
  const float* csrc[CH];
  int64_t cts[CH];
  uint32_t cdst[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int ci = threadIdx.x + i * NTHR;
    int n, c4, kind;                                          // kind 0 gates, 1 dh, 2 c
    if (ci < NR * H) { kind = 0; n = ci / H; c4 = ci - n * H; }
    else if (ci < NR * H + NR * (H / 4)) { kind = 1; n = (ci - NR * H) / (H / 4); c4 = (ci - NR * H) - n * (H / 4); }
    else { kind = 2; n = (ci - NR * H - NR * (H / 4)) / (H / 4); c4 = (ci - NR * H - NR * (H / 4)) - n * (H / 4); }
    const int rr = min(r0 + n, R - 1);
    const int64_t row0 = (int64_t)(rr / Bp) * T * Bp + (rr % Bp);
    if (kind == 0) {
      csrc[i] = gates + row0 * G + 4 * c4; cts[i] = (int64_t)Bp * G;
      cdst[i] = smem_u32(op_ring) + (uint32_t)(n * PG + 4 * c4) * 4u;
    } else {
      csrc[i] = (kind == 1 ? dh_out : c_all) + row0 * H + 4 * c4; cts[i] = (int64_t)Bp * H;
      cdst[i] = smem_u32(op_ring) + (uint32_t)(NR * PG + (kind - 1) * NR * PH + n * PH + 4 * c4) * 4u;
    }
  }
  auto prefetch = [&](int sidx) {                             // step sidx handles time T - 1 - sidx
    const int tt = T - 1 - sidx;
    if (tt >= 0) {
      const uint32_t so = (uint32_t)(sidx % kLstmStages) * (uint32_t)(STG * 4);
#pragma unroll
      for (int i = 0; i < CH; ++i)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(cdst[i] + so), "l"(csrc[i] + (int64_t)tt * cts[i])
                     : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int d = 0; d < kLstmDepth; ++d) prefetch(d);

  for (int t = T - 1; t >= 0; --t) {
    const int sidx = T - 1 - t;
    const int cur = sidx & 1, nxt = cur ^ 1;
    // steps sidx and sidx + 1 (c_{t-1}) have landed; the barrier also publishes the previous step's gate gradients
    asm volatile("cp.async.wait_group %0;" ::"n"(kLstmDepth - 2) : "memory");
    __syncthreads();
    prefetch(sidx + kLstmDepth);
    const float* sg = op_ring + (size_t)(sidx % kLstmStages) * STG;
    const float* sc_prev = op_ring + (size_t)((sidx + 1) % kLstmStages) * STG + NR * PG + NR * PH;
    // dh_rec from the previous (t+1) step's gate gradients: four interleaved accumulators shorten the MMA chain
    float acc[NT][4][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[nt][a][q] = 0.f;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int n = nt * 8 + lq, m = kt * 16 + 2 * lr;
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(&dg[cur][0][n][m]);
        const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(&dg[cur][0][n][m + 8]);
        const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(&dg[cur][1][n][m]);
        const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(&dg[cur][1][n][m + 8]);
        mma_bf16_16816(acc[nt][(2 * kt) & 3], afr[kt], bh0, bh1);
        mma_bf16_16816(acc[nt][(2 * kt + 1) & 3], afr[kt], bl0, bl1);
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int u = warp * 16 + lq + 8 * j;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int q = 2 * j + e;
          const int nn = nt * 8 + 2 * lr + e;
          const float* gp = sg + nn * PG + u;
          const float dh = sg[NR * PG + nn * PH + u] + (acc[nt][0][q] + acc[nt][1][q]) + (acc[nt][2][q] + acc[nt][3][q]);
          const float gi = gp[0], gf = gp[H], gg = gp[2 * H], go = gp[3 * H];
          const float ct = sg[NR * PG + NR * PH + nn * PH + u];
          const float cprev = t > 0 ? sc_prev[nn * PH + u] : 0.f;
          const float tc = tanh_fast(ct);
          float dout = dh * tc * go * (1.f - go);
          const float dc = dh * go * (1.f - tc * tc) + dc_next[j][nt][e];
          float di = dc * gg * gi * (1.f - gi);
          float df = dc * cprev * gf * (1.f - gf);
          float dgg = dc * gi * (1.f - gg * gg);
          dc_next[j][nt][e] = dc * gf;
          if (!rok[nt][e]) { di = 0.f; df = 0.f; dgg = 0.f; dout = 0.f; }
          const int n = nt * 8 + 2 * lr + e;
          const float v[4] = {di, df, dgg, dout};
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(v[g]);
            dg[nxt][0][n][g * H + u] = hi;
            dg[nxt][1][n][g * H + u] = __float2bfloat16_rn(v[g] - __bfloat162float(hi));
          }
          if (rok[nt][e]) {
            float* dp = dpre + prow[nt][e] + (int64_t)t * pre_tstride + u;
            dp[0] = di;
            dp[H] = df;
            dp[2 * H] = dgg;
            dp[3 * H] = dout;
          }
        }
    }
  }
}

constexpr size_t kSmemLimit = 200 * 1024;

}  // namespace
}  // namespace clskd

using namespace clskd;
#define ST ((cudaStream_t)stream)

extern "C" int clskd_lstm_fwd(const float* pre, const float* whh_t, int T, int R, int Bp, int H,
                              int nsets, int64_t pre_pstride, int64_t pre_tstride, int64_t pre_ld,
                              int64_t pre_set_stride, int64_t whh_set_stride, int w_bf16, float* h,
                              float* gates, float* c, void* stream) {
  return clskd_lstm_fwd_state(pre, whh_t, T, R, Bp, H, nsets, pre_pstride, pre_tstride, pre_ld, pre_set_stride,
                              whh_set_stride, w_bf16, h, gates, c, nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int clskd_lstm_fwd_state(const float* pre, const float* whh_t, int T, int R, int Bp, int H,
                                    int nsets, int64_t pre_pstride, int64_t pre_tstride, int64_t pre_ld,
                                    int64_t pre_set_stride, int64_t whh_set_stride, int w_bf16, float* h,
                                    float* gates, float* c, const float* h0, const float* c0, float* hN,
                                    float* cN, void* stream) {
  CLSKD_CHECK_ARG(pre && whh_t && h, "clskd_lstm_fwd: null pointer");
  CLSKD_CHECK_ARG(H >= 4 && H % 4 == 0 && H <= 1024, "clskd_lstm_fwd: H=%d unsupported (4..1024, multiple of 4)", H);
  CLSKD_CHECK_ARG(T >= 0 && R >= 1 && nsets >= 1 && Bp >= 1 && R % Bp == 0, "clskd_lstm_fwd: bad extents");
  if (T == 0) return CLSKD_OK;
  const bool pre_vec = (uintptr_t)pre % 16 == 0 && pre_pstride % 4 == 0 && pre_tstride % 4 == 0 && pre_ld % 4 == 0 &&
                       pre_set_stride % 4 == 0;                      // 16-byte cp.async of the pre-activation rows
  if (w_bf16 && (H == 32 || H == 64 || H == 128) && pre_vec && !g_lstm_legacy) {
    // tensor-core recurrence: 8 batch rows per CTA, 16 once 8 would not give every CTA its own SM
    const bool nt2 = (int64_t)cdiv(R, 8) * nsets > (int64_t)sm_count();
    dim3 grid(cdiv(R, nt2 ? 16 : 8), nsets);
    const size_t ring = (size_t)kLstmStages * (nt2 ? 16 : 8) * (4 * H + 4) * sizeof(float);
#define LSTM_MMA_ARGS pre, whh_t, T, R, Bp, pre_pstride, pre_tstride, pre_ld, pre_set_stride, whh_set_stride, h, gates, c, h0, c0, hN, cN
#define LSTM_MMA_LAUNCH(HH, NTT, THR)                                                                               \
  do {                                                                                                              \
    cudaError_t e_ = cudaFuncSetAttribute(lstm_fwd_mma_kernel<HH, NTT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)ring);                                                               \
    if (e_ != cudaSuccess) { set_error("clskd_lstm_fwd: smem attr: %s", cudaGetErrorString(e_)); return CLSKD_ERR_CUDA; } \
    lstm_fwd_mma_kernel<HH, NTT><<<grid, THR, ring, ST>>>(LSTM_MMA_ARGS);                                            \
  } while (0)
    if (H == 128) {
      if (nt2) LSTM_MMA_LAUNCH(128, 2, 256);
      else LSTM_MMA_LAUNCH(128, 1, 256);
    } else if (H == 64) {
      if (nt2) LSTM_MMA_LAUNCH(64, 2, 256);
      else LSTM_MMA_LAUNCH(64, 1, 256);
    } else {
      if (nt2) LSTM_MMA_LAUNCH(32, 2, 128);
      else LSTM_MMA_LAUNCH(32, 1, 128);
    }
#undef LSTM_MMA_LAUNCH
#undef LSTM_MMA_ARGS
    CLSKD_CHECK_LAUNCH("clskd_lstm_fwd");
    return CLSKD_OK;
  }
  size_t base = sizeof(float4) * 2 * (size_t)H;
  size_t w32 = sizeof(float4) * (size_t)H * H, w16 = w32 / 2;
  // 2 rows per CTA while that still leaves every CTA its own SM (the recurrence is FMA-issue bound
  // per SM), 4 rows per CTA for large batches
  const bool rb2 = (int64_t)cdiv(R, 2) * nsets <= (int64_t)sm_count();
  dim3 grid(cdiv(R, rb2 ? 2 : 4), nsets);
  const bool ks2 = 2 * H <= 1024;      // two threads per hidden unit (k range split in halves)
  cudaError_t e;
#define LSTM_FWD_ARGS pre, whh_t, T, R, Bp, H, pre_pstride, pre_tstride, pre_ld, pre_set_stride, whh_set_stride, h, gates, c, h0, c0, hN, cN
  if (!w_bf16 && base + w32 <= kSmemLimit) {
    e = cudaFuncSetAttribute(lstm_fwd_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(base + w32));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_fwd_kernel<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(base + w32));
    if (e != cudaSuccess) { set_error("clskd_lstm_fwd: smem attr: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
    if (ks2) {
      e = cudaFuncSetAttribute(lstm_fwd_ks2_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(base + w32));
      if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_fwd_ks2_kernel<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(base + w32));
      if (e != cudaSuccess) { set_error("clskd_lstm_fwd: smem attr: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
      if (rb2) lstm_fwd_ks2_kernel<0, 2><<<grid, 2 * H, base + w32, ST>>>(LSTM_FWD_ARGS);
      else lstm_fwd_ks2_kernel<0, 4><<<grid, 2 * H, base + w32, ST>>>(LSTM_FWD_ARGS);
    } else if (rb2) lstm_fwd_kernel<0, 2><<<grid, H, base + w32, ST>>>(LSTM_FWD_ARGS);
    else lstm_fwd_kernel<0, 4><<<grid, H, base + w32, ST>>>(LSTM_FWD_ARGS);
  } else if (w_bf16 && base + w16 <= kSmemLimit) {
    e = cudaFuncSetAttribute(lstm_fwd_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(base + w16));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_fwd_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(base + w16));
    if (e != cudaSuccess) { set_error("clskd_lstm_fwd: smem attr: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
    if (ks2) {
      e = cudaFuncSetAttribute(lstm_fwd_ks2_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(base + w16));
      if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_fwd_ks2_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(base + w16));
      if (e != cudaSuccess) { set_error("clskd_lstm_fwd: smem attr: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
      if (rb2) lstm_fwd_ks2_kernel<1, 2><<<grid, 2 * H, base + w16, ST>>>(LSTM_FWD_ARGS);
      else lstm_fwd_ks2_kernel<1, 4><<<grid, 2 * H, base + w16, ST>>>(LSTM_FWD_ARGS);
    } else if (rb2) lstm_fwd_kernel<1, 2><<<grid, H, base + w16, ST>>>(LSTM_FWD_ARGS);
    else lstm_fwd_kernel<1, 4><<<grid, H, base + w16, ST>>>(LSTM_FWD_ARGS);
  } else {
    if (rb2) lstm_fwd_kernel<2, 2><<<grid, H, base, ST>>>(LSTM_FWD_ARGS);
    else lstm_fwd_kernel<2, 4><<<grid, H, base, ST>>>(LSTM_FWD_ARGS);
  }
#undef LSTM_FWD_ARGS
  CLSKD_CHECK_LAUNCH("clskd_lstm_fwd");
  return CLSKD_OK;
}

extern "C" int clskd_lstm_bwd(const float* dh_out, const float* whh, const float* gates,
                              const float* c, int T, int R, int Bp, int H, int nsets,
                              int64_t whh_set_stride, int64_t pre_pstride, int64_t pre_tstride,
                              int64_t pre_ld, int64_t pre_set_stride, float* dpre, void* stream) {
  return clskd_lstm_bwd_policy(dh_out, whh, gates, c, T, R, Bp, H, nsets, whh_set_stride, pre_pstride, pre_tstride,
                               pre_ld, pre_set_stride, dpre, 0, stream);
}

extern "C" int clskd_lstm_bwd_policy(const float* dh_out, const float* whh, const float* gates,
                                     const float* c, int T, int R, int Bp, int H, int nsets,
                                     int64_t whh_set_stride, int64_t pre_pstride, int64_t pre_tstride,
                                     int64_t pre_ld, int64_t pre_set_stride, float* dpre, int w_bf16, void* stream) {
  CLSKD_CHECK_ARG(dh_out && whh && gates && c && dpre, "clskd_lstm_bwd: null pointer");
  CLSKD_CHECK_ARG(H >= 4 && H % 4 == 0 && 4 * H <= 1024, "clskd_lstm_bwd: H=%d unsupported", H);
  CLSKD_CHECK_ARG(R >= 1 && Bp >= 1 && R % Bp == 0, "clskd_lstm_bwd: bad extents");
  if (T == 0) return CLSKD_OK;
  const bool vec = (uintptr_t)dh_out % 16 == 0 && (uintptr_t)gates % 16 == 0 && (uintptr_t)c % 16 == 0;
  if (w_bf16 && (H == 32 || H == 64 || H == 128) && vec && !g_lstm_legacy) {
    // 8 rows per CTA; 16 once 8 would not give every CTA its own SM (not at H = 128: the 16-row variant spills and
    // needs > 48 KB of static shared memory)
    const bool nt2 = H != 128 && (int64_t)cdiv(R, 8) * nsets > (int64_t)sm_count();
    dim3 grid(cdiv(R, nt2 ? 16 : 8), nsets);
    const size_t ring = (size_t)kLstmStages * (nt2 ? 16 : 8) * (4 * H + 4 + 2 * (H + 4)) * sizeof(float);
#define LSTM_MMA_ARGS dh_out, whh, gates, c, T, R, Bp, whh_set_stride, pre_pstride, pre_tstride, pre_ld, pre_set_stride, dpre
#define LSTM_MMA_LAUNCH(HH, NTT, THR)                                                                               \
  do {                                                                                                              \
    cudaError_t e_ = cudaFuncSetAttribute(lstm_bwd_mma_kernel<HH, NTT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)ring);                                                               \
    if (e_ != cudaSuccess) { set_error("clskd_lstm_bwd: smem attr: %s", cudaGetErrorString(e_)); return CLSKD_ERR_CUDA; } \
    lstm_bwd_mma_kernel<HH, NTT><<<grid, THR, ring, ST>>>(LSTM_MMA_ARGS);                                            \
  } while (0)
    if (H == 128) {
      LSTM_MMA_LAUNCH(128, 1, 256);
    } else if (H == 64) {
      if (nt2) LSTM_MMA_LAUNCH(64, 2, 128);
      else LSTM_MMA_LAUNCH(64, 1, 128);
    } else {
      if (nt2) LSTM_MMA_LAUNCH(32, 2, 64);
      else LSTM_MMA_LAUNCH(32, 1, 64);
    }
#undef LSTM_MMA_LAUNCH
#undef LSTM_MMA_ARGS
    CLSKD_CHECK_LAUNCH("clskd_lstm_bwd");
    return CLSKD_OK;
  }
  const int G = 4 * H;
  size_t base = sizeof(float) * ((size_t)RB * G + 4 * (size_t)RB * H);
  size_t w32 = sizeof(float) * (size_t)H * G;
  dim3 grid(cdiv(R, RB), nsets);
#define LSTM_BWD_ARGS dh_out, whh, gates, c, T, R, Bp, H, whh_set_stride, pre_pstride, pre_tstride, pre_ld, pre_set_stride, dpre
  if (base + w32 <= kSmemLimit) {
    cudaError_t e = cudaFuncSetAttribute(lstm_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(base + w32));
    if (e != cudaSuccess) { set_error("clskd_lstm_bwd: smem attr: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
    lstm_bwd_kernel<0><<<grid, G, base + w32, ST>>>(LSTM_BWD_ARGS);
  } else {
    lstm_bwd_kernel<2><<<grid, G, base, ST>>>(LSTM_BWD_ARGS);
  }
#undef LSTM_BWD_ARGS
  CLSKD_CHECK_LAUNCH("clskd_lstm_bwd");
  return CLSKD_OK;
}
