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
  CLSKD_CHECK_ARG(dh_out && whh && gates && c && dpre, "clskd_lstm_bwd: null pointer");
  CLSKD_CHECK_ARG(H >= 4 && H % 4 == 0 && 4 * H <= 1024, "clskd_lstm_bwd: H=%d unsupported", H);
  CLSKD_CHECK_ARG(R >= 1 && Bp >= 1 && R % Bp == 0, "clskd_lstm_bwd: bad extents");
  if (T == 0) return CLSKD_OK;
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
