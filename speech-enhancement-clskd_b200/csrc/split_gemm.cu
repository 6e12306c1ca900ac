// fp32 -> split-bf16 operand staging for "fp32-accurate" GEMMs on the tcgen05 tensor cores.
//
// x = hi + lo + r with hi = bf16(x), lo = bf16(x - hi), |r| <= 2^-17 |x|.  A GEMM  Y = X W  is run as ONE
// bf16 contraction over a 3x longer K axis:
//     [ X_hi | X_lo | X_hi ] (M x 3Kp)   times   [ W_hi ; W_hi ; W_lo ] (3Kp x N)
//   = X_hi W_hi + X_lo W_hi + X_hi W_lo   (the dropped lo*lo term is ~2^-18 relative)
// with fp32 accumulation in TMEM, i.e. the windowed-DFT / inverse-DFT GEMMs of the STFT front end
// (tools_for_model.py:57,100; framework.py:27) and the fp32 LSTM / Linear projections
// (tools_for_model.py:164-172) keep ~1e-5 relative accuracy while running on the tensor pipe.
//
// This file only stages the operand: rows addressed by three strides (so the overlapping STFT frames
// xpad[b, t*hop + k] are gathered straight from the padded waveform), K zero-padded to Kp, and rows
// >= M (up to Mp) zero-filled (weight staging pads N).
#include "common.cuh"

namespace clskd {
namespace {

// one thread = 8 consecutive k of one row: two or three 16-byte stores
template <int ORDER>
__global__ void __launch_bounds__(256)
split_bf16x3_kernel(const float* __restrict__ x, int64_t s0, int64_t s1, int64_t s2, int64_t sk,
                    int n1, int n2, int64_t M, int64_t Mp, int K, int Kp, int nseg,
                    __nv_bfloat16* __restrict__ out) {
  const int groups = Kp >> 3;
  const int64_t total = Mp * groups;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / groups;
    const int k0 = (int)(i - m * groups) << 3;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (m < M) {
      const int64_t i2 = m % n2, r = m / n2;
      const int64_t i1 = r % n1, i0 = r / n1;
      const float* p = x + i0 * s0 + i1 * s1 + i2 * s2 + (int64_t)k0 * sk;
      if (sk == 1 && k0 + 8 <= K && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (k0 + j < K) v[j] = p[(int64_t)j * sk];
      }
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
      __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * j] - __bfloat162float(h0));
      __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * j + 1] - __bfloat162float(h1));
      hi[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      lo[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    const uint4 H = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    const uint4 L = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    __nv_bfloat16* o = out + m * ((int64_t)nseg * Kp) + k0;
    // ORDER 0 (activations): [hi | lo | (hi)]     ORDER 1 (weights): [hi | hi | lo]
    *reinterpret_cast<uint4*>(o) = H;
    if (ORDER == 0) {
      *reinterpret_cast<uint4*>(o + Kp) = L;
      if (nseg == 3) *reinterpret_cast<uint4*>(o + 2 * Kp) = H;
    } else {
      *reinterpret_cast<uint4*>(o + Kp) = H;
      *reinterpret_cast<uint4*>(o + 2 * Kp) = L;
    }
  }
}

}  // namespace
}  // namespace clskd

using namespace clskd;

extern "C" int clskd_split_bf16x3(const float* x, int64_t s0, int64_t s1, int64_t s2, int64_t sk,
                                  int64_t n0, int n1, int n2, int K, int Kp, int64_t Mp, int order,
                                  int nseg, void* out, void* stream) {
  CLSKD_CHECK_ARG(x && out, "clskd_split_bf16x3: null pointer");
  CLSKD_CHECK_ARG(K >= 1 && Kp >= K && Kp % 8 == 0, "clskd_split_bf16x3: Kp must be a multiple of 8 and >= K");
  CLSKD_CHECK_ARG(n0 >= 0 && n1 >= 1 && n2 >= 1, "clskd_split_bf16x3: row counts");
  CLSKD_CHECK_ARG((order == 0 && (nseg == 2 || nseg == 3)) || (order == 1 && nseg == 3),
                  "clskd_split_bf16x3: order 0 takes 2 or 3 segments, order 1 takes 3");
  CLSKD_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "clskd_split_bf16x3: out must be 16-byte aligned");
  const int64_t M = n0 * n1 * n2;
  if (Mp < M) Mp = M;
  if (Mp == 0) return CLSKD_OK;
  const int64_t total = Mp * (Kp >> 3);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (order == 0)
    split_bf16x3_kernel<0><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        x, s0, s1, s2, sk, n1, n2, M, Mp, K, Kp, nseg, (__nv_bfloat16*)out);
  else
    split_bf16x3_kernel<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        x, s0, s1, s2, sk, n1, n2, M, Mp, K, Kp, nseg, (__nv_bfloat16*)out);
  CLSKD_CHECK_LAUNCH("clskd_split_bf16x3");
  return CLSKD_OK;
}
