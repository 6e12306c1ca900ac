// Shared helpers for the clskd sm_100a kernels: error reporting for the C ABI,
// dtype tags, vector load/store and block reductions.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/clskd.h"

namespace clskd {

// thread-local last-error string, read through clskd_last_error()
void set_error(const char* fmt, ...);

#define CLSKD_CHECK_ARG(cond, ...)                 \
  do {                                             \
    if (!(cond)) {                                 \
      ::clskd::set_error(__VA_ARGS__);             \
      return CLSKD_ERR_ARG;                        \
    }                                              \
  } while (0)

#define CLSKD_CHECK_LAUNCH(name)                                              \
  do {                                                                        \
    cudaError_t e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess) {                                                 \
      ::clskd::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return CLSKD_ERR_CUDA;                                                  \
    }                                                                         \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Zero up to four output spans before a reduction kernel; spans that are adjacent in memory (the Python side carves its
// sums out of one buffer) become ONE memset node - every extra node is 3-4 us of launch-bound stream time in front of
// kernels that take 40-100 us (about 100 such calls per step).
static inline cudaError_t zero_spans(cudaStream_t st, void* p0, size_t n0, void* p1 = nullptr, size_t n1 = 0,
                                     void* p2 = nullptr, size_t n2 = 0, void* p3 = nullptr, size_t n3 = 0) {
  void* ps[4] = {p0, p1, p2, p3};
  size_t ns[4] = {n0, n1, n2, n3};
  cudaError_t e = cudaSuccess;
  int i = 0;
  while (i < 4 && e == cudaSuccess) {
    if (!ps[i] || !ns[i]) { ++i; continue; }
    uint8_t* beg = static_cast<uint8_t*>(ps[i]);
    size_t len = ns[i];
    int j = i + 1;
    while (j < 4 && ps[j] && ns[j] && static_cast<uint8_t*>(ps[j]) == beg + len) { len += ns[j]; ++j; }
    e = cudaMemsetAsync(beg, 0, len, st);
    i = j;
  }
  return e;
}

int sm_count();

// ---------------------------------------------------------------- dtype access
template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void st_f(T* p, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

// load 4 consecutive elements (caller guarantees alignment: 16 B for float, 8 B for bf16)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; result valid in thread 0. `sh` must hold >= 32 elements.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* sh) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : T(0);
  if (w == 0) v = warp_sum(v);
  return v;
}

// 128-bit vectorised fast paths (elementwise_vec.cu); each returns false when the shape / alignment
// does not qualify and the caller falls back to the scalar kernel.
namespace vec {
bool colstats(const void* x, const void* dy, int dtype, int mode, int64_t M, int C, const float* mean,
              const float* invstd, const float* gamma, const float* beta, const float* slope, double* o0,
              double* o1, double* o2, cudaStream_t st);
bool bn_act_fwd(const void* x, int x_dtype, int64_t M, int C, const float* mean, const float* invstd,
                const float* gamma, const float* beta, const float* slope, void* y, int y_dtype, cudaStream_t st);
bool bn_act_bwd_apply(const void* x, const void* dy, int dtype, int64_t M, int C, const float* mean,
                      const float* invstd, const float* gamma, const float* beta, const float* slope,
                      const double* sum_dz, const double* sum_dz_xhat, int training, void* dx, cudaStream_t st);
bool resize_f(const void* src, int dtype, int64_t BT, int Fi, int Fo, int C, void* dst, bool backward, cudaStream_t st);
bool att_blend_fwd(const void* x, const void* y, int dtype, const float* z, int64_t M, int C, void* out, cudaStream_t st);
bool att_blend_bwd(const void* x, const void* y, int dtype, const float* z, const void* dout, int64_t M, int C,
                   void* dx, void* dy, float* dz, cudaStream_t st);
}  // namespace vec

// pointwise (1x1, dense) fast paths of the tap-list contraction (pointwise.cu); true = handled
namespace pw {
bool try_fwd(const ClskdTapConv* d, cudaStream_t st);
bool try_wgrad(const ClskdTapConv* d, cudaStream_t st);
}  // namespace pw

// first encoder layer (2-channel fp32 spectrogram -> bf16 maps) on mma.sync tensor cores (tapconv_c2_mma.cu); true = handled
namespace c2mma {
bool try_fwd(const ClskdTapConv* d, cudaStream_t st);
bool try_wgrad(const ClskdTapConv* d, cudaStream_t st);
}  // namespace c2mma

// dispatch on a runtime dtype tag
#define CLSKD_DISPATCH_DTYPE(tag, T, ...)                         \
  do {                                                            \
    if ((tag) == CLSKD_F32) { using T = float; __VA_ARGS__; }     \
    else { using T = __nv_bfloat16; __VA_ARGS__; }                \
  } while (0)

}  // namespace clskd
