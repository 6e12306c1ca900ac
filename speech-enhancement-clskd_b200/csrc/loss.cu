// Objectives of the distillation step: time-domain losses (SI-SNR / SDR / SI-SDR / MSE), the
// STFT-magnitude losses, and SPKD (Gram matrix + L1 row-normalised Frobenius distance).
// Reductions accumulate in fp64 so the loss values track the fp32 reference to ~1e-7 relative.
#include "common.cuh"

namespace clskd {
namespace {

// ------------------------------------------------------------------------------- wave losses
// pass 1 / pass 2 partial sums per utterance into part[b][0..3] (fp64 atomics)
template <int PASS>
__global__ void wave_loss_pass_kernel(const float* __restrict__ s1, const float* __restrict__ s2,
                                      int L, int kind, float eps, double* __restrict__ part) {
  __shared__ double sh[32];
  const int b = blockIdx.y;
  const float* p1 = s1 + (int64_t)b * L;
  const float* p2 = s2 + (int64_t)b * L;
  double* pp = part + (int64_t)b * 4;
  double alpha = 0.;
  if (PASS == 2) {
    if (kind == 0) alpha = pp[0] / (pp[1] + (double)eps);
    else alpha = pp[1] / pp[0] + (double)eps;  // kind 2
  }
  const float al = (float)alpha;
  double a0 = 0., a1 = 0.;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
    float u = p1[i], v = p2[i];
    if (PASS == 1) {
      if (kind == 0) { a0 += (double)(u * v); a1 += (double)(v * v); }
      else if (kind == 1) { float d = u - v; a0 += (double)(u * u); a1 += (double)(d * d); }
      else if (kind == 2) { a0 += (double)(u * u); a1 += (double)(u * v); }
      else { float d = u - v; a0 += (double)(d * d); }
    } else {
      if (kind == 0) { float tg = al * v, e = u - tg; a0 += (double)(tg * tg); a1 += (double)(e * e); }
      else { float pr = al * u, e = v - pr; a0 += (double)(pr * pr); a1 += (double)(e * e); }
    }
  }
  a0 = block_sum(a0, sh);
  a1 = block_sum(a1, sh);
  if (threadIdx.x == 0) {
    int o = PASS == 1 ? 0 : 2;
    atomicAdd(pp + o, a0);
    atomicAdd(pp + o + 1, a1);
  }
}

__global__ void wave_loss_final_kernel(const double* __restrict__ part, int B, int L, int kind,
                                       float eps, float* __restrict__ out) {
  __shared__ double sh[32];
  double acc = 0.;
  const double e = (double)eps;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const double* p = part + (int64_t)b * 4;
    if (kind == 0) acc += 10. * log10(p[2] / (p[3] + e) + e);
    else if (kind == 1) acc += 10. * log10(p[0] * p[0] / (p[1] * p[1] + e));
    else if (kind == 2) acc += p[2] / p[3] + e;
    else acc += p[0];
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    if (kind == 0 || kind == 1) out[0] = (float)(acc / B);
    else if (kind == 2) out[0] = (float)(10. * log10(acc / B + e));
    else out[0] = (float)(acc / ((double)B * L));
  }
}

__global__ void wave_loss_bwd_kernel(const float* __restrict__ s1, const float* __restrict__ s2,
                                     int B, int L, int kind, float eps,
                                     const double* __restrict__ part,
                                     const float* __restrict__ gout, float* __restrict__ ds1,
                                     float* __restrict__ ds2) {
  const int b = blockIdx.y;
  const float* p1 = s1 + (int64_t)b * L;
  const float* p2 = s2 + (int64_t)b * L;
  const double* p = part + (int64_t)b * 4;
  const double e = (double)eps;
  const double g = (double)gout[0];
  const double k10 = 10. / log(10.);
  // coefficients so that grad = cu*u + cv*v  (u = s1[i], v = s2[i])
  double cu1 = 0., cv1 = 0., cu2 = 0., cv2 = 0.;
  if (kind == 0) {
    // snr = k10*ln(R+e), R = tn/(nn+e); tn = alpha^2 b, nn = |s1 - alpha s2|^2, alpha = a/(b+e)
    double a = p[0], bb = p[1], tn = p[2], nn = p[3];
    double alpha = a / (bb + e);
    double R = tn / (nn + e);
    double c = g * k10 / (R + e) / B;
    double q = a - alpha * bb;  // <e_noise, s2>
    // d tn/d s1 = 2 alpha bb/(bb+e) * s2 ; d nn/d s1 = 2(s1 - alpha s2) - 2 q/(bb+e) * s2
    double ct = c / (nn + e), cn = -c * tn / ((nn + e) * (nn + e));
    cu1 = cn * 2.;
    cv1 = ct * 2. * alpha * bb / (bb + e) + cn * (-2. * alpha - 2. * q / (bb + e));
  } else if (kind == 1) {
    // sdr(s1=labels, s2=est) = k10*ln(sn^2/(d^2+e)); d = |s1-s2|^2 ; gradient wrt s2 (and s1)
    double sn = p[0], d = p[1];
    double c = g * k10 / B;
    double cd = -c * 2. * d / (d * d + e);  // d out / d d
    // d d / d s2 = -2 (s1 - s2); d d / d s1 = 2 (s1 - s2); d sn/d s1 = 2 s1
    cu2 = cd * -2.; cv2 = cd * 2.;
    cu1 = c * 2. / sn * 2. + cd * 2.; cv1 = cd * -2.;
  } else if (kind == 2) {
    // si_sdr(reference=s1, estimation=s2); gradient wrt s2 only
    double re = p[0], ab = p[1], P = p[2], Nn = p[3];
    double alpha = ab / re + e;
    // R = mean_b(P/Nn + e) needs all rows: recompute here (B is small)
    double Rm = 0.;
    for (int i = 0; i < B; ++i) Rm += part[i * 4 + 2] / part[i * 4 + 3] + e;
    Rm /= B;
    double c = g * k10 / (Rm + e) / B;
    double nr = ab - alpha * re;  // <noise, ref>
    // dP/d est = 2 alpha ref ; dNn/d est = 2 noise - 2 nr/re * ref, noise = est - alpha ref
    double cP = c / Nn, cN = -c * P / (Nn * Nn);
    cv2 = cN * 2.;
    cu2 = cP * 2. * alpha + cN * (-2. * alpha - 2. * nr / re);
  } else {
    double c = g * 2. / ((double)B * L);
    cu1 = c; cv1 = -c; cu2 = -c; cv2 = c;
  }
  const float fu1 = (float)cu1, fv1 = (float)cv1, fu2 = (float)cu2, fv2 = (float)cv2;
  float* d1 = ds1 ? ds1 + (int64_t)b * L : nullptr;
  float* d2 = ds2 ? ds2 + (int64_t)b * L : nullptr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
    float u = p1[i], v = p2[i];
    if (d1) d1[i] = fu1 * u + fv1 * v;
    if (d2) d2[i] = fu2 * u + fv2 * v;
  }
}

// ------------------------------------------------------------------------------- stft-mag loss
__global__ void stftmag_loss_fwd_kernel(const float* __restrict__ xs, const float* __restrict__ ys,
                                        int64_t n, double* __restrict__ part) {
  __shared__ double sh[32];
  double a0 = 0., a1 = 0., a2 = 0.;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float2 x = *reinterpret_cast<const float2*>(xs + 2 * i);
    float2 y = *reinterpret_cast<const float2*>(ys + 2 * i);
    float xm = sqrtf(fmaxf(x.x * x.x + x.y * x.y, 1e-7f));
    float ym = sqrtf(fmaxf(y.x * y.x + y.y * y.y, 1e-7f));
    a0 += (double)fabsf(logf(ym) - logf(xm));
    float d = ym - xm;
    a1 += (double)(d * d);
    a2 += (double)(ym * ym);
  }
  a0 = block_sum(a0, sh);
  a1 = block_sum(a1, sh);
  a2 = block_sum(a2, sh);
  if (threadIdx.x == 0) {
    atomicAdd(part, a0);
    atomicAdd(part + 1, a1);
    atomicAdd(part + 2, a2);
  }
}

__global__ void stftmag_final_kernel(const double* __restrict__ part, int64_t n,
                                     float* __restrict__ out) {
  out[0] = (float)(sqrt(part[1]) / sqrt(part[2]));  // spectral convergence
  out[1] = (float)(part[0] / (double)n);            // mean |log y - log x|
}

__global__ void stftmag_loss_bwd_kernel(const float* __restrict__ xs, const float* __restrict__ ys,
                                        int64_t n, const double* __restrict__ part,
                                        const float* __restrict__ gmag, const float* __restrict__ gsc,
                                        float scale_mag, float scale_sc, float* __restrict__ dxs) {
  const float cm = gmag ? gmag[0] * scale_mag : 0.f;
  float cs = 0.f;
  if (gsc) {
    double den = sqrt(part[1]) * sqrt(part[2]);
    cs = den > 0. ? (float)((double)gsc[0] * scale_sc / den) : 0.f;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float2 x = *reinterpret_cast<const float2*>(xs + 2 * i);
    float2 y = *reinterpret_cast<const float2*>(ys + 2 * i);
    float px = x.x * x.x + x.y * x.y;
    float2 o = make_float2(0.f, 0.f);
    if (px >= 1e-7f) {
      float xm = sqrtf(px);
      float ym = sqrtf(fmaxf(y.x * y.x + y.y * y.y, 1e-7f));
      float dl = logf(xm) - logf(ym);
      float sg = dl > 0.f ? 1.f : (dl < 0.f ? -1.f : 0.f);
      float dm = cm * sg / xm + cs * (xm - ym);  // d loss / d xm
      o.x = dm * x.x / xm;
      o.y = dm * x.y / xm;
    }
    *reinterpret_cast<float2*>(dxs + 2 * i) = o;
  }
}

// ------------------------------------------------------------------------------- Gram (SPKD)
constexpr int GT = 64, GK = 16, GNT = 256;

template <typename T>
__global__ void __launch_bounds__(GNT) gram_fwd_kernel(const T* __restrict__ z, int B, int64_t K,
                                                       int64_t ldz, int64_t k_per_block, int nbt,
                                                       bool vec, float* __restrict__ G) {
  __shared__ float As[GK][GT + 4];
  __shared__ float Bs[GK][GT + 4];
  // decode the (bi <= bj) tile pair
  int pair = blockIdx.y, bi = 0, bj = 0;
  {
    int p = pair;
    for (bi = 0; bi < nbt; ++bi) {
      int cnt = nbt - bi;
      if (p < cnt) { bj = bi + p; break; }
      p -= cnt;
    }
  }
  const int tid = threadIdx.x;
  const int l_row = tid >> 2, l_k4 = (tid & 3) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  const int64_t kbeg = (int64_t)blockIdx.x * k_per_block;
  const int64_t kend = min(K, kbeg + k_per_block);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ri = bi * GT + l_row, rj = bj * GT + l_row;
  for (int64_t kk = kbeg; kk < kend; kk += GK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    int64_t k = kk + l_k4;
    if (vec && k + 3 < kend) {
      if (ri < B) { float4 v = ld4(z + (int64_t)ri * ldz + k); av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w; }
      if (bi != bj && rj < B) { float4 v = ld4(z + (int64_t)rj * ldz + k); bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w; }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (k + e < kend) {
          if (ri < B) av[e] = ld_f(z + (int64_t)ri * ldz + k + e);
          if (bi != bj && rj < B) bv[e] = ld_f(z + (int64_t)rj * ldz + k + e);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      As[l_k4 + e][l_row] = av[e];
      Bs[l_k4 + e][l_row] = (bi != bj) ? bv[e] : av[e];
    }
    __syncthreads();
#pragma unroll
    for (int k2 = 0; k2 < GK; ++k2) {
      float4 a = *reinterpret_cast<const float4*>(&As[k2][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k2][tx * 4]);
      float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = bi * GT + ty * 4 + i;
    if (r >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = bj * GT + tx * 4 + j;
      if (c >= B) continue;
      atomicAdd(G + (int64_t)r * B + c, acc[i][j]);
      if (bi != bj) atomicAdd(G + (int64_t)c * B + r, acc[i][j]);
    }
  }
}

// dZ[i,k] = g * sum_j (dG[i,j]+dG[j,i]) Z[j,k]
template <typename T, typename TD>
__global__ void __launch_bounds__(GNT) gram_bwd_kernel(const T* __restrict__ z, int B, int64_t K,
                                                       int64_t ldz, const float* __restrict__ dG,
                                                       const float* __restrict__ gout,
                                                       TD* __restrict__ dz, int64_t lddz,
                                                       int accumulate) {
  __shared__ float As[GK][GT + 4];  // [j][i]
  __shared__ float Bs[GK][GT + 4];  // [j][k]
  const int tid = threadIdx.x;
  const int64_t k0 = (int64_t)blockIdx.x * GT;
  const int i0 = blockIdx.y * GT;
  const int ty = tid >> 4, tx = tid & 15;
  const int a_i = tid >> 2, a_j4 = (tid & 3) * 4;  // S tile loader: row i, 4 consecutive j
  const int b_j = tid >> 4, b_k4 = (tid & 15) * 4; // Z tile loader: row j, 4 consecutive k
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const float g = gout ? gout[0] : 1.f;

  for (int j0 = 0; j0 < B; j0 += GK) {
    float av[4], bv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int i = i0 + a_i, j = j0 + a_j4 + e;
      av[e] = (i < B && j < B) ? dG[(int64_t)i * B + j] + dG[(int64_t)j * B + i] : 0.f;
      int jj = j0 + b_j;
      int64_t k = k0 + b_k4 + e;
      bv[e] = (jj < B && k < K) ? ld_f(z + (int64_t)jj * ldz + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      As[a_j4 + e][a_i] = av[e];
      Bs[b_j][b_k4 + e] = bv[e];
    }
    __syncthreads();
#pragma unroll
    for (int k2 = 0; k2 < GK; ++k2) {
      float4 a = *reinterpret_cast<const float4*>(&As[k2][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k2][tx * 4]);
      float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = i0 + ty * 4 + i;
    if (r >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t k = k0 + tx * 4 + j;
      if (k >= K) continue;
      TD* p = dz + (int64_t)r * lddz + k;
      float v = g * acc[i][j];
      if (accumulate) v += ld_f(p);
      st_f(p, v);
    }
  }
}

// single block; one WARP per row with coalesced reads (a thread per row walked its row at stride B: 34 us per launch
// for a 64 x 64 pair, 14 launches per step)
__global__ void spkd_loss_kernel(const float* __restrict__ Gt, const float* __restrict__ Gs, int B,
                                 float scale, float* __restrict__ loss, float* __restrict__ dGs) {
  extern __shared__ double shd[];  // 1/nt[B], 1/ns[B] (0 when the norm was clamped: see below), ns[B], red[32]
  double* int_ = shd;
  double* ins = shd + B;
  double* nsv = shd + 2 * B;
  double* red = shd + 3 * B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  // row L1 norms (F.normalize(p=1, dim=1, eps=1e-12), framework.py:159)
  for (int i = warp; i < B; i += nw) {
    double a = 0., b = 0.;
    for (int j = lane; j < B; j += 32) {
      a += fabs((double)Gt[(int64_t)i * B + j]);
      b += fabs((double)Gs[(int64_t)i * B + j]);
    }
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      int_[i] = 1. / fmax(a, 1e-12);
      ins[i] = 1. / fmax(b, 1e-12);
      nsv[i] = fmax(b, 1e-12);
    }
  }
  __syncthreads();
  // loss = scale * sum D^2, D = Gt/nt - Gs/ns; rowdot_i = sum_j E_ij Gs_ij with E = -2 scale D
  double acc = 0.;
  for (int i = warp; i < B; i += nw) {
    const double it = int_[i], is = ins[i];
    double s = 0.;
    for (int j = lane; j < B; j += 32) {
      const double gs = (double)Gs[(int64_t)i * B + j];
      const double d = (double)Gt[(int64_t)i * B + j] * it - gs * is;
      acc += d * d;
      s += -2. * scale * d * gs;
    }
    if (dGs) {
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      // E_ij = dL/dGhat_s = -2*scale*D_ij ; dGs_ij = E_ij/n_i - sign(Gs_ij) * (sum_k E_ik Gs_ik)/n_i^2
      const bool clamped = !(nsv[i] > 1e-12);
      for (int j = lane; j < B; j += 32) {
        const double gs = (double)Gs[(int64_t)i * B + j];
        const double d = (double)Gt[(int64_t)i * B + j] * it - gs * is;
        const double e = -2. * scale * d;
        const double sg = gs > 0. ? 1. : (gs < 0. ? -1. : 0.);
        double v = e * is;
        if (!clamped) v -= sg * s * is * is;
        dGs[(int64_t)i * B + j] = (float)v;
      }
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) loss[0] = (float)(acc * (double)scale);
}

}  // namespace
}  // namespace clskd

using namespace clskd;
#define ST ((cudaStream_t)stream)

extern "C" int clskd_wave_loss_fwd(const float* s1, const float* s2, int B, int L, int kind,
                                   float eps, double* part, float* out, void* stream) {
  CLSKD_CHECK_ARG(s1 && s2 && part && out, "clskd_wave_loss_fwd: null pointer");
  CLSKD_CHECK_ARG(kind >= 0 && kind <= 3 && B > 0 && L > 0, "clskd_wave_loss_fwd: bad arguments");
  cudaError_t e = cudaMemsetAsync(part, 0, sizeof(double) * 4 * B, ST);
  if (e != cudaSuccess) { set_error("clskd_wave_loss_fwd: memset: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
  int gx = cdiv(L, 256 * 16);
  if (gx < 1) gx = 1;
  dim3 grid(gx, B);
  wave_loss_pass_kernel<1><<<grid, 256, 0, ST>>>(s1, s2, L, kind, eps, part);
  if (kind == 0 || kind == 2) wave_loss_pass_kernel<2><<<grid, 256, 0, ST>>>(s1, s2, L, kind, eps, part);
  wave_loss_final_kernel<<<1, 256, 0, ST>>>(part, B, L, kind, eps, out);
  CLSKD_CHECK_LAUNCH("clskd_wave_loss_fwd");
  return CLSKD_OK;
}

extern "C" int clskd_wave_loss_bwd(const float* s1, const float* s2, int B, int L, int kind,
                                   float eps, const double* part, const float* gout, float* ds1,
                                   float* ds2, void* stream) {
  CLSKD_CHECK_ARG(s1 && s2 && part && gout && (ds1 || ds2), "clskd_wave_loss_bwd: null pointer");
  CLSKD_CHECK_ARG(!(kind == 0 && ds2), "clskd_wave_loss_bwd: si_snr gradient wrt the reference s2 is not provided");
  CLSKD_CHECK_ARG(!(kind == 2 && ds1), "clskd_wave_loss_bwd: si_sdr gradient wrt the reference s1 is not provided");
  int gx = cdiv(L, 256 * 8);
  if (gx < 1) gx = 1;
  wave_loss_bwd_kernel<<<dim3(gx, B), 256, 0, ST>>>(s1, s2, B, L, kind, eps, part, gout, ds1, ds2);
  CLSKD_CHECK_LAUNCH("clskd_wave_loss_bwd");
  return CLSKD_OK;
}

extern "C" int clskd_stftmag_loss_fwd(const float* xs, const float* ys, int64_t n, double* part,
                                      float* out, void* stream) {
  CLSKD_CHECK_ARG(xs && ys && part && out && n > 0, "clskd_stftmag_loss_fwd: bad arguments");
  cudaError_t e = cudaMemsetAsync(part, 0, sizeof(double) * 3, ST);
  if (e != cudaSuccess) { set_error("clskd_stftmag_loss_fwd: memset: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
  if (n == 0) return CLSKD_OK;
  int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
  int cap = sm_count() * 8;
  int grid = (int)(blocks < cap ? blocks : cap);
  stftmag_loss_fwd_kernel<<<grid, 256, 0, ST>>>(xs, ys, n, part);
  stftmag_final_kernel<<<1, 1, 0, ST>>>(part, n, out);
  CLSKD_CHECK_LAUNCH("clskd_stftmag_loss_fwd");
  return CLSKD_OK;
}

extern "C" int clskd_stftmag_loss_bwd(const float* xs, const float* ys, int64_t n,
                                      const double* part, const float* gout_mag,
                                      const float* gout_sc, float scale_mag, float scale_sc,
                                      float* dxs, void* stream) {
  CLSKD_CHECK_ARG(xs && ys && part && dxs, "clskd_stftmag_loss_bwd: null pointer");
  if (n == 0) return CLSKD_OK;
  int64_t blocks = (n + 255) / 256;
  int cap = sm_count() * 16;
  int grid = (int)(blocks < cap ? blocks : cap);
  stftmag_loss_bwd_kernel<<<grid, 256, 0, ST>>>(xs, ys, n, part, gout_mag, gout_sc, scale_mag,
                                                scale_sc, dxs);
  CLSKD_CHECK_LAUNCH("clskd_stftmag_loss_bwd");
  return CLSKD_OK;
}

extern "C" int clskd_gram_fwd(const void* z, int dtype, int B, int64_t K, int64_t ldz, float* G,
                              int accumulate, void* stream) {
  CLSKD_CHECK_ARG(z && G && B >= 1 && K >= 0, "clskd_gram_fwd: bad arguments");
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(G, 0, sizeof(float) * (size_t)B * B, ST);
    if (e != cudaSuccess) { set_error("clskd_gram_fwd: memset: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
  }
  if (K == 0) return CLSKD_OK;
  int nbt = cdiv(B, GT);
  int npairs = nbt * (nbt + 1) / 2;
  int64_t splits = (int64_t)sm_count() * 8 / npairs;
  int64_t max_splits = (K + 1023) / 1024;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t kpb = (K + splits - 1) / splits;
  kpb = (kpb + GK - 1) / GK * GK;
  splits = (K + kpb - 1) / kpb;
  int es = dtype == CLSKD_F32 ? 4 : 2;
  bool vec = ((uintptr_t)z % (4 * es) == 0) && (ldz % 4 == 0);
  dim3 grid((unsigned)splits, npairs);
  CLSKD_DISPATCH_DTYPE(dtype, T,
                       (gram_fwd_kernel<T><<<grid, GNT, 0, ST>>>((const T*)z, B, K, ldz, kpb, nbt,
                                                                   vec, G)));
  CLSKD_CHECK_LAUNCH("clskd_gram_fwd");
  return CLSKD_OK;
}

extern "C" int clskd_spkd_loss(const float* Gt, const float* Gs, int B, float scale, float* loss,
                               float* dGs, void* stream) {
  CLSKD_CHECK_ARG(Gt && Gs && loss && B >= 1 && B <= 1024, "clskd_spkd_loss: bad arguments");
  size_t sh = (3 * (size_t)B + 32) * sizeof(double);
  spkd_loss_kernel<<<1, 256, sh, ST>>>(Gt, Gs, B, scale, loss, dGs);
  CLSKD_CHECK_LAUNCH("clskd_spkd_loss");
  return CLSKD_OK;
}

extern "C" int clskd_gram_bwd(const void* z, int dtype, int B, int64_t K, int64_t ldz,
                              const float* dG, const float* gout, void* dz, int dz_dtype,
                              int64_t lddz, int accumulate, void* stream) {
  CLSKD_CHECK_ARG(z && dG && dz && B >= 1, "clskd_gram_bwd: bad arguments");
  if (K == 0) return CLSKD_OK;
  int64_t gx = (K + GT - 1) / GT;
  CLSKD_CHECK_ARG(gx <= 2147483647LL, "clskd_gram_bwd: K too large");
  dim3 grid((unsigned)gx, cdiv(B, GT));
#define L(T, TD)                                                                                 \
  gram_bwd_kernel<T, TD><<<grid, GNT, 0, ST>>>((const T*)z, B, K, ldz, dG, gout, (TD*)dz, lddz,  \
                                               accumulate)
  if (dtype == CLSKD_F32 && dz_dtype == CLSKD_F32) L(float, float);
  else if (dtype == CLSKD_F32) L(float, __nv_bfloat16);
  else if (dz_dtype == CLSKD_F32) L(__nv_bfloat16, float);
  else L(__nv_bfloat16, __nv_bfloat16);
#undef L
  CLSKD_CHECK_LAUNCH("clskd_gram_bwd");
  return CLSKD_OK;
}


// ---------------------------------------------------------------------------------------------------------
// Per-utterance second moments of a pair of waveforms (fp64): out[b] = (sum a, sum b, sum a*a, sum a*b, sum b*b).
// Everything the batched validation metrics need (SI-SDR with mean removal, SNR, improvements: distill.py:150-200
// computes them one utterance at a time on the CPU through asteroid's get_metrics) in one pass over the batch.
// ---------------------------------------------------------------------------------------------------------
namespace clskd {
namespace {
__global__ void __launch_bounds__(256) pair_moments_kernel(const float* __restrict__ a, const float* __restrict__ b, int L,
                                                           int64_t a_sB, int64_t b_sB, double* __restrict__ out) {
  __shared__ double sh[32];
  const int bi = blockIdx.y;
  const float* ap = a + (int64_t)bi * a_sB;
  const float* bp = b + (int64_t)bi * b_sB;
  double s[5] = {0., 0., 0., 0., 0.};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
    const double x = ap[i], y = bp[i];
    s[0] += x; s[1] += y; s[2] += x * x; s[3] += x * y; s[4] += y * y;
  }
  for (int k = 0; k < 5; ++k) {
    const double t = block_sum<double>(s[k], sh);
    if (threadIdx.x == 0) atomicAdd(out + (int64_t)bi * 5 + k, t);
    __syncthreads();
  }
}
}  // namespace
}  // namespace clskd

extern "C" int clskd_pair_moments(const float* a, const float* b, int B, int L, int64_t a_sB, int64_t b_sB, double* out,
                                  void* stream) {
  CLSKD_CHECK_ARG(a && b && out && B >= 0 && L >= 0, "clskd_pair_moments: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return CLSKD_OK;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double) * 5 * (size_t)B, st);
  if (e != cudaSuccess) { clskd::set_error("clskd_pair_moments: memset: %s", cudaGetErrorString(e)); return CLSKD_ERR_CUDA; }
  if (L == 0) return CLSKD_OK;
  int bx = (L + 256 * 8 - 1) / (256 * 8);
  if (bx > 64) bx = 64;
  clskd::pair_moments_kernel<<<dim3(bx, B), 256, 0, st>>>(a, b, L, a_sB, b_sB, out);
  CLSKD_CHECK_LAUNCH("clskd_pair_moments");
  return CLSKD_OK;
}
