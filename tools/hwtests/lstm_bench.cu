// Standalone timing of the LSTM recurrence kernels (csrc/lstm.cu compiled in) with probe switches:
//   nvcc -DLSTM_PROBE=<n> ...   0 full, 1 no global stores, 2 no gate math (h = acc), 3 no MMAs
#include <cstdarg>
#include <cstdio>
#include <vector>
#include "../../speech-enhancement-clskd_b200/csrc/lstm.cu"
namespace clskd {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
int sm_count() { return 148; }
}
int main() {
  const int T = 643, B = 64, P = 2, nsets = 2;
  for (int H : {64, 128}) {
    const int G = 4 * H, R = P * B;
    float *pre, *whh, *h, *gates, *c, *dh, *dpre;
    size_t npre = (size_t)P * T * B * nsets * G, nh = (size_t)nsets * P * T * B * H;
    cudaMalloc(&pre, npre * 4); cudaMalloc(&dpre, npre * 4); cudaMalloc(&whh, (size_t)nsets * H * G * 4);
    cudaMalloc(&h, nh * 4); cudaMalloc(&c, nh * 4); cudaMalloc(&dh, nh * 4); cudaMalloc(&gates, nh * 4 * 4);
    std::vector<float> hp(npre), hw((size_t)nsets * H * G);
    for (size_t i = 0; i < npre; ++i) hp[i] = ((i * 2654435761u) % 2001) / 1000.f - 1.f;
    for (size_t i = 0; i < hw.size(); ++i) hw[i] = (((i * 40503u) % 2001) / 1000.f - 1.f) / 8.f;
    cudaMemcpy(pre, hp.data(), npre * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(whh, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dh, 0, nh * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int train = 0; train < 2; ++train)
      for (int legacy = 0; legacy < 2; ++legacy) {
        clskd::g_lstm_legacy = legacy;
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(e0);
          int rc = clskd_lstm_fwd(pre, whh, T, R, B, H, nsets, (int64_t)T * B * nsets * G, (int64_t)B * nsets * G, nsets * G, G,
                                  (int64_t)H * G, 1, h, train ? gates : nullptr, train ? c : nullptr, nullptr);
          cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
          if (rc) return 1;
        }
        printf("probe %d H=%3d fwd train=%d %s: %.3f ms (%.2f us/step)\n", LSTM_PROBE, H, train, legacy ? "cuda-core" : "mma      ", ms, ms * 1e3 / T);
      }
    for (int legacy = 0; legacy < 2; ++legacy) {
      clskd::g_lstm_legacy = legacy;
      float ms = 0;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        int rc = clskd_lstm_bwd_policy(dh, whh, gates, c, T, R, B, H, nsets, (int64_t)G * H, (int64_t)T * B * nsets * G,
                                       (int64_t)B * nsets * G, nsets * G, G, dpre, 1, nullptr);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        if (rc) return 1;
      }
      printf("probe %d H=%3d bwd         %s: %.3f ms (%.2f us/step)\n", LSTM_PROBE, H, legacy ? "cuda-core" : "mma      ", ms, ms * 1e3 / T);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaFree(pre); cudaFree(dpre); cudaFree(whh); cudaFree(h); cudaFree(c); cudaFree(dh); cudaFree(gates);
  }
  return 0;
}
