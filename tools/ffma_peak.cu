// Microbenchmark: sustained fp32 FMA rate per SM for register-operand FFMA and packed FFMA2 (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma_peak tools/ffma_peak.cu && /tmp/ffma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s0, float s1) {
  float a[16], b0 = s0 + threadIdx.x, b1 = s1 + threadIdx.x * 0.5f;
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = i * 0.001f + threadIdx.x;
  float w[4] = {s0, s1, s0 * 0.5f, s1 * 0.25f};
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // 16 independent 3-register FFMAs, two different multiplier pairs
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(w[i & 3], (i & 1) ? b0 : b1, a[i]);
    } else {          // 8 packed FFMA2 = 16 FMAs
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float2 r = __ffma2_rn(make_float2(w[i & 3], w[(i + 1) & 3]), make_float2(b0, b1), make_float2(a[2 * i], a[2 * i + 1]));
        a[2 * i] = r.x;
        a[2 * i + 1] = r.y;
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* out;
  cudaMalloc(&out, sms * 8 * 256 * sizeof(float));
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<sms * 8, 256>>>(out, iters, 1.0001f, 0.9999f);
      else k<1><<<sms * 8, 256>>>(out, iters, 1.0001f, 0.9999f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      double fma = (double)sms * 8 * 256 * iters * 16;
      printf("%s: %.3f ms  %.1f TFLOP/s  %.1f FMA/clk/SM (at %.0f MHz nominal)\n", mode ? "FFMA2" : "FFMA ", ms,
             2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / (khz * 1e3), khz / 1e3);
    }
  }
  return 0;
}
