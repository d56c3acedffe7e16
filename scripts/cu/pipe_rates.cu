// Per-SM throughput of the instructions the focal kernels lean on: MUFU.EX2 / RCP / LG2, FFMA, FFMA2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates.bin pipe_rates.cu && ./pipe_rates.bin
// Each thread runs 8 independent chains so that latency is covered; 148 x 8 CTAs of 256 threads.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(256) rate_kernel(float* out, int iters, float seed) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed + 0.001f * (threadIdx.x + i);
  float2 w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = make_float2(v[2 * i], v[2 * i + 1]);
  const float2 c2 = make_float2(0.999f, 0.999f), d2 = make_float2(1e-3f, 1e-3f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(0.999f), "f"(1e-3f));
    }
    if (OP == 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        w[i] = __ffma2_rn(w[i], c2, d2);
        w[i] = __ffma2_rn(w[i], c2, d2);          // 8 FFMA2 = 16 fp32 FMAs per iteration
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += w[i].x + w[i].y;
  if (s == 123.456f) out[0] = s;
}

template <int OP>
void run(const char* name, int lanes_per_iter, int sms, float* out) {
  const int iters = 4096, ctas = sms * 8;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  rate_kernel<OP><<<ctas, 256>>>(out, iters, 0.7f);
  cudaEventRecord(a);
  rate_kernel<OP><<<ctas, 256>>>(out, iters, 0.7f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  int khz;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ops = (double)ctas * 256 * iters * lanes_per_iter;
  printf("%-10s %8.3f ms  %7.2f lane-ops / clk / SM (at the %d MHz attribute clock)\n", name, ms,
         ops / (ms * 1e-3) / (khz * 1e3) / sms, khz / 1000);
}

int main() {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, 4);
  run<0>("MUFU.EX2", 8, sms, out);
  run<1>("MUFU.RCP", 8, sms, out);
  run<2>("MUFU.LG2", 8, sms, out);
  run<3>("FFMA", 8, sms, out);
  run<4>("FFMA2", 16, sms, out);
  return cudaDeviceSynchronize() != cudaSuccess;
}
