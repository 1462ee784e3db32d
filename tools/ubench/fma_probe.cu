// Micro-benchmark (GPU box): issue rates of scalar FFMA, packed FFMA2 and mixes of the two on sm_100a,
// to size k_verify / k_vote_join's FP32 pre-filters.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// NP packed chains + NS scalar chains per thread, `iters` rounds
template <int NP, int NS>
__global__ void k(float *out, int iters, float seed) {
  f32x2 p[NP > 0 ? NP : 1];
  float s[NS > 0 ? NS : 1];
  const float a = seed + threadIdx.x * 1e-9f;
  f32x2 a2;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
  for (int i = 0; i < NP; ++i) p[i] = a2 + i;
  for (int i = 0; i < NS; ++i) s[i] = a + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < NP; ++i) p[i] = fma2(p[i], a2, a2);
#pragma unroll
      for (int i = 0; i < NS; ++i) s[i] = fma1(s[i], a, a);
    }
  }
  float acc = 0;
  for (int i = 0; i < NP; ++i) acc += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  for (int i = 0; i < NS; ++i) acc += s[i];
  if (acc == 12345.678f) out[0] = acc;
}

template <int NP, int NS>
void run(const char *name, int sms, float *d) {
  const int iters = 4096, threads = 512, blocks = sms * 4;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NP, NS><<<blocks, threads>>>(d, 16, 1.0f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<NP, NS><<<blocks, threads>>>(d, iters, 1.0f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double inst = (double)blocks * threads / 32 * iters * 8 * (NP + NS);  // warp instructions
  const double lanes = (double)blocks * threads * iters * 8 * (2.0 * NP + NS);  // FP32 FMAs
  printf("%-28s %8.3f ms  %7.2f warp-inst/clk/SM (at 1.965 GHz)  %7.2f TFMA/s\n", name, ms,
         inst / (ms * 1e-3) / 1.965e9 / sms, lanes / (ms * 1e-3) / 1e12);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  float *d;
  cudaMalloc(&d, 4);
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, sms);
  run<0, 12>("scalar FFMA x12", sms, d);
  run<12, 0>("packed FFMA2 x12", sms, d);
  run<8, 4>("FFMA2 x8 + FFMA x4", sms, d);
  run<6, 6>("FFMA2 x6 + FFMA x6", sms, d);
  run<4, 8>("FFMA2 x4 + FFMA x8", sms, d);
  run<8, 8>("FFMA2 x8 + FFMA x8", sms, d);
  run<10, 2>("FFMA2 x10 + FFMA x2", sms, d);
  return 0;
}
