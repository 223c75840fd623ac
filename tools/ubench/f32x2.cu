// Microbenchmark: issue rate of packed fp32 (add / mul / fma .f32x2 -> FADD2 / FMUL2 / FFMA2 on sm_100a) against the
// scalar instructions, 8 independent accumulator chains per thread.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

template <int OP> __device__ __forceinline__ unsigned long long op2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  if (OP == 0) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  if (OP == 1) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  if (OP == 2) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
template <int OP> __device__ __forceinline__ float op1(float a, float b, float c) {
  float r;
  if (OP == 0) asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  if (OP == 1) asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  if (OP == 2) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
template <int OP> __global__ void k_packed(unsigned long long *out, int iters, float seed) {
  unsigned long long acc[8], b, c;
  float s = seed + threadIdx.x * 1e-9f;
  asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(1.0f + 1e-7f * s));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(1e-9f * s));
#pragma unroll
  for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(s + i));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = op2<OP>(acc[i], b, c);
  }
  unsigned long long r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r ^= acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int OP> __global__ void k_scalar(float *out, int iters, float seed) {
  float acc[8];
  float s = seed + threadIdx.x * 1e-9f;
  const float b = 1.0f + 1e-7f * s, c = 1e-9f * s;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = s + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = op1<OP>(acc[i], b, c);
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <typename F> float timed(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int blocks = sms * 4, threads = 512, iters = 1 << 16;
  void *buf; cudaMalloc(&buf, (size_t)blocks * threads * 8);
  const double instr = (double)blocks * threads / 32 * 8.0 * iters;  // warp instructions per launch
  const char *names[3] = {"add", "mul", "fma"};
  printf("%d SMs, %d threads x %d CTAs, %d iterations x 8 chains; clock attribute %.0f MHz\n", sms, threads, blocks, iters, khz / 1e3);
#define RUN(OP)                                                                                                          \
  {                                                                                                                      \
    float ms1 = timed([&] { k_scalar<OP><<<blocks, threads>>>((float *)buf, iters, 1.f); });                            \
    float ms2 = timed([&] { k_packed<OP><<<blocks, threads>>>((unsigned long long *)buf, iters, 1.f); });               \
    printf("%s: scalar %.3f ms (%.1f warp-instr/clk/SM at the clock attribute), packed .f32x2 %.3f ms (%.1f): packed / scalar time %.2f\n", \
           names[OP], ms1, instr / (ms1 * 1e-3) / (khz * 1e3) / sms, ms2, instr / (ms2 * 1e-3) / (khz * 1e3) / sms, ms2 / ms1); \
  }
  RUN(0) RUN(1) RUN(2)
  return 0;
}
