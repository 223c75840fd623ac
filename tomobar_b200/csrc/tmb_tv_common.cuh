// Pieces shared by the two TV translation units (tmb_tv.cu: PD_TV, tmb_tv_rof.cu: ROF_TV): tile constants,
// fp16 / fp32 load-store helpers, the kernel-selection test hook.
#pragma once
#include <cuda_fp16.h>

#include <cstddef>

#include "tmb_common.h"
#include "tmb_tv_fused.cuh"

namespace tmb {

template <typename T> __device__ __forceinline__ float ldp(const T *p, size_t i);
template <> __device__ __forceinline__ float ldp<float>(const float *p, size_t i) { return __ldg(p + i); }
template <> __device__ __forceinline__ float ldp<__half>(const __half *p, size_t i) { return __half2float(p[i]); }
template <typename T> __device__ __forceinline__ void stp(T *p, size_t i, float v);
template <> __device__ __forceinline__ void stp<float>(float *p, size_t i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void stp<__half>(__half *p, size_t i, float v) { p[i] = __float2half(v); }

constexpr int TV_BX = 128, TV_BY = 2, TV_ZRUN = 8;
constexpr int PT_TX = 64, PT_TY = 8, PT_THREADS = PT_TX * PT_TY;
constexpr int PT_HX = PT_TX + 2, PT_HY = PT_TY + 2, PT_PLANE = PT_HX * PT_HY;
constexpr int PW_RY = 4, PW_WARPS = 4, PW_TX = 128, PW_STAGES = 4;

__device__ __forceinline__ float ldg1(const float *p) { return __ldg(p); }
__device__ __forceinline__ float ldg1(const __half *p) { return __half2float(*p); }
__device__ __forceinline__ float4 cvt4(const uint2 raw) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 ldv4(const __half *p) { return cvt4(__ldg(reinterpret_cast<const uint2 *>(p))); }
// the same from shared memory
__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 lds4(const __half *p) { return cvt4(*reinterpret_cast<const uint2 *>(p)); }
__device__ __forceinline__ void stv4(__half *p, const float4 &v) {
  const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 raw;
  raw.x = *reinterpret_cast<const unsigned *>(&a);
  raw.y = *reinterpret_cast<const unsigned *>(&b);
  *reinterpret_cast<uint2 *>(p) = raw;
}

// test hook (tmb_tv_set_simple_kernels), see tmb_tv.cu
extern int g_tv_simple;
inline dim3 tv_grid(int dx, int dy, int dz) {
  return dim3((dx + TV_BX - 1) / TV_BX, (dy + TV_BY - 1) / TV_BY, (dz + TV_ZRUN - 1) / TV_ZRUN);
}

}  // namespace tmb
