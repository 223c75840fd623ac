// Just enough of the CUDA device environment to compile tomobar_b200/csrc/tmb_tv_fused.cuh with g++ and run
// its kernels on the CPU, source unchanged: one OS thread per lane, 32 lanes per warp, warp shuffles as a
// barrier-synchronised exchange, shared memory as a plain array, launch indices as thread-locals.
// Test infrastructure only (tests/test_warp_shim_fused_tv.py); it checks index logic and pointer
// arithmetic, not speed and not the last bit (the host has no MUFU unit: tolerance 2e-6).
#pragma once

#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__

struct alignas(16) float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct shim_uint3 { unsigned x, y, z; };
extern thread_local shim_uint3 threadIdx, blockIdx;

inline float __ldg(const float *p) { return *p; }
inline float4 __ldg(const float4 *p) { return *p; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }

struct ShimWarp {
  std::barrier<> bar{32};
  float slot[32];
};
extern thread_local ShimWarp *shim_warp;
extern thread_local int shim_lane;

inline float __shfl_down_sync(unsigned, float v, int d) {
  ShimWarp &w = *shim_warp;
  w.slot[shim_lane] = v;
  w.bar.arrive_and_wait();
  const float r = shim_lane + d < 32 ? w.slot[shim_lane + d] : v;
  w.bar.arrive_and_wait();
  return r;
}
inline float __shfl_up_sync(unsigned, float v, int d) {
  ShimWarp &w = *shim_warp;
  w.slot[shim_lane] = v;
  w.bar.arrive_and_wait();
  const float r = shim_lane - d >= 0 ? w.slot[shim_lane - d] : v;
  w.bar.arrive_and_wait();
  return r;
}
