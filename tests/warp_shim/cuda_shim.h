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

// ---- TMA ring of k_pd_tv3d_f2t: all warps of a CTA (consumers + the producer warp) run concurrently; mbarriers
// are emulated (phase bit, pending arrivals, pending transaction bytes) under one global mutex; a bulk copy is a
// memcpy followed by complete_tx.  This checks the program logic of the ring (stage / parity bookkeeping, who
// waits for whom, no deadlock, every packet read from the right stage); what it cannot show is an ordering
// mistake of the async proxy against pending LDS reads.
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>
extern std::barrier<> *shim_cta_bar;
inline std::mutex &shim_mbar_mutex() { static std::mutex m; return m; }
struct ShimMbar { uint32_t phase : 1, init : 15, pending : 16; int32_t tx; };
static_assert(sizeof(ShimMbar) == 8, "an emulated mbarrier lives in the kernel's uint64_t");
inline void shim_mbar_check(ShimMbar *b) {  // call with the mutex held
  if (b->pending == 0 && b->tx == 0) { b->phase ^= 1u; b->pending = b->init; }
}
inline void __syncwarp() { shim_warp->bar.arrive_and_wait(); }
inline void __syncthreads() { shim_cta_bar->arrive_and_wait(); }
inline void mbar_init(uint64_t *bar, uint32_t count) {
  ShimMbar *b = reinterpret_cast<ShimMbar *>(bar);
  b->phase = 0; b->init = count; b->pending = count; b->tx = 0;
}
inline void mbar_fence_init() {}
inline void mbar_arrive(uint64_t *bar) {
  std::lock_guard<std::mutex> g(shim_mbar_mutex());
  ShimMbar *b = reinterpret_cast<ShimMbar *>(bar);
  b->pending -= 1;
  shim_mbar_check(b);
}
inline void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  std::lock_guard<std::mutex> g(shim_mbar_mutex());
  ShimMbar *b = reinterpret_cast<ShimMbar *>(bar);
  b->tx += (int32_t)bytes;
  b->pending -= 1;
  shim_mbar_check(b);
}
inline void mbar_wait_spin(uint64_t *bar, uint32_t parity) {
  // mbarrier.try_wait.parity: true once the phase with that parity has completed, i.e. the current phase bit differs
  for (long spins = 0;; ++spins) {
    {
      std::lock_guard<std::mutex> g(shim_mbar_mutex());
      const ShimMbar *b = reinterpret_cast<ShimMbar *>(bar);
      if (b->phase != (parity & 1u)) return;
      if (spins == 300000) {  // a deadlock of the ring protocol: say where and give up
        std::fprintf(stderr, "shim: thread %u stuck on mbarrier %p parity %u (phase %u pending %u tx %d)\n", threadIdx.x,
                     (void *)bar, parity, (unsigned)b->phase, (unsigned)b->pending, b->tx);
        std::abort();
      }
    }
    std::this_thread::yield();
  }
}
inline void mbar_wait(uint64_t *bar, uint32_t parity) { mbar_wait_spin(bar, parity); }
inline void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  std::memcpy(dst, src, bytes);
  std::lock_guard<std::mutex> g(shim_mbar_mutex());
  ShimMbar *b = reinterpret_cast<ShimMbar *>(bar);
  b->tx -= (int32_t)bytes;
  shim_mbar_check(b);
}
inline int __shfl_sync(unsigned, int v, int src) {  // broadcast of an int
  ShimWarp &w = *shim_warp;
  w.slot[shim_lane] = (float)v;
  w.bar.arrive_and_wait();
  const int r = (int)w.slot[src];
  w.bar.arrive_and_wait();
  return r;
}
