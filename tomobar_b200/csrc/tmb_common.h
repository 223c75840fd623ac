// Internal helpers shared by the libtmb translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/tmb.h"

namespace tmb {

void set_error(const std::string &msg);

#define TMB_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::tmb::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));               \
      return TMB_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define TMB_REQUIRE(cond, msg)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::tmb::set_error(msg);                                                              \
      return TMB_ERR_ARG;                                                                 \
    }                                                                                     \
  } while (0)

// Launch-error check that does not synchronise.
inline int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return TMB_ERR_CUDA;
  }
  return TMB_OK;
}

// ---- interior layouts ---------------------------------------------------------------------
// z is blocked in chunks of 4 slices so that one smem/global element is a float4 (16 B):
//   S_int[zc][a][SPAD + u + SPAD]   sinogram lines, zero borders of SPAD elements
//   V1  [zc][r][VPAD + c + VPAD]    volume rows    (lines = rows,    interpolate along columns)
//   V0  [zc][c][VPAD + r + VPAD]    volume columns (lines = columns, interpolate along rows)
// The zero borders implement ASTRA's "zero outside" addressing and make every TMA window a
// plain in-bounds contiguous bulk copy.
constexpr int ZC = 4;         // slices per z-chunk (float4)
constexpr int NZC = 2;        // z-chunks per CTA
constexpr int BP_W = 48;      // sinogram window (elements) of a 32x32 voxel tile
constexpr int SPAD = BP_W;    // zero border of S_int
constexpr int FP_K = 128;     // detector bins per CTA in the forward projector
constexpr int FP_W = 184;     // volume-line window of FP_K bins: ceil(127*sqrt(2)) + 4
constexpr int VPAD = FP_W;    // zero border of V0 / V1
constexpr int MAX_ANGLES = 2000;  // constant-memory table capacity (2 x float4 per angle)

// "Q" layouts of the forward projector for volumes of more than FQ_MIN_NZ slices: z is blocked in
// groups of 32 slices that sit next to each other per in-plane position (128 B = one shared-memory
// wavefront), so the 8 lanes of a quarter-warp read 8 z-chunks of ONE position: bank-conflict free
//   VQ1[zg][r][QPAD + c + QPAD][8][4]   VQ0[zg][c][QPAD + r + QPAD][8][4]
constexpr int FQ_K = 64;          // detector bins per CTA
constexpr int FQ_CG = 8;          // z-chunks (of 4 slices) per position
constexpr int FQ_W = 96;          // volume-line window of FQ_K bins: ceil(63*sqrt(2)) + 4, rounded up
constexpr int QPAD = FQ_W;        // zero border
constexpr int FQ_MIN_NZ = 17;     // smaller stacks keep the 8-slice kernel

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// cudaFuncSetAttribute applies to the CURRENT device only: one flag per device for each call site
// (a process that reconstructs on cuda:0 and then on cuda:1 must opt in to the large shared memory twice)
struct PerDeviceOnce {
  unsigned long long done = 0;  // bit d: device d (mod 64) has been set up
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    const unsigned long long bit = 1ull << (d & 63);
    if (done & bit) return false;
    done |= bit;
    return true;
  }
};

struct GeomDims {
  int nz, n, nu, na;
  int nzc;   // z-chunks allocated (multiple of NZC)
  int up;    // nu + 2*SPAD
  int qp;    // n + 2*VPAD
  int nzg;   // 32-slice groups of the Q layouts
  int qpq;   // n + 2*QPAD
};

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------
// mbarrier / bulk-copy primitives (PTX)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait suspends the thread for a hardware time slice per attempt.  A barrier that has not completed after
// 30 SECONDS of wall clock (%globaltimer, looked at every 4096 attempts) means a lost copy (bad size / alignment):
// trap instead of hanging the device.  The bound is a time, not an attempt count, so that time-slicing, MPS, a
// debugger or a profiler replay stretching a healthy copy cannot trip it.
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint64_t t0 = 0;
  for (uint32_t spins = 1; !mbar_try_wait(bar, parity); ++spins) {
    if ((spins & 0xfffu) == 0) {
      const uint64_t t = global_timer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 30000000000ull) __trap();
    }
  }
}
// plain spin on an mbarrier phase (no watchdog: used where the wait sits in a hot unrolled loop)
__device__ __forceinline__ void mbar_wait_spin(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// L2 eviction policy "keep": for rows that a neighbouring warp / CTA re-reads a few microseconds later
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
template <bool HINT>
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                              uint64_t pol) {
  if constexpr (HINT) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
  } else {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
  }
}
// orders earlier generic-proxy accesses of shared memory (LDS / STS) before later async-proxy ones (TMA writes)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// global -> shared bulk copy (TMA engine), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

#endif  // __CUDACC__

}  // namespace tmb

// Geometry object behind the opaque handle.
struct tmb_geom {
  tmb::GeomDims d;
  int os_number;
  int quant8;
  int bins;                 // ceil(na / os)
  float *table;             // host [na][8]
  uint64_t id;              // identity for the constant-memory cache
  int fp_q;                 // 1: forward projector runs on the Q layouts (k_fpq), 0: k_fp
  // workspace carve-up (bytes offsets)
  size_t off_v0, off_v1, off_s, off_part, ws_bytes;
  int seg_len, nseg;        // k_fpq: volume lines per L2 segment, number of segments
  int part_angles;          // angles per launch the partial-sum buffer holds
};
