// 3-D parallel-beam projector pair for sm_100a.
//
// Replaces astra-toolbox's par3d kernels that the reference reaches through
// astra_wrappers/astra_base.py:554 (direct_BP3D) and :601 (direct_FP3D).
//
// Both operators have the same shape:
//     out[item][z] = sum over lines  lerp( IN[line][ pos(item, line) ][z] )
//   back-projection : items = voxels of a 32x32 tile, lines = angles,       IN = sinogram
//   forward (Joseph): items = 128 detector bins of one angle, lines = the N volume rows
//                     (or columns) the ray marches through,                 IN = volume
// The interpolation index and weight depend on (item, line) only -- never on z -- so each
// thread computes them once and applies them to 8 slices (two float4 z-chunks).  The part
// of every line a CTA needs is a contiguous window of the z-blocked interior layout; a
// producer warp streams these windows into a shared-memory ring with cp.async.bulk (TMA,
// SASS UBLKCP) completing on mbarriers, the consumer warps read them with LDS.128.
// Per-angle constants live in __constant__ memory.
#include "tmb_common.h"

#include <algorithm>
#include <cmath>
#include <mutex>
#include <utility>
#include <vector>

namespace tmb {

// ------------------------------------------------------------------------------------------
// constant-memory angle table (uploaded once per geometry, see ensure_table)
//   c_bp[a] = (cos, sin, bp_off, 0)
//   c_fp[a] = (alpha, b0, bstep, +-scale)   sign(scale) < 0 <=> dir 0 (march along columns)
// ------------------------------------------------------------------------------------------
__constant__ float4 c_bp[MAX_ANGLES];
__constant__ float4 c_fp[MAX_ANGLES];

// One table per device, shared by every geometry, stream and host thread that uses the device.  Who may overwrite
// it and when is tracked here: an upload waits -- on the device, through events, never on the host --
// for every stream that launched kernels reading the current table, and a stream that finds "its" table already
// resident waits for the upload if another stream issued it.
struct TableState {
  std::mutex mu;
  uint64_t key = 0;                // geometry id * 4096 + chunk; 0 = nothing loaded
  cudaStream_t upload_stream = nullptr;
  cudaEvent_t uploaded = nullptr;  // recorded after the two symbol copies
  std::vector<std::pair<cudaStream_t, cudaEvent_t>> users;  // last kernel of each stream reading the table
};
static TableState g_table[64];

int ensure_table(const tmb_geom *g, int a_begin, int a_count, cudaStream_t st) {
  // table chunk [a_begin, a_begin + a_count) of geometry g -> constant memory slots [0, a_count)
  static thread_local float4 hbp[MAX_ANGLES], hfp[MAX_ANGLES];
  int dev = 0;
  TMB_CUDA_CHECK(cudaGetDevice(&dev));
  TableState &t = g_table[dev & 63];
  std::lock_guard<std::mutex> lock(t.mu);
  const uint64_t key = g->id * 4096u + (uint64_t)(a_begin / MAX_ANGLES);
  if (t.key == key) {
    if (st != t.upload_stream && t.uploaded) TMB_CUDA_CHECK(cudaStreamWaitEvent(st, t.uploaded, 0));
    return TMB_OK;
  }
  // kernels of other streams may still read the table that is about to be replaced
  for (auto &u : t.users)
    if (u.first != st) TMB_CUDA_CHECK(cudaStreamWaitEvent(st, u.second, 0));
  // (the copies below are synchronous w.r.t. the host for pageable memory, so the staging arrays can be reused)
  for (int i = 0; i < a_count; ++i) {
    const float *tb = g->table + (size_t)(a_begin + i) * 8;
    hbp[i] = make_float4(tb[0], tb[1], tb[2], 0.f);
    hfp[i] = make_float4(tb[3], tb[4], tb[5], tb[7] == 0.f ? -tb[6] : tb[6]);
  }
  TMB_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_bp, hbp, sizeof(float4) * a_count, 0, cudaMemcpyHostToDevice, st));
  TMB_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_fp, hfp, sizeof(float4) * a_count, 0, cudaMemcpyHostToDevice, st));
  if (!t.uploaded) TMB_CUDA_CHECK(cudaEventCreateWithFlags(&t.uploaded, cudaEventDisableTiming));
  TMB_CUDA_CHECK(cudaEventRecord(t.uploaded, st));
  t.key = key;
  t.upload_stream = st;
  return TMB_OK;
}

// the kernels launched on `st` so far read the resident table: remember it for the next upload
int table_used(cudaStream_t st) {
  int dev = 0;
  TMB_CUDA_CHECK(cudaGetDevice(&dev));
  TableState &t = g_table[dev & 63];
  std::lock_guard<std::mutex> lock(t.mu);
  for (auto &u : t.users)
    if (u.first == st) return cudaEventRecord(u.second, st) == cudaSuccess ? TMB_OK : TMB_ERR_CUDA;
  if (t.users.size() >= 16) {  // more streams than anybody uses: fold the oldest into a device-wide wait
    TMB_CUDA_CHECK(cudaEventSynchronize(t.users.front().second));
    cudaEventDestroy(t.users.front().second);
    t.users.erase(t.users.begin());
  }
  cudaEvent_t ev;
  TMB_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  TMB_CUDA_CHECK(cudaEventRecord(ev, st));
  t.users.emplace_back(st, ev);
  return TMB_OK;
}

__device__ __forceinline__ void lerp_acc(float4 &acc, float g, float f, const float4 &s0, const float4 &s1) {
  acc.x = fmaf(g, s0.x, acc.x); acc.x = fmaf(f, s1.x, acc.x);
  acc.y = fmaf(g, s0.y, acc.y); acc.y = fmaf(f, s1.y, acc.y);
  acc.z = fmaf(g, s0.z, acc.z); acc.z = fmaf(f, s1.z, acc.z);
  acc.w = fmaf(g, s0.w, acc.w); acc.w = fmaf(f, s1.w, acc.w);
}

__device__ __forceinline__ float frac_weight(float f, int quant) {
  // CUDA texture units hold the interpolation fraction in 1.8 fixed point
  return quant ? rintf(f * 256.0f) * (1.0f / 256.0f) : f;
}

// ==========================================================================================
// layout conversion kernels (HBM-bound; a few ms next to seconds of projector work)
// ==========================================================================================
// sino[nz][na][nu] -> S_int[nzc][na][up] (float4 over 4 slices), interior only
__global__ void k_sino_to_int(const float *__restrict__ sino, float4 *__restrict__ sint, int nz, int na, int nu,
                              int up) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  const int a = blockIdx.y;
  const int zc = blockIdx.z;
  if (u >= nu) return;
  float v[ZC];
#pragma unroll
  for (int j = 0; j < ZC; ++j) {
    const int z = zc * ZC + j;
    v[j] = z < nz ? sino[((size_t)z * na + a) * nu + u] : 0.f;
  }
  sint[((size_t)zc * na + a) * up + SPAD + u] = make_float4(v[0], v[1], v[2], v[3]);
}

// vol[nz][n][n] -> V1[nzc][r][qp] (rows) and V0[nzc][c][qp] (columns, via a smem transpose)
__global__ void k_vol_to_int(const float *__restrict__ vol, float4 *__restrict__ v0, float4 *__restrict__ v1, int nz,
                             int n, int qp, int want0, int want1) {
  __shared__ float4 tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, zc = blockIdx.z;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    float v[ZC] = {0.f, 0.f, 0.f, 0.f};
    if (r < n && c < n) {
#pragma unroll
      for (int k = 0; k < ZC; ++k) {
        const int z = zc * ZC + k;
        if (z < nz) v[k] = vol[((size_t)z * n + r) * n + c];
      }
    }
    const float4 q = make_float4(v[0], v[1], v[2], v[3]);
    tile[j][tx] = q;
    if (want1 && r < n && c < n) v1[((size_t)zc * n + r) * qp + VPAD + c] = q;
  }
  if (!want0) return;
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (r < n && c < n) v0[((size_t)zc * n + c) * qp + VPAD + r] = tile[tx][j];
  }
}

// ==========================================================================================
// back-projection  (voxel-driven)
// ==========================================================================================
constexpr int BP_G = 8;       // angles per pipeline stage
constexpr int BP_STAGES = 3;
constexpr int BP_CONSUMERS = 256;
constexpr int BP_VPT = 4;     // voxels per thread (32x32 tile / 256 threads)

struct BpArgs {
  const float4 *sint;  // S_int of the angles being back-projected: [nzc][na_loc][up]
  float *vol;          // [nz][n][n]
  int n, nu, up, nz, na_loc;
  int a_first, a_stride;  // constant-table index of local angle j: a_first + j*a_stride
  int j_begin, j_count;   // local angle range handled by this launch
  int accumulate;         // 1: start from the values already in vol
  int quant;
};

__global__ void __launch_bounds__(BP_CONSUMERS + 32) k_bp(const BpArgs p) {
  __shared__ __align__(128) float4 buf[BP_STAGES][BP_G][NZC][BP_W];
  __shared__ int wst[BP_STAGES][BP_G];
  __shared__ __align__(8) uint64_t full_bar[BP_STAGES], empty_bar[BP_STAGES];

  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32, zc0 = blockIdx.z * NZC;
  const float half = 0.5f * (float)p.n;

  if (tid == 0) {
    for (int s = 0; s < BP_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], BP_CONSUMERS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int n_iter = (p.j_count + BP_G - 1) / BP_G;

  if (tid >= BP_CONSUMERS) {
    // ---------------- producer warp: one elected lane drives the TMA engine -----------------
    if (tid == BP_CONSUMERS) {
      const float xa = (float)x0 - half + 0.5f, xb = (float)(x0 + 31) - half + 0.5f;
      const float ya = (float)y0 - half + 0.5f, yb = (float)(y0 + 31) - half + 0.5f;
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % BP_STAGES;
        const uint32_t ph = (it / BP_STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int j0 = p.j_begin + it * BP_G;
        const int ng = min(BP_G, p.j_begin + p.j_count - j0);
        for (int ga = 0; ga < ng; ++ga) {
          const float4 t = c_bp[p.a_first + (j0 + ga) * p.a_stride];
          const float mn = t.z + fminf(xa * t.x, xb * t.x) + fminf(ya * t.y, yb * t.y);
          int ws = (int)floorf(mn) - 1;
          ws = max(-SPAD, min(ws, p.nu));
          wst[s][ga] = ws;
        }
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(ng * NZC * BP_W * sizeof(float4)));
        for (int ga = 0; ga < ng; ++ga) {
          const int ws = wst[s][ga];
#pragma unroll
          for (int c = 0; c < NZC; ++c) {
            const float4 *src = p.sint + ((size_t)(zc0 + c) * p.na_loc + (j0 + ga)) * p.up + (SPAD + ws);
            bulk_g2s(&buf[s][ga][c][0], src, BP_W * sizeof(float4), &full_bar[s]);
          }
        }
      }
    }
    return;
  }

  // ---------------- consumers: 8 warps, each lane owns 4 voxels x 8 slices -------------------
  const int warp = tid >> 5, lane = tid & 31;
  float vx[BP_VPT], vy[BP_VPT];
  int ix[BP_VPT], iy[BP_VPT];
  float4 acc[BP_VPT][NZC];
#pragma unroll
  for (int j = 0; j < BP_VPT; ++j) {
    const int patch = warp * BP_VPT + j;  // 32 patches of 8(x) x 4(y) voxels
    ix[j] = x0 + (patch & 3) * 8 + (lane & 7);
    iy[j] = y0 + (patch >> 2) * 4 + (lane >> 3);
    vx[j] = (float)ix[j] - half + 0.5f;
    vy[j] = (float)iy[j] - half + 0.5f;
#pragma unroll
    for (int c = 0; c < NZC; ++c) acc[j][c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (p.accumulate) {
#pragma unroll
    for (int j = 0; j < BP_VPT; ++j) {
      if (ix[j] < p.n && iy[j] < p.n) {
#pragma unroll
        for (int c = 0; c < NZC; ++c) {
          float *a4 = reinterpret_cast<float *>(&acc[j][c]);
#pragma unroll
          for (int k = 0; k < ZC; ++k) {
            const int z = (zc0 + c) * ZC + k;
            if (z < p.nz) a4[k] = p.vol[((size_t)z * p.n + iy[j]) * p.n + ix[j]];
          }
        }
      }
    }
  }

  for (int it = 0; it < n_iter; ++it) {
    const int s = it % BP_STAGES;
    const uint32_t ph = (it / BP_STAGES) & 1;
    mbar_wait(&full_bar[s], ph);
    const int j0 = p.j_begin + it * BP_G;
    const int ng = min(BP_G, p.j_begin + p.j_count - j0);
    for (int ga = 0; ga < ng; ++ga) {
      const float4 t = c_bp[p.a_first + (j0 + ga) * p.a_stride];
      const int ws = wst[s][ga];
#pragma unroll
      for (int j = 0; j < BP_VPT; ++j) {
        const float u = fmaf(vx[j], t.x, fmaf(vy[j], t.y, t.z));
        const float fl = floorf(u);
        const float f = frac_weight(u - fl, p.quant);
        const float g = 1.0f - f;
        int i = (int)fl - ws;
        i = max(0, min(i, BP_W - 2));
#pragma unroll
        for (int c = 0; c < NZC; ++c) {
          const float4 s0 = buf[s][ga][c][i];
          const float4 s1 = buf[s][ga][c][i + 1];
          lerp_acc(acc[j][c], g, f, s0, s1);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }

  // epilogue: each warp stores 4 rows x 8 columns (one 32-B sector per row) per slice
#pragma unroll
  for (int j = 0; j < BP_VPT; ++j) {
    if (ix[j] < p.n && iy[j] < p.n) {
#pragma unroll
      for (int c = 0; c < NZC; ++c) {
        const float *a4 = reinterpret_cast<const float *>(&acc[j][c]);
#pragma unroll
        for (int k = 0; k < ZC; ++k) {
          const int z = (zc0 + c) * ZC + k;
          if (z < p.nz) p.vol[((size_t)z * p.n + iy[j]) * p.n + ix[j]] = a4[k];
        }
      }
    }
  }
}

// ==========================================================================================
// forward projection  (Joseph)
// ==========================================================================================
constexpr int FP_G = 4;       // volume lines per pipeline stage
constexpr int FP_STAGES = 3;

struct FpArgs {
  const float4 *v0;  // [nzc][n][qp]  lines = columns
  const float4 *v1;  // [nzc][n][qp]  lines = rows
  float *sino;       // mode 0: API layout [nz][na_loc][nu]
  float4 *sint;      // mode 1: residual written straight into S_int [nzc][na_loc][up]
  const float *b;    // mode 1: full data sinogram [nz][na_tot][nu]
  const float *w;    // mode 1: PWLS weights (same layout) or nullptr
  int n, nu, up, qp, nz, na_loc, na_tot;
  int nzc_alloc;          // z-chunks S_int holds (k_fpq's 32-slice groups may reach past it)
  int zg_first;           // k_fpq: z-group of blockIdx.z == 0
  // k_fpq line segments (L2 blocking): blockIdx.z = z-group * nseg + segment; a CTA marches
  // seg_len lines and, when nseg > 1, leaves its raw partial sums in part[seg][zc][angle][nu_pad]
  int seg_len, nseg, nu_pad, jc;
  float4 *part;
  int a_first, a_stride;  // constant-table slot of local angle j
  int g_first, g_stride;  // global angle index (row of b / w) of local angle j
  int j_begin;            // local angle of blockIdx.y == 0
  int mode;               // 0 plain projection, 1 fused residual epilogue
  int fidelity;
  int quant;
};

__device__ __forceinline__ void fp_epilogue(const FpArgs &p, const float4 &acc, float scale, int j, int k, int zc);

__global__ void __launch_bounds__(FP_K + 32) k_fp(const FpArgs p) {
  extern __shared__ __align__(128) unsigned char fp_smem[];
  float4(*buf)[FP_G][NZC][FP_W] = reinterpret_cast<float4(*)[FP_G][NZC][FP_W]>(fp_smem);
  __shared__ int wst[FP_STAGES][FP_G];
  __shared__ __align__(8) uint64_t full_bar[FP_STAGES], empty_bar[FP_STAGES];

  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * FP_K;
  const int j = p.j_begin + blockIdx.y;  // local angle
  const int zc0 = blockIdx.z * NZC;
  const float4 t = c_fp[p.a_first + j * p.a_stride];
  const float alpha = t.x, b0 = t.y, bstep = t.z;
  const float scale = fabsf(t.w);
  const float4 *vsrc = (t.w < 0.f) ? p.v0 : p.v1;
  const float half = 0.5f * (float)p.n;

  if (tid == 0) {
    for (int s = 0; s < FP_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], FP_K / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int n_iter = (p.n + FP_G - 1) / FP_G;
  // elements of a volume line the CTA's FP_K bins can touch at this angle: the bins advance by
  // |bstep| in [1, sqrt 2] per bin, so only 45-degree rays need the whole FP_W window
  const int win = min(FP_W, (int)ceilf((float)(FP_K - 1) * fabsf(bstep)) + 4);

  if (tid >= FP_K) {
    if (tid == FP_K) {
      const float beta_a = fmaf((float)k0, bstep, b0);
      const float beta_b = fmaf((float)(k0 + FP_K - 1), bstep, b0);
      const float beta_min = fminf(beta_a, beta_b);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % FP_STAGES;
        const uint32_t ph = (it / FP_STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int m0 = it * FP_G;
        const int ng = min(FP_G, p.n - m0);
        for (int gm = 0; gm < ng; ++gm) {
          const float xm = (float)(m0 + gm) - half + 0.5f;
          int ws = (int)floorf(fmaf(alpha, xm, beta_min)) - 1;
          ws = max(-VPAD, min(ws, p.n));
          wst[s][gm] = ws;
        }
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(ng * NZC * win * sizeof(float4)));
        for (int gm = 0; gm < ng; ++gm) {
          const int ws = wst[s][gm];
#pragma unroll
          for (int c = 0; c < NZC; ++c) {
            const float4 *src = vsrc + ((size_t)(zc0 + c) * p.n + (m0 + gm)) * p.qp + (VPAD + ws);
            bulk_g2s(&buf[s][gm][c][0], src, win * sizeof(float4), &full_bar[s]);
          }
        }
      }
    }
    return;
  }

  const int lane = tid & 31;
  const int k = k0 + tid;
  const float beta = fmaf((float)k, bstep, b0);
  float4 acc[NZC];
#pragma unroll
  for (int c = 0; c < NZC; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int it = 0; it < n_iter; ++it) {
    const int s = it % FP_STAGES;
    const uint32_t ph = (it / FP_STAGES) & 1;
    mbar_wait(&full_bar[s], ph);
    const int m0 = it * FP_G;
    const int ng = min(FP_G, p.n - m0);
    for (int gm = 0; gm < ng; ++gm) {
      const float xm = (float)(m0 + gm) - half + 0.5f;
      const float rho = fmaf(alpha, xm, beta);
      const float fl = floorf(rho);
      const float f = frac_weight(rho - fl, p.quant);
      const float g = 1.0f - f;
      int i = (int)fl - wst[s][gm];
      i = max(0, min(i, win - 2));
#pragma unroll
      for (int c = 0; c < NZC; ++c) {
        const float4 s0 = buf[s][gm][c][i];
        const float4 s1 = buf[s][gm][c][i + 1];
        lerp_acc(acc[c], g, f, s0, s1);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }

  if (k >= p.nu) return;
#pragma unroll
  for (int c = 0; c < NZC; ++c) fp_epilogue(p, acc[c], scale, j, k, zc0 + c);
}

// Epilogue of the forward projectors for the 4 slices of z-chunk zc of (local angle j, bin k):
// plain projection into the API layout, or the fused residual (data_fidelities.py:28-39) written
// directly in the back-projector's layout.  Explicit roundings: identical to the unfused sequence
// FP -> subtract -> weight.
__device__ __forceinline__ void fp_epilogue(const FpArgs &p, const float4 &acc, float scale, int j, int k, int zc) {
  const float *a4 = reinterpret_cast<const float *>(&acc);
  if (p.mode == 0) {
#pragma unroll
    for (int e = 0; e < ZC; ++e) {
      const int z = zc * ZC + e;
      if (z < p.nz) p.sino[((size_t)z * p.na_loc + j) * p.nu + k] = __fmul_rn(a4[e], scale);
    }
  } else if (zc < p.nzc_alloc) {
    const int ga = p.g_first + j * p.g_stride;
    float r[ZC];
#pragma unroll
    for (int e = 0; e < ZC; ++e) {
      const int z = zc * ZC + e;
      float v = 0.f;
      if (z < p.nz) {
        const size_t idx = ((size_t)z * p.na_tot + ga) * p.nu + k;
        const float ax = __fmul_rn(a4[e], scale);
        if (p.fidelity == TMB_FID_KL) {
          v = __fsub_rn(1.0f, __fdiv_rn(p.b[idx], fmaxf(ax, 1e-8f)));
        } else {
          v = __fsub_rn(ax, p.b[idx]);
          if (p.w != nullptr) v = __fmul_rn(v, p.w[idx]);
        }
      }
      r[e] = v;
    }
    p.sint[((size_t)zc * p.na_loc + j) * p.up + SPAD + k] = make_float4(r[0], r[1], r[2], r[3]);
  }
}

// ==========================================================================================
// forward projection on the Q layouts (stacks of >= FQ_MIN_NZ slices)
//
// Same Joseph march as k_fp, re-mapped so that the shared-memory reads are bank-conflict free:
// a thread owns ONE detector bin and ONE z-chunk (4 slices); the 8 lanes of a quarter-warp are
// the 8 z-chunks of one bin, and the Q layouts keep those 8 chunks next to each other (128 B per
// in-plane position), so every LDS.128 wavefront is one contiguous 128-byte row.  (In k_fp a
// quarter-warp is 8 consecutive bins whose positions span up to 11 elements at 45 degrees: 32 % of
// its wavefronts are bank conflicts.)  CTA = FQ_K bins x 32 slices = 512 consumer threads + one
// producer warp that streams FQ_G lines per stage with one bulk copy (TMA) per line.
// ==========================================================================================
constexpr int FQ_G = 3;       // volume lines per pipeline stage (one z-group per CTA)
constexpr int FQ_G2 = 2;      // ... with two z-groups per CTA (twice the bytes per line)
constexpr int FQ_STAGES = 3;
constexpr int FQ_THREADS = FQ_K * FQ_CG;

// vol[nz][n][n] -> VQ1 / VQ0.  A block converts a 32 (columns) x 8 (rows) in-plane tile of one whole
// z-group (32 slices): reads are 128-byte rows of the volume, and each position of the Q layouts --
// 8 chunks x 16 B = one 128-byte line -- is written by 8 consecutive lanes, so every store
// instruction of a warp fills four whole lines.
__global__ void __launch_bounds__(256) k_vol_to_intq(const float *__restrict__ vol, float4 *__restrict__ v0,
                                                      float4 *__restrict__ v1, int nz, int n, int qpq) {
  // [slice][row][column]; the slice stride is odd so that the 8 chunks x 4 positions a warp writes at
  // a time read 32 different banks
  constexpr int TS = 8 * 33 + 1;
  __shared__ float tile[FQ_CG * ZC * TS];
  auto T = [&](int zz, int rr, int cl) -> float & { return tile[zz * TS + rr * 33 + cl]; };
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 8, zg = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;  // 8 warps
  // load: warp w takes row r0 + w of every slice of the group
  for (int zz = 0; zz < FQ_CG * ZC; ++zz) {
    const int z = zg * FQ_CG * ZC + zz, r = r0 + w, c = c0 + lane;
    T(zz, w, lane) = (z < nz && r < n && c < n) ? vol[((size_t)z * n + r) * n + c] : 0.f;
  }
  __syncthreads();
  const int cc = tid & (FQ_CG - 1);  // chunk written by this thread
  // VQ1[zg][r][QPAD + c][cc]: positions run along the columns
  for (int i = tid >> 3; i < 32 * 8; i += 32) {  // i = row * 32 + column
    const int rr = i >> 5, cl = i & 31;
    const int r = r0 + rr, c = c0 + cl;
    if (r < n && c < n)
      v1[(((size_t)zg * n + r) * qpq + QPAD + c) * FQ_CG + cc] =
          make_float4(T(cc * ZC, rr, cl), T(cc * ZC + 1, rr, cl), T(cc * ZC + 2, rr, cl), T(cc * ZC + 3, rr, cl));
  }
  // VQ0[zg][c][QPAD + r][cc]: positions run along the rows (8 consecutive rows = 1 KB per column)
  for (int i = tid >> 3; i < 32 * 8; i += 32) {  // i = column * 8 + row
    const int cl = i >> 3, rr = i & 7;
    const int r = r0 + rr, c = c0 + cl;
    if (r < n && c < n)
      v0[(((size_t)zg * n + c) * qpq + QPAD + r) * FQ_CG + cc] =
          make_float4(T(cc * ZC, rr, cl), T(cc * ZC + 1, rr, cl), T(cc * ZC + 2, rr, cl), T(cc * ZC + 3, rr, cl));
  }
}

// NG = z-groups (of 32 slices) per CTA.  With NG = 2 a thread carries two accumulators (8 slices)
// and its index / weight arithmetic is amortised over twice the updates (the kernel is otherwise
// issue-bound: ncu, profiles/); NG = 1 serves stacks of at most 32 slices and an odd last group.
template <int NG>
__global__ void __launch_bounds__(FQ_THREADS + 32, NG == 1 ? 2 : 1) k_fpq(const FpArgs p) {
  constexpr int G = NG == 1 ? FQ_G : FQ_G2;  // volume lines per stage
  extern __shared__ __align__(128) unsigned char fp_smem[];
  // buf[stage][line][group][position][chunk]
  float4(*buf)[G][NG][FQ_W][FQ_CG] = reinterpret_cast<float4(*)[G][NG][FQ_W][FQ_CG]>(fp_smem);
  __shared__ int wst[FQ_STAGES][G];
  __shared__ __align__(8) uint64_t full_bar[FQ_STAGES], empty_bar[FQ_STAGES];

  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * FQ_K;
  const int j = p.j_begin + blockIdx.y;  // local angle
  const int seg = (int)blockIdx.z % p.nseg;
  const int zg0 = p.zg_first + ((int)blockIdx.z / p.nseg) * NG;
  const int m_lo = seg * p.seg_len, m_hi = min(p.n, m_lo + p.seg_len);  // this CTA's volume lines
  const float4 t = c_fp[p.a_first + j * p.a_stride];
  const float alpha = t.x, b0 = t.y, bstep = t.z;
  const float scale = fabsf(t.w);
  const float4 *vsrc = (t.w < 0.f) ? p.v0 : p.v1;
  const float half = 0.5f * (float)p.n;

  if (tid == 0) {
    for (int s = 0; s < FQ_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], FQ_THREADS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int n_iter = (m_hi - m_lo + G - 1) / G;
  const int win = min(FQ_W, (int)ceilf((float)(FQ_K - 1) * fabsf(bstep)) + 4);

  if (tid >= FQ_THREADS) {
    if (tid == FQ_THREADS) {
      const float beta_a = fmaf((float)k0, bstep, b0);
      const float beta_b = fmaf((float)(k0 + FQ_K - 1), bstep, b0);
      const float beta_min = fminf(beta_a, beta_b);
      const uint32_t line_bytes = (uint32_t)(win * FQ_CG * sizeof(float4));
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % FQ_STAGES;
        const uint32_t ph = (it / FQ_STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int m0 = m_lo + it * G;
        const int ng = min(G, m_hi - m0);
        for (int gm = 0; gm < ng; ++gm) {
          const float xm = (float)(m0 + gm) - half + 0.5f;
          int ws = (int)floorf(fmaf(alpha, xm, beta_min)) - 1;
          ws = max(-QPAD, min(ws, p.n));
          wst[s][gm] = ws;
        }
        // the arrive releases the window starts to the consumers that acquire the completed phase
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(ng * NG) * line_bytes);
        for (int gm = 0; gm < ng; ++gm) {
#pragma unroll
          for (int q = 0; q < NG; ++q) {
            const float4 *src =
                vsrc + (((size_t)(zg0 + q) * p.n + (m0 + gm)) * p.qp + (QPAD + wst[s][gm])) * FQ_CG;
            bulk_g2s(&buf[s][gm][q][0][0], src, line_bytes, &full_bar[s]);
          }
        }
      }
    }
    return;
  }

  const int lane = tid & 31;
  const int cc = tid & (FQ_CG - 1);  // z-chunk inside the group
  const int k = k0 + (tid >> 3);     // detector bin
  const float beta = fmaf((float)k, bstep, b0);
  float4 acc[NG];
#pragma unroll
  for (int q = 0; q < NG; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);

  const bool quant = p.quant != 0;
  for (int it = 0; it < n_iter; ++it) {
    const int s = it % FQ_STAGES;
    const uint32_t ph = (it / FQ_STAGES) & 1;
    mbar_wait(&full_bar[s], ph);
    const int m0 = m_lo + it * G;
    const int ng = min(G, m_hi - m0);
    // (float)(m0 + gm) - half + 0.5f: integers and halves below 2^24 are exact, so base + gm is identical
    const float xbase = (float)m0 - half + 0.5f;
    const float4 *sbuf = &buf[s][0][0][0][cc];
#pragma unroll
    for (int gm = 0; gm < G; ++gm) {
      if (gm < ng) {
        const float rho = fmaf(alpha, xbase + (float)gm, beta);
        // floor and round-to-nearest-even without the quarter-rate conversion pipe: one F2I, the
        // rest are adds (|rho| < 2^22; f * 256 in [0, 256])
        const int ifl = __float2int_rd(rho);
        float f = rho - (float)ifl;
        if (quant) f = ((f * 256.0f + 12582912.0f) - 12582912.0f) * (1.0f / 256.0f);
        const float g = 1.0f - f;
        const int i = max(0, min(ifl - wst[s][gm], win - 2));
#pragma unroll
        for (int q = 0; q < NG; ++q) {
          const float4 *e = sbuf + ((gm * NG + q) * FQ_W + i) * FQ_CG;
          lerp_acc(acc[q], g, f, e[0], e[FQ_CG]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }

#pragma unroll
  for (int q = 0; q < NG; ++q) {
    const int zc = (zg0 + q) * FQ_CG + cc;
    if (p.part != nullptr) {
      // raw partial sum of this line segment; k_fp_finish adds the segments up in a fixed order
      const int nzc8 = (int)(gridDim.z / p.nseg) * NG * FQ_CG + p.zg_first * FQ_CG;
      p.part[(((size_t)seg * nzc8 + zc) * p.jc + blockIdx.y) * p.nu_pad + k] = acc[q];
    } else if (k < p.nu) {
      fp_epilogue(p, acc[q], scale, j, k, zc);
    }
  }
}

// adds up the line-segment partial sums of k_fpq (fixed order: deterministic) and applies the epilogue
__global__ void k_fp_finish(const FpArgs p, int nzc8) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int jl = blockIdx.y, zc = blockIdx.z;
  if (k >= p.nu) return;
  const int j = p.j_begin + jl;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int sg = 0; sg < p.nseg; ++sg) {
    const float4 v = p.part[(((size_t)sg * nzc8 + zc) * p.jc + jl) * p.nu_pad + k];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const float scale = fabsf(c_fp[p.a_first + j * p.a_stride].w);
  fp_epilogue(p, acc, scale, j, k, zc);
}

// ==========================================================================================
// k_fpm: the Joseph march of k_fpq with NA neighbouring angles of the subset sharing ONE window.
//
// k_fpq is bound by the L2 -> shared-memory traffic (4.7 B per update: every volume sample a CTA stages
// serves one angle and only the ~1 / |bstep| bins whose rays pass next to it).  Here a CTA stages one
// (wider) window per volume line and marches up to NA angles through it, so a staged byte serves NA times
// as many updates.  The angles of an ordered subset are degrees apart and their rays diverge along the
// march, so the bins an angle contributes to a CTA are chosen PER ANGLE, such that all of them cross the
// centre line of the CTA's line segment inside the same interval of D = (FQ_K - 0.5) * min |bstep|
// positions ("tile" t of the group: r in [R0 + t D, R0 + (t + 1) D), r = position of a ray on the centre
// line); away from the centre line the window grows by |m - mc| * (max alpha - min alpha), which the short
// L2 segments keep small (host check: fpm_fits).  An angle's tiles partition its bins (same boundary
// expression on both sides), each with at most FQ_K of them.  Angles of a group that march along the other
// axis (the group straddles |sin| = |cos|) go through a second pass of the same CTA.
// Arithmetic per (angle, bin, line) and the order of the line / segment sums are those of k_fpq:
// the result is bit-identical (tests/test_gpu_projector.py).
// ==========================================================================================
constexpr int FM_W = 144;      // positions of a staged line (18 KB)
constexpr int FM_G = 2;        // volume lines per pipeline stage
constexpr int FM_STAGES = 3;
constexpr size_t FM_SMEM = sizeof(float4) * FM_STAGES * FM_G * FM_W * FQ_CG;

// what thread 0 works out per pass and every thread of the CTA then reads (keeps the consumers' registers
// for the accumulators: 544 threads x 2 CTAs per SM leave 56 each)
template <int NA> struct FmPass {
  float alpha[NA], b0[NA], bstep[NA];
  int k0[NA], cnt[NA];
  unsigned act;   // angles of the pass
  int total;      // bins of all angles in this tile (0: nothing to do)
  int dir0;       // the pass marches along columns (V0)
};

// one (line, angle) update of a consumer thread: index and weight exactly as k_fpq computes them (QUANT: ASTRA's
// 8-bit texture weights, rounded with the 1.5 * 2^23 trick instead of a conversion-pipe instruction).
// (Taking both from one conversion, q = rint(256 rho), saves two instructions but is not the same number for
// rho in (-1, 0), the one interval where rho - floor(rho) rounds -- measured: one bin off by 1/256 of a sample.)
template <bool QUANT>
__device__ __forceinline__ void fm_update(float4 &acc, float al, float be, float x, int wsl, int wl2,
                                          const float4 *line) {
  const float rho = fmaf(al, x, be);
  const int ifl = __float2int_rd(rho);
  float f = rho - (float)ifl;
  if constexpr (QUANT) f = ((f * 256.0f + 12582912.0f) - 12582912.0f) * (1.0f / 256.0f);
  const float g = 1.0f - f;
  const int i = max(0, min(ifl - wsl, wl2));
  const float4 *e = line + i * FQ_CG;
  lerp_acc(acc, g, f, e[0], e[FQ_CG]);
}

template <int NA, bool QUANT>
__global__ void __launch_bounds__(FQ_THREADS + 32, 2) k_fpm(const FpArgs p) {
  extern __shared__ __align__(128) unsigned char fp_smem[];
  // buf[stage][line][position][chunk]
  float4(*buf)[FM_G][FM_W][FQ_CG] = reinterpret_cast<float4(*)[FM_G][FM_W][FQ_CG]>(fp_smem);
  __shared__ int2 wsl_s[FM_STAGES][FM_G];  // per staged line: first position, positions - 2
  __shared__ __align__(8) uint64_t full_bar[FM_STAGES], empty_bar[FM_STAGES];
  __shared__ FmPass<NA> ps;

  const int tid = threadIdx.x;
  const int jg0 = (int)blockIdx.y * NA;      // first angle of the group, relative to j_begin
  const int seg = (int)blockIdx.z % p.nseg;
  const int zg = p.zg_first + (int)blockIdx.z / p.nseg;
  const int m_lo = seg * p.seg_len, m_hi = min(p.n, m_lo + p.seg_len);  // this CTA's volume lines
  const float half = 0.5f * (float)p.n;

  if (tid == 0) {
    for (int s = 0; s < FM_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], FQ_THREADS / 32);
    }
    mbar_fence_init();
  }

  const int lane = tid & 31;
  const int cc = tid & (FQ_CG - 1);  // z-chunk inside the group
  const int kk = tid >> 3;           // bin of the tile (consumers)
  const int n_iter = (m_hi - m_lo + FM_G - 1) / FM_G;
  int it_base = 0;  // pipeline iterations of the earlier pass (stages and phases carry on)

  for (int pass = 0; pass < 2; ++pass) {
    if (tid == 0) {
      const int nag = min(NA, p.jc - jg0);  // angles of the group
      const float xmc = (float)((m_lo + m_hi) >> 1) - half + 0.5f;  // centre line of the segment
      float4 t[NA];
#pragma unroll
      for (int a = 0; a < NA; ++a) t[a] = c_fp[p.a_first + (p.j_begin + jg0 + min(a, nag - 1)) * p.a_stride];
      // angles of this pass: those marching along the same axis as the first (pass 0) / the others (pass 1)
      unsigned act = 0;
#pragma unroll
      for (int a = 0; a < NA; ++a)
        if (a < nag && ((t[a].w < 0.f) == (t[0].w < 0.f)) == (pass == 0)) act |= 1u << a;
      // tile of the pass: rays that cross the centre line at r in [lo, hi)
      float R0 = 3.0e38f, bmin = 3.0e38f;
#pragma unroll
      for (int a = 0; a < NA; ++a)
        if (act & (1u << a)) {
          const float c = fmaf(t[a].x, xmc, t[a].y);
          R0 = fminf(R0, fminf(c, fmaf((float)(p.nu - 1), t[a].z, c)));
          bmin = fminf(bmin, fabsf(t[a].z));
        }
      R0 -= 0.5f;  // no ray sits on the first boundary (the quotients below round)
      const float D = ((float)FQ_K - 0.5f) * bmin;
      const float lo = fmaf((float)blockIdx.x, D, R0), hi = fmaf((float)(blockIdx.x + 1), D, R0);
      int total = 0;
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        int k0 = 0, cnt = 0;
        if (act & (1u << a)) {
          const float c = fmaf(t[a].x, xmc, t[a].y), b = t[a].z;
          float fa, fb;  // bins [fa, fb)
          if (b > 0.f) { fa = ceilf((lo - c) / b); fb = ceilf((hi - c) / b); }
          else { fa = floorf((c - hi) / -b) + 1.0f; fb = floorf((c - lo) / -b) + 1.0f; }
          k0 = (int)fminf(fmaxf(fa, 0.f), (float)p.nu);
          cnt = max((int)fminf(fmaxf(fb, 0.f), (float)p.nu) - k0, 0);
        }
        ps.alpha[a] = t[a].x; ps.b0[a] = t[a].y; ps.bstep[a] = t[a].z;
        ps.k0[a] = k0; ps.cnt[a] = cnt;
        if (cnt == 0) act &= ~(1u << a);
        total += cnt;
      }
      ps.act = act; ps.total = total;
      ps.dir0 = ((t[0].w < 0.f) == (pass == 0)) ? 1 : 0;
    }
    __syncthreads();  // (first pass: also publishes the mbarrier inits)
    const unsigned act = ps.act;
    if (ps.total > 0) {  // CTA-uniform
      if (tid >= FQ_THREADS) {
        if (tid == FQ_THREADS) {
          const float4 *vsrc = ps.dir0 ? p.v0 : p.v1;
          // per angle: beta of its first and last bin of the tile (the window of a line spans their positions)
          float al[NA], ba[NA], bb[NA];
#pragma unroll
          for (int a = 0; a < NA; ++a) {
            al[a] = ps.alpha[a];
            ba[a] = fmaf((float)ps.k0[a], ps.bstep[a], ps.b0[a]);
            bb[a] = fmaf((float)(ps.k0[a] + max(ps.cnt[a], 1) - 1), ps.bstep[a], ps.b0[a]);
          }
          for (int it = 0; it < n_iter; ++it) {
            const int gi = it_base + it;
            const int s = gi % FM_STAGES;
            const uint32_t ph = (gi / FM_STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            const int m0 = m_lo + it * FM_G;
            const int ng = min(FM_G, m_hi - m0);
            uint32_t bytes = 0;
            for (int gm = 0; gm < ng; ++gm) {
              const float xm = (float)(m0 + gm) - half + 0.5f;
              float wmin = 3.0e38f, wmax = -3.0e38f;
#pragma unroll
              for (int a = 0; a < NA; ++a)
                if (act & (1u << a)) {
                  const float ra = fmaf(al[a], xm, ba[a]), rb = fmaf(al[a], xm, bb[a]);
                  wmin = fminf(wmin, fminf(ra, rb));
                  wmax = fmaxf(wmax, fmaxf(ra, rb));
                }
              int ws = (int)floorf(fmaxf(wmin, -4.0e6f)) - 1;
              ws = max(-QPAD, min(ws, p.n));
              int wl = (int)floorf(fminf(wmax, 4.0e6f)) - ws + 2;  // floor(wmax) + 1 is the last sample read
              wl = max(2, min(wl, min(FM_W, p.n + QPAD - ws)));
              wsl_s[s][gm] = make_int2(ws, wl - 2);
              bytes += (uint32_t)(wl * FQ_CG * sizeof(float4));
            }
            // the arrive releases the window starts to the consumers that acquire the completed phase
            mbar_arrive_expect_tx(&full_bar[s], bytes);
            for (int gm = 0; gm < ng; ++gm) {
              const int2 w = wsl_s[s][gm];
              const float4 *src = vsrc + (((size_t)zg * p.n + (m0 + gm)) * p.qp + (QPAD + w.x)) * FQ_CG;
              bulk_g2s(&buf[s][gm][0][0], src, (uint32_t)((w.y + 2) * FQ_CG * sizeof(float4)), &full_bar[s]);
            }
          }
        }
      } else {
        float al[NA], beta[NA];
        float4 acc[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          al[a] = ps.alpha[a];
          beta[a] = fmaf((float)(ps.k0[a] + kk), ps.bstep[a], ps.b0[a]);
          acc[a] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const bool full = act == (1u << NA) - 1u;
        for (int it = 0; it < n_iter; ++it) {
          const int gi = it_base + it;
          const int s = gi % FM_STAGES;
          const uint32_t ph = (gi / FM_STAGES) & 1;
          if (!mbar_try_wait(&full_bar[s], ph)) mbar_wait(&full_bar[s], ph);
          const int m0 = m_lo + it * FM_G;
          const int ng = min(FM_G, m_hi - m0);
          // (float)(m0 + gm) - half + 0.5f: integers and halves below 2^24 are exact, so base + gm is identical
          const float xbase = (float)m0 - half + 0.5f;
          const float4 *sbuf = &buf[s][0][0][cc];
          if (full && ng == FM_G) {  // the common case as one basic block: all angles of the group, all lines
#pragma unroll
            for (int gm = 0; gm < FM_G; ++gm) {
              const int2 w = wsl_s[s][gm];
              const float x = xbase + (float)gm;
#pragma unroll
              for (int a = 0; a < NA; ++a) fm_update<QUANT>(acc[a], al[a], beta[a], x, w.x, w.y, sbuf + gm * FM_W * FQ_CG);
            }
          } else {
#pragma unroll
            for (int gm = 0; gm < FM_G; ++gm) {
              if (gm < ng) {
                const int2 w = wsl_s[s][gm];
                const float x = xbase + (float)gm;
#pragma unroll
                for (int a = 0; a < NA; ++a)
                  if (act & (1u << a)) fm_update<QUANT>(acc[a], al[a], beta[a], x, w.x, w.y, sbuf + gm * FM_W * FQ_CG);
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
        const int zc = zg * FQ_CG + cc;
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          if ((act & (1u << a)) && kk < ps.cnt[a]) {
            const int k = ps.k0[a] + kk, jl = jg0 + a;
            // raw partial sum of this line segment (k_fpm always runs segmented; k_fp_finish has the epilogue)
            const int nzc8 = (int)(gridDim.z / p.nseg) * FQ_CG + p.zg_first * FQ_CG;
            p.part[(((size_t)seg * nzc8 + zc) * p.jc + jl) * p.nu_pad + k] = acc[a];
          }
        }
      }
      it_base += n_iter;
    }
    __syncthreads();  // everybody is done with `ps` before thread 0 rewrites it
  }
}

// ==========================================================================================
// residual post-pass for the robust / ring-artefact data terms (extension, see DESIGN.md: the
// reference snapshot only keeps their call sites, Demos/methods_IR_legacy/DemoFISTA_artifacts2D.py:
// 197,307-309).  Runs on the residual the forward projector's epilogue left in S_int, one thread
// per (z-chunk, detector bin) column walking the subset's angles:
//   1. Group-Huber ring model: res += alpha * r_x[z][u];  vec[z][u] = sum over angles of res
//   2. Huber:                   res *= min(1, delta / |res|)
//   3. weights:  PWLS  res *= w        SWLS  res = w res - w * (sum_a w res) / (sum_a w + beta)
// ==========================================================================================
struct PostArgs {
  float4 *sint;        // [nzc][na_loc][up]
  const float *w;      // full weights [nz][na_tot][nu] or nullptr
  const float *rx;     // [nz][nu] or nullptr
  float *vec;          // [nz][nu] or nullptr
  int nz, nu, up, na_loc, na_tot, g_first, g_stride;
  int weight_mode;     // 0 none, 1 PWLS, 2 SWLS
  float alpha, delta, beta;
  float sigma2;        // Student's-t scale squared (0: off)
};

__device__ __forceinline__ float huber_w(float r, float delta) {
  const float a = fabsf(r);
  return a > delta ? __fmul_rn(r, __fdiv_rn(delta, a)) : r;
}

__global__ void k_resid_post(const PostArgs p) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  const int zc = blockIdx.y;
  if (u >= p.nu) return;
  float rx[ZC] = {0.f, 0.f, 0.f, 0.f};
  bool zin[ZC];
#pragma unroll
  for (int q = 0; q < ZC; ++q) {
    zin[q] = zc * ZC + q < p.nz;
    if (p.rx && zin[q]) rx[q] = __fmul_rn(p.alpha, p.rx[(size_t)(zc * ZC + q) * p.nu + u]);
  }
  float vec[ZC] = {0.f, 0.f, 0.f, 0.f}, s[ZC] = {0.f, 0.f, 0.f, 0.f}, sw[ZC] = {0.f, 0.f, 0.f, 0.f};
  float4 *col = p.sint + (size_t)zc * p.na_loc * p.up + SPAD + u;
  for (int j = 0; j < p.na_loc; ++j) {
    float4 r4 = col[(size_t)j * p.up];
    float *r = reinterpret_cast<float *>(&r4);
    const int ga = p.g_first + j * p.g_stride;
#pragma unroll
    for (int q = 0; q < ZC; ++q) {
      if (!zin[q]) continue;
      float v = r[q];
      if (p.rx) {
        v = __fadd_rn(v, rx[q]);
        vec[q] = __fadd_rn(vec[q], v);
      }
      if (p.delta > 0.f) v = huber_w(v, p.delta);
      if (p.sigma2 > 0.f) v = __fdiv_rn(__fmul_rn(2.f, v), __fadd_rn(p.sigma2, __fmul_rn(v, v)));  // Student's t
      if (p.weight_mode) {
        const float wv = p.w[((size_t)(zc * ZC + q) * p.na_tot + ga) * p.nu + u];
        v = __fmul_rn(v, wv);
        if (p.weight_mode == 2) {
          s[q] = __fadd_rn(s[q], v);
          sw[q] = __fadd_rn(sw[q], wv);
        }
      }
      r[q] = v;
    }
    col[(size_t)j * p.up] = r4;
  }
  if (p.vec) {
#pragma unroll
    for (int q = 0; q < ZC; ++q)
      if (zin[q]) p.vec[(size_t)(zc * ZC + q) * p.nu + u] = vec[q];
  }
  if (p.weight_mode == 2) {
    float c[ZC];
#pragma unroll
    for (int q = 0; q < ZC; ++q) c[q] = __fdiv_rn(s[q], __fadd_rn(sw[q], p.beta));
    for (int j = 0; j < p.na_loc; ++j) {
      float4 r4 = col[(size_t)j * p.up];
      float *r = reinterpret_cast<float *>(&r4);
      const int ga = p.g_first + j * p.g_stride;
#pragma unroll
      for (int q = 0; q < ZC; ++q)
        if (zin[q])
          r[q] = __fsub_rn(r[q], __fmul_rn(p.w[((size_t)(zc * ZC + q) * p.na_tot + ga) * p.nu + u], c[q]));
      col[(size_t)j * p.up] = r4;
    }
  }
}

// ==========================================================================================
// host-side launchers
// ==========================================================================================
// test hook (tmb_fp_set_kernel): 0 / 2 = k_fpq<1> with line segments, 3 = k_fpq<2> (two z-groups per
// CTA, no segments), 4 = k_fpq<1> without segments
// 5 / 6 / 7 = k_fpm (groups of at most 2 / 3 / 4 angles sharing a window) where its windows fit
int g_fpq_mode = 0;
static int g_fpm_default = 4;  // largest group of k_fpm when no hook is set (measured: profiles/fp_multi_angle_r02.txt)
static int subset_first(const tmb_geom *g, int subset) { return subset < 0 ? 0 : subset; }
static int subset_stride(const tmb_geom *g, int subset) { return subset < 0 ? 1 : g->os_number; }
int subset_size(const tmb_geom *g, int subset) {
  if (subset < 0) return g->d.na;
  return (g->d.na - subset + g->os_number - 1) / g->os_number;
}

static int launch_sino_to_int(const tmb_geom *g, const float *sino, float4 *sint, int na_loc, cudaStream_t st) {
  dim3 grid((g->d.nu + 255) / 256, na_loc, g->d.nzc);
  k_sino_to_int<<<grid, 256, 0, st>>>(sino, sint, g->d.nz, na_loc, g->d.nu, g->d.up);
  return check_launch("k_sino_to_int");
}

static int launch_vol_to_int(const tmb_geom *g, const float *vol, float4 *v0, float4 *v1, cudaStream_t st) {
  if (g->fp_q) {
    dim3 gridq((g->d.n + 31) / 32, (g->d.n + 7) / 8, g->d.nzg);
    k_vol_to_intq<<<gridq, 256, 0, st>>>(vol, v0, v1, g->d.nz, g->d.n, g->d.qpq);
    return check_launch("k_vol_to_intq");
  }
  dim3 grid((g->d.n + 31) / 32, (g->d.n + 31) / 32, g->d.nzc);
  k_vol_to_int<<<grid, dim3(32, 8), 0, st>>>(vol, v0, v1, g->d.nz, g->d.n, g->d.qp, 1, 1);
  return check_launch("k_vol_to_int");
}

// back-project S_int (na_loc local angles of `subset`) into vol
static int launch_bp(const tmb_geom *g, int subset, const float4 *sint, float *vol, cudaStream_t st) {
  const int na_loc = subset_size(g, subset);
  const int first = subset_first(g, subset), stride = subset_stride(g, subset);
  BpArgs a;
  a.sint = sint; a.vol = vol;
  a.n = g->d.n; a.nu = g->d.nu; a.up = g->d.up; a.nz = g->d.nz; a.na_loc = na_loc;
  a.quant = g->quant8;
  dim3 grid((g->d.n + 31) / 32, (g->d.n + 31) / 32, g->d.nzc / NZC);
  // the constant table holds MAX_ANGLES global angles per upload; chunk the angle loop on it
  int j = 0;
  while (j < na_loc) {
    const int gl = first + j * stride;                   // global angle of local j
    const int chunk_base = (gl / MAX_ANGLES) * MAX_ANGLES;
    int cnt = 0;
    while (j + cnt < na_loc && first + (j + cnt) * stride < chunk_base + MAX_ANGLES) ++cnt;
    int rc = ensure_table(g, chunk_base, min(MAX_ANGLES, g->d.na - chunk_base), st);
    if (rc) return rc;
    a.a_first = gl - chunk_base; a.a_stride = stride;
    a.j_begin = j; a.j_count = cnt; a.accumulate = (j > 0);
    // a_first + (j0+ga)*stride is evaluated with j0 relative to j_begin inside the kernel
    a.a_first -= j * stride;
    k_bp<<<grid, BP_CONSUMERS + 32, 0, st>>>(a);
    rc = check_launch("k_bp");
    if (rc) return rc;
    rc = table_used(st);
    if (rc) return rc;
    j += cnt;
  }
  return TMB_OK;
}

// k_fpm: does a group size of NA angles fit the staged window for local angles [jbeg, jbeg + jc) of the
// subset, and how many tiles does the widest group need?  Mirrors the kernel's tile geometry in double
// precision with a margin (a tile too many costs an empty CTA; a window too narrow would lose samples).
static bool fpm_fits(const tmb_geom *g, int first, int stride, int jbeg, int jc, int NA, int seg_len, int nseg,
                     int *tiles_out) {
  const int n = g->d.n, nu = g->d.nu;
  const double half = 0.5 * n;
  int tiles = 1;
  for (int g0 = 0; g0 < jc; g0 += NA) {
    const int nag = std::min(NA, jc - g0);
    const float *t0 = g->table + (size_t)(first + (jbeg + g0) * stride) * 8;
    for (int pass = 0; pass < 2; ++pass) {
      double amin = 1e30, amax = -1e30, bmin = 1e30;
      int nact = 0;
      for (int a = 0; a < nag; ++a) {
        const float *t = g->table + (size_t)(first + (jbeg + g0 + a) * stride) * 8;
        if (((t[7] == t0[7]) ? 0 : 1) != pass) continue;
        ++nact;
        amin = std::min(amin, (double)t[3]); amax = std::max(amax, (double)t[3]);
        bmin = std::min(bmin, std::fabs((double)t[5]));
      }
      if (!nact) continue;
      const double D = (FQ_K - 0.5) * bmin;
      for (int sg = 0; sg < nseg; ++sg) {
        const int m_lo = sg * seg_len, m_hi = std::min(n, m_lo + seg_len);
        const double xmc = (double)((m_lo + m_hi) >> 1) - half + 0.5;
        double rmin = 1e30, rmax = -1e30;
        for (int a = 0; a < nag; ++a) {
          const float *t = g->table + (size_t)(first + (jbeg + g0 + a) * stride) * 8;
          if (((t[7] == t0[7]) ? 0 : 1) != pass) continue;
          const double c = t[3] * xmc + t[4], e = c + (nu - 1) * (double)t[5];
          rmin = std::min(rmin, std::min(c, e)); rmax = std::max(rmax, std::max(c, e));
        }
        tiles = std::max(tiles, (int)std::floor((rmax - rmin + 0.5) / D) + 2);
        const double reach = std::max(xmc - (m_lo - half + 0.5), (m_hi - 1 - half + 0.5) - xmc);
        if (D + reach * (amax - amin) + 6.0 > FM_W) return false;
      }
    }
  }
  *tiles_out = tiles;
  return true;
}

// group size of k_fpm for these angles: the largest allowed one whose windows fit (0: k_fpq)
static int fpm_group(const tmb_geom *g, int first, int stride, int jbeg, int jc, int seg_len, int nseg, int *tiles) {
  // test hook (tmb_fp_set_kernel): 5 / 6 / 7 = k_fpm with groups of at most 2 / 3 / 4 angles, 2 = k_fpq only
  const int na_max = g_fpq_mode >= 5 ? g_fpq_mode - 3 : (g_fpq_mode == 0 ? g_fpm_default : 0);
  for (int na = na_max; na >= 2; --na)
    if (fpm_fits(g, first, stride, jbeg, jc, na, seg_len, nseg, tiles)) return na;
  return 0;
}

template <int NA>
static void launch_fpm(const FpArgs &a, int tiles, int jc, int gz, cudaStream_t st) {
  static PerDeviceOnce attr;
  if (attr.first()) {
    cudaFuncSetAttribute(k_fpm<NA, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FM_SMEM);
    cudaFuncSetAttribute(k_fpm<NA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FM_SMEM);
  }
  const dim3 grid(tiles, (jc + NA - 1) / NA, gz);
  if (a.quant) k_fpm<NA, true><<<grid, FQ_THREADS + 32, FM_SMEM, st>>>(a);
  else k_fpm<NA, false><<<grid, FQ_THREADS + 32, FM_SMEM, st>>>(a);
}

// k_fpq<1> or, where its windows fit, the multi-angle k_fpm for local angles [a.j_begin, + jc)
static void launch_fpq_or_fpm(const tmb_geom *g, FpArgs &a, int first, int stride, int jc, int gz, size_t smem_q1,
                              cudaStream_t st) {
  int tiles = 0;
  const int na = fpm_group(g, first, stride, a.j_begin, jc, a.seg_len, a.nseg, &tiles);
  if (na == 4) launch_fpm<4>(a, tiles, jc, gz, st);
  else if (na == 3) launch_fpm<3>(a, tiles, jc, gz, st);
  else if (na == 2) launch_fpm<2>(a, tiles, jc, gz, st);
  else k_fpq<1><<<dim3((g->d.nu + FQ_K - 1) / FQ_K, jc, gz), FQ_THREADS + 32, smem_q1, st>>>(a);
}

static int launch_fp(const tmb_geom *g, int subset, const float4 *v0, const float4 *v1, float *sino, float4 *sint,
                     const float *b, const float *w, int mode, int fidelity, cudaStream_t st) {
  const int na_loc = subset_size(g, subset);
  const int first = subset_first(g, subset), stride = subset_stride(g, subset);
  FpArgs a;
  a.v0 = v0; a.v1 = v1; a.sino = sino; a.sint = sint; a.b = b; a.w = w;
  a.n = g->d.n; a.nu = g->d.nu; a.up = g->d.up; a.qp = g->fp_q ? g->d.qpq : g->d.qp; a.nz = g->d.nz;
  a.na_loc = na_loc; a.na_tot = g->d.na; a.nzc_alloc = g->d.nzc;
  a.g_first = first; a.g_stride = stride;
  a.mode = mode; a.fidelity = fidelity; a.quant = g->quant8;
  const size_t smem = sizeof(float4) * FP_STAGES * FP_G * NZC * FP_W;
  const size_t smem_q1 = sizeof(float4) * FQ_STAGES * FQ_G * FQ_W * FQ_CG;
  const size_t smem_q2 = sizeof(float4) * FQ_STAGES * FQ_G2 * 2 * FQ_W * FQ_CG;
  static PerDeviceOnce attr_set;
  if (attr_set.first()) {
    TMB_CUDA_CHECK(cudaFuncSetAttribute(k_fp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TMB_CUDA_CHECK(cudaFuncSetAttribute(k_fpq<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q1));
    TMB_CUDA_CHECK(cudaFuncSetAttribute(k_fpq<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q2));
  }
  int j = 0;
  while (j < na_loc) {
    const int gl = first + j * stride;
    const int chunk_base = (gl / MAX_ANGLES) * MAX_ANGLES;
    int cnt = 0;
    while (j + cnt < na_loc && first + (j + cnt) * stride < chunk_base + MAX_ANGLES) ++cnt;
    int rc = ensure_table(g, chunk_base, min(MAX_ANGLES, g->d.na - chunk_base), st);
    if (rc) return rc;
    a.a_first = gl - chunk_base - j * stride; a.a_stride = stride;
    a.j_begin = j;
    // grid.y is limited to 65535: far above any angle count
    if (g->fp_q) {
      const int tiles = (g->d.nu + FQ_K - 1) / FQ_K;
      a.seg_len = g->seg_len; a.nseg = g->nseg; a.nu_pad = tiles * FQ_K;
      if (g_fpq_mode == 3) {
        // experiment: two z-groups per CTA (8 slices per thread), no line segments
        a.nseg = 1; a.seg_len = g->d.n; a.part = nullptr; a.jc = cnt;
        const int pairs = g->d.nzg / 2;
        if (pairs > 0) {
          a.zg_first = 0;
          k_fpq<2><<<dim3(tiles, cnt, pairs), FQ_THREADS + 32, smem_q2, st>>>(a);
        }
        if (g->d.nzg - 2 * pairs > 0) {
          a.zg_first = 2 * pairs;
          k_fpq<1><<<dim3(tiles, cnt, g->d.nzg - 2 * pairs), FQ_THREADS + 32, smem_q1, st>>>(a);
        }
      } else if (g->nseg == 1 || g_fpq_mode == 4) {
        a.nseg = 1; a.seg_len = g->d.n; a.part = nullptr; a.jc = cnt; a.zg_first = 0;
        k_fpq<1><<<dim3(tiles, cnt, g->d.nzg), FQ_THREADS + 32, smem_q1, st>>>(a);
      } else {
        // locality blocking: every (bin tile, angle) CTA of one (z-group, line segment) runs back to
        // back, so the CTAs in flight work on one ~190 MB segment instead of the whole multi-GB layout
        // (L2 hit rate 57 % -> 97 % at config-2 size); the angle chunk is what the partial-sum buffer holds
        const int nzc8 = g->d.nzg * FQ_CG;
        a.zg_first = 0;
        a.part = reinterpret_cast<float4 *>(const_cast<char *>(reinterpret_cast<const char *>(v0)) - g->off_v0 +
                                            g->off_part);
        for (int c0 = 0; c0 < cnt; c0 += g->part_angles) {
          const int jc = min(g->part_angles, cnt - c0);
          a.j_begin = j + c0; a.jc = jc;
          launch_fpq_or_fpm(g, a, first, stride, jc, g->d.nzg * g->nseg, smem_q1, st);
          k_fp_finish<<<dim3((g->d.nu + 127) / 128, jc, nzc8), 128, 0, st>>>(a, nzc8);
        }
        a.j_begin = j;
      }
    } else {
      dim3 grid((g->d.nu + FP_K - 1) / FP_K, cnt, g->d.nzc / NZC);
      k_fp<<<grid, FP_K + 32, smem, st>>>(a);
    }
    rc = check_launch("k_fp");
    if (rc) return rc;
    rc = table_used(st);
    if (rc) return rc;
    j += cnt;
  }
  return TMB_OK;
}

struct Ws {
  float4 *v0, *v1, *s;
};
static Ws carve(const tmb_geom *g, void *workspace) {
  char *base = static_cast<char *>(workspace);
  return Ws{reinterpret_cast<float4 *>(base + g->off_v0), reinterpret_cast<float4 *>(base + g->off_v1),
            reinterpret_cast<float4 *>(base + g->off_s)};
}

}  // namespace tmb

using namespace tmb;

extern "C" int tmb_fp3d(tmb_geom *g, int subset, const float *vol, float *sino, void *workspace, void *stream) {
  TMB_REQUIRE(g && vol && sino && workspace, "tmb_fp3d: null argument");
  TMB_REQUIRE(subset >= -1 && subset < g->os_number, "tmb_fp3d: subset out of range");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Ws ws = carve(g, workspace);
  int rc = launch_vol_to_int(g, vol, ws.v0, ws.v1, st);
  if (rc) return rc;
  return launch_fp(g, subset, ws.v0, ws.v1, sino, nullptr, nullptr, nullptr, 0, 0, st);
}

extern "C" int tmb_bp3d(tmb_geom *g, int subset, const float *sino, float *vol, void *workspace, void *stream) {
  TMB_REQUIRE(g && vol && sino && workspace, "tmb_bp3d: null argument");
  TMB_REQUIRE(subset >= -1 && subset < g->os_number, "tmb_bp3d: subset out of range");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Ws ws = carve(g, workspace);
  int rc = launch_sino_to_int(g, sino, ws.s, subset_size(g, subset), st);
  if (rc) return rc;
  return launch_bp(g, subset, ws.s, vol, st);
}

extern "C" int tmb_grad(tmb_geom *g, int subset, int fidelity, const float *x, const float *b, const float *w,
                        float *grad, void *workspace, void *stream) {
  TMB_REQUIRE(g && x && b && grad && workspace, "tmb_grad: null argument");
  TMB_REQUIRE(subset >= -1 && subset < g->os_number, "tmb_grad: subset out of range");
  TMB_REQUIRE(fidelity >= TMB_FID_LS && fidelity <= TMB_FID_KL, "tmb_grad: unknown fidelity");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Ws ws = carve(g, workspace);
  int rc = launch_vol_to_int(g, x, ws.v0, ws.v1, st);
  if (rc) return rc;
  rc = launch_fp(g, subset, ws.v0, ws.v1, nullptr, ws.s, b, fidelity == TMB_FID_PWLS ? w : nullptr, 1, fidelity, st);
  if (rc) return rc;
  return launch_bp(g, subset, ws.s, grad, st);
}

// Gradient of the robust / ring-artefact data terms (extension of tmb_grad; see k_resid_post).
extern "C" int tmb_grad_ext(tmb_geom *g, int subset, const float *x, const float *b, const float *w,
                            int weight_mode, float huber_delta, float studentst_sigma, const float *ring_rx,
                            float ring_alpha, float beta_swls, float *ring_vec, float *grad, void *workspace,
                            void *stream) {
  TMB_REQUIRE(g && x && b && grad && workspace, "tmb_grad_ext: null argument");
  TMB_REQUIRE(subset >= -1 && subset < g->os_number, "tmb_grad_ext: subset out of range");
  TMB_REQUIRE(weight_mode >= 0 && weight_mode <= 2, "tmb_grad_ext: weight_mode must be 0 (none), 1 (PWLS), 2 (SWLS)");
  TMB_REQUIRE(weight_mode == 0 || w, "tmb_grad_ext: weights missing");
  TMB_REQUIRE(!ring_rx == !ring_vec, "tmb_grad_ext: ring_rx and ring_vec go together");
  TMB_REQUIRE(!(huber_delta > 0.f && studentst_sigma > 0.f), "tmb_grad_ext: Huber and Student's-t exclude each other");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Ws ws = carve(g, workspace);
  int rc = launch_vol_to_int(g, x, ws.v0, ws.v1, st);
  if (rc) return rc;
  // plain residual A x - b into S_int, then the post-pass, then the back-projection
  rc = launch_fp(g, subset, ws.v0, ws.v1, nullptr, ws.s, b, nullptr, 1, TMB_FID_LS, st);
  if (rc) return rc;
  PostArgs a;
  a.sint = ws.s; a.w = weight_mode ? w : nullptr; a.rx = ring_rx; a.vec = ring_vec;
  a.nz = g->d.nz; a.nu = g->d.nu; a.up = g->d.up; a.na_loc = subset_size(g, subset); a.na_tot = g->d.na;
  a.g_first = subset < 0 ? 0 : subset; a.g_stride = subset < 0 ? 1 : g->os_number;
  a.weight_mode = weight_mode; a.alpha = ring_alpha; a.delta = huber_delta; a.beta = beta_swls;
  a.sigma2 = studentst_sigma > 0.f ? studentst_sigma * studentst_sigma : 0.f;
  dim3 grid((g->d.nu + 127) / 128, (g->d.nz + ZC - 1) / ZC);
  k_resid_post<<<grid, 128, 0, st>>>(a);
  rc = check_launch("k_resid_post");
  if (rc) return rc;
  return launch_bp(g, subset, ws.s, grad, st);
}

extern "C" int tmb_geom_fp_group(const tmb_geom *g, int subset) {
  if (!g || subset < -1 || subset >= g->os_number) return TMB_ERR_ARG;
  if (!g->fp_q || g->nseg <= 1) return 0;
  const int na_loc = subset_size(g, subset);
  const int first = subset < 0 ? 0 : subset, stride = subset < 0 ? 1 : g->os_number;
  int cnt = 0;  // first constant-table / partial-buffer chunk, as launch_fp cuts it
  const int chunk_base = (first / MAX_ANGLES) * MAX_ANGLES;
  while (cnt < na_loc && first + cnt * stride < chunk_base + MAX_ANGLES) ++cnt;
  int tiles = 0;
  return fpm_group(g, first, stride, 0, std::min(cnt, g->part_angles), g->seg_len, g->nseg, &tiles);
}

// Number of kernels the forward projection of `subset` launches (k_fp / k_fpq per constant-table and
// partial-buffer chunk, plus k_fp_finish when the march is segmented): for launch accounting.
extern "C" int tmb_geom_fp_launches(const tmb_geom *g, int subset) {
  if (!g || subset < -1 || subset >= g->os_number) return TMB_ERR_ARG;
  const int na_loc = subset_size(g, subset);
  const int first = subset < 0 ? 0 : subset, stride = subset < 0 ? 1 : g->os_number;
  int launches = 0, j = 0;
  while (j < na_loc) {
    const int chunk_base = ((first + j * stride) / MAX_ANGLES) * MAX_ANGLES;
    int cnt = 0;
    while (j + cnt < na_loc && first + (j + cnt) * stride < chunk_base + MAX_ANGLES) ++cnt;
    if (g->fp_q && g->nseg > 1) launches += 2 * ((cnt + g->part_angles - 1) / g->part_angles);
    else launches += 1;
    j += cnt;
  }
  return launches;
}
