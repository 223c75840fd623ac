// FOURIER_INV (USFFT gridding reconstruction) kernels for sm_100a.
//
// Replaces the default ("centre gather") path of RecToolsDIRCuPy.FOURIER_INV
// (methodsDIR_CuPy.py:152-447) whose kernels live in cuda_kernels/fft_us_kernels.cu:
//   r2c_c1dfftshift (:529-557)  -> k_fi_pack
//   c1dfftshift     (:559-586)  -> k_fi_scale_sign
//   gather_kernel_center_angle_based_prune (:193-319) + gather_kernel_center (:468-527)
//                               -> k_fi_gather (angle ranges found on the fly, no uint16 table)
//   c2dfftshift     (:588-609)  -> k_fi_sign2d
//   unpadding_mul_phi (:611-657)-> k_fi_unpad
// The FFTs themselves run in cuFFT (called by the host through torch.fft).
//
// k_fi_gather: one thread owns one point of the 2n x 2n Cartesian frequency grid and a chunk of
// FI_SC complex slices (the loads of a warp are coalesced along the radial sample index).  A polar line (projection) contributes to the point when it passes within
// r = sqrt(2)(m + 1/2)/(2n) of it, i.e. when its angle lies in phi +- asin(r/|p|) (mod pi); the
// thread finds those index ranges in the sorted angle list by binary search, then, per line,
// walks the chord inside the disc and accumulates Gaussian-weighted samples.  The Gaussian
// weight is computed once per sample and applied to every slice of the chunk (the reference
// launches one thread per (point, slice) and recomputes it for each slice).
#include "tmb_common.h"

namespace tmb {

constexpr float FI_PI = 3.14159265358979323846f;
constexpr int FI_SC = 4;  // complex slices per thread in the gather (measured: 4 -> 81 ms, 32 -> 148 ms at config 4:
                          // more slices per thread cost parallelism and cache locality across 32 planes)

__global__ void k_fi_pack(const float *__restrict__ in, float2 *__restrict__ out, int n, int nproj, int nz2) {
  const int tx = blockIdx.x * blockDim.x + threadIdx.x;
  const int ty = blockIdx.y * blockDim.y + threadIdx.y;
  const int tz = blockIdx.z;
  if (tx >= n || ty >= nproj || tz >= nz2) return;
  const size_t plane = (size_t)n * nproj;
  const size_t i = (size_t)ty * n + tx;
  const float sgn = (tx & 1) ? 1.f : -1.f;
  // slices 2t and 2t+1 become the real and imaginary part of complex slice t
  out[tz * plane + i] = make_float2(in[(2 * (size_t)tz) * plane + i] * sgn, in[(2 * (size_t)tz + 1) * plane + i] * sgn);
}

// the same from rows of pitch `row_pitch` floats inside slices of pitch `slice_pitch` floats: packs straight out of
// the oversampled filter output (crop + pack in one pass; the filtered projections are never materialised)
__global__ void k_fi_pack_rows(const float *__restrict__ in, size_t row_pitch, size_t slice_pitch,
                               float2 *__restrict__ out, int n, int nproj, int nz2) {
  const int tx = blockIdx.x * blockDim.x + threadIdx.x;
  const int ty = blockIdx.y * blockDim.y + threadIdx.y;
  const int tz = blockIdx.z;
  if (tx >= n || ty >= nproj || tz >= nz2) return;
  const float sgn = (tx & 1) ? 1.f : -1.f;
  const float *a = in + (2 * (size_t)tz) * slice_pitch + (size_t)ty * row_pitch + tx;
  out[((size_t)tz * nproj + ty) * n + tx] = make_float2(a[0] * sgn, a[slice_pitch] * sgn);
}

// crop + sign from COMPLEX rows (the slice pairs were filtered as complex rows): out[t][row][x] = in[t][row][off + x] * (-1)^(x+1)
__global__ void k_fi_crop_sign(const float2 *__restrict__ in, size_t row_pitch, float2 *__restrict__ out, int n,
                               size_t rows) {
  const int tx = blockIdx.x * blockDim.x + threadIdx.x;
  if (tx >= n) return;
  const float sgn = (tx & 1) ? 1.f : -1.f;
  for (size_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const float2 v = in[r * row_pitch + tx];
    out[r * (size_t)n + tx] = make_float2(v.x * sgn, v.y * sgn);
  }
}

__global__ void k_fi_scale_sign(float2 *__restrict__ d, float c, int n, size_t rows) {
  const int tx = blockIdx.x * blockDim.x + threadIdx.x;
  if (tx >= n) return;
  const float sgn = (tx & 1) ? 1.f : -1.f;
  for (size_t r = blockIdx.y; r < rows; r += gridDim.y) {
    float2 v = d[r * n + tx];
    if (c == 1.f) {
      v.x = v.x * sgn;
      v.y = v.y * sgn;
    } else {
      v.x = v.x * c * sgn;
      v.y = v.y * c * sgn;
    }
    d[r * n + tx] = v;
  }
}

// c1dfftshift (:559-586) out of place into the slice-PAIR layout of the gather: out[t][row][x] = (in[2t][row][x],
// in[2t + 1][row][x]) * c * (-1)^(x+1).  Same traffic as the in-place pass; the products are rounded as there.
__global__ void k_fi_scale_sign_pairs(const float2 *__restrict__ in, float4 *__restrict__ out, float c, int n,
                                      size_t rows_per_slice, size_t rows) {
  const int tx = blockIdx.x * blockDim.x + threadIdx.x;
  if (tx >= n) return;
  const float sgn = (tx & 1) ? 1.f : -1.f;
  for (size_t r = blockIdx.y; r < rows; r += gridDim.y) {  // r = t * rows_per_slice + row
    const size_t t = r / rows_per_slice, row = r - t * rows_per_slice;
    float2 a = in[((2 * t) * rows_per_slice + row) * n + tx], b = in[((2 * t + 1) * rows_per_slice + row) * n + tx];
    if (c == 1.f) {
      a.x = a.x * sgn; a.y = a.y * sgn; b.x = b.x * sgn; b.y = b.y * sgn;
    } else {
      a.x = a.x * c * sgn; a.y = a.y * c * sgn; b.x = b.x * c * sgn; b.y = b.y * c * sgn;
    }
    out[r * n + tx] = make_float4(a.x, a.y, b.x, b.y);
  }
}

__global__ void k_fi_sign2d(float2 *__restrict__ f, int n2, int nz2) {
  const int tx = blockIdx.x * blockDim.x + threadIdx.x;
  const int ty = blockIdx.y * blockDim.y + threadIdx.y;
  if (tx >= n2 || ty >= n2) return;
  if (((tx ^ ty) & 1) == 0) return;  // sign +1
  for (int z = blockIdx.z; z < nz2; z += gridDim.z) {
    float2 *p = f + (size_t)z * n2 * n2 + (size_t)ty * n2 + tx;
    float2 v = *p;
    v.x = -v.x;
    v.y = -v.y;
    *p = v;
  }
}

// __expf without its denormal-result handling (three instructions around MUFU.EX2): the same bits for every normal
// result, zero instead of a denormal below 2^-126 -- Gaussian weights that small are 1e-34 of the largest one
__device__ __forceinline__ float exp_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950216293334961f));
  return y;
}

__device__ __forceinline__ int lower_bound_f(const float *__restrict__ a, int n, float v) {
  int lo = 0, hi = n;  // first index with a[i] >= v
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int upper_bound_f(const float *__restrict__ a, int n, float v) {
  int lo = 0, hi = n;  // first index with a[i] > v
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// contribution of polar line `proj` to the grid point (fft_us_kernels.cu:379-466)
// FULL: all SC slices of the chunk exist (no per-slice predicate: 46 instead of 70 instructions per sample at SC = 8 --
// the predicated loads cost a compare and two register moves each)
// Z2: the polar samples are stored as slice PAIRS, g4[nz2 / 2][nproj][n] of (slice 2t, slice 2t + 1): one 128-bit load
// and one address per two slices (needs FULL and an even z0)
template <int SC = FI_SC, bool FULL = false, bool Z2 = false>
__device__ __forceinline__ void fi_line(const float2 *__restrict__ g, float theta, float2 (&acc)[SC], float px,
                                        float py, float radius_2, int proj, int z0, int nzc, float coeff0,
                                        float coeff1, int n, int nproj) {
  float st, ct;
  __sincosf(theta, &st, &ct);
  const float pr = 0.5f, pr2 = 0.25f;
  const float vx = pr * ct, vy = pr * st;
  const float dot = vx * px + vy * py;
  const float mx = dot * vx / pr2, my = dot * vy / pr2;
  const float d2 = (mx - px) * (mx - px) + (my - py) * (my - py);
  if (!(radius_2 >= d2)) return;
  const float reach = __fsqrt_rn(radius_2 - d2);
  int rmin, rmax;
  if (fabsf(vx) > fabsf(vy)) {
    rmin = n / 2 - 1 + (int)floorf((mx - reach * vx / pr) / (2.f * vx / n));
    rmax = n / 2 + 1 + (int)floorf((mx + reach * vx / pr) / (2.f * vx / n));
  } else {
    rmin = n / 2 - 1 + (int)floorf((my - reach * vy / pr) / (2.f * vy / n));
    rmax = n / 2 + 1 + (int)floorf((my + reach * vy / pr) / (2.f * vy / n));
  }
  if (rmin > rmax) { const int t = rmax; rmax = rmin; rmin = t; }
  rmin = min(max(rmin, 0), n - 1);
  rmax = min(max(rmax, 0), n - 1);
  const size_t plane = (size_t)n * nproj;
  const float2 *row = g + (size_t)proj * n + (size_t)z0 * plane;
  const float inv_n = 1.0f / (float)n;
  for (int ri = rmin; ri < rmax; ++ri) {  // exclusive upper bound, like the reference
    // (ri - n/2) / n as a product with 1/n: the same number for the power-of-two n of the padded detector, one
    // rounding (1e-7 relative in the Gaussian's argument) otherwise -- and ~10 instructions fewer per sample
    const float t = (float)(ri - n / 2) * inv_n;
    float x0 = t * ct;
    float y0 = t * st;
    // the reference's "if (x0 >= 0.5f) x0 = 0.5f - 1e-5": |t| <= 1/2 in steps of 1/n, so below n = 50000 no product
    // lies in (0.5 - 1e-5, 0.5) and the minimum is the same number
    x0 = fminf(x0, (float)(0.5f - 1e-5));
    y0 = fminf(y0, (float)(0.5f - 1e-5));
    const float w0 = px - x0, w1 = py - y0;
    const float w = coeff0 * exp_ftz(coeff1 * (w0 * w0 + w1 * w1));
    if constexpr (Z2) {
      const float4 *row4 = reinterpret_cast<const float4 *>(g) + (size_t)proj * n + (size_t)(z0 >> 1) * plane;
#pragma unroll
      for (int s = 0; s < SC / 2; ++s) {
        const float4 v = __ldg(row4 + (size_t)s * plane + ri);
        acc[2 * s].x += v.x * w;
        acc[2 * s].y += v.y * w;
        acc[2 * s + 1].x += v.z * w;
        acc[2 * s + 1].y += v.w * w;
      }
    } else {
#pragma unroll
      for (int s = 0; s < SC; ++s) {
        if (FULL || s < nzc) {
          const float2 v = __ldg(row + (size_t)s * plane + ri);
          acc[s].x += v.x * w;
          acc[s].y += v.y * w;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(128)
    k_fi_gather(const float2 *__restrict__ g, float2 *__restrict__ f, const float *__restrict__ theta,
                const float *__restrict__ sth, const int *__restrict__ sidx, int m, float mu, int n, int nproj,
                int nz2, int center_size) {
  const int n2 = 2 * n;
  // the centre square of the grid (gather_kernel_center, :468-527): center_size == 2n is the whole grid
  const int c0 = max(0, n - center_size / 2);
  const int lx = blockIdx.x * blockDim.x + threadIdx.x, ly = blockIdx.y * blockDim.y + threadIdx.y;
  const int tx = c0 + lx, ty = c0 + ly;
  const int z0 = blockIdx.z * FI_SC;
  if (lx >= center_size || ly >= center_size || tx >= n2 || ty >= n2) return;
  const int nzc = min(FI_SC, nz2 - z0);
  const float coeff0 = FI_PI / mu;
  const float coeff1 = -FI_PI * FI_PI / mu;
  const int fs2 = n2 * n2;
  const float radius_2 = 2.f * ((float)m + 0.5f) * ((float)m + 0.5f) / fs2;
  const float px = (float)(tx - n) / (float)n2, py = (float)(n - ty) / (float)n2;
  const float len2 = px * px + py * py;

  float2 acc[FI_SC];
#pragma unroll
  for (int s = 0; s < FI_SC; ++s) acc[s] = make_float2(0.f, 0.f);

  if (radius_2 >= len2) {
    for (int j = 0; j < nproj; ++j) {
      const int proj = __ldg(sidx + j);
      fi_line(g, __ldg(theta + proj), acc, px, py, radius_2, proj, z0, nzc, coeff0, coeff1, n, nproj);
    }
  } else {
    // angles whose line passes within the disc: phi +- delta (mod pi); a small slack keeps the
    // range a superset, fi_line applies the exact test
    const float len = __fsqrt_rn(len2);
    const float delta = asinf(fminf(1.f, __fsqrt_rn(radius_2) / len)) + 2e-3f;
    const float phi = atan2f(py, px);
    const float tmin = __ldg(sth), tmax = __ldg(sth + nproj - 1);
    const int kmin = (int)ceilf((tmin - phi - delta) / FI_PI);
    const int kmax = (int)floorf((tmax - phi + delta) / FI_PI);
    int done = 0;  // first sorted index not yet consumed (ranges may touch near the centre)
    for (int k = kmin; k <= kmax; ++k) {
      const float a = phi - delta + k * FI_PI, b = phi + delta + k * FI_PI;
      int lo = lower_bound_f(sth, nproj, a);
      const int hi = upper_bound_f(sth, nproj, b);
      lo = max(lo, done);
      for (int j = lo; j < hi; ++j) {
        const int proj = __ldg(sidx + j);
        fi_line(g, __ldg(theta + proj), acc, px, py, radius_2, proj, z0, nzc, coeff0, coeff1, n, nproj);
      }
      done = max(done, hi);
    }
  }
  // the (-1)^(x+y) of the centred inverse 2-D FFT (c2dfftshift, :588-609) is applied on the way out
  const float sg = ((tx ^ ty) & 1) ? -1.f : 1.f;
  const size_t o = (size_t)ty * n2 + tx;
#pragma unroll
  for (int s = 0; s < FI_SC; ++s)
    if (s < nzc) f[(size_t)(z0 + s) * fs2 + o] = make_float2(acc[s].x * sg, acc[s].y * sg);
}

// ------------------------------------------------------------------------------------------
// k_fi_gather_s: the same gather with the polar samples of a TILE of grid points staged in shared memory.
//
// k_fi_gather is bound by the memory transactions of its loads, not by the weight arithmetic (profiles/
// gather_z_interleaved_r02.txt): every polar sample is read by the ~40 grid points within the Gaussian's support,
// each time as its own L1 / L2 access.  Here the 128 threads of a CTA (a 32 x 4 tile of grid points, FI_SC complex
// slices) walk the angle range of the WHOLE tile in batches of FS_B lines; per line they copy the one contiguous
// run of FS_SEG samples that the tile's points can touch (tile centre projected onto the line +- half diagonal +-
// chord) into shared memory with coalesced loads, and every thread then evaluates the line exactly as fi_line
// does, reading the samples from shared memory (a sample outside the staged run -- never for the sizes the bound
// below is derived for -- is read from global memory, so the result does not depend on that bound).
// The tile's angle range is a superset of each of its points' ranges; the exact distance test of fi_line decides.
// ------------------------------------------------------------------------------------------
constexpr int FS_B = 32;    // lines per batch
constexpr int FS_SEG = 32;  // staged samples per line
constexpr int FS_TX = 16, FS_TY = 8;  // the tile of grid points (128 threads)

__device__ __forceinline__ void fi_line_s(const float2 *__restrict__ g, const float2 (*sm)[FS_SEG], int r0,
                                          float theta, float2 (&acc)[FI_SC], float px, float py, float radius_2,
                                          int proj, int z0, int nzc, float coeff0, float coeff1, int n, int nproj) {
  float st, ct;
  __sincosf(theta, &st, &ct);
  const float pr = 0.5f, pr2 = 0.25f;
  const float vx = pr * ct, vy = pr * st;
  const float dot = vx * px + vy * py;
  const float mx = dot * vx / pr2, my = dot * vy / pr2;
  const float d2 = (mx - px) * (mx - px) + (my - py) * (my - py);
  if (!(radius_2 >= d2)) return;
  const float reach = __fsqrt_rn(radius_2 - d2);
  int rmin, rmax;
  if (fabsf(vx) > fabsf(vy)) {
    rmin = n / 2 - 1 + (int)floorf((mx - reach * vx / pr) / (2.f * vx / n));
    rmax = n / 2 + 1 + (int)floorf((mx + reach * vx / pr) / (2.f * vx / n));
  } else {
    rmin = n / 2 - 1 + (int)floorf((my - reach * vy / pr) / (2.f * vy / n));
    rmax = n / 2 + 1 + (int)floorf((my + reach * vy / pr) / (2.f * vy / n));
  }
  if (rmin > rmax) { const int t = rmax; rmax = rmin; rmin = t; }
  rmin = min(max(rmin, 0), n - 1);
  rmax = min(max(rmax, 0), n - 1);
  const size_t plane = (size_t)n * nproj;
  const float2 *row = g + (size_t)proj * n + (size_t)z0 * plane;
  const float inv_n = 1.0f / (float)n;
  for (int ri = rmin; ri < rmax; ++ri) {  // exclusive upper bound, like the reference
    // (ri - n/2) / n as a product with 1/n: the same number for the power-of-two n of the padded detector, one
    // rounding (1e-7 relative in the Gaussian's argument) otherwise -- and ~10 instructions fewer per sample
    const float t = (float)(ri - n / 2) * inv_n;
    float x0 = t * ct;
    float y0 = t * st;
    // the reference's "if (x0 >= 0.5f) x0 = 0.5f - 1e-5": |t| <= 1/2 in steps of 1/n, so below n = 50000 no product
    // lies in (0.5 - 1e-5, 0.5) and the minimum is the same number
    x0 = fminf(x0, (float)(0.5f - 1e-5));
    y0 = fminf(y0, (float)(0.5f - 1e-5));
    const float w0 = px - x0, w1 = py - y0;
    const float w = coeff0 * exp_ftz(coeff1 * (w0 * w0 + w1 * w1));
    const unsigned k = (unsigned)(ri - r0);
    if (k < (unsigned)FS_SEG) {
#pragma unroll
      for (int s = 0; s < FI_SC; ++s) {
        const float2 v = sm[s][k];
        acc[s].x += v.x * w;
        acc[s].y += v.y * w;
      }
    } else {
#pragma unroll
      for (int s = 0; s < FI_SC; ++s) {
        if (s < nzc) {
          const float2 v = __ldg(row + (size_t)s * plane + ri);
          acc[s].x += v.x * w;
          acc[s].y += v.y * w;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(128)
    k_fi_gather_s(const float2 *__restrict__ g, float2 *__restrict__ f, const float *__restrict__ theta,
                  const float *__restrict__ sth, const int *__restrict__ sidx, int m, float mu, int n, int nproj,
                  int nz2, int center_size) {
  __shared__ float2 sm[FS_B][FI_SC][FS_SEG];  // 32 KB
  __shared__ int s_proj[FS_B], s_r0[FS_B];
  __shared__ float s_theta[FS_B];
  __shared__ float2 s_cs[FS_B];  // (cos, sin) of the line: the cheap rejection test below
  const int n2 = 2 * n;
  const int c0 = max(0, n - center_size / 2);
  const int tid = threadIdx.y * 32 + threadIdx.x;
  // the CTA's tile: FS_TX x FS_TY grid points (a warp = two rows of 16)
  const int lx = blockIdx.x * FS_TX + (tid & (FS_TX - 1)), ly = blockIdx.y * FS_TY + tid / FS_TX;
  const int tx = c0 + lx, ty = c0 + ly;
  const int z0 = blockIdx.z * FI_SC;
  const bool on = lx < center_size && ly < center_size && tx < n2 && ty < n2;
  const int nzc = min(FI_SC, nz2 - z0);
  const float coeff0 = FI_PI / mu;
  const float coeff1 = -FI_PI * FI_PI / mu;
  const int fs2 = n2 * n2;
  const float radius_2 = 2.f * ((float)m + 0.5f) * ((float)m + 0.5f) / fs2;
  const float radius = __fsqrt_rn(radius_2);
  const float px = (float)(tx - n) / (float)n2, py = (float)(n - ty) / (float)n2;

  float2 acc[FI_SC];
#pragma unroll
  for (int s = 0; s < FI_SC; ++s) acc[s] = make_float2(0.f, 0.f);

  // the tile: centre and half diagonal (normalised frequency units), the same for every thread of the CTA
  const float pcx = ((float)(c0 + (int)blockIdx.x * FS_TX - n) + 0.5f * (FS_TX - 1)) / (float)n2;
  const float pcy = ((float)(n - (c0 + (int)blockIdx.y * FS_TY)) - 0.5f * (FS_TY - 1)) / (float)n2;
  // half diagonal of the tile in cells, with a margin
  const float hd = (0.5f * sqrtf((float)((FS_TX - 1) * (FS_TX - 1) + (FS_TY - 1) * (FS_TY - 1))) + 0.6f) / (float)n2;
  const float lenc = __fsqrt_rn(pcx * pcx + pcy * pcy);
  const float reachc = radius + hd;  // a line farther than this from the tile centre touches none of its points
  const size_t plane = (size_t)n * nproj;

  // sorted-index ranges of the lines that may touch the tile (at most 3 pieces, in ascending order, deduplicated)
  int rlo[3], rhi[3], nr = 0;
  if (reachc >= lenc) {
    rlo[0] = 0; rhi[0] = nproj; nr = 1;
  } else {
    const float delta = asinf(fminf(1.f, reachc / lenc)) + 2e-3f;
    const float phi = atan2f(pcy, pcx);
    const float tmin = __ldg(sth), tmax = __ldg(sth + nproj - 1);
    const int kmin = (int)ceilf((tmin - phi - delta) / FI_PI);
    const int kmax = (int)floorf((tmax - phi + delta) / FI_PI);
    int done = 0;
    for (int k = kmin; k <= kmax && nr < 3; ++k) {
      const float a = phi - delta + k * FI_PI, b = phi + delta + k * FI_PI;
      int lo = lower_bound_f(sth, nproj, a);
      const int hi = upper_bound_f(sth, nproj, b);
      lo = max(lo, done);
      if (hi > lo) { rlo[nr] = lo; rhi[nr] = hi; ++nr; }
      done = max(done, hi);
    }
    if (kmax - kmin + 1 > 3) { rlo[0] = 0; rhi[0] = nproj; nr = 1; }  // (angles spanning more than 3 pi: take all)
  }

  for (int piece = 0; piece < nr; ++piece) {
    for (int j0 = rlo[piece]; j0 < rhi[piece]; j0 += FS_B) {
      const int nb = min(FS_B, rhi[piece] - j0);
      // ---- stage: thread b < nb sets up line b of the batch, then everybody copies
      if (tid < nb) {
        const int proj = __ldg(sidx + j0 + tid);
        const float th = __ldg(theta + proj);
        float st, ct;
        __sincosf(th, &st, &ct);
        const float tc = pcx * ct + pcy * st;  // tile centre along the line
        // first sample any point of the tile can ask for: (tc - hd - radius) n + n/2, minus the kernel's own margins
        int r0 = (int)floorf((tc - hd - radius) * (float)n) + n / 2 - 3;
        r0 = max(0, min(r0, n - FS_SEG));
        s_proj[tid] = proj; s_theta[tid] = th; s_r0[tid] = r0; s_cs[tid] = make_float2(ct, st);
      }
      __syncthreads();
      {
        const int k = tid & (FS_SEG - 1), s = tid >> 5;  // 128 threads = FS_SEG samples x FI_SC slices
#pragma unroll 8
        for (int b = 0; b < nb; ++b) {
          float2 v = make_float2(0.f, 0.f);
          const int ri = s_r0[b] + k;
          if (s < nzc && ri < n) v = __ldg(g + (size_t)(z0 + s) * plane + (size_t)s_proj[b] * n + ri);
          sm[b][s][k] = v;
        }
      }
      __syncthreads();
      // ---- accumulate
      if (on) {
        for (int b = 0; b < nb; ++b) {
          // cheap rejection with a margin (distance of the point from the line through the origin); fi_line_s
          // applies the reference's own test to what is left
          const float2 cs = s_cs[b];
          const float dq = py * cs.x - px * cs.y;
          if (dq * dq > radius_2 * 1.01f + 1e-12f) continue;
          fi_line_s(g, sm[b], s_r0[b], s_theta[b], acc, px, py, radius_2, s_proj[b], z0, nzc, coeff0, coeff1, n, nproj);
        }
      }
      __syncthreads();
    }
  }
  if (!on) return;
  // the (-1)^(x+y) of the centred inverse 2-D FFT (c2dfftshift, :588-609) is applied on the way out
  const float sg = ((tx ^ ty) & 1) ? -1.f : 1.f;
  const size_t o = (size_t)ty * n2 + tx;
#pragma unroll
  for (int s = 0; s < FI_SC; ++s)
    if (s < nzc) f[(size_t)(z0 + s) * fs2 + o] = make_float2(acc[s].x * sg, acc[s].y * sg);
}

// ------------------------------------------------------------------------------------------
// k_fi_gather_w: the gather with a WARP walking the polar lines of its patch of grid points in lock step.
//
// In k_fi_gather every thread walks its own list of lines, so at any instant the lanes of a warp read samples of
// different lines: 13.5 sectors per LDG.64 request, and the L1 data pipe (78 % of peak) bounds the kernel.  Here a warp
// owns a compact 8 x 4 patch of grid points and walks the angle range of the PATCH (a superset of each point's range,
// found once per warp); all lanes look at the same line at the same time, a lane that is too far from it (cheap test
// against the line's cos / sin, then fi_line's own exact test) sits the line out, and the lanes that take it read
// neighbouring samples of one polar row: a few sectors per request.  Same (point, line, sample) visits in the same
// order as k_fi_gather: bit-identical grids.
// ------------------------------------------------------------------------------------------
constexpr int FW_PX = 8, FW_PY = 4;  // the patch of a warp
#ifndef FW_MINB
#define FW_MINB 4  // CTAs per SM the register allocation aims at (5: 96 registers with spills, 56 instead of 50 ms at config 4)
#endif

template <int SC, bool FULL, bool Z2 = false>
__global__ void __launch_bounds__(128, FW_MINB)
    k_fi_gather_w(const float2 *__restrict__ g, float2 *__restrict__ f, const float *__restrict__ theta,
                  const float *__restrict__ sth, const int *__restrict__ sidx, int m, float mu, int n, int nproj,
                  int nz2, int center_size) {
  const int n2 = 2 * n;
  const int c0 = max(0, n - center_size / 2);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // CTA = 4 patches side by side: 32 x 4 grid points
#ifndef FW_LINEAR_ROWS  // (A/B builds: -DFW_LINEAR_ROWS, -DFW_NO_SKIP; profiles/fourier_chunks_r02.txt)
  // rows of patches are taken from the centre of the grid outwards: a point at distance |p| from the centre is crossed by
  // ~1/|p| of the lines (all of them at the centre), so the centre's CTAs run ~100 x longer than the average one and
  // must not be among the last to start (16 complex slices per launch, 8 per thread: 15.5 -> 13.9 ms)
  const int by = blockIdx.y;
  const int yrow = (int)(gridDim.y >> 1) + ((by & 1) ? -((by >> 1) + 1) : (by >> 1));
#else
  const int yrow = blockIdx.y;
#endif
  const int px0 = blockIdx.x * 32 + warp * FW_PX, py0 = yrow * FW_PY;
  const int lx = px0 + (lane & (FW_PX - 1)), ly = py0 + lane / FW_PX;
  const int tx = c0 + lx, ty = c0 + ly;
  const int z0 = blockIdx.z * SC;
  const bool on = lx < center_size && ly < center_size && tx < n2 && ty < n2;
  const int nzc = min(SC, nz2 - z0);
  const float coeff0 = FI_PI / mu;
  const float coeff1 = -FI_PI * FI_PI / mu;
  const int fs2 = n2 * n2;
  const float radius_2 = 2.f * ((float)m + 0.5f) * ((float)m + 0.5f) / fs2;
  const float radius = __fsqrt_rn(radius_2);
  const float px = (float)(tx - n) / (float)n2, py = (float)(n - ty) / (float)n2;

  float2 acc[SC];
#pragma unroll
  for (int s = 0; s < SC; ++s) acc[s] = make_float2(0.f, 0.f);

  // the patch: centre and half diagonal (normalised frequency units), the same for every lane of the warp
  const float pcx = ((float)(c0 + px0 - n) + 0.5f * (FW_PX - 1)) / (float)n2;
  const float pcy = ((float)(n - (c0 + py0)) - 0.5f * (FW_PY - 1)) / (float)n2;
  const float hd = (0.5f * sqrtf((float)((FW_PX - 1) * (FW_PX - 1) + (FW_PY - 1) * (FW_PY - 1))) + 0.6f) / (float)n2;
  const float lenc = __fsqrt_rn(pcx * pcx + pcy * pcy);
  const float reachc = radius + hd;  // a line farther than this from the patch centre touches none of its points

  // sorted-index ranges of the lines that may touch the patch (at most 3 pieces, ascending, deduplicated)
  int rlo0 = 0, rhi0 = 0, rlo1 = 0, rhi1 = 0, rlo2 = 0, rhi2 = 0;
  // The polar samples end at |p| = 1/2 and fi_line clamps its sample range to the line: a patch whose nearest point
  // lies farther out than the Gaussian's reach plus fi_line's two samples of slack (2/n; 4/n taken) receives nothing
  // -- the corners of the grid, 21 % of it, store their zeros without looking for a line.
#ifndef FW_NO_SKIP
  if (lenc - hd > 0.5f + radius + 4.f / (float)n) {
  } else
#endif
  if (reachc >= lenc) {
    rhi0 = nproj;
  } else {
    const float delta = asinf(fminf(1.f, reachc / lenc)) + 2e-3f;
    const float phi = atan2f(pcy, pcx);
    const float tmin = __ldg(sth), tmax = __ldg(sth + nproj - 1);
    const int kmin = (int)ceilf((tmin - phi - delta) / FI_PI);
    const int kmax = (int)floorf((tmax - phi + delta) / FI_PI);
    if (kmax - kmin + 1 > 3) {  // (angles spanning more than 3 pi: take all)
      rhi0 = nproj;
    } else {
      int done = 0, nr = 0;
      for (int k = kmin; k <= kmax; ++k) {
        const float a = phi - delta + k * FI_PI, b = phi + delta + k * FI_PI;
        int lo = lower_bound_f(sth, nproj, a);
        const int hi = upper_bound_f(sth, nproj, b);
        lo = max(lo, done);
        if (hi > lo) {
          if (nr == 0) { rlo0 = lo; rhi0 = hi; } else if (nr == 1) { rlo1 = lo; rhi1 = hi; } else { rlo2 = lo; rhi2 = hi; }
          ++nr;
        }
        done = max(done, hi);
      }
    }
  }
  if (on) {  // (a warp is either entirely inside the centre square's rows or its off lanes simply skip)
#pragma unroll 1
    for (int piece = 0; piece < 3; ++piece) {
      const int jlo = piece == 0 ? rlo0 : (piece == 1 ? rlo1 : rlo2), jhi = piece == 0 ? rhi0 : (piece == 1 ? rhi1 : rhi2);
      for (int j = jlo; j < jhi; ++j) {
        const int proj = __ldg(sidx + j);
        const float th = __ldg(theta + proj);
        float st, ct;
        __sincosf(th, &st, &ct);
        // cheap rejection with a margin (distance of the point from the line through the origin); fi_line applies
        // the reference's own test to what is left
        const float dq = py * ct - px * st;
        if (dq * dq > radius_2 * 1.01f + 1e-12f) continue;
        fi_line<SC, FULL, Z2>(g, th, acc, px, py, radius_2, proj, z0, nzc, coeff0, coeff1, n, nproj);
      }
    }
    // the (-1)^(x+y) of the centred inverse 2-D FFT (c2dfftshift, :588-609) is applied on the way out
    const float sg = ((tx ^ ty) & 1) ? -1.f : 1.f;
    const size_t o = (size_t)ty * n2 + tx;
#pragma unroll
    for (int s = 0; s < SC; ++s)
      if (FULL || s < nzc) f[(size_t)(z0 + s) * fs2 + o] = make_float2(acc[s].x * sg, acc[s].y * sg);
  }
}

// The scatter ("gather_kernel" / "gather_kernel_partial", fft_us_kernels.cu:44-109) of the non-default branches
// (methodsDIR_CuPy.py:761-779, 818-835: center_size < 192, or a centre square smaller than the grid): every polar
// sample spreads its (2m+1)^2 Gaussian footprint onto the grid with atomic adds; PARTIAL skips the targets inside
// the centre square, which the centre gather fills.  One thread owns one sample of FI_SC complex slices (one weight
// evaluation serves the chunk); the (-1)^(x+y) of the centred inverse 2-D FFT is applied to what is added, as in
// k_fi_gather's store.  The grid must be zero on entry where the scatter adds.
template <bool PARTIAL>
__global__ void __launch_bounds__(256)
    k_fi_scatter(const float2 *__restrict__ g, float2 *__restrict__ f, const float *__restrict__ theta, int m, float mu,
                 int center_size, int n, int nproj, int nz2) {
  const int tx = blockIdx.x * blockDim.x + threadIdx.x;
  const int ty = blockIdx.y * blockDim.y + threadIdx.y;
  const int z0 = blockIdx.z * FI_SC;
  if (tx >= n || ty >= nproj) return;
  const int nzc = min(FI_SC, nz2 - z0);
  const int ch = center_size / 2;
  const float coeff0 = FI_PI / mu;
  const float coeff1 = -FI_PI * FI_PI / mu;
  float st, ct;
  __sincosf(__ldg(theta + ty), &st, &ct);
  float x0 = (tx - n / 2) / (float)n * ct;
  float y0 = -(tx - n / 2) / (float)n * st;
  if (x0 >= 0.5f) x0 = 0.5f - 1e-5;
  if (y0 >= 0.5f) y0 = 0.5f - 1e-5;
  const int n2 = 2 * n;
  const size_t fs2 = (size_t)n2 * n2, plane = (size_t)n * nproj;
  float2 g0[FI_SC];
#pragma unroll
  for (int s = 0; s < FI_SC; ++s)
    g0[s] = s < nzc ? __ldg(g + (size_t)(z0 + s) * plane + (size_t)ty * n + tx) : make_float2(0.f, 0.f);
  const int bx = (int)floorf(2 * n * x0) - m, by = (int)floorf(2 * n * y0) - m;
  for (int i1 = 0; i1 < 2 * m + 1; ++i1) {
    const int ell1 = by + i1;
    for (int i0 = 0; i0 < 2 * m + 1; ++i0) {
      const int ell0 = bx + i0;
      if (PARTIAL && !(ell0 < -ch || ell0 >= ch || ell1 < -ch || ell1 >= ch)) continue;
      const float w0 = ell0 / (float)(2 * n) - x0, w1 = ell1 / (float)(2 * n) - y0;
      float w = coeff0 * __expf(coeff1 * (w0 * w0 + w1 * w1));
      // target (column, row) of the 2n x 2n grid: the reference offsets f by (n, n) and wraps (:20, :40)
      const int col = (ell0 + 3 * n) % n2, row = (ell1 + 3 * n) % n2;
      if ((col ^ row) & 1) w = -w;
      float2 *dst = f + (size_t)z0 * fs2 + (size_t)row * n2 + col;
#pragma unroll
      for (int s = 0; s < FI_SC; ++s) {
        if (s < nzc) {
          atomicAdd(&dst[(size_t)s * fs2].x, w * g0[s].x);
          atomicAdd(&dst[(size_t)s * fs2].y, w * g0[s].y);
        }
      }
    }
  }
}

// `scale` multiplies the grid values first: 1 when the inverse 2-D FFT was normalised, 1 / (2n)^2 when it was not (the
// product is rounded where the normalisation pass would have rounded it: same bits)
__global__ void k_fi_unpad(float *__restrict__ recon, const float2 *__restrict__ f, float mu, float scale, int nproj,
                           int up, int unpad_z, int um, int n, int nz2) {
  const int rxu = blockIdx.x * blockDim.x + threadIdx.x;
  const int ryu = blockIdx.y * blockDim.y + threadIdx.y;
  const int rz = blockIdx.z;
  const int rx = um + rxu, ry = um + ryu;
  if (rx >= up || ry >= up || rz >= nz2) return;
  const int n2 = 2 * n;
  const int rs = up - um;
  const size_t rs2 = (size_t)rs * rs;
  float2 v = f[(size_t)rz * n2 * n2 + (size_t)(n / 2 + ry) * n2 + (n / 2 + rx)];
  v.x = __fmul_rn(v.x, scale);
  v.y = __fmul_rn(v.y, scale);
  if (((n / 2 + ry) ^ (n / 2 + rx)) & 1) {  // the second c2dfftshift, fused
    v.x = -v.x;
    v.y = -v.y;
  }
  const float ddx = -0.5f + rx * 1.f / n;
  const float ddy = -0.5f + ry * 1.f / n;
  const float phi = expf(mu * (n * n) * (ddx * ddx + ddy * ddy)) * ((float)(1 - n % 4) / nproj);
  const size_t ri = (size_t)ryu * rs + rxu;
  // complex slice t carries output slices 2t (real) and 2t+1 (imaginary)
  recon[(size_t)rz * 2 * rs2 + ri] = v.x * phi;
  if (2 * rz + 1 < unpad_z) recon[((size_t)rz * 2 + 1) * rs2 + ri] = v.y * phi;
}

}  // namespace tmb

using namespace tmb;

extern "C" int tmb_fi_pack(const float *in, float *datac, int n, int nproj, int nz2, void *stream) {
  TMB_REQUIRE(in && datac && n > 0 && nproj > 0 && nz2 > 0, "tmb_fi_pack: bad argument");
  dim3 block(32, 8), grid((n + 31) / 32, (nproj + 7) / 8, nz2);
  k_fi_pack<<<grid, block, 0, (cudaStream_t)stream>>>(in, reinterpret_cast<float2 *>(datac), n, nproj, nz2);
  return check_launch("k_fi_pack");
}

extern "C" int tmb_fi_pack_rows(const float *in, size_t row_pitch, size_t slice_pitch, float *datac, int n, int nproj,
                                int nz2, void *stream) {
  TMB_REQUIRE(in && datac && n > 0 && nproj > 0 && nz2 > 0 && row_pitch >= (size_t)n && slice_pitch >= row_pitch * nproj,
              "tmb_fi_pack_rows: bad argument");
  dim3 block(32, 8), grid((n + 31) / 32, (nproj + 7) / 8, nz2);
  k_fi_pack_rows<<<grid, block, 0, (cudaStream_t)stream>>>(in, row_pitch, slice_pitch, reinterpret_cast<float2 *>(datac), n,
                                                           nproj, nz2);
  return check_launch("k_fi_pack_rows");
}

extern "C" int tmb_fi_crop_sign(const float *in, size_t row_pitch, float *datac, int n, size_t rows, void *stream) {
  TMB_REQUIRE(in && datac && n > 0 && rows > 0 && row_pitch >= (size_t)n, "tmb_fi_crop_sign: bad argument");
  dim3 grid((n + 127) / 128, (unsigned)(rows < 8192 ? rows : 8192));
  k_fi_crop_sign<<<grid, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2 *>(in), row_pitch,
                                                         reinterpret_cast<float2 *>(datac), n, rows);
  return check_launch("k_fi_crop_sign");
}

extern "C" int tmb_fi_scale_sign(float *datac, float c, int n, int nproj, int nz2, void *stream) {
  TMB_REQUIRE(datac && n > 0 && nproj > 0 && nz2 > 0, "tmb_fi_scale_sign: bad argument");
  const size_t rows = (size_t)nproj * nz2;
  dim3 grid((n + 127) / 128, (unsigned)(rows < 4096 ? rows : 4096));
  k_fi_scale_sign<<<grid, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<float2 *>(datac), c, n, rows);
  return check_launch("k_fi_scale_sign");
}

extern "C" int tmb_fi_scale_sign_pairs(const float *datac, float *dataz, float c, int n, int nproj, int nz2, void *stream) {
  TMB_REQUIRE(datac && dataz && datac != dataz && n > 0 && nproj > 0 && nz2 > 0 && nz2 % 2 == 0 &&
                  reinterpret_cast<uintptr_t>(dataz) % 16 == 0,
              "tmb_fi_scale_sign_pairs: bad argument");
  const size_t rows = (size_t)nproj * (nz2 / 2);
  dim3 grid((n + 127) / 128, (unsigned)(rows < 4096 ? rows : 4096));
  k_fi_scale_sign_pairs<<<grid, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2 *>(datac),
                                                                reinterpret_cast<float4 *>(dataz), c, n, (size_t)nproj, rows);
  return check_launch("k_fi_scale_sign_pairs");
}

// test hook: 1 = k_fi_gather (every thread walks its own lines), 2 = k_fi_gather_s (a tile's samples staged in shared
// memory), 3 = k_fi_gather_w (a warp walks its patch's lines in lock step); 0 = the measured best (3)
static int g_fi_gather_mode = 0;
// test hook: complex slices per thread of k_fi_gather_w (2, 4, 8, 16; 0 = default)
static int g_fi_sc = 0;
constexpr int FW_SC_DEFAULT = 4;
static bool g_fi_no_full = false;  // test hook: sc + 100 keeps the per-slice predicates on full chunks too (A/B timing)
extern "C" int tmb_fi_set_slices_per_thread(int sc) {
  const int old = g_fi_sc + (g_fi_no_full ? 100 : 0);
  g_fi_no_full = sc >= 100;
  if (sc >= 100) sc -= 100;
  g_fi_sc = (sc == 2 || sc == 4 || sc == 8 || sc == 16) ? sc : 0;
  return old;
}
extern "C" int tmb_fi_set_gather(int mode) {
  const int old = g_fi_gather_mode;
  g_fi_gather_mode = (mode >= 1 && mode <= 3) ? mode : 0;
  return old;
}

static int fi_gather_launch(const float *datac, float *fde, const float *theta, const float *sorted_theta,
                            const int *sorted_idx, int m, float mu, int n, int nproj, int nz2, int center_size,
                            void *stream) {
  dim3 block(32, 4), grid((center_size + 31) / 32, (center_size + 3) / 4, (nz2 + FI_SC - 1) / FI_SC);
  if (g_fi_gather_mode == 3 || g_fi_gather_mode == 0) {
    // complex slices per thread: g_fi_sc (test hook) or the measured best
    // (measured at 2048^2, 2000 angles, 16 complex slices per launch as FOURIER_INV launches it, with the per-slice
    // predicates: 4 -> 15.3, 8 -> 13.9, 16 -> 18.3 ... 22 ms (the predicated 128-register build of 16 is the one variant
    // whose time moves between boxes); FULL: 8 -> 11.5, 16 -> 10.2 ms; 32 per thread spill; at 5 slices 4 is best)
    const int sc = g_fi_sc ? g_fi_sc : (nz2 % 16 == 0 ? 16 : (nz2 >= 16 ? 8 : FW_SC_DEFAULT));
    const dim3 wgrid((center_size + 31) / 32, (center_size + FW_PY - 1) / FW_PY, (nz2 + sc - 1) / sc);
    const bool full = nz2 % sc == 0 && !g_fi_no_full;  // every z-block holds sc slices: the kernel without per-slice predicates
#define TMB_FW(SC_, FULL_)                                                                                           \
  k_fi_gather_w<SC_, FULL_><<<wgrid, 128, 0, (cudaStream_t)stream>>>(                                                \
      reinterpret_cast<const float2 *>(datac), reinterpret_cast<float2 *>(fde), theta, sorted_theta, sorted_idx, m, mu, \
      n, nproj, nz2, center_size)
#define TMB_FW2(SC_) do { if (full) TMB_FW(SC_, true); else TMB_FW(SC_, false); } while (0)
    if (sc == 16) TMB_FW2(16);
    else if (sc == 8) TMB_FW2(8);
    else if (sc == 2) TMB_FW2(2);
    else TMB_FW2(4);
#undef TMB_FW2
#undef TMB_FW
    return check_launch("k_fi_gather_w");
  }
  if (g_fi_gather_mode == 2) {
    const dim3 sgrid((center_size + FS_TX - 1) / FS_TX, (center_size + FS_TY - 1) / FS_TY, (nz2 + FI_SC - 1) / FI_SC);
    k_fi_gather_s<<<sgrid, block, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2 *>(datac),
                                                            reinterpret_cast<float2 *>(fde), theta, sorted_theta,
                                                            sorted_idx, m, mu, n, nproj, nz2, center_size);
    return check_launch("k_fi_gather_s");
  }
  k_fi_gather<<<grid, block, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2 *>(datac),
                                                        reinterpret_cast<float2 *>(fde), theta, sorted_theta,
                                                        sorted_idx, m, mu, n, nproj, nz2, center_size);
  return check_launch("k_fi_gather");
}

extern "C" int tmb_fi_gather(const float *datac, float *fde, const float *theta, const float *sorted_theta,
                             const int *sorted_idx, int m, float mu, int n, int nproj, int nz2, void *stream) {
  TMB_REQUIRE(datac && fde && theta && sorted_theta && sorted_idx, "tmb_fi_gather: null argument");
  TMB_REQUIRE(n > 0 && nproj > 0 && nz2 > 0 && m > 0 && mu > 0.f, "tmb_fi_gather: bad argument");
  return fi_gather_launch(datac, fde, theta, sorted_theta, sorted_idx, m, mu, n, nproj, nz2, 2 * n, stream);
}

// the whole-grid gather from polar samples in the slice-pair layout of tmb_fi_scale_sign_pairs (one 128-bit load per two
// slices: 16 -> 8 loads and addresses per sample at 16 slices per thread); nz2 must be a multiple of 8
extern "C" int tmb_fi_gather_pairs(const float *dataz, float *fde, const float *theta, const float *sorted_theta,
                                   const int *sorted_idx, int m, float mu, int n, int nproj, int nz2, void *stream) {
  TMB_REQUIRE(dataz && fde && theta && sorted_theta && sorted_idx && m >= 0 && n > 0 && nproj > 0 && nz2 > 0 &&
                  nz2 % 8 == 0 && reinterpret_cast<uintptr_t>(dataz) % 16 == 0,
              "tmb_fi_gather_pairs: bad argument");
  const int center_size = 2 * n;
  const int sc = (g_fi_sc == 8 || nz2 % 16 != 0) ? 8 : 16;
  const dim3 wgrid((center_size + 31) / 32, (center_size + FW_PY - 1) / FW_PY, nz2 / sc);
  if (sc == 16)
    k_fi_gather_w<16, true, true><<<wgrid, 128, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2 *>(dataz), reinterpret_cast<float2 *>(fde), theta, sorted_theta, sorted_idx, m, mu, n,
        nproj, nz2, center_size);
  else
    k_fi_gather_w<8, true, true><<<wgrid, 128, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2 *>(dataz), reinterpret_cast<float2 *>(fde), theta, sorted_theta, sorted_idx, m, mu, n,
        nproj, nz2, center_size);
  return check_launch("k_fi_gather_w (slice pairs)");
}

extern "C" int tmb_fi_gather_center(const float *datac, float *fde, const float *theta, const float *sorted_theta,
                                    const int *sorted_idx, int m, float mu, int n, int nproj, int nz2, int center_size,
                                    void *stream) {
  TMB_REQUIRE(datac && fde && theta && sorted_theta && sorted_idx, "tmb_fi_gather_center: null argument");
  TMB_REQUIRE(n > 0 && nproj > 0 && nz2 > 0 && m > 0 && mu > 0.f, "tmb_fi_gather_center: bad argument");
  TMB_REQUIRE(center_size > 0 && center_size <= 2 * n && center_size % 2 == 0, "tmb_fi_gather_center: bad centre size");
  return fi_gather_launch(datac, fde, theta, sorted_theta, sorted_idx, m, mu, n, nproj, nz2, center_size, stream);
}

extern "C" int tmb_fi_scatter(const float *datac, float *fde, const float *theta, int m, float mu, int center_size,
                              int n, int nproj, int nz2, void *stream) {
  TMB_REQUIRE(datac && fde && theta, "tmb_fi_scatter: null argument");
  TMB_REQUIRE(n > 0 && nproj > 0 && nz2 > 0 && m > 0 && mu > 0.f && center_size >= 0 && center_size <= 2 * n,
              "tmb_fi_scatter: bad argument");
  dim3 block(16, 16), grid((n + 15) / 16, (nproj + 15) / 16, (nz2 + FI_SC - 1) / FI_SC);
  if (center_size > 0)
    k_fi_scatter<true><<<grid, block, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2 *>(datac),
                                                                  reinterpret_cast<float2 *>(fde), theta, m, mu,
                                                                  center_size, n, nproj, nz2);
  else
    k_fi_scatter<false><<<grid, block, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2 *>(datac),
                                                                   reinterpret_cast<float2 *>(fde), theta, m, mu, 0, n,
                                                                   nproj, nz2);
  return check_launch("k_fi_scatter");
}

extern "C" int tmb_fi_sign2d(float *fde, int n, int nz2, void *stream) {
  TMB_REQUIRE(fde && n > 0 && nz2 > 0, "tmb_fi_sign2d: bad argument");
  dim3 block(32, 8), grid((2 * n + 31) / 32, (2 * n + 7) / 8, nz2 < 64 ? nz2 : 64);
  k_fi_sign2d<<<grid, block, 0, (cudaStream_t)stream>>>(reinterpret_cast<float2 *>(fde), 2 * n, nz2);
  return check_launch("k_fi_sign2d");
}

extern "C" int tmb_fi_unpad(float *recon, const float *fde, float mu, float scale, int nproj, int unpad_recon_p,
                            int unpad_z, int unpad_recon_m, int n, int nz2, void *stream) {
  TMB_REQUIRE(recon && fde && n > 0 && nz2 > 0 && unpad_recon_p > unpad_recon_m, "tmb_fi_unpad: bad argument");
  const int rs = unpad_recon_p - unpad_recon_m;
  dim3 block(32, 8), grid((rs + 31) / 32, (rs + 7) / 8, nz2);
  k_fi_unpad<<<grid, block, 0, (cudaStream_t)stream>>>(recon, reinterpret_cast<const float2 *>(fde), mu, scale, nproj,
                                                       unpad_recon_p, unpad_z, unpad_recon_m, n, nz2);
  return check_launch("k_fi_unpad");
}
