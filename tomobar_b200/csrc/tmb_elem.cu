// Fused elementwise steps of the iterative loops, FBP filter construction, circular mask.
// All of these are single-pass HBM streams; 128-bit accesses where alignment allows.
#include "tmb_common.h"

namespace tmb {

constexpr int EL_THREADS = 256;

static inline bool aligned16(const void *a, const void *b, const void *c) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}

static inline int el_blocks(size_t count) {
  size_t b = (count + EL_THREADS - 1) / EL_THREADS;
  const size_t cap = 148 * 16;  // grid-stride: 16 resident CTAs on each of the 148 SMs
  return (int)(b < cap ? (b ? b : 1) : cap);
}

// X = X_t - Linv*grad, optional clamp   (methodsIR_CuPy.py:463-468)
// (x may alias g: every element is read before it is written)
__device__ __forceinline__ float grad_step1(float xt, float g, float linv, int nonneg) {
  // separate multiply and subtract (two roundings) like the reference's two array ops
  const float v = __fsub_rn(xt, __fmul_rn(linv, g));
  return nonneg ? fmaxf(v, 0.f) : v;
}
__global__ void k_fista_grad_step(const float *__restrict__ xt, const float *g, float *x, size_t n, float linv,
                                  int nonneg) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] = grad_step1(xt[i], g[i], linv, nonneg);
}
// 128-bit variant (count % 4 == 0, 16-byte aligned arrays)
__global__ void k_fista_grad_step4(const float4 *__restrict__ xt, const float4 *g, float4 *x, size_t n4, float linv,
                                   int nonneg) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = xt[i], b = g[i];
    x[i] = make_float4(grad_step1(a.x, b.x, linv, nonneg), grad_step1(a.y, b.y, linv, nonneg),
                       grad_step1(a.z, b.z, linv, nonneg), grad_step1(a.w, b.w, linv, nonneg));
  }
}

// X_t = X + coef*(X - X_old)   (methodsIR_CuPy.py:475)
__device__ __forceinline__ float momentum1(float a, float o, float coef) {
  return __fadd_rn(a, __fmul_rn(coef, __fsub_rn(a, o)));
}
__global__ void k_fista_momentum(const float *__restrict__ x, const float *__restrict__ xo, float *__restrict__ xt,
                                 size_t n, float coef) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    xt[i] = momentum1(x[i], xo[i], coef);
}
__global__ void k_fista_momentum4(const float4 *__restrict__ x, const float4 *__restrict__ xo,
                                  float4 *__restrict__ xt, size_t n4, float coef) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = x[i], o = xo[i];
    xt[i] = make_float4(momentum1(a.x, o.x, coef), momentum1(a.y, o.y, coef), momentum1(a.z, o.z, coef),
                        momentum1(a.w, o.w, coef));
  }
}

// ADMM z-update, relaxation, z_old copy and prox input in one pass (methodsIR_CuPy.py:545-557)
__global__ void k_admm_z(float *__restrict__ z, float *__restrict__ zo, const float *__restrict__ x,
                         const float *__restrict__ u, const float *__restrict__ g, float *__restrict__ xp, size_t n,
                         float tau, float rho, int nonneg, int relax, float alpha, float oma) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float zi = z[i];
    const float ui = u[i];
    const float ga = __fmul_rn(rho, __fadd_rn(__fsub_rn(zi, x[i]), ui));
    zi = __fsub_rn(zi, __fmul_rn(tau, __fadd_rn(g[i], ga)));
    if (nonneg) zi = fmaxf(zi, 0.f);
    if (relax) zi = __fadd_rn(__fmul_rn(oma, zo[i]), __fmul_rn(alpha, zi));
    z[i] = zi;
    zo[i] = zi;
    xp[i] = __fadd_rn(zi, ui);
  }
}

__global__ void k_admm_u(float *__restrict__ u, const float *__restrict__ z, const float *__restrict__ x, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    u[i] = __fadd_rn(u[i], __fsub_rn(z[i], x[i]));
}

__global__ void k_axpy(float a, const float *__restrict__ x, float *__restrict__ y, size_t n, int nonneg) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = __fadd_rn(y[i], __fmul_rn(a, x[i]));
    if (nonneg) v = fmaxf(v, 0.f);
    y[i] = v;
  }
}

// ---- sinc filter (generate_filtersync.cu:5-82) ----------------------------------------------
__device__ float block_sum(float v, float *sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
  return t;
}

__global__ void k_sinc_filter(float a, float *__restrict__ f, int n, float multiplier) {
  __shared__ float sh[32];
  const float pi = 3.1415926535897932384626433832795f;
  const float dw = 2 * pi / n;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float rd = a * (-pi + i * dw) / 2.0f;
    s += rd * rd;
  }
  const float sum_sq = block_sum(s, sh);
  float d = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float rd = a * (-pi + i * dw) / 2.0f;
    d += sinf(rd) * rd / sum_sq;
  }
  const float dot = block_sum(d, sh);
  const float dot_sq = dot * dot;
  const int shift = n / 2;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int o = (i + shift) % n;  // ifftshifted position; keep the rfft half
    if (o >= n / 2 + 1) continue;
    const float rd = a * (-pi + i * dw) / 2.0f;
    const float r = fabsf((float)(2.0 / (double)a * (double)sinf(rd))) * dot_sq;
    f[o] = r * multiplier;
  }
}

__global__ void k_apply_filter(float2 *__restrict__ spec, const float *__restrict__ f, size_t rows, int nbins) {
  const size_t total = rows * (size_t)nbins;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const float w = __ldg(f + (i % nbins));
    float2 v = spec[i];
    v.x *= w;
    v.y *= w;
    spec[i] = v;
  }
}

// edge padding of the last axis (supp/suppTools.py:425-459; methodsDIR_CuPy.py:505-521):
// out[row][j] = in[row][clamp(j - pad_left, 0, w - 1)]
__global__ void k_edge_pad(const float *__restrict__ in, float *__restrict__ out, size_t rows, int w, int wout,
                           int pad_left) {
  const size_t total = rows * (size_t)wout;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / wout;
    const int j = (int)(i - r * wout) - pad_left;
    out[i] = __ldg(in + r * w + min(max(j, 0), w - 1));
  }
}

// the same with 128-bit stores (wout % 4 == 0, 16-byte aligned out): a thread writes four consecutive outputs of a
// row; the output is 4 x the input for FOURIER_INV's oversampled detector and mostly replicated edge values, so the
// kernel is a store stream (round 2: 5.1 -> ~2 ms per call at config 4; the scalar version above divided a 64-bit
// index per element)
__global__ void k_edge_pad4(const float *__restrict__ in, float4 *__restrict__ out, size_t rows, int w, int wout4,
                            int pad_left) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;  // which float4 of the row
  if (q >= wout4) return;
  const int j = 4 * q - pad_left;
  const int i0 = min(max(j, 0), w - 1), i1 = min(max(j + 1, 0), w - 1), i2 = min(max(j + 2, 0), w - 1),
            i3 = min(max(j + 3, 0), w - 1);
  for (size_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const float *row = in + r * (size_t)w;
    out[r * (size_t)wout4 + q] = make_float4(__ldg(row + i0), __ldg(row + i1), __ldg(row + i2), __ldg(row + i3));
  }
}

// edge padding of TWO real slices into one complex row: out[t][row][j] = (in[2t][row][c], in[2t+1][row][c]),
// c = clamp(j - pad_left, 0, w - 1).  FOURIER_INV filters slice pairs as complex rows (one c2c transform instead of two
// r2c / c2r: the filter has a real impulse response, so real and imaginary part are filtered independently), which is
// also the pairing its gridding step needs.  A thread writes two complex samples (one 128-bit store).
__global__ void k_edge_pad_pair(const float *__restrict__ in, float4 *__restrict__ out, size_t rows, size_t slice_elems,
                                int nzc, int w, int wout2, int pad_left) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;  // which pair of output samples of the row
  if (q >= wout2) return;
  const int j = 2 * q - pad_left;
  const int i0 = min(max(j, 0), w - 1), i1 = min(max(j + 1, 0), w - 1);
  for (size_t r = blockIdx.y; r < rows * (size_t)nzc; r += gridDim.y) {
    const size_t t = r / rows, row = r - t * rows;
    const float *a = in + (2 * t) * slice_elems + row * (size_t)w;
    const float *b = a + slice_elems;
    out[r * (size_t)wout2 + q] = make_float4(__ldg(a + i0), __ldg(b + i0), __ldg(a + i1), __ldg(b + i1));
  }
}

// circular mask (supp/suppTools.py:364-396)
__global__ void k_mask(float *__restrict__ vol, int nz, int n, double limit) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (c >= n) return;
  const int h = n / 2;
  const double dist = sqrt((double)((c - h) * (c - h) + (r - h) * (r - h)));
  if (dist <= limit) return;
  for (int z = blockIdx.z; z < nz; z += gridDim.z) vol[((size_t)z * n + r) * n + c] = 0.f;
}


// flat / dark-field normalisation and negative log (supp/suppTools.py:187-264, "mean" / "median"
// branch): one pass from the raw uint16 (or fp32) projections to the fp32 sinogram.
// data is [n0][n1][n2] with the ANGLE axis 0 or 1; flat / dark are the averaged [.][n2] fields.
template <typename TI>
__global__ void k_normalise(const TI *__restrict__ data, const float *__restrict__ flat,
                            const float *__restrict__ dark, float *__restrict__ out, size_t total, int n1, int n2,
                            int angle_axis, int take_log) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t i2 = i % n2, i01 = i / n2;
    const size_t row = angle_axis == 0 ? i01 % n1 : i01 / n1;  // index of the non-angle slow axis
    const float d = __ldg(dark + row * n2 + i2);
    float denom = __fsub_rn(__ldg(flat + row * n2 + i2), d);
    if (denom <= 0.f) denom = 1.f;
    float nomin = __fsub_rn((float)data[i], d);
    if (nomin < 0.f) nomin = 1.f;
    float v = __fdiv_rn(nomin, denom);
    if (take_log) {
      if (v > 0.f) v = -logf(v);
      if (v < 0.f) v = 0.f;
    }
    out[i] = v;
  }
}

}  // namespace tmb

using namespace tmb;

extern "C" int tmb_fista_grad_step(const float *x_t, const float *grad, float *x, size_t count, float l_inv,
                                   int nonneg, void *stream) {
  TMB_REQUIRE(x_t && grad && x, "tmb_fista_grad_step: null argument");
  if (count % 4 == 0 && aligned16(x_t, grad, x))
    k_fista_grad_step4<<<el_blocks(count / 4), EL_THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(x_t), reinterpret_cast<const float4 *>(grad), reinterpret_cast<float4 *>(x),
        count / 4, l_inv, nonneg);
  else
    k_fista_grad_step<<<el_blocks(count), EL_THREADS, 0, (cudaStream_t)stream>>>(x_t, grad, x, count, l_inv, nonneg);
  return check_launch("k_fista_grad_step");
}

extern "C" int tmb_fista_momentum(const float *x, const float *x_old, float *x_t, size_t count, float coef,
                                  void *stream) {
  TMB_REQUIRE(x && x_old && x_t, "tmb_fista_momentum: null argument");
  if (count % 4 == 0 && aligned16(x, x_old, x_t))
    k_fista_momentum4<<<el_blocks(count / 4), EL_THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(x), reinterpret_cast<const float4 *>(x_old), reinterpret_cast<float4 *>(x_t),
        count / 4, coef);
  else
    k_fista_momentum<<<el_blocks(count), EL_THREADS, 0, (cudaStream_t)stream>>>(x, x_old, x_t, count, coef);
  return check_launch("k_fista_momentum");
}

extern "C" int tmb_admm_z_step(float *z, float *z_old, const float *x, const float *u, const float *grad,
                               float *xprox, size_t count, float tau, float rho, int nonneg, int relax, float alpha,
                               void *stream) {
  TMB_REQUIRE(z && z_old && x && u && grad && xprox, "tmb_admm_z_step: null argument");
  k_admm_z<<<el_blocks(count), EL_THREADS, 0, (cudaStream_t)stream>>>(z, z_old, x, u, grad, xprox, count, tau, rho,
                                                                      nonneg, relax, alpha,
                                                                      (float)(1.0 - (double)alpha));
  return check_launch("k_admm_z");
}

extern "C" int tmb_admm_u_step(float *u, const float *z, const float *x, size_t count, void *stream) {
  TMB_REQUIRE(u && z && x, "tmb_admm_u_step: null argument");
  k_admm_u<<<el_blocks(count), EL_THREADS, 0, (cudaStream_t)stream>>>(u, z, x, count);
  return check_launch("k_admm_u");
}

extern "C" int tmb_axpy(float a, const float *x, float *y, size_t count, int nonneg, void *stream) {
  TMB_REQUIRE(x && y, "tmb_axpy: null argument");
  k_axpy<<<el_blocks(count), EL_THREADS, 0, (cudaStream_t)stream>>>(a, x, y, count, nonneg);
  return check_launch("k_axpy");
}

extern "C" int tmb_sinc_filter(float cutoff, float *f, int n, float multiplier, void *stream) {
  TMB_REQUIRE(f && n >= 2, "tmb_sinc_filter: bad argument");
  k_sinc_filter<<<1, 256, 0, (cudaStream_t)stream>>>(cutoff, f, n, multiplier);
  return check_launch("k_sinc_filter");
}

extern "C" int tmb_apply_filter(float *spec, const float *f, size_t rows, int nbins, void *stream) {
  TMB_REQUIRE(spec && f && nbins >= 1, "tmb_apply_filter: bad argument");
  k_apply_filter<<<el_blocks(rows * (size_t)nbins), EL_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<float2 *>(spec), f, rows, nbins);
  return check_launch("k_apply_filter");
}

extern "C" int tmb_circular_mask(float *vol, int nz, int n, float radius, void *stream) {
  TMB_REQUIRE(vol && nz >= 1 && n >= 1 && radius > 0.f, "tmb_circular_mask: bad argument");
  const int h = n / 2;
  // python: h - abs(h - h/radius)  (radius <= 1)   or   h + abs(h - h/radius)
  const double delta = fabs((double)h - (double)h / (double)radius);
  const double limit = radius <= 1.0f ? (double)h - delta : (double)h + delta;
  dim3 grid((n + 127) / 128, n, nz < 64 ? nz : 64);
  k_mask<<<grid, 128, 0, (cudaStream_t)stream>>>(vol, nz, n, limit);
  return check_launch("k_mask");
}

extern "C" int tmb_normalise(const void *data, int data_is_u16, const float *flat_mean, const float *dark_mean,
                             float *out, int n0, int n1, int n2, int angle_axis, int take_log, void *stream) {
  TMB_REQUIRE(data && flat_mean && dark_mean && out, "tmb_normalise: null argument");
  TMB_REQUIRE(n0 >= 1 && n1 >= 1 && n2 >= 1 && (angle_axis == 0 || angle_axis == 1), "tmb_normalise: bad argument");
  const size_t total = (size_t)n0 * n1 * n2;
  if (data_is_u16)
    k_normalise<unsigned short><<<el_blocks(total), EL_THREADS, 0, (cudaStream_t)stream>>>(
        static_cast<const unsigned short *>(data), flat_mean, dark_mean, out, total, n1, n2, angle_axis, take_log);
  else
    k_normalise<float><<<el_blocks(total), EL_THREADS, 0, (cudaStream_t)stream>>>(
        static_cast<const float *>(data), flat_mean, dark_mean, out, total, n1, n2, angle_axis, take_log);
  return check_launch("k_normalise");
}

extern "C" int tmb_edge_pad_pair(const float *in, float *out, int nzc, size_t rows, int w, int wout, int pad_left,
                                 void *stream) {
  TMB_REQUIRE(in && out && nzc >= 1 && rows >= 1 && w >= 1 && wout >= w && wout % 2 == 0 && pad_left >= 0 &&
                  pad_left + w <= wout && reinterpret_cast<uintptr_t>(out) % 16 == 0,
              "tmb_edge_pad_pair: bad argument");
  const int wout2 = wout / 2;
  const size_t total = rows * (size_t)nzc;
  const dim3 grid((wout2 + 255) / 256, (unsigned)(total < 16384 ? total : 16384));
  k_edge_pad_pair<<<grid, 256, 0, (cudaStream_t)stream>>>(in, reinterpret_cast<float4 *>(out), rows, rows * (size_t)w, nzc, w,
                                                          wout2, pad_left);
  return check_launch("k_edge_pad_pair");
}

extern "C" int tmb_edge_pad(const float *in, float *out, size_t rows, int w, int wout, int pad_left, void *stream) {
  TMB_REQUIRE(in && out && in != out && w >= 1 && wout >= w && pad_left >= 0 && pad_left + w <= wout,
              "tmb_edge_pad: bad argument");
  if (wout % 4 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0) {
    const int wout4 = wout / 4;
    const dim3 grid((wout4 + 255) / 256, (unsigned)(rows < 16384 ? rows : 16384));
    k_edge_pad4<<<grid, 256, 0, (cudaStream_t)stream>>>(in, reinterpret_cast<float4 *>(out), rows, w, wout4, pad_left);
    return check_launch("k_edge_pad4");
  }
  k_edge_pad<<<el_blocks(rows * (size_t)wout), EL_THREADS, 0, (cudaStream_t)stream>>>(in, out, rows, w, wout, pad_left);
  return check_launch("k_edge_pad");
}
