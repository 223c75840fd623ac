/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see proj_oracle.c).  Multi-threaded (OpenMP) twins of
 * oracle.py's numpy restatements of the two TV proximal operators, so that the CPU baseline of
 * bench.py uses all host cores for the whole sub-step and not only for the projector pair:
 *   PD_TV  : tomobar/regularisersCuPy.py:170-296 + cuda_kernels/primal_dual_for_total_variation.cu:126-261 (3-D),
 *            :361-492 (2-D)
 *   ROF_TV : tomobar/regularisersCuPy.py:41-167 + cuda_kernels/rudin_osher_fatemi_total_variation.cu:70-148, 157-248
 * Every float operation is written in the order oracle.pd_tv / oracle.rof_tv evaluate it (compiled with
 * -ffp-contract=off), so the results are BIT-IDENTICAL to the numpy versions
 * (tests/test_oracle_tv_c.py); fp32 dual variables only.
 *
 * Layout: vol[dz][dy][dx], x fastest; dz == 1 means a 2-D image (two dual components).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

/* forward difference, U[prev] - U at the last index (primal_dual_for_total_variation.cu:214-222) */
static inline float fwd(const float *U, size_t i, int k, int n, size_t stride) {
  if (n == 1) return 0.0f - U[i];
  return (k == n - 1 ? U[i - stride] : U[i + stride]) - U[i];
}

int oracle_pd_tv(const float *in, float *out, int dz, int dy, int dx, float tau, float sigma, float lt, float theta,
                 int iterations, int methodTV, int nonneg) {
  const size_t nv = (size_t)dz * dy * dx, sy = (size_t)dx, sz = (size_t)dx * dy;
  const int is3d = dz > 1;
  /* tau, sigma, lt, theta: the host scalars of regularisersCuPy.py:208-212, computed by the caller (oracle.pd_tv) */
  float *U = (float *)malloc(nv * sizeof(float)), *V = (float *)malloc(nv * sizeof(float));
  float *P1 = (float *)calloc(nv, sizeof(float)), *P2 = (float *)calloc(nv, sizeof(float));
  float *P3 = (float *)calloc(is3d ? nv : 1, sizeof(float));
  if (!U || !V || !P1 || !P2 || !P3) return -1;
  memcpy(U, in, nv * sizeof(float));
  for (int it = 0; it < iterations; ++it) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < dz; ++z)
      for (int y = 0; y < dy; ++y)
        for (int x = 0; x < dx; ++x) {
          const size_t i = (size_t)z * sz + (size_t)y * sy + x;
          float p1 = P1[i] + sigma * fwd(U, i, x, dx, 1);
          float p2 = P2[i] + sigma * fwd(U, i, y, dy, sy);
          float p3 = is3d ? P3[i] + sigma * fwd(U, i, z, dz, sz) : 0.0f;
          if (methodTV == 0) {
            float den = p1 * p1 + p2 * p2;
            if (is3d) den = den + p3 * p3;
            const float sc = den > 1.0f ? 1.0f / sqrtf(den) : 1.0f;
            p1 = p1 * sc; p2 = p2 * sc; p3 = p3 * sc;
          } else {
            p1 = p1 / fmaxf(fabsf(p1), 1.0f);
            p2 = p2 / fmaxf(fabsf(p2), 1.0f);
            p3 = p3 / fmaxf(fabsf(p3), 1.0f);
          }
          P1[i] = p1; P2[i] = p2;
          if (is3d) P3[i] = p3;
        }
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < dz; ++z)
      for (int y = 0; y < dy; ++y)
        for (int x = 0; x < dx; ++x) {
          const size_t i = (size_t)z * sz + (size_t)y * sy + x;
          /* div = sum_d -(P_d - P_d[-1]), P_d[-1] := 0 at index 0 (:147-162) */
          float div = -(P1[i] - (x > 0 ? P1[i - 1] : 0.0f));
          div = div + -(P2[i] - (y > 0 ? P2[i - sy] : 0.0f));
          if (is3d) div = div + -(P3[i] - (z > 0 ? P3[i - sz] : 0.0f));
          const float ub = nonneg ? fmaxf(U[i], 0.0f) : U[i];
          const float nu = ((ub - tau * div) + lt * in[i]) / (1.0f + lt);
          V[i] = nu + theta * (nu - ub);
        }
    float *t = U; U = V; V = t;
  }
  memcpy(out, U, nv * sizeof(float));
  free(U); free(V); free(P1); free(P2); free(P3);
  return 0;
}

static inline float sgnf(float v) { return (float)((v > 0.0f) - (v < 0.0f)); }

int oracle_rof_tv(const float *in, float *out, int dz, int dy, int dx, float lambda, int iterations, float tau) {
  const size_t nv = (size_t)dz * dy * dx, sy = (size_t)dx, sz = (size_t)dx * dy;
  const int is3d = dz > 1;
  float *U = (float *)malloc(nv * sizeof(float)), *V = (float *)malloc(nv * sizeof(float));
  float *Dx = (float *)malloc(nv * sizeof(float)), *Dy = (float *)malloc(nv * sizeof(float));
  float *Dz = (float *)malloc((is3d ? nv : 1) * sizeof(float));
  if (!U || !V || !Dx || !Dy || !Dz) return -1;
  memcpy(U, in, nv * sizeof(float));
  for (int it = 0; it < iterations; ++it) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < dz; ++z)
      for (int y = 0; y < dy; ++y)
        for (int x = 0; x < dx; ++x) {
          const size_t i = (size_t)z * sz + (size_t)y * sy + x;
          const float u = U[i];
          /* neighbours reflect at the boundary (rudin_osher_fatemi_total_variation.cu:170-181) */
          const float xp = dx > 1 ? U[x == dx - 1 ? i - 1 : i + 1] : u, xm = dx > 1 ? U[x == 0 ? i + 1 : i - 1] : u;
          const float yp = dy > 1 ? U[y == dy - 1 ? i - sy : i + sy] : u, ym = dy > 1 ? U[y == 0 ? i + sy : i - sy] : u;
          const float n1x = xp - u, n0x = u - xm, n1y = yp - u, n0y = u - ym;
          /* calculate_denominator (:51-55): 0.5 is a double literal, result stored as float */
          const float ex = (float)(0.5 * (double)(sgnf(n1x) + sgnf(n0x)) * (double)fminf(fabsf(n1x), fabsf(n0x)));
          const float ey = (float)(0.5 * (double)(sgnf(n1y) + sgnf(n0y)) * (double)fminf(fabsf(n1y), fabsf(n0y)));
          const float mx = ex * ex, my = ey * ey;
          if (is3d) {
            const float zp = U[z == dz - 1 ? i - sz : i + sz], zm = U[z == 0 ? i + sz : i - sz];
            const float n1z = zp - u, n0z = u - zm;
            const float ez = (float)(0.5 * (double)(sgnf(n1z) + sgnf(n0z)) * (double)fminf(fabsf(n1z), fabsf(n0z)));
            const float mz = ez * ez;
            /* normalize_difference (:57-61): float sums in the kernels' argument order (middle, fast, slow
             * axis), + EPS in double */
            float s = (my + n1x * n1x) + mz;
            Dx[i] = n1x / sqrtf((float)((double)s + 1.0e-8));
            s = (n1y * n1y + mx) + mz;
            Dy[i] = n1y / sqrtf((float)((double)s + 1.0e-8));
            s = (my + mx) + n1z * n1z;
            Dz[i] = n1z / sqrtf((float)((double)s + 1.0e-8));
          } else {
            float s = my + n1x * n1x;
            Dx[i] = n1x / sqrtf((float)((double)s + 1.0e-8));
            s = n1y * n1y + mx;
            Dy[i] = n1y / sqrtf((float)((double)s + 1.0e-8));
          }
        }
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < dz; ++z)
      for (int y = 0; y < dy; ++y)
        for (int x = 0; x < dx; ++x) {
          const size_t i = (size_t)z * sz + (size_t)y * sy + x;
          /* backward differences of D, index 0 reads index 1 (:235) */
          const float tx = Dx[i] - (dx > 1 ? Dx[x == 0 ? i + 1 : i - 1] : Dx[i]);
          const float ty = Dy[i] - (dy > 1 ? Dy[y == 0 ? i + sy : i - sy] : Dy[i]);
          float dv = ty + tx;
          if (is3d) dv = dv + (Dz[i] - Dz[z == 0 ? i + sz : i - sz]);
          V[i] = U[i] + tau * (lambda * dv - (U[i] - in[i]));
        }
    float *t = U; U = V; V = t;
  }
  memcpy(out, U, nv * sizeof(float));
  free(U); free(V); free(Dx); free(Dy); free(Dz);
  return 0;
}
