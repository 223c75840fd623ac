/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see proj_oracle.c).
 *
 * The back-projection half of the reference's CPU direct method for 2-D data (BASELINE.json config 1):
 *   RecToolsDIR.FBP, device "cpu":  self.Atools._backproj(_filtersinc2D(data))      tomobar/methodsDIR.py:161-168
 *   -> AstraTools2D._backproj -> _runAstraBackproj2D(method "BP", projector "line")   astra_wrappers/astra_tools2d.py:88-92,
 *                                                                                     astra_base.py:224-232, 311-370
 * The arithmetic lives in astra-toolbox==2.4.* (pyproject.toml:41, un-vendored): its CPU "line" projector gives
 * every (ray, pixel) pair the exact length of the ray inside the pixel, walking the rows (|ray_y| > |ray_x|) or
 * the columns of the image and splitting the per-row length between the at most two pixels the ray crosses there;
 * the BP algorithm adds weight * sinogram value into the image, angle by angle, detector by detector, in float.
 * Restated here from that published algorithm and pinned on the reference's own golden for this path
 * (tests/test_RecToolsDIR.py:198-218, see tests/test_oracle_c1_cpu_fbp.py).
 *
 * ASTRA's 2-D image is y-up: row 0 is the TOP row (y = +n/2 - 1/2), i.e. vertically flipped with respect to the
 * 3-D path of proj_oracle.c ("orientation").
 *
 * vec[na][6] (doubles) = (ray_x, ray_y, det_centre_x, det_centre_y, u_x, u_y) per angle, the "parallel" geometry's
 * toVectorGeometry(): ray (sin t, -cos t), u (cos t, sin t), detector centred (the CPU path ignores the CoR offset,
 * astra_base.py:224-231).  threads <= 1: one thread, ASTRA's accumulation order (parity); threads > 1: angles are
 * dealt to OpenMP threads with private images that are summed at the end (timing on all cores).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static void bp_angle(const float *srow, float *vol, const double *v, int n, int nu) {
  const double rayX = v[0], rayY = v[1], uX = v[4], uY = v[5];
  const double detSX = v[2] - 0.5 * (double)nu * uX, detSY = v[3] - 0.5 * (double)nu * uY;
  const int vertical = fabs(rayX) < fabs(rayY);
  float S, T, lengthPer, invTminS, delta, ratio;
  if (vertical) {
    ratio = (float)(rayX / rayY);
    lengthPer = (float)(sqrt(rayY * rayY + rayX * rayX) / fabs(rayY));
    delta = -ratio;
  } else {
    ratio = (float)(rayY / rayX);
    lengthPer = (float)(sqrt(rayY * rayY + rayX * rayX) / fabs(rayX));
    delta = -ratio;
  }
  S = 0.5f - 0.5f * fabsf(ratio);
  T = 0.5f + 0.5f * fabsf(ratio);
  invTminS = lengthPer / (T - S);
  const float Ex = -0.5f * (float)n + 0.5f, Ey = 0.5f * (float)n - 0.5f;
  for (int d = 0; d < nu; ++d) {
    const float val = srow[d];
    const float Dx = (float)(detSX + ((double)d + 0.5) * uX), Dy = (float)(detSY + ((double)d + 0.5) * uY);
    int isin = 0;
    if (vertical) {
      float c = (Dx + (Ey - Dy) * ratio - Ex);
      for (int row = 0; row < n; ++row, c += delta) {
        const int col = (int)floorf(c + 0.5f);
        if (col < -1 || col > n) { if (!isin) continue; else break; }
        const float offset = c - (float)col;
        float *vr = vol + (size_t)row * n;
        if (offset < -S) {
          const float w = (offset + T) * invTminS;
          if (col > 0) vr[col - 1] += (lengthPer - w) * val;
          if (col >= 0 && col < n) vr[col] += w * val;
        } else if (S < offset) {
          const float w = (offset - S) * invTminS;
          if (col >= 0 && col < n) vr[col] += (lengthPer - w) * val;
          if (col + 1 < n) vr[col + 1] += w * val;
        } else if (col >= 0 && col < n) {
          vr[col] += lengthPer * val;
        }
        isin = 1;
      }
    } else {
      float r = -(Dy + (Ex - Dx) * ratio - Ey);
      for (int col = 0; col < n; ++col, r += delta) {
        const int row = (int)floorf(r + 0.5f);
        if (row < -1 || row > n) { if (!isin) continue; else break; }
        const float offset = r - (float)row;
        if (offset < -S) {
          const float w = (offset + T) * invTminS;
          if (row > 0) vol[(size_t)(row - 1) * n + col] += (lengthPer - w) * val;
          if (row >= 0 && row < n) vol[(size_t)row * n + col] += w * val;
        } else if (S < offset) {
          const float w = (offset - S) * invTminS;
          if (row >= 0 && row < n) vol[(size_t)row * n + col] += (lengthPer - w) * val;
          if (row + 1 < n) vol[(size_t)(row + 1) * n + col] += w * val;
        } else if (row >= 0 && row < n) {
          vol[(size_t)row * n + col] += lengthPer * val;
        }
        isin = 1;
      }
    }
  }
}

void oracle_bp2d_line(const float *sino, float *vol, const double *vec, int n, int nu, int na, int threads) {
  memset(vol, 0, (size_t)n * n * sizeof(float));
  if (threads <= 1) {
    for (int a = 0; a < na; ++a) bp_angle(sino + (size_t)a * nu, vol, vec + (size_t)a * 6, n, nu);
    return;
  }
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
  {
    float *mine = (float *)calloc((size_t)n * n, sizeof(float));
#pragma omp for schedule(static)
    for (int a = 0; a < na; ++a) bp_angle(sino + (size_t)a * nu, mine, vec + (size_t)a * 6, n, nu);
#pragma omp critical
    for (size_t i = 0; i < (size_t)n * n; ++i) vol[i] += mine[i];
    free(mine);
  }
#else
  for (int a = 0; a < na; ++a) bp_angle(sino + (size_t)a * nu, vol, vec + (size_t)a * 6, n, nu);
#endif
}
