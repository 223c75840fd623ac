/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path
 * (tomobar_b200/); used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the checker / CPU baseline.
 *
 * CPU restatement of the parallel-beam 3-D projector pair that the reference
 * reaches through astra-toolbox==2.4.* (pyproject.toml:41; un-vendored):
 *   forward  : tomobar/astra_wrappers/astra_base.py:560-606 (direct_FP3D, :601)
 *   backward : tomobar/astra_wrappers/astra_base.py:518-558 (direct_BP3D, :554)
 * ASTRA's published par3d model (Joseph line integrals for A, voxel-driven
 * linear interpolation for A^T, both through the CUDA texture unit whose
 * interpolation fraction is an 8-bit fixed-point number) is restated here as
 * specified in SURVEY.md Appendix A.  Pinned against the reference's own
 * goldens in tests/test_oracle_goldens.py.
 *
 * The per-angle table (8 floats per angle) is produced by the host from
 * supp/funcs.py:45-81 vector geometry; layout (see oracle/oracle.py:angle_table):
 *   [0] cos  [1] sin  [2] bp_off  [3] fp_alpha  [4] fp_b0  [5] fp_bstep
 *   [6] fp_scale  [7] dir (0: march along x/columns, interpolate along rows;
 *                          1: march along y/rows,    interpolate along columns)
 *
 * Layouts: vol[nz][n][n] (row r <-> +y, col c <-> +x), sino[nz][na][nu].
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#define TBL 8

/* quant: bit 0 = 8-bit texture weights; bit 1 = ASTRA's accumulation order as well (hard part 1:
 * the forward projector sums 32-line slabs, scales each slab sum and adds it to the output; the back-projector sums
 * groups of 32 angles per launch and adds each group sum to the volume).  Bit 1 exists to MEASURE how much of the
 * residual golden drift that order explains (profiles/golden_report_r02.txt); the product kernels do not use it.
 * Bit 2 (probe as well): the interpolation coordinates evaluated without fused multiply-adds / in another order, i.e.
 * another equally valid fp32 rounding of the same geometry: measures how far a result moves when ~0.2 % of the 8-bit
 * weights flip to the neighbouring 1/256 step, which is what separates ANY restatement from ASTRA's own kernels. */
#define SLAB 32
static inline float quant8(float f, int quant) {
  if (quant & 64) return floorf(f * 256.0f) * (1.0f / 256.0f); /* probe: truncation instead of round-to-nearest */
  if (quant & 128) return floorf(f * 256.0f + 0.5f) * (1.0f / 256.0f); /* probe: round half up */
  return (quant & 1) ? rintf(f * 256.0f) * (1.0f / 256.0f) : f;
}

/* A^T : voxel-driven back-projection, scale 1 */
void oracle_bp3d(const float *sino, float *vol, const float *tbl, int nz, int n,
                 int nu, int na, int quant) {
  const float half = 0.5f * (float)n;
  if (quant & 32) quant = (quant & 7) | 64; /* probe: truncating weights in the back-projector only */
  else if (quant & 256) quant = (quant & 7) | 128; /* probe: round-half-up weights in the back-projector */
  else quant &= 7;
#pragma omp parallel for collapse(2) schedule(static)
  for (int z = 0; z < nz; ++z) {
    for (int r = 0; r < n; ++r) {
      const float y = (float)r - half + 0.5f;
      const float *sz = sino + (size_t)z * na * nu;
      float *out = vol + ((size_t)z * n + r) * n;
      for (int c = 0; c < n; ++c) out[c] = 0.0f;
      for (int a = 0; a < na; ++a) {
        const float ca = tbl[a * TBL + 0], sa = tbl[a * TBL + 1], off = tbl[a * TBL + 2];
        const float *row = sz + (size_t)a * nu;
        const float ys = fmaf(y, sa, off);
        for (int c = 0; c < n; ++c) {
          const float x = (float)c - half + 0.5f;
          /* bit 2 (sensitivity probe only): the same coordinate with the products rounded separately and added in
           * another order -- a different, equally valid fp32 evaluation of x cos + y sin + off */
          const float u = (quant & 4) ? (x * ca + y * sa) + off : fmaf(x, ca, ys);
          const float fl = floorf(u);
          const float f = quant8(u - fl, quant);
          const float g = 1.0f - f;
          const int i = (int)fl;
          const float s0 = (i >= 0 && i < nu) ? row[i] : 0.0f;
          const float s1 = (i + 1 >= 0 && i + 1 < nu) ? row[i + 1] : 0.0f;
          float acc = out[c];
          acc = fmaf(g, s0, acc);
          acc = fmaf(f, s1, acc);
          out[c] = acc;
        }
      }
    }
  }
  if (quant & 2) { /* the same sum associated as ASTRA's launches of 32 angles: vol += (sum over the group) */
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < nz; ++z) {
      for (int r = 0; r < n; ++r) {
        const float y = (float)r - half + 0.5f;
        const float *sz = sino + (size_t)z * na * nu;
        float *out = vol + ((size_t)z * n + r) * n;
        for (int c = 0; c < n; ++c) {
          const float x = (float)c - half + 0.5f;
          float total = 0.0f;
          for (int a0 = 0; a0 < na; a0 += SLAB) {
            float acc = 0.0f;
            for (int a = a0; a < na && a < a0 + SLAB; ++a) {
              const float ca = tbl[a * TBL + 0], sa = tbl[a * TBL + 1], off = tbl[a * TBL + 2];
              const float *row = sz + (size_t)a * nu;
              const float u = fmaf(x, ca, fmaf(y, sa, off));
              const float fl = floorf(u);
              const float f = quant8(u - fl, quant);
              const int i = (int)fl;
              const float s0 = (i >= 0 && i < nu) ? row[i] : 0.0f;
              const float s1 = (i + 1 >= 0 && i + 1 < nu) ? row[i + 1] : 0.0f;
              acc += s0 + f * (s1 - s0); /* the texture unit's lerp */
            }
            total += acc;
          }
          out[c] = total;
        }
      }
    }
  }
}

/* A : Joseph forward projection */
void oracle_fp3d(const float *vol, float *sino, const float *tbl, int nz, int n,
                 int nu, int na, int quant) {
  const float half = 0.5f * (float)n;
  if (quant & 8) quant = (quant & 7) | 64;        /* probe: truncating weights in the forward projector only */
  else if (quant & 16) quant = (quant & 7) | 128; /* probe: round-half-up weights in the forward projector only */
  else quant &= 7;
#pragma omp parallel for collapse(2) schedule(static)
  for (int z = 0; z < nz; ++z) {
    for (int a = 0; a < na; ++a) {
      const float alpha = tbl[a * TBL + 3], b0 = tbl[a * TBL + 4], bstep = tbl[a * TBL + 5];
      const float scale = tbl[a * TBL + 6];
      const int dir = (int)tbl[a * TBL + 7];
      const float *vz = vol + (size_t)z * n * n;
      float *out = sino + ((size_t)z * na + a) * nu;
      for (int k = 0; k < nu; ++k) {
        const float beta = fmaf((float)k, bstep, b0);
        float acc = 0.0f, total = 0.0f;
        for (int m = 0; m < n; ++m) {
          const float xm = (float)m - half + 0.5f;
          const float rho = (quant & 4) ? alpha * xm + beta : fmaf(alpha, xm, beta);
          const float fl = floorf(rho);
          const float f = quant8(rho - fl, quant);
          const float g = 1.0f - f;
          const int i = (int)fl;
          float v0 = 0.0f, v1 = 0.0f;
          if (dir == 0) { /* line m = column m, i = row */
            if (i >= 0 && i < n) v0 = vz[(size_t)i * n + m];
            if (i + 1 >= 0 && i + 1 < n) v1 = vz[(size_t)(i + 1) * n + m];
          } else { /* line m = row m, i = column */
            if (i >= 0 && i < n) v0 = vz[(size_t)m * n + i];
            if (i + 1 >= 0 && i + 1 < n) v1 = vz[(size_t)m * n + i + 1];
          }
          if (quant & 2) {
            acc += v0 + f * (v1 - v0); /* the texture unit's lerp */
            if ((m + 1) % SLAB == 0 || m + 1 == n) { /* slab sum, scaled, added to the output */
              total += acc * scale;
              acc = 0.0f;
            }
          } else {
            acc = fmaf(g, v0, acc);
            acc = fmaf(f, v1, acc);
          }
        }
        out[k] = (quant & 2) ? total : acc * scale;
      }
    }
  }
}
