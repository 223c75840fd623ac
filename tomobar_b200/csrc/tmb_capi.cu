// Geometry handle, error plumbing and host-buffer entry points of libtmb.so.
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "tmb_common.h"

namespace tmb {
static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int subset_size(const tmb_geom *g, int subset);
extern int g_fpq_mode;
static std::atomic<uint64_t> g_next_id{1};
// test hook: which forward-projector kernel geometries created from now on use
// (0 = by stack height, 1 = k_fp, 2 = k_fpq)
static int g_fp_kernel = 0;
static int g_fp_segment = 0;  // test hook: forced k_fpq segment length in lines (0 = sized for L2)
}  // namespace tmb

using namespace tmb;

extern "C" int tmb_version(void) { return 100; }
extern "C" int tmb_fp_set_kernel(int mode) {
  const int old = g_fp_kernel == 2 ? g_fpq_mode : g_fp_kernel;
  g_fp_kernel = (mode == 1) ? 1 : ((mode >= 2 && mode <= 7) ? 2 : 0);
  g_fpq_mode = (mode >= 2 && mode <= 7) ? mode : 0;
  return old;
}
extern "C" const char *tmb_last_error(void) { return g_err.c_str(); }
extern "C" int tmb_fp_set_segment(int lines) {
  const int old = g_fp_segment;
  g_fp_segment = lines > 0 ? lines : 0;
  return old;
}

// Per-angle fp32 table derived in double from the parallel3d_vec vectors of
// supp/funcs.py:45-81: ray (sin, -cos, 0), detector centre CoR*(cos, sin, 0), u (cos, sin, 0).
//   [0] cos [1] sin [2] bp_off = -CoR + nu/2 - 1/2
//   [3] alpha = -minor/major  [4] b0  [5] bstep = 1/major  [6] scale = sqrt(1+alpha^2)
//   [7] dir: 0 when |sin| > |cos| (march along columns, interpolate along rows), else 1
static void fill_table(float *tbl, int n, int nu, int na, const double *c, const double *s, const double *cor) {
  for (int a = 0; a < na; ++a) {
    const double ca = c[a], sa = s[a], co = cor[a];
    const bool dirx = std::fabs(sa) > std::fabs(ca);
    const double major = dirx ? sa : ca, minor = dirx ? ca : sa;
    const double alpha = -minor / major;
    float *t = tbl + (size_t)a * 8;
    t[0] = (float)ca;
    t[1] = (float)sa;
    t[2] = (float)(-co + (nu / 2.0 - 0.5));
    t[3] = (float)alpha;
    t[4] = (float)((-nu / 2.0 + 0.5 + co) / major + (n / 2.0 - 0.5));
    t[5] = (float)(1.0 / major);
    t[6] = (float)std::sqrt(1.0 + alpha * alpha);
    t[7] = dirx ? 0.f : 1.f;
  }
}

extern "C" tmb_geom *tmb_geom_create(int nz, int n, int nu, int na, const double *cos_t, const double *sin_t,
                                     const double *cor, int os_number, int quant8) {
  if (nz <= 0 || n <= 0 || nu <= 0 || na <= 0 || !cos_t || !sin_t || !cor || os_number <= 0) {
    set_error("tmb_geom_create: sizes must be positive and tables non-null");
    return nullptr;
  }
  if (os_number > na) {
    set_error("tmb_geom_create: more ordered subsets than angles");
    return nullptr;
  }
  tmb_geom *g = new tmb_geom();
  g->d.nz = nz; g->d.n = n; g->d.nu = nu; g->d.na = na;
  g->d.nzc = round_up((nz + ZC - 1) / ZC, NZC);
  g->d.up = nu + 2 * SPAD;
  g->d.qp = n + 2 * VPAD;
  g->os_number = os_number;
  g->quant8 = quant8 ? 1 : 0;
  g->bins = (na + os_number - 1) / os_number;
  g->table = static_cast<float *>(std::malloc(sizeof(float) * 8 * (size_t)na));
  fill_table(g->table, n, nu, na, cos_t, sin_t, cor);
  g->id = g_next_id.fetch_add(1);
  g->d.nzg = (nz + ZC * FQ_CG - 1) / (ZC * FQ_CG);
  g->d.qpq = n + 2 * QPAD;
  g->fp_q = g_fp_kernel == 1 ? 0 : (g_fp_kernel == 2 ? 1 : (nz >= FQ_MIN_NZ ? 1 : 0));
  const size_t vbytes = g->fp_q ? sizeof(float4) * FQ_CG * (size_t)g->d.nzg * n * g->d.qpq
                                : sizeof(float4) * (size_t)g->d.nzc * n * g->d.qp;
  const size_t sbytes = sizeof(float4) * (size_t)g->d.nzc * na * g->d.up;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  g->off_v0 = 0;
  g->off_v1 = al(vbytes);
  g->off_s = g->off_v1 + al(vbytes);
  g->off_part = g->off_s + al(sbytes);
  // k_fpq line segments: the marching range is cut so that one segment (both marching directions of
  // one z-group) is ~190 MB.  Measured optimum (profiles/fp_segments_r01.txt): at N = 2048 segments of
  // 243-324 lines (139-186 MB) give 77-78 ms per 75-angle subset against 92 ms unsegmented and 83 ms
  // with 81-line (46 MB) segments; at N = 1024 anything from 200 MB up is within 0.5 % of unsegmented.
  g->seg_len = n; g->nseg = 1; g->part_angles = 0;
  size_t pbytes = 0;
  if (g->fp_q) {
    const double seg_bytes_per_line = 2.0 * g->d.qpq * FQ_CG * sizeof(float4);
    int sl = g_fp_segment > 0 ? g_fp_segment : (int)(190.0e6 / seg_bytes_per_line);
    sl = sl < 24 ? 24 : sl;
    sl -= sl % 3;  // multiple of the lines per pipeline stage
    if (sl < n) {
      g->seg_len = sl;
      g->nseg = (n + sl - 1) / sl;
      const int tiles = (nu + FQ_K - 1) / FQ_K;
      const size_t per_angle = sizeof(float4) * (size_t)g->nseg * g->d.nzg * FQ_CG * tiles * FQ_K;
      const int sub_max = (na + os_number - 1) / os_number;
      size_t cap = (size_t)1 << 31;  // 2 GiB of partial sums at most
      int pa = (int)(cap / per_angle);
      pa = pa < 1 ? 1 : (pa > sub_max ? sub_max : pa);
      g->part_angles = pa;
      pbytes = per_angle * pa;
    }
  }
  g->ws_bytes = g->off_part + al(pbytes);
  return g;
}

extern "C" void tmb_geom_destroy(tmb_geom *g) {
  if (!g) return;
  std::free(g->table);
  delete g;
}

extern "C" int tmb_geom_subset_size(const tmb_geom *g, int subset) {
  if (!g || subset < -1 || subset >= g->os_number) return TMB_ERR_ARG;
  return subset_size(g, subset);
}

extern "C" int tmb_geom_subset_row(const tmb_geom *g, int subset, int *out_bins) {
  TMB_REQUIRE(g && out_bins && subset >= 0 && subset < g->os_number, "tmb_geom_subset_row: bad argument");
  for (int p = 0; p < g->bins; ++p) {
    const int idx = subset + p * g->os_number;
    out_bins[p] = idx < g->d.na ? idx : 0;
  }
  return TMB_OK;
}

extern "C" int tmb_geom_table(const tmb_geom *g, float *out) {
  TMB_REQUIRE(g && out, "tmb_geom_table: null argument");
  std::memcpy(out, g->table, sizeof(float) * 8 * (size_t)g->d.na);
  return TMB_OK;
}

extern "C" size_t tmb_geom_workspace_bytes(const tmb_geom *g) { return g ? g->ws_bytes : 0; }

// ---- host-buffer entry points ---------------------------------------------------------------
namespace {
struct DevBuf {
  void *p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) {
    TMB_CUDA_CHECK(cudaMalloc(&p, bytes));
    return TMB_OK;
  }
};
}  // namespace

static int host_op(tmb_geom *g, int subset, const float *in_host, float *out_host, bool forward) {
  TMB_REQUIRE(g && in_host && out_host, "host op: null argument");
  TMB_REQUIRE(subset >= -1 && subset < g->os_number, "host op: subset out of range");
  const size_t nvol = (size_t)g->d.nz * g->d.n * g->d.n;
  const size_t nsino = (size_t)g->d.nz * subset_size(g, subset) * g->d.nu;
  DevBuf dvol, dsino, dws;
  int rc;
  if ((rc = dvol.alloc(nvol * 4)) || (rc = dsino.alloc(nsino * 4)) || (rc = dws.alloc(g->ws_bytes))) return rc;
  TMB_CUDA_CHECK(cudaMemset(dws.p, 0, g->ws_bytes));
  if (forward) {
    TMB_CUDA_CHECK(cudaMemcpy(dvol.p, in_host, nvol * 4, cudaMemcpyHostToDevice));
    if ((rc = tmb_fp3d(g, subset, (const float *)dvol.p, (float *)dsino.p, dws.p, nullptr))) return rc;
    TMB_CUDA_CHECK(cudaMemcpy(out_host, dsino.p, nsino * 4, cudaMemcpyDeviceToHost));
  } else {
    TMB_CUDA_CHECK(cudaMemcpy(dsino.p, in_host, nsino * 4, cudaMemcpyHostToDevice));
    if ((rc = tmb_bp3d(g, subset, (const float *)dsino.p, (float *)dvol.p, dws.p, nullptr))) return rc;
    TMB_CUDA_CHECK(cudaMemcpy(out_host, dvol.p, nvol * 4, cudaMemcpyDeviceToHost));
  }
  return TMB_OK;
}

extern "C" int tmb_fp3d_host(tmb_geom *g, int subset, const float *vol_host, float *sino_host) {
  return host_op(g, subset, vol_host, sino_host, true);
}
extern "C" int tmb_bp3d_host(tmb_geom *g, int subset, const float *sino_host, float *vol_host) {
  return host_op(g, subset, sino_host, vol_host, false);
}
