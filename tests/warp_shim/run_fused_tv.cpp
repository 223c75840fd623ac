// Host runner of the fused PD_TV kernels under the warp shim (see cuda_shim.h).
#define TMB_HOST_SHIM 1
#include "cuda_shim.h"

#include <thread>
#include <vector>

thread_local shim_uint3 threadIdx, blockIdx;
thread_local ShimWarp *shim_warp = nullptr;
thread_local int shim_lane = 0;
std::barrier<> *shim_cta_bar = nullptr;

#include "../../tomobar_b200/csrc/tmb_tv_fused.cuh"

namespace tmb {
// lane-private slots of every warp, then (k_pd_tv3d_f2t) the TMA ring of every warp
alignas(16) unsigned char f2_smem[f2t_smem_bytes(F2_WARPS, 8)];
}

namespace {
struct Args {
  const float *in, *U;
  float *Uo;
  const float *P1, *P2, *P3;
  float *Q1, *Q2, *Q3;
  float sigma, tau, lt, theta;
  int dx, dy, dz, zrun;
  tmb::F2Ghost<true> gh;
};

template <bool NN, bool AN> void lane_entry(int variant, const Args &a) {
  using namespace tmb;
  switch (variant) {
    case 0:
      k_pd_tv3d_f2<NN, AN>(a.in, a.U, a.Uo, a.P1, a.P2, a.P3, a.Q1, a.Q2, a.Q3, a.sigma, a.tau, a.lt, a.theta, a.dx, a.dy,
                           a.dz, a.zrun);
      break;
    case 1:
      k_pd_tv3d_f2s<NN, AN, false, 3>(a.in, a.U, a.Uo, a.P1, a.P2, a.P3, a.Q1, a.Q2, a.Q3, a.sigma, a.tau, a.lt, a.theta,
                                      a.dx, a.dy, a.dz, a.zrun, F2Ghost<false>{});
      break;
    case 2:
      k_pd_tv3d_f2s<NN, AN, false, 4>(a.in, a.U, a.Uo, a.P1, a.P2, a.P3, a.Q1, a.Q2, a.Q3, a.sigma, a.tau, a.lt, a.theta,
                                      a.dx, a.dy, a.dz, a.zrun, F2Ghost<false>{});
      break;
    case 4:
      k_pd_tv3d_f2s<NN, AN, false, 3, 2>(a.in, a.U, a.Uo, a.P1, a.P2, a.P3, a.Q1, a.Q2, a.Q3, a.sigma, a.tau, a.lt,
                                         a.theta, a.dx, a.dy, a.dz, a.zrun, F2Ghost<false>{});
      break;
    case 5:
      k_pd_tv3d_f2s<NN, AN, false, 3, 1, true>(a.in, a.U, a.Uo, a.P1, a.P2, a.P3, a.Q1, a.Q2, a.Q3, a.sigma, a.tau, a.lt,
                                               a.theta, a.dx, a.dy, a.dz, a.zrun, F2Ghost<false>{});
      break;
    case 6:
      k_pd_tv3d_f2s<NN, AN, false, 3, 1, false, true>(a.in, a.U, a.Uo, a.P1, a.P2, a.P3, a.Q1, a.Q2, a.Q3, a.sigma, a.tau,
                                                      a.lt, a.theta, a.dx, a.dy, a.dz, a.zrun, F2Ghost<false>{});
      break;
    case 7:
      k_pd_tv3d_f2t<NN, AN, false, F2_WARPS, 4>(a.in, a.U, a.Uo, a.P1, a.P2, a.P3, a.Q1, a.Q2, a.Q3, a.sigma, a.tau, a.lt,
                                                a.theta, a.dx, a.dy, a.dz, a.zrun, F2Ghost<false>{});
      break;
    case 8:
      k_pd_tv3d_f2t<NN, AN, true, F2_WARPS, 4>(a.in, a.U, a.Uo, a.P1, a.P2, a.P3, a.Q1, a.Q2, a.Q3, a.sigma, a.tau, a.lt,
                                               a.theta, a.dx, a.dy, a.dz, a.zrun, a.gh);
      break;
    case 9:
      k_pd_tv3d_f2t<NN, AN, false, F2_WARPS, 2, true>(a.in, a.U, a.Uo, a.P1, a.P2, a.P3, a.Q1, a.Q2, a.Q3, a.sigma, a.tau,
                                                      a.lt, a.theta, a.dx, a.dy, a.dz, a.zrun, F2Ghost<false>{});
      break;
    case 10:
      k_pd_tv3d_f2t<NN, AN, false, F2_WARPS, 8>(a.in, a.U, a.Uo, a.P1, a.P2, a.P3, a.Q1, a.Q2, a.Q3, a.sigma, a.tau, a.lt,
                                                a.theta, a.dx, a.dy, a.dz, a.zrun, F2Ghost<false>{});
      break;
    case 11:
      k_pd_tv3d_f2s<NN, AN, true, 3, 1, true>(a.in, a.U, a.Uo, a.P1, a.P2, a.P3, a.Q1, a.Q2, a.Q3, a.sigma, a.tau, a.lt,
                                              a.theta, a.dx, a.dy, a.dz, a.zrun, a.gh);
      break;
    case 12:  // strips of 8 rows, two CTAs of four warps per SM
      k_pd_tv3d_f2s<NN, AN, false, 2, 1, false, false, 0, 0, 8, 4>(a.in, a.U, a.Uo, a.P1, a.P2, a.P3, a.Q1, a.Q2, a.Q3,
                                                                  a.sigma, a.tau, a.lt, a.theta, a.dx, a.dy, a.dz,
                                                                  a.zrun, F2Ghost<false>{});
      break;
    case 13:  // strips of 6 rows, five warps per CTA, first pass of a prox call
      k_pd_tv3d_f2s<NN, AN, false, 2, 1, true, false, 0, 0, 6, 5>(a.in, a.U, a.Uo, a.P1, a.P2, a.P3, a.Q1, a.Q2, a.Q3,
                                                                 a.sigma, a.tau, a.lt, a.theta, a.dx, a.dy, a.dz,
                                                                 a.zrun, F2Ghost<false>{});
      break;
    case 14:  // strips of 8 rows, packets two rows ahead
      k_pd_tv3d_f2s<NN, AN, false, 2, 2, false, false, 0, 0, 8, 4>(a.in, a.U, a.Uo, a.P1, a.P2, a.P3, a.Q1, a.Q2, a.Q3,
                                                                  a.sigma, a.tau, a.lt, a.theta, a.dx, a.dy, a.dz,
                                                                  a.zrun, F2Ghost<false>{});
      break;
    default:
      k_pd_tv3d_f2s<NN, AN, true, 3>(a.in, a.U, a.Uo, a.P1, a.P2, a.P3, a.Q1, a.Q2, a.Q3, a.sigma, a.tau, a.lt, a.theta,
                                     a.dx, a.dy, a.dz, a.zrun, a.gh);
  }
}
}  // namespace

// variant: 0 k_pd_tv3d_f2, 1 k_pd_tv3d_f2s, 2 k_pd_tv3d_f2s at four CTAs per SM, 3 k_pd_tv3d_f2s<GHOST>,
// 4 k_pd_tv3d_f2s with packets two rows ahead, 5 k_pd_tv3d_f2s<PZERO> (dual variable zero on entry, not read),
// 6 k_pd_tv3d_f2s<L2PF> (prefetches are no-ops on the host: this checks their address arithmetic compiles and the
//   rest of the kernel is untouched), 7 k_pd_tv3d_f2t (TMA-fed ring of 4 stages), 8 k_pd_tv3d_f2t<GHOST>,
// 9 k_pd_tv3d_f2t<PZERO> with 2 stages, 10 k_pd_tv3d_f2t with a ring of 8 stages,
// 11 k_pd_tv3d_f2s<GHOST, PZERO> (first pair of a sharded prox call: no dual variable read, here or in the ghosts),
// 12 / 14 k_pd_tv3d_f2s on strips of 8 rows (14: packets two rows ahead), 13 strips of 6 rows x 5 warps with PZERO
extern "C" int shim_run_fused_tv(int variant, int nonneg, int aniso, const float *in, const float *U, float *Uo,
                                 const float *P1, const float *P2, const float *P3, float *Q1, float *Q2, float *Q3,
                                 float sigma, float tau, float lt, float theta, int dx, int dy, int dz, int zrun,
                                 int ghost_lo, int ghost_hi, const float *U_lo, const float *P1_lo, const float *P2_lo,
                                 const float *P3_lo, const float *in_lo, const float *U_hi, const float *P1_hi,
                                 const float *P2_hi, const float *P3_hi, const float *in_hi) {
  Args a{in, U, Uo, P1, P2, P3, Q1, Q2, Q3, sigma, tau, lt, theta, dx, dy, dz, zrun, {}};
  a.gh.lo = ghost_lo; a.gh.hi = ghost_hi;
  a.gh.U_lo = U_lo; a.gh.P1_lo = P1_lo; a.gh.P2_lo = P2_lo; a.gh.P3_lo = P3_lo; a.gh.in_lo = in_lo;
  a.gh.U_hi = U_hi; a.gh.P1_hi = P1_hi; a.gh.P2_hi = P2_hi; a.gh.P3_hi = P3_hi; a.gh.in_hi = in_hi;
  // rows per strip and warps per CTA of the variant
  const int vs = (variant == 12 || variant == 14) ? 8 : (variant == 13 ? 6 : tmb::F2_S);
  const int vw = variant == 13 ? 5 : tmb::F2_WARPS;
  const int gx = (dx + tmb::F2_OUT - 1) / tmb::F2_OUT, gy = (dy + vs * vw - 1) / (vs * vw);
  const int gz = (dz + zrun - 1) / zrun;
  auto lane_body = [&](int warp, int lane, int bx, int by, int bz, ShimWarp *w) {
    threadIdx = {unsigned(warp * 32 + lane), 0, 0};
    blockIdx = {unsigned(bx), unsigned(by), unsigned(bz)};
    shim_warp = w;
    shim_lane = lane;
    if (nonneg) { if (aniso) lane_entry<true, true>(variant, a); else lane_entry<true, false>(variant, a); }
    else { if (aniso) lane_entry<false, true>(variant, a); else lane_entry<false, false>(variant, a); }
  };
  const bool cta_wide = variant >= 7 && variant <= 10;  // k_pd_tv3d_f2t: consumer warps + the producer warp together
  for (int bz = 0; bz < gz; ++bz)
    for (int by = 0; by < gy; ++by)
      for (int bx = 0; bx < gx; ++bx) {
        if (cta_wide) {
          constexpr int NW = tmb::F2_WARPS + 1;
          std::memset(tmb::f2_smem, 0xff, sizeof(tmb::f2_smem));  // NaN-poison the slots and the ring
          std::barrier<> cta_bar(NW * 32);
          shim_cta_bar = &cta_bar;
          std::vector<ShimWarp> warps(NW);
          std::vector<std::thread> lanes;
          for (int warp = 0; warp < NW; ++warp)
            for (int lane = 0; lane < 32; ++lane)
              lanes.emplace_back([&, warp, lane] { lane_body(warp, lane, bx, by, bz, &warps[warp]); });
          for (auto &t : lanes) t.join();
          continue;
        }
        for (int warp = 0; warp < vw; ++warp) {
          std::memset(tmb::f2_smem, 0xff, sizeof(tmb::f2_smem));  // NaN-poison the slots
          ShimWarp w;
          std::vector<std::thread> lanes;
          for (int lane = 0; lane < 32; ++lane)
            lanes.emplace_back([&, warp, lane] { lane_body(warp, lane, bx, by, bz, &w); });
          for (auto &t : lanes) t.join();
        }
      }
  return 0;
}
