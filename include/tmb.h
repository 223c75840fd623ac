/*
 * tmb.h -- C ABI of libtmb.so: the B200 (sm_100a) parallel-beam reconstruction hot path.
 *
 * This is the drop-in boundary for dkazanc/ToMoBAR's GPU hot path.  Every entry point names
 * the reference interface it replaces (paths relative to the reference checkout).  All data
 * pointers are DEVICE pointers unless the name ends in `_host`; `stream` is a cudaStream_t
 * passed as void* (NULL = legacy default stream).  Functions return 0 on success and a
 * negative code on failure; tmb_last_error() returns the message (thread-local).
 *
 * Layouts (reference conventions, C-contiguous fp32):
 *   volume   vol [nz][n][n]      row r <-> +y, column c <-> +x   (astra_base.py:215-222)
 *   sinogram sino[nz][na][nu]    ["detY","angles","detX"]        (astra_base.py:244-255)
 * There is no CPU fallback anywhere in this library.
 */
#ifndef TMB_H
#define TMB_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMB_OK 0
#define TMB_ERR_ARG (-1)
#define TMB_ERR_CUDA (-2)
#define TMB_ERR_UNSUPPORTED (-3)

/* data-fidelity selector of tmb_residual (data_fidelities.py:28-39) */
#define TMB_FID_LS 0
#define TMB_FID_PWLS 1
#define TMB_FID_KL 2

typedef struct tmb_geom tmb_geom; /* opaque geometry + launch plan */

int tmb_version(void);
const char *tmb_last_error(void);

/* ---- geometry --------------------------------------------------------------------------
 * Replaces AstraBase._set_vol3d_geometry / _set_gpu_projection3d_parallel_geometry /
 * _setOS_indices / _set_projection3d_OS_parallel_geometry (astra_wrappers/astra_base.py:
 * 195-222, 244-255, 287-308) and supp/funcs.py:45-81 (_vec_geom_init3D).
 *   cos_t, sin_t : host, length na; the caller evaluates them in the dtype of AnglesVec
 *                  (the reference does np.cos(theta) on the user's array, funcs.py:74-81)
 *   cor          : host, length na (a scalar CoR is broadcast by the caller)
 *   os_number    : number of ordered subsets (>=1); subsets are interleaved s, s+os, ...
 *   quant8       : 1 = round the interpolation fraction to 1/256 like the CUDA texture unit
 *                  ASTRA samples through (needed for <=1e-4 parity), 0 = exact fp32 weights
 */
tmb_geom *tmb_geom_create(int nz, int n, int nu, int na, const double *cos_t, const double *sin_t,
                          const double *cor, int os_number, int quant8);
void tmb_geom_destroy(tmb_geom *g);
/* number of angles in `subset` (-1 = all angles) */
int tmb_geom_subset_size(const tmb_geom *g, int subset);
/* reference-format subset table row: newInd_Vec[subset,:] zero padded to bins (astra_base.py:195-209) */
int tmb_geom_subset_row(const tmb_geom *g, int subset, int *out_bins /* [ceil(na/os)] */);
/* the fp32 per-angle table the kernels read from constant memory: out[na][8] (host) */
int tmb_geom_table(const tmb_geom *g, float *out);
/* number of kernels the forward projection of `subset` launches (launch accounting of benchmarks) */
int tmb_geom_fp_launches(const tmb_geom *g, int subset);
/* angles per CTA the forward projection of `subset` shares one staged window between (k_fpm); 0 = one angle per
 * CTA (k_fp / k_fpq).  For tests and benchmarks: which kernel family a call runs. */
int tmb_geom_fp_group(const tmb_geom *g, int subset);
/* bytes of device scratch tmb_fp3d / tmb_bp3d / tmb_grad need.  The caller allocates it ONCE,
 * zero-fills it ONCE (the kernels keep the zero borders intact) and passes it to every call. */
size_t tmb_geom_workspace_bytes(const tmb_geom *g);

/* Debug/test switch read by tmb_geom_create: 0 = choose the forward-projector kernel by stack height
 * (k_fpq with the 32-slice-blocked layouts from 17 slices up, else k_fp), 1 = k_fp, 2 = k_fpq
 * (line-segmented for L2 reuse), 3 = k_fpq with two 32-slice groups per CTA and no segments,
 * 4 = k_fpq without segments.  Returns the old value. */
int tmb_fp_set_kernel(int mode);
/* Debug/test switch read by tmb_geom_create: forced k_fpq line-segment length (0 = sized so that a
 * segment of both marching directions fits the L2).  Returns the old value. */
int tmb_fp_set_segment(int lines);

/* ---- projector pair --------------------------------------------------------------------
 * tmb_fp3d replaces AstraBase.runAstraProj3DCuPy -> astra direct_FP3D (astra_base.py:560-606)
 *          i.e. AstraTools3D._forwprojCuPy/_forwprojOSCuPy (astra_tools3d.py:78-86)
 * tmb_bp3d replaces AstraBase.runAstraBackproj3DCuPy -> direct_BP3D (astra_base.py:518-558)
 *          i.e. AstraTools3D._backprojCuPy/_backprojOSCuPy (astra_tools3d.py:102-110)
 * subset = -1: all angles, sino is [nz][na][nu]; otherwise sino is [nz][subset_size][nu].
 */
int tmb_fp3d(tmb_geom *g, int subset, const float *vol, float *sino, void *workspace, void *stream);
int tmb_bp3d(tmb_geom *g, int subset, const float *sino, float *vol, void *workspace, void *stream);

/* Fused gradient of the data term (data_fidelities.py:7-40):
 *   grad = A_s^T ( W .* (A_s x - b_s) )          LS / PWLS
 *   grad = A_s^T ( 1 - b_s / max(A_s x, 1e-8) )  KL
 * b (and w, or NULL) are the FULL sinograms [nz][na][nu]; the subset rows are picked inside
 * the forward projector's epilogue (replaces the b[:, indVec, :] gather copy of
 * methodsIR_CuPy.py:454-457 and the temporaries of data_fidelities.py:30-39).             */
int tmb_grad(tmb_geom *g, int subset, int fidelity, const float *x, const float *b, const float *w,
             float *grad, void *workspace, void *stream);

/* Extension (no code in the reference snapshot; its legacy call sites are
 * Demos/methods_IR_legacy/DemoFISTA_artifacts2D.py:197,307-309, Demo_RealData.py:157-158,219):
 * gradient of the robust / ring-artefact data terms.  With res = A_s x - b_s:
 *   ring_rx != NULL : Group-Huber ring model, res += ring_alpha * ring_rx[z][u] (one offset per
 *                     detector pixel), ring_vec[z][u] = sum over the subset's angles of res
 *   huber_delta > 0 : res *= min(1, huber_delta / |res|)
 *   studentst_sigma > 0 : Student's-t penalty log(1 + res^2 / sigma^2): res = 2 res / (sigma^2 + res^2)
 *                     (excludes Huber)
 *   weight_mode 1   : PWLS, res *= w          weight_mode 2 : SWLS,
 *                     res = w res - w * (sum_a w res) / (sum_a w + beta_swls)
 * then grad = A_s^T res.  b, w are the full sinograms [nz][na][nu]; ring_rx / ring_vec are [nz][nu]. */
int tmb_grad_ext(tmb_geom *g, int subset, const float *x, const float *b, const float *w, int weight_mode,
                 float huber_delta, float studentst_sigma, const float *ring_rx, float ring_alpha, float beta_swls,
                 float *ring_vec, float *grad, void *workspace, void *stream);

/* ---- TV proximal operators ---------------------------------------------------------------
 * tmb_pd_tv  replaces PD_TV_cupy  (regularisersCuPy.py:170-296 + primal_dual_for_total_variation.cu)
 * tmb_rof_tv replaces ROF_TV_cupy (regularisersCuPy.py:41-167 + rudin_osher_fatemi_total_variation.cu)
 * dims: dz = 1 selects the 2-D kernels (regularisersCuPy.py:299-315 squeezes unit axes on the host).
 * `out` receives the result (may not alias `in`).  scratch: tmb_tv_workspace_bytes().           */
size_t tmb_tv_workspace_bytes(int method /*0 PD, 1 ROF*/, int dz, int dy, int dx, int half_precision);
int tmb_pd_tv(const float *in, float *out, int dz, int dy, int dx, float regularisation_parameter,
              int iterations, int methodTV, int nonneg, float lipschitz_const, int half_precision,
              void *workspace, void *stream);
int tmb_rof_tv(const float *in, float *out, int dz, int dy, int dx, float regularisation_parameter,
               int iterations, float time_marching_parameter, int half_precision, void *workspace,
               void *stream);

/* One Chambolle-Pock iteration on caller-owned ping-pong buffers: the kernel-level seam of
 * regularisersCuPy.py:255-292 (one launch per inner iteration).  p*_in / p*_out are fp32, or fp16
 * when half_precision.  For a z-SHARD of a larger volume set ghost_lo / ghost_hi: with ghost_hi
 * plane dz of u_in must exist (the next shard's first plane); with ghost_lo plane -1 of u_in and of
 * p1_in..p3_in must exist (the previous shard's last plane).  The caller refreshes those ghost planes
 * between iterations; the sharded result is then bit-identical to the whole-volume one.
 * u_lo / p*_lo / u_hi name the ghost planes explicitly (NULL = the memory adjacent to the arrays as
 * described above).  They may be PEER pointers into the neighbouring GPU's own buffers (NVLink /
 * NVSwitch P2P mapping, e.g. torch symmetric memory): the kernel then pulls its halos itself and the
 * caller only places a cross-GPU barrier between iterations.
 * Ghost planes need dx % 4 == 0 and 16-byte aligned arrays (TMB_ERR_UNSUPPORTED otherwise).       */
int tmb_pd_tv_iter(const float *in, const float *u_in, float *u_out, const void *p1_in, const void *p2_in,
                   const void *p3_in, void *p1_out, void *p2_out, void *p3_out, int dz, int dy, int dx,
                   float regularisation_parameter, int methodTV, int nonneg, float lipschitz_const,
                   int half_precision, int ghost_lo, int ghost_hi, const float *u_lo, const void *p1_lo,
                   const void *p2_lo, const void *p3_lo, const float *u_hi, void *stream);

/* TWO Chambolle-Pock iterations (two trips of the loop of regularisersCuPy.py:255-292) in one pass over
 * the same caller-owned buffers, fp32 duals: the intermediate iterate is never written out, so a pair
 * of iterations costs 36 B/voxel of HBM traffic instead of 72, and a z-shard needs its ghost planes
 * refreshed / its neighbours synchronised once per PAIR.  The pass reaches two planes deep: with
 * ghost_lo planes -2 and -1 of u_in and p1_in..p3_in and plane -1 of `in` must exist, with ghost_hi
 * planes dz and dz+1 of u_in and plane dz of p1_in..p3_in and `in`.  u_lo / p*_lo point at plane -2,
 * in_lo at plane -1, u_hi / p*_hi / in_hi at plane dz (NULL = adjacent memory; peer pointers allowed
 * as for tmb_pd_tv_iter).  Needs dx % 4 == 0, 16-byte aligned arrays and shards of >= 2 planes
 * (TMB_ERR_UNSUPPORTED otherwise).  Same arithmetic as two tmb_pd_tv_iter calls.
 * p1_in == p2_in == p3_in == NULL: the dual variable is zero everywhere (the first pair of a prox call,
 * regularisersCuPy.py:219-223 allocates it as zeros): it is read neither here nor in the ghost planes, and u_in may
 * then be the prox input itself (u_lo / u_hi: the neighbours' inputs), which saves the caller the copy of the input
 * into the primal buffer and the three memsets.                                                                    */
int tmb_pd_tv_iter2(const float *in, const float *u_in, float *u_out, const float *p1_in, const float *p2_in,
                    const float *p3_in, float *p1_out, float *p2_out, float *p3_out, int dz, int dy, int dx,
                    float regularisation_parameter, int methodTV, int nonneg, float lipschitz_const,
                    int ghost_lo, int ghost_hi, const float *u_lo, const float *p1_lo, const float *p2_lo,
                    const float *p3_lo, const float *in_lo, const float *u_hi, const float *p1_hi,
                    const float *p2_hi, const float *p3_hi, const float *in_hi, void *stream);

/* One ROF iteration on caller-owned ping-pong buffers (rudin_osher_fatemi_total_variation.cu:157-248,
 * both kernels fused; regularisersCuPy.py:112-162 launches them per iteration).  z-SHARDS: with
 * ghost_hi plane dz of u_in must exist; with ghost_lo planes -2 and -1 must exist (the normalised z
 * difference of plane -1 enters the divergence at plane 0).  Needs dx % 4 == 0 and aligned arrays. */
int tmb_rof_tv_iter(const float *in, const float *u_in, float *u_out, int dz, int dy, int dx,
                    float regularisation_parameter, float time_marching_parameter, int half_precision,
                    int ghost_lo, int ghost_hi, const float *u_lo /* planes -2,-1 */, const float *u_hi,
                    void *stream);

/* Debug/test switch: 1 routes 3-D TV through the simple one-thread-per-voxel kernels instead of
 * the z-marching ones, 2 through the CTA-tiled z-marching kernels, 3 / 4 force the register-fed / TMA-fed
 * warp-strip PD_TV kernel (one iteration per launch), 5 the kernel that does two PD_TV iterations per
 * pass, 6 its compile-time-specialised variant, 7 that variant at four CTAs per SM, 8 with its loads two rows
 * ahead, 9 = 6 without the memset / copy at the start of a prox call, 10 = 6 with
 * the next plane prefetched into L2 (0 picks the measured best; same arithmetic, used by the parity tests).  Returns the old value. */
int tmb_tv_set_simple_kernels(int enable);
/* Debug/test switch: consumer warps per CTA and ring depth (rows in flight per warp) of the TMA-fed fused PD_TV
 * kernel k_pd_tv3d_f2t (modes 11 / 12 of tmb_tv_set_simple_kernels; instantiated: 2x4, 3x4, 4x2, 4x4, 4x8, 5x2,
 * anything else selects 4x4).  Returns the old setting as warps * 10 + stages. */
int tmb_tv_set_f2t(int warps, int stages);
/* number of kernels tmb_pd_tv launches for `iterations` iterations (PD_TV_cupy's loop,
 * regularisersCuPy.py:262-294, is one launch per iteration in the reference; here pairs of
 * iterations share a launch where the fused kernel applies).  Launch accounting of benchmarks. */
int tmb_pd_tv_launches(int dz, int dy, int dx, int iterations, int half_precision);

/* ---- fused elementwise steps of the iterative loops ---------------------------------------
 * FISTA (methodsIR_CuPy.py:463-468):  X = X_t - Linv * grad ; optional max(X, 0)             */
int tmb_fista_grad_step(const float *x_t, const float *grad, float *x, size_t count, float l_inv,
                        int nonneg, void *stream);
/* FISTA (methodsIR_CuPy.py:475):  X_t = X + coef * (X - X_old)                              */
int tmb_fista_momentum(const float *x, const float *x_old, float *x_t, size_t count, float coef,
                       void *stream);
/* ADMM z-update (methodsIR_CuPy.py:545-557): z -= tau*(grad + rho*(z - x + u)); optional
 * max(z,0); optional relaxation z = (1-alpha) z_old + alpha z; z_old = z; xprox = z + u     */
int tmb_admm_z_step(float *z, float *z_old, const float *x, const float *u, const float *grad,
                    float *xprox, size_t count, float tau, float rho, int nonneg, int relax,
                    float alpha, void *stream);
/* ADMM dual update (methodsIR_CuPy.py:566): u += z - x                                        */
int tmb_admm_u_step(float *u, const float *z, const float *x, size_t count, void *stream);
/* y = a*x + y-like helpers used by Landweber/SIRT/CGLS (methodsIR_CuPy.py:164-166,211-216,279-289) */
int tmb_axpy(float a, const float *x, float *y, size_t count, int nonneg, void *stream);

/* ---- FBP filter ---------------------------------------------------------------------------
 * Builds the half-spectrum sinc filter of generate_filtersync.cu:5-82 (fourier.py:52-66):
 * f[n/2+1], device pointer.                                                                   */
int tmb_sinc_filter(float cutoff, float *f, int n, float multiplier, void *stream);
/* spectrum *= filter, in place; spec is interleaved complex64 [rows][n/2+1] (fourier.py:69)    */
int tmb_apply_filter(float *spec, const float *f, size_t rows, int nbins, void *stream);
/* edge padding of the detector axis (supp/suppTools.py:425-459 _apply_horiz_detector_padding;
 * methodsDIR_CuPy.py:505-521): out[rows][wout], out[r][j] = in[r][clamp(j - pad_left, 0, w - 1)]   */
int tmb_edge_pad(const float *in, float *out, size_t rows, int w, int wout, int pad_left, void *stream);
/* circular mask (supp/suppTools.py:364-396), in place on vol[nz][n][n]                        */
int tmb_circular_mask(float *vol, int nz, int n, float radius, void *stream);

/* flat / dark-field normalisation with negative log (supp/suppTools.py:187-264, methods "mean" /
 * "median": the caller averages the flats / darks): data[n0][n1][n2] raw projections (uint16 when
 * data_is_u16, else fp32) with the angle axis 0 or 1; flat_mean / dark_mean [.][n2]; out fp32.     */
int tmb_normalise(const void *data, int data_is_u16, const float *flat_mean, const float *dark_mean,
                  float *out, int n0, int n1, int n2, int angle_axis, int take_log, void *stream);

/* ---- FOURIER_INV (USFFT gridding) kernels ----------------------------------------------------
 * Replace the default centre-gather path of RecToolsDIRCuPy.FOURIER_INV (methodsDIR_CuPy.py:152-447)
 * and cuda_kernels/fft_us_kernels.cu; the FFTs themselves are cuFFT calls made by the host.
 * Complex arrays are interleaved float pairs.  nz2 = number of complex slices (= slices / 2).
 *   tmb_fi_pack         : tmp_p[2*nz2][nproj][n] -> datac[nz2][nproj][n], x (-1)^(x+1)  (r2c_c1dfftshift :529-557)
 *   tmb_fi_scale_sign   : datac *= c * (-1)^(x+1), in place                              (c1dfftshift :559-586)
 *   tmb_fi_gather       : polar samples datac -> fde[nz2][2n][2n], ALREADY multiplied by (-1)^(x+y) (the first
 *                         c2dfftshift of :860-868 is fused); theta = -angles (device), sorted_theta / sorted_idx =
 *                         ascending sort of theta and its permutation (int32)
 *                         (gather_kernel_center_angle_based_prune :193-319 + gather_kernel_center :468-527)
 *   tmb_fi_sign2d       : fde *= (-1)^(x+y)                                              (c2dfftshift :588-609)
 *   tmb_fi_unpad        : (-1)^(x+y) of the second c2dfftshift (:888-896, fused), crop, de-apodise, unpack re/im
 *                         -> recon[unpad_z][R][R]; fde is multiplied by `scale` first (1, or 1 / (2n)^2 after an
 *                         unnormalised inverse 2-D FFT)                                   (unpadding_mul_phi :611-657) */
int tmb_fi_pack(const float *tmp_p, float *datac, int n, int nproj, int nz2, void *stream);
/* tmb_fi_pack reading rows of pitch row_pitch (floats) inside slices of pitch slice_pitch: packs a chunk of slices straight
 * out of the oversampled filter output (`in` points at the first kept detector sample), so that the crop of
 * methodsDIR_CuPy.py:541-545 and the pack are one pass */
int tmb_fi_pack_rows(const float *in, size_t row_pitch, size_t slice_pitch, float *datac, int n, int nproj, int nz2,
                     void *stream);
/* FOURIER_INV filters slice PAIRS as complex rows (real impulse response: real and imaginary part are filtered
 * independently; one c2c transform instead of an r2c / c2r pair per slice):
 *   tmb_edge_pad_pair : in[2 nzc][rows][w] -> out[nzc][rows][wout] complex, (slice 2t, slice 2t+1) edge-padded like tmb_edge_pad
 *   tmb_fi_crop_sign  : complex rows of pitch row_pitch (complex samples; `in` points at the first kept one) ->
 *                       datac[rows][n] * (-1)^(x+1): the crop of :541-545 and r2c_c1dfftshift (:529-557) in one pass */
int tmb_edge_pad_pair(const float *in, float *out, int nzc, size_t rows, int w, int wout, int pad_left, void *stream);
int tmb_fi_crop_sign(const float *in, size_t row_pitch, float *datac, int n, size_t rows, void *stream);
int tmb_fi_scale_sign(float *datac, float c, int n, int nproj, int nz2, void *stream);
/* tmb_fi_scale_sign out of place into the slice-PAIR layout dataz[nz2 / 2][nproj][n] of (slice 2t, slice 2t + 1) (nz2 even),
 * and the whole-grid gather reading it (tmb_fi_gather with one 128-bit load per two slices; nz2 a multiple of 8):
 * c1dfftshift :559-586 and gather_kernel_center :468-527 as above, same bits */
int tmb_fi_scale_sign_pairs(const float *datac, float *dataz, float c, int n, int nproj, int nz2, void *stream);
int tmb_fi_gather_pairs(const float *dataz, float *fde, const float *theta, const float *sorted_theta,
                        const int *sorted_idx, int m, float mu, int n, int nproj, int nz2, void *stream);
/* test hook: 1 = k_fi_gather (every thread walks its own polar lines), 2 = k_fi_gather_s (samples of a tile staged in
 * shared memory), 3 = k_fi_gather_w (a warp walks the lines of its 8 x 4 patch in lock step), 0 = the measured best (3).
 * Returns the old value. */
int tmb_fi_set_gather(int mode);
/* test hook: complex slices per thread of k_fi_gather_w (2, 4, 8 or 16; 0 = the default).  Returns the old value. */
int tmb_fi_set_slices_per_thread(int sc);
int tmb_fi_gather(const float *datac, float *fde, const float *theta, const float *sorted_theta,
                  const int *sorted_idx, int m, float mu, int n, int nproj, int nz2, void *stream);
/* The non-default branches of FOURIER_INV (methodsDIR_CuPy.py:759-835, taken for the keyword center_size < 2n):
 *   tmb_fi_gather_center : tmb_fi_gather restricted to the centre square of center_size x center_size grid points
 *                          (gather_kernel_center with center_size < 2n, fft_us_kernels.cu:468-527)
 *   tmb_fi_scatter       : every polar sample spreads its (2m+1)^2 Gaussian footprint onto the grid with atomic adds
 *                          (gather_kernel :104-109 when center_size == 0: the whole grid, the reference's branch for
 *                          center_size < 192; gather_kernel_partial :98-102 otherwise: only outside the centre square).
 *                          fde must be zero where it adds; what is added carries the (-1)^(x+y) like tmb_fi_gather.   */
int tmb_fi_gather_center(const float *datac, float *fde, const float *theta, const float *sorted_theta,
                         const int *sorted_idx, int m, float mu, int n, int nproj, int nz2, int center_size, void *stream);
int tmb_fi_scatter(const float *datac, float *fde, const float *theta, int m, float mu, int center_size, int n,
                   int nproj, int nz2, void *stream);
int tmb_fi_sign2d(float *fde, int n, int nz2, void *stream);
int tmb_fi_unpad(float *recon, const float *fde, float mu, float scale, int nproj, int unpad_recon_p, int unpad_z,
                 int unpad_recon_m, int n, int nz2, void *stream);

/* ---- host-buffer entry points (what a non-CUDA caller binds; H2D/D2H inside) --------------- */
int tmb_fp3d_host(tmb_geom *g, int subset, const float *vol_host, float *sino_host);
int tmb_bp3d_host(tmb_geom *g, int subset, const float *sino_host, float *vol_host);

#ifdef __cplusplus
}
#endif
#endif /* TMB_H */
