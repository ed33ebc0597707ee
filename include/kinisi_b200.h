/*
 * kinisi_b200.h -- C ABI of the B200-native displacement-statistics path.
 *
 * The reference (bjmorgan/kinisi 1.1.1) is pure Python/numpy and has no FFI layer; its boundary is
 * the Python API of kinisi/parser.py and kinisi/diffusion.py.  This header declares the entry points
 * a ctypes binding for that path calls (see INTEGRATION.md for the reference-side stub).  Each entry
 * point cites the reference code it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - Every pointer is a DEVICE pointer owned by the caller unless the name ends in _host.
 *   - `stream` is a cudaStream_t passed as void*.  Calls only enqueue work on that stream; none of
 *     them synchronises, allocates or keeps state between calls (re-entrant per stream).
 *   - Return value: 0 on success, non-zero on error; kb2_last_error() returns a thread-local
 *     message for the last non-zero return on the calling thread.
 *   - Trajectory arrays: cumulative unwrapped displacements `dc` are float64, component-major
 *     ("SoA"): dc[(atom*3 + component)*pitch + frame], pitch >= n_frames, pitch % 2 == 0, the
 *     pad [n_frames, pitch) is zero.
 */
#ifndef KINISI_B200_H
#define KINISI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KB2_ABI_VERSION 11

/* coordinate element types */
#define KB2_F32 0
#define KB2_F64 1

/* input meaning for the stage-1 kernels */
#define KB2_MODE_FRACTIONAL 0 /* fractional coordinates + lattices: unwrap, then cumulative sum     */
#define KB2_MODE_DISPLACEMENT 1 /* already-unwrapped cumulative displacement (Parser(disp, ...))      */

/* dimension mask bits (kinisi/diffusion.py:23-38 DIMENSIONALITY) */
#define KB2_DIM_X 1
#define KB2_DIM_Y 2
#define KB2_DIM_Z 4

int kb2_abi_version(void);
const char* kb2_last_error(void);

/* Device properties the host side sizes its launches with (SM count, max opt-in smem). */
int kb2_device_info(int* sm_count, int* max_smem_optin, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * Stage 1 -- displacement construction.  Replaces Parser.get_disp (kinisi/parser.py:101-129),
 * Parser.correct_drift (:131-148) and the layout change feeding Parser.get_disps (:170-222).
 *
 * coords: [n_atoms_total][n_frames_in][3] (atom-major, as np.concatenate(coords, axis=1) yields at
 *         parser.py:120), element type `dtype`.  In KB2_MODE_FRACTIONAL n_frames_in includes the
 *         duplicated first frame (parser.py:323-324) and the result has n_frames_in-1 frames; in
 *         KB2_MODE_DISPLACEMENT the result has n_frames_in frames.
 * latt / latt_inv: [n_frames_in][3][3] float64 row-vector convention (cart = frac @ latt),
 *         `latt_stride` = 9 for per-frame lattices or 0 when every frame shares latt[0].
 * ------------------------------------------------------------------------------------------- */

/* Frame-major -> atom-major (replaces np.concatenate(coords, axis=1), parser.py:120, and the per-frame
 * loops of the front-ends, :309-325, 458-473): frames[n_frames][n_atoms][3] goes to
 * coords_out[(a * n_frames_out + frame_offset + f) * 3 + k]; duplicate_first != 0 also writes frame 0 to
 * frame_offset - 1 (the front-ends insert a copy of the first frame, parser.py:323-324).  Blocks of frames may be
 * transposed by successive calls (frame_offset).  Both arrays have the element type `dtype`. */
int kb2_frames_to_atoms(const void* frames, int dtype, int n_atoms, int n_frames, void* coords_out, int64_t n_frames_out,
                        int64_t frame_offset, int duplicate_first, void* stream);

/* The same transposition for CARTESIAN frames (the MDAnalysis front-end, parser.py:620-638, whose per-frame
 * np.dot(positions, np.linalg.inv(cell)) runs on the host in the reference): coords_out (always float64) receives
 * frames[f][a] . cell_inv[f] (row vector times matrix); cell_inv[n_frames][3][3] with cell_stride = 9, or one matrix
 * with cell_stride = 0. */
int kb2_frames_to_atoms_fractional(const void* frames, int dtype, int n_atoms, int n_frames, const double* cell_inv,
                                   int cell_stride, double* coords_out, int64_t n_frames_out, int64_t frame_offset,
                                   int duplicate_first, void* stream);

/* Pseudo-centre-of-mass of groups of atoms in fractional coordinates (_get_molecules, parser.py:674-738,
 * arXiv:2501.14578): members[n_molecules][molecule_size] are 0-based atom rows of coords[n_atoms_total][n_frames_in][3],
 * weights[molecule_size] the masses (NULL: centre of geometry); centres[n_molecules][n_frames_in][3] float64 in [0, 1). */
int kb2_molecule_centres(const void* coords, int dtype, int n_atoms_total, int64_t n_frames_in, const int32_t* members,
                         int n_molecules, int molecule_size, const double* weights, double* centres, void* stream);

/* inv[f] = latt[f]^-1 for f < n (np.linalg.inv at parser.py:121). */
int kb2_invert_lattice(const double* latt, double* latt_inv, int n, void* stream);

/* frame_sums[c*pitch + t] = sum_{i < n_idx} w[i] * u[idx[i]][t][c], where u is the per-frame
 * unwrapped displacement (MODE_FRACTIONAL; its cumulative sum over t is the summed displacement)
 * or the displacement itself (MODE_DISPLACEMENT).  w == NULL means unit weights.  The buffer
 * [3][pitch] is zeroed by the call.  Used for the framework drift (parser.py:143-145) and for the
 * collective / charge-weighted sums of MSTD / MSCD (diffusion.py:659, 751). */
int kb2_frame_sums(const void* coords, int dtype, int mode, int n_atoms_total, int n_frames_in, const double* latt,
                   const double* latt_inv, int latt_stride, const int32_t* idx, const double* w, int n_idx,
                   double* frame_sums, int64_t pitch, void* stream);

/* In place on buf[3][pitch] (n_frames valid entries per row):
 *   cumulative != 0:  buf <- cumsum_t(buf)
 *   then              buf <- buf / divide - sub_scale * sub      (sub may be NULL)
 * Turns frame sums into the framework mean displacement (divide = n_framework) or into the
 * drift-corrected collective sum (divide = 1, sub = framework mean, sub_scale = sum of weights). */
int kb2_scan_frames(double* buf, int n_frames, int64_t pitch, int cumulative, double divide, const double* sub,
                    double sub_scale, void* stream);

/* dc_out[(i*3 + c)*pitch + t] = cumulative unwrapped displacement of atom idx[i], drift-corrected: the
 * per-frame quantity (unwrapped step in MODE_FRACTIONAL, displacement in MODE_DISPLACEMENT) minus
 * step_scale * step_sums[c*pitch + t], summed over time in MODE_FRACTIONAL.  With step_sums = kb2_frame_sums of the
 * framework atoms and step_scale = 1/n_framework this is dc - mean(dc[framework]) of parser.py:143-145 (the sum over
 * atoms commutes with the sum over time); step_sums == NULL: no correction.  dc[:, 0] == 0 in MODE_FRACTIONAL
 * exactly as in the reference.  `workspace`: kb2_unwrap_cumsum_workspace_bytes(n_idx, pitch) bytes of device memory
 * (the tiles of a trajectory exchange their totals through it); its contents need not be preserved between calls. */
size_t kb2_unwrap_cumsum_workspace_bytes(int n_idx, int64_t pitch);
int kb2_unwrap_cumsum(const void* coords, int dtype, int mode, int n_atoms_total, int n_frames_in, const double* latt,
                      const double* latt_inv, int latt_stride, const int32_t* idx, int n_idx, const double* step_sums,
                      double step_scale, double* dc_out, int64_t pitch, void* workspace, void* stream);

/* out[c*pitch + t] = sum_{a < n_atoms} w[a] * dc[(a*3 + c)*pitch + t]  (w == NULL: unit weights).
 * The collective displacement of MSTD (np.sum(disp_slice, axis=0), kinisi/diffusion.py:659) and the
 * charge-weighted one of MSCD (:751), formed once per trajectory instead of once per dt.  The buffer
 * [3][pitch] is zeroed by the call. */
int kb2_atom_sum(const double* dc, int n_atoms, int64_t pitch, const double* w, double* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 2 -- moment reductions, fused with Parser.get_disps so the O(N*T*n_dt) displacement list
 * (parser.py:208-216) is never materialised.  Replaces the default branch of
 * MSDBootstrap.__init__ (kinisi/diffusion.py:571-602) and the per-atom part of MSTD/MSCD.
 *
 * For every lag k = lags[i] (ascending, 1 <= k <= n_frames):
 *   sums[2i]   = sum over atoms a and observations of d^2
 *   sums[2i+1] = sum of d^4
 * where the observations of atom a are dc[a][k-1] (the extra first observation, parser.py:209) and
 * dc[a][j+k] - dc[a][j] for 0 <= j < n_frames-k, and d^2 sums the components in dim_mask.
 * Calling it with n_atoms = 1 on a collective sum [3][pitch] yields the MSTD/MSCD sums
 * (diffusion.py:659, 683-684, 751, 775-776).
 *
 * variant: 0 = direct kernel (L1-cached partner loads), 1 = TMA-staged ring-buffer kernel
 *          (the arithmetic-run kernel, variant 2, needs the lags on the host: kb2_self_moments_runs).
 * workspace: kb2_self_moments_workspace_bytes(n_lags) bytes, contents irrelevant on entry.
 * ------------------------------------------------------------------------------------------- */
size_t kb2_self_moments_workspace_bytes(int n_lags);
int kb2_self_moments(const double* dc, int n_atoms, int n_frames, int64_t pitch, const int32_t* lags, int n_lags,
                     int dim_mask, int variant, void* workspace, double* sums, void* stream);

/* The same sums through the arithmetic-run kernel (variant 2).  `lags_host` is the lag grid in HOST memory
 * (ascending): the call splits it into arithmetic progressions k0 + i*delta (kinisi's default grid,
 * np.linspace(min_dt, max_dt, n_steps, dtype=int) at kinisi/parser.py:150-168, is one up to integer
 * truncation, handled as a one-sample drift inside the run), runs the register-window kernel on every progression
 * (one partner sample per step serves up to 8 terms; lags that fit no progression become runs of one), and enqueues
 * one small plan copy on `stream`.  workspace: kb2_self_moments_runs_workspace_bytes(n_lags, n_frames) bytes. */
size_t kb2_self_moments_runs_workspace_bytes(int n_lags, int n_frames);
int kb2_self_moments_runs_blocks(int n_lags); /* launches (blocks of 192 lags) = 256-byte counter slots of the _planned form */
int kb2_self_moments_runs(const double* dc, int n_atoms, int n_frames, int64_t pitch, const int32_t* lags_host,
                          int n_lags, int dim_mask, void* workspace, double* sums, void* stream);
/* The same in two steps, for callers that keep the plan: it depends on (lags, n_frames) only.
 *   kb2_self_moments_runs_plan_pack   plans on the host into plan_host (kb2_self_moments_runs_workspace_bytes bytes of
 *                                     HOST memory, pinned if its upload is to be asynchronous); no device work.
 *   kb2_self_moments_runs_planned     launches the kernels of a packed plan: plan_host = the host copy (its headers are
 *                                     read by the call), plan_dev = the same bytes in device memory (uploaded by the caller
 *                                     earlier in stream order; never written), counters = one 256-byte device slot per
 *                                     block of 192 lags (zeroed by the call; not shared between concurrent calls).
 * The convenience form above re-plans and uploads from pageable memory on every call, which synchronises the stream. */
int kb2_self_moments_runs_plan_pack(const int32_t* lags_host, int n_lags, int n_frames, void* plan_host);
int kb2_self_moments_runs_planned(const double* dc, int n_atoms, int n_frames, int64_t pitch, int n_lags, int dim_mask,
                                  const void* plan_host, const void* plan_dev, void* counters, double* sums, void* stream);
/* The plan kb2_self_moments_runs makes for (the first 192 lags of -- one launch) a grid, for tests and diagnostics: returns the
 * number of runs of >= 3 lags, fills up to max_runs entries of run_k0 / run_delta / run_len / run_n_chg (host arrays,
 * may be NULL; run_n_chg = signed number of one-sample drift changes inside the run, lags k0 + i*delta + e_i) and
 * stores the number of lags that became runs of one in *n_rest. */
int kb2_self_moments_plan(const int32_t* lags_host, int n_lags, int n_frames, int32_t* run_k0, int32_t* run_delta,
                          int32_t* run_len, int32_t* run_n_chg, int max_runs, int* n_rest);

/* The same sums through the lag-window kernel (variant 3, the default of the Python host side).  The grid
 * (kinisi/parser.py:150-168) is cut into WINDOWS of up to 8 lags in arithmetic progression, kappa + w*D; a thread walks the
 * origins j = t + a*D of one residue class downwards with the 8 most recent partner samples and the 16 accumulators of
 * the window in registers: one origin and one partner sample per step serve 8 terms, the accumulators are folded across
 * the warp once per work item, no step is predicated (csrc/k2_moments_win.cu).  Lags that fit no progression of >= 3
 * become windows of one.  Same two-step protocol as the arithmetic-run kernel:
 *   kb2_self_moments_win_workspace_bytes  size of a packed plan (host and device copy) for n_lags lags
 *   kb2_self_moments_win_blocks           number of launches (blocks of 256 lags) = 256-byte counter slots needed
 *   kb2_self_moments_win_plan_pack        plans on the host into plan_host; no device work; depends on (lags, n_frames) only
 *   kb2_self_moments_win_planned          launches: plan_host = the host copy (headers are read by the call), plan_dev = the
 *                                         same bytes in device memory (uploaded earlier in stream order; never written),
 *                                         counters = kb2_self_moments_win_blocks(n_lags) * 256 bytes of device memory
 *                                         (zeroed by the call; not shared between concurrent calls); sums[n_lags][2] is
 *                                         zeroed by the call.
 *   kb2_self_moments_win_plan             the windows of (the first 256 lags of) a grid, for tests and diagnostics: fills up
 *                                         to max_wins entries of win_kappa / win_delta / win_len (host arrays, may be
 *                                         NULL), returns the number of windows. */
size_t kb2_self_moments_win_workspace_bytes(int n_lags);
int kb2_self_moments_win_blocks(int n_lags);
int kb2_self_moments_win_plan_pack(const int32_t* lags_host, int n_lags, int n_frames, void* plan_host);
int kb2_self_moments_win_planned(const double* dc, int n_atoms, int n_frames, int64_t pitch, int n_lags, int dim_mask,
                                 const void* plan_host, const void* plan_dev, void* counters, double* sums, void* stream);
int kb2_self_moments_win_plan(const int32_t* lags_host, int n_lags, int n_frames, int32_t* win_kappa, int32_t* win_delta,
                              int32_t* win_len, int max_wins);

/* The same sums through the tiled lag-window kernel (variant 4, the default of the Python host side; replaces the per-dt
 * loop of kinisi/diffusion.py:571-602 over the list of kinisi/parser.py:208-216).  One component of a trajectory is read as
 * a matrix with rows of D frames (D a multiple of the common difference of a progression of the grid,
 * kinisi/parser.py:150-168); a lag is k = I*D + r, a WINDOW up to 8 lags with consecutive I and one r.  The rows a group
 * of windows needs for a block of origin rows of a strip of 32 columns are copied into shared memory by the TMA engine
 * (cp.async.bulk + mbarrier, three buffers in flight, one producer warp); consumer warps walk the windows out of shared
 * memory with the partner window and the accumulators in registers: one origin and one partner sample per step serve 8
 * terms (csrc/k2_moments_tile.cu).  dc must be 16-byte aligned and pitch a multiple of 16.  Same two-step protocol:
 *   kb2_self_moments_tile_workspace_bytes  size of the packed plan (host and device copy); runs the planner (host only)
 *   kb2_self_moments_tile_blocks           number of launches (blocks of 256 lags) = 256-byte counter slots needed
 *   kb2_self_moments_tile_plan_pack        plans on the host into plan_host; depends on (lags, n_frames) only
 *   kb2_self_moments_tile_planned          launches (plan_host: headers are read by the call; plan_dev: the same bytes on the
 *                                          device, uploaded earlier in stream order; counters, sums: zeroed by the call)
 *   kb2_self_moments_tile_plan             the windows of (the first 256 lags of) a grid, for tests and diagnostics: first
 *                                          lag, lag spacing (= row length D) and number of lags per window; *n_tiles (may
 *                                          be NULL) receives the number of tile templates
 *   kb2_self_moments_tile_plan_cost        what the plan costs per atom (host only): out[0] = bytes copied into shared
 *                                          memory, out[1] = consumer issue cycles (steps x (16 m + 24) over all strips),
 *                                          out[2] = terms of the column walks.  The host side compares out[0] with the
 *                                          24 n_frames bytes of an atom to choose the kernel for a grid. */
size_t kb2_self_moments_tile_workspace_bytes(const int32_t* lags_host, int n_lags, int n_frames);
int kb2_self_moments_tile_blocks(int n_lags);
int kb2_self_moments_tile_plan_pack(const int32_t* lags_host, int n_lags, int n_frames, void* plan_host);
int kb2_self_moments_tile_planned(const double* dc, int n_atoms, int n_frames, int64_t pitch, int n_lags, int dim_mask,
                                  const void* plan_host, const void* plan_dev, void* counters, double* sums, void* stream);
int kb2_self_moments_tile_plan(const int32_t* lags_host, int n_lags, int n_frames, int32_t* win_first, int32_t* win_delta,
                               int32_t* win_len, int max_wins, int* n_tiles);
int kb2_self_moments_tile_plan_cost(const int32_t* lags_host, int n_lags, int n_frames, double* out);

/* Which of the kernels above for a lag grid (host only; what the Python host side does when no variant is forced).  The
 * reference has one code path for every grid (kinisi/diffusion.py:571-602); here the tiled kernel is taken while its
 * producers copy <= 2 bytes per term (kinisi's default 100-point grid on trajectories of >= 20 000 frames), the run kernel
 * for long runs of one spacing, the window kernel while the windows stay long, the direct kernel otherwise -- thresholds
 * from a measured survey of 68 grids (csrc/k2_moments_tile.cu, profiles/r02_k2_grid_survey.txt).  Returns the variant
 * number (0 direct, 2 runs, 3 windows, 4 tiles) or -1 on invalid input; features[4] (may be NULL) receives the bytes
 * per term of the tile plan, the run model's fraction, the share of terms in windows of <= 3 lags and the mean window
 * length. */
int kb2_self_moments_choose(const int32_t* lags_host, int n_lags, int n_frames, double* features);

/* The same sums from ONE materialised displacement array disp[n_atoms][n_obs][3] (float64, the
 * reference seam: an entry of disp_3d as passed to the *Bootstrap constructors,
 * diffusion.py:552-564).  out[0..1] = sum d^2, sum d^4 over atoms x observations (dim_mask);
 * out[2..3] = sum_j c_j, sum_j c_j^2 with c_j = |sum_a w_a disp[a][j]|^2 over coll_mask components
 * (w == NULL: unit weights; coll_mask == 0: skip).  out[4] is zeroed by the call. */
int kb2_disp_moments(const double* disp, int n_atoms, int n_obs, int dim_mask, const double* w, int coll_mask,
                     double* out, void* stream);

/* n, v, s, ngp from raw sums (diffusion.py:597-600, 683-686, 291-303):
 *   n   = m1/cnt,  var = (m2 - m1^2/cnt)/(cnt-1),  v = var/divisor,  s = sqrt(v)
 *   ngp = 3 (g2/gcnt) / (5 (g1/gcnt)^2) - 1
 * main_sums / ngp_sums are [n][2] (they may alias for MSD); cnt, divisor, gcnt are [n] float64. */
int kb2_moment_stats(const double* main_sums, const double* cnt, const double* divisor, const double* ngp_sums,
                     const double* gcnt, int n, double* out_n, double* out_v, double* out_s, double* out_ngp,
                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 3 -- covariance and Bayesian regression.
 * ------------------------------------------------------------------------------------------- */

/* cov[i][j] = cov[j][i] = (n_o[i] / n_o[j]) * v[i] for i <= j
 * (_populate_covariance_matrix, kinisi/diffusion.py:838-854). */
int kb2_cov_build(const double* v, const double* n_o, int n, double* cov, void* stream);

/* Turns the symmetric eigendecomposition (evals[n] ascending, evecs[n][n] row-major with
 * eigenvectors in columns) of the reconditioned covariance into the truncated whitening factor that
 * scipy.linalg.pinvh implies (diffusion.py:352: directions with |lambda| <= n*eps*max|lambda| are
 * dropped), applied to the straight-line model of diffusion.py:362-364:
 *   proj[0][r] = s_r V_r.x     proj[1][r] = s_r V_r.1     proj[2][r] = s_r V_r.y
 *   proj[3][r] = sign(lambda_r) or 0 if dropped,          s_r = |lambda_r|^-1/2
 * so that diff^T pinvh(C) diff = sum_r proj[3][r] (g proj[0][r] + c proj[1][r] - proj[2][r])^2.
 * gls[8] receives {g0, c0, AA, AB, BB, RA, RB, Q0}: the generalised-least-squares optimum (slope clamped
 * at >= 1e-20 like diffusion.py:370-371; the start point the reference finds with scipy.optimize.minimize,
 * :380-383) and the exact expansion of the quadratic form about it,
 *   Q(g0+dg, c0+dc) = Q0 + 2 (dg RA + dc RB) + dg^2 AA + 2 dg dc AB + dc^2 BB. */
int kb2_gls_prepare(const double* evals, const double* evecs, const double* x, const double* y, int n,
                    int fit_intercept, double* proj, double* gls, void* stream);

/* The spectral step of the regression in one call, for n <= kb2_spectral_prepare_max_n(): a cyclic Jacobi
 * eigen-iteration of the model covariance cov[n][n] in shared memory that yields, without forming eigenvectors,
 *   evals[n]       the raw spectrum (unsorted),
 *   proj[4][n]     the whitened model of kb2_gls_prepare for the spectrum clipped at lambda_max / cond_max
 *                  (minimum_eigenvalue_method, kinisi/diffusion.py:870-888) and cut at n*eps (pinvh, :352),
 *   gls[8]         as kb2_gls_prepare,
 *   logdet_term    sum_r log(clipped lambda_r) + n ln(2 pi)  (diffusion.py:350-351),
 *   info[0]        (optional) the number of Jacobi sweeps.
 * Equivalent, to rounding, to eigh + clip + rebuild + cov_nearest + slogdet + pinvh of the reference. */
int kb2_spectral_prepare_max_n(void);
int kb2_spectral_prepare(const double* cov, const double* x, const double* y, int n, double cond_max,
                         int fit_intercept, double* evals, double* proj, double* gls, double* logdet_term, int* info,
                         void* stream);

/* out[w] = log_likelihood(theta[w]) of diffusion.py:354-365 for n_walkers parameter vectors
 * theta[n_walkers][ndim] (ndim = 1 or 2); -inf where theta[w][0] < 0.
 * logdet_term is a device scalar holding slogdet(C) + n ln(2 pi) (diffusion.py:350-351). */
int kb2_mvn_loglike(const double* theta, int n_walkers, int ndim, const double* proj, int n,
                    const double* logdet_term, double* out, void* stream);

/* The whole ensemble MCMC of diffusion.py:384-390 in one launch: walkers start at
 * theta_ml + theta_ml * 1e-3 * N(0,1) (:384), then n_steps affine-invariant stretch moves (a = 2,
 * red/blue split with a random balanced partition per step) are run on the log-likelihood
 * -0.5 (logdet_term + Q(theta)), -inf for a negative slope, Q given by gls[8] from kb2_gls_prepare.
 * chain[n_steps][n_walkers][ndim], logp[n_steps][n_walkers].
 * workspace: kb2_sampler_workspace_bytes(n_steps, n_walkers) bytes. */
size_t kb2_sampler_workspace_bytes(int n_steps, int n_walkers);
int kb2_ensemble_sample(const double* gls, int ndim, int n_walkers, int n_steps, const double* logdet_term,
                        uint64_t seed, void* workspace, double* chain, double* logp, void* stream);

/* ---------------------------------------------------------------------------------------------
 * `bootstrap=True`: the resampled means of diffusion.py:258-289, 783-801.
 * ------------------------------------------------------------------------------------------- */
/* The population of one lag from the cumulative displacements (the array d_squared of diffusion.py:574, without the
 * list of parser.py:205-207): out[a * (n_frames - lag + 1) + m] = |dc[a][lag - 1 + m] - dc[a][m - 1]|^2 over dim_mask,
 * the m = 0 entry being the first-observation term |dc[a][lag - 1]|^2.  With n_atoms = 1 on the atom sum of
 * kb2_atom_sum this is coll_motion (diffusion.py:659, 751). */
int kb2_lag_squares(const double* dc, int n_atoms, int n_frames, int64_t pitch, int lag, int dim_mask, double* out,
                    void* stream);

/* The same populations from ONE materialised displacement array disp[n_atoms][n_obs][3] (an entry of disp_3d):
 * collective == 0: out[a * n_obs + j] = |disp[a][j]|^2 over dim_mask (diffusion.py:574);
 * collective != 0: out[j] = |sum_a w_a disp[a][j]|^2 over dim_mask (diffusion.py:659, 751; w == NULL: unit weights). */
int kb2_disp_squares(const double* disp, int n_atoms, int n_obs, int dim_mask, const double* w, int collective,
                     double* out, void* stream);

/* _bootstrap (diffusion.py:783-801): means[i] = mean of n_draws values drawn with replacement from
 * values[n_values], for the resamples first_resample + i, i < n_resamples.  Draw d of resample r takes the index
 * floor(u * n_values / 2^64) with u the 64-bit half (d & 1) of Philox4x32-10(counter {d >> 1 (64 bit), r, stream_id},
 * key seed): a pure function of its arguments, so a later call continues the sequence of an earlier one
 * (sample_until_normal, diffusion.py:258-289) and the same seed reproduces the same bits. */
int kb2_resample_means(const double* values, int64_t n_values, int64_t n_draws, int n_resamples, int first_resample,
                       uint64_t seed, int stream_id, double* means, void* stream);

/* `block=True` (diffusion.py:584-595, 670-681, 762-773): the statistics of pyblock.blocking.reblock (pyblock is not part
 * of the reference tree; Flyvbjerg & Petersen, J. Chem. Phys. 91, 461 (1989)) for one population.  Level 0 is the data;
 * level l + 1 holds the means of neighbouring pairs of level l, the last point of an odd level being dropped.  For
 * each of the kb2_reblock_levels(n_values) levels (those with at least two points) stats[2 l] = sum of the level and
 * stats[2 l + 1] = sum of squared deviations from its mean; the level has n_values >> l points.  stats holds
 * 2 * (levels + 1) doubles (device), workspace kb2_reblock_workspace_bytes(n_values) bytes (device, 16-byte aligned,
 * as `values`). */
int kb2_reblock_levels(int64_t n_values);
size_t kb2_reblock_workspace_bytes(int64_t n_values);
int kb2_reblock(const double* values, int64_t n_values, double* workspace, double* stats, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU joins through peer memory (csrc/k7_peer.cu).  With the atoms of a trajectory sharded over
 * ranks, the reference's sums over atoms -- np.mean(disp[:, framework], axis=1) in Parser.correct_drift
 * (kinisi/parser.py:143-145), np.sum(disp, axis=0) in MSTDBootstrap / MSCDBootstrap (kinisi/diffusion.py:659, 751) and
 * the per-lag sums behind np.mean / np.var in MSDBootstrap (:571-602) -- end with a sum over ranks of a [3][pitch] or
 * [K][2] float64 array.  One launch per rank does it over NVLink: every rank stores its array into a window of every
 * peer, raises a flag, and adds the W copies in rank order (bit-identical on all ranks).
 *
 *   kb2_peer_window_bytes    bytes of one rank's window for arrays of up to `capacity` doubles
 *   kb2_peer_window_create   allocates and zeroes this rank's window (cudaMalloc, synchronises) and returns the 64-byte
 *                            CUDA IPC handle its peers open; collective in spirit: every rank creates one, the caller
 *                            exchanges the handles (the reference-side binding uses torch.distributed / MPI for that)
 *   kb2_peer_window_open     maps a peer's window from its handle (enables peer access); _close unmaps it
 *   kb2_peer_window_destroy  frees this rank's own window (after every peer has closed it)
 *   kb2_peer_allreduce_f64   buf[0..n) <- sum over ranks, in place, enqueued on `stream`.  windows_host: HOST array of
 *                            `world` device pointers, entry `rank` being this rank's own window; `seq` is the number of
 *                            earlier calls on these windows -- every rank must issue the same calls (same n) in the same
 *                            order, on any streams; at most kb2_peer_slots() calls are buffered, later ones wait on the
 *                            device.  buf must be 16-byte aligned, n <= capacity.
 *   kb2_peer_status          0, or the code of the last call whose wait for a peer ran out (KB2_PEER_TIMEOUT_S seconds,
 *                            default 120): 1 a slot was never released, 2 a peer's copy never arrived.  Readable at any
 *                            time (mapped host memory), no synchronisation.
 * ------------------------------------------------------------------------------------------- */
int kb2_peer_slots(void);
size_t kb2_peer_window_bytes(int world, int64_t capacity);
int kb2_peer_window_create(int world, int64_t capacity, void** window, unsigned char* handle64);
int kb2_peer_window_open(const unsigned char* handle64, void** window);
int kb2_peer_window_close(void* window);
int kb2_peer_window_destroy(void* window);
int kb2_peer_allreduce_f64(double* buf, int64_t n, void* const* windows_host, int world, int rank, int64_t capacity,
                           uint64_t seq, void* stream);
/* The same join of a [rows][row_len] array cut by COLUMNS: slice c is columns [1024 c, 1024 (c + 1)) of every row.  This is
 * how the [3][pitch] frame sums are joined, because it is how kb2_frame_sums_joined produces them: a rank may take either
 * entry point for one and the same join (same windows, same seq).  kb2_peer_rows_slices(row_len): the number of slices, 0
 * if row_len is odd or needs more than 128. */
int kb2_peer_rows_slices(int64_t row_len);
int kb2_peer_allreduce_rows_f64(double* buf, int rows, int64_t row_len, void* const* windows_host, int world, int rank,
                                int64_t capacity, uint64_t seq, void* stream);
/* kb2_frame_sums followed by the sum over ranks, in ONE kernel (np.mean(disp[:, framework], axis=1), kinisi/parser.py:143-145,
 * with the framework atoms sharded over GPUs): the CTA that completes a tile of 1024 frames pushes it to the peers' windows,
 * waits for their copies and writes the all-rank sums, while the CTAs of other tiles still compute.  On return of the
 * kernel frame_sums holds the joined array.  tile_counters: 128 unsigned ints of device memory, zero on entry (the kernel
 * leaves them zero), not shared with another call in flight.  Needs ceil(pitch / 1024) == ceil(n_out_frames / 1024) and
 * kb2_peer_rows_slices(pitch) > 0; windows / world / rank / capacity / seq as for kb2_peer_allreduce_f64. */
int kb2_frame_sums_joined(const void* coords, int dtype, int mode, int n_atoms_total, int n_frames_in, const double* latt,
                          const double* latt_inv, int latt_stride, const int32_t* idx, const double* w, int n_idx,
                          double* frame_sums, int64_t pitch, void* const* windows_host, int world, int rank, int64_t capacity,
                          uint64_t seq, unsigned int* tile_counters, void* stream);
/* kb2_moment_stats with the sum over ranks of ngp_sums [n][2] (the per-lag sums of the atoms' own displacements, local to
 * this rank on entry: np.mean / np.var over atoms x origins, kinisi/diffusion.py:571-602) done first, in place, by the same
 * kernel: n <= 512.  main_sums may be ngp_sums itself (MSD) or an array that is already equal on all ranks (MSTD / MSCD).  The
 * join is the one kb2_peer_allreduce_f64(ngp_sums, 2 n, ...) would do (same windows, same seq). */
int kb2_moment_stats_joined(const double* main_sums, const double* cnt, const double* divisor, double* ngp_sums,
                            const double* gcnt, int n, double* out_n, double* out_v, double* out_s, double* out_ngp,
                            void* const* windows_host, int world, int rank, int64_t capacity, uint64_t seq, void* stream);
int kb2_peer_status(void);

/* ---------------------------------------------------------------------------------------------
 * Measurement helpers (bench.py): a dependent-free DFMA loop and a streaming copy, to measure the
 * FP64-pipe and HBM rooflines on the device the bench runs on.
 * ------------------------------------------------------------------------------------------- */
/* Launches grid x 256 threads, each doing iters * 16 DFMAs (counted: grid*256*iters*16). */
int kb2_fp64_peak(int grid, int iters, double* sink, void* stream);
int kb2_stream_copy(const double* src, double* dst, size_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KINISI_B200_H */
