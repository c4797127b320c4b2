/* mlb.h - C ABI of the B200-native constrained-HMC hot path (libmlb.so).
 *
 * Drop-in boundary for the hot path of thiery-lab/manifold_lifting (`mlift`): each entry
 * point replaces one of the reference's XLA:CPU-jitted callables (the jit table at
 * mlift/systems.py:328-346) or the Mici-driven loop around them, batched over
 * independent chains.  All file:line citations are relative to the reference tree.
 *
 * Conventions
 *   - All array arguments are DEVICE pointers unless the function name ends in `_host`.
 *   - Numeric type is fp64 (`double`) everywhere (systems.py:8); counters and status
 *     codes are `int32_t`.
 *   - Chain-state arrays are component-major ("SoA"): element (k, chain) of a
 *     [dim][n_chain] array lives at `ptr[k * ld + chain]`, ld >= n_chain.  A torch
 *     tensor of shape [n_chain, dim] with strides (1, ld) is a view of it.
 *   - The library never allocates or frees caller memory; outputs are written into
 *     caller-provided buffers.  No global mutable state; every call is re-entrant given
 *     distinct buffers; `stream` is a `cudaStream_t` passed as `void*` (NULL = default).
 *   - Return value: 0 on success, negative `MLB_ERR_*` for argument / CUDA errors.
 *     Per-chain numerical failures are NOT errors: they are reported in `status[chain]`
 *     (the reference raises `ConvergenceError` / Mici raises `NonReversibleStepError`,
 *     solvers.py:86-95; a batched kernel reports a mask instead).
 *   - `generate_y` is an arbitrary Python closure in the reference (systems.py:520-527);
 *     it cannot cross a C ABI, so a model is described by `mlb_model_desc`.
 */
#ifndef MLB_H_
#define MLB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLB_VERSION 100

/* ---- process-level error codes */
enum {
  MLB_OK = 0,
  MLB_ERR_ARG = -1,       /* bad argument (null pointer, unsupported dims, ...) */
  MLB_ERR_CUDA = -2,      /* CUDA runtime error; see mlb_last_error() */
  MLB_ERR_UNSUPPORTED = -3 /* model / shape combination not compiled in */
};

/* ---- per-chain status codes (written to `status[chain]`) */
enum {
  MLB_ST_OK = 0,
  MLB_ST_DIVERGED = 1,       /* |c| > divergence_tol or NaN  (solvers.py:86-90)  */
  MLB_ST_NOT_CONVERGED = 2,  /* max_iters reached            (solvers.py:91-95)  */
  MLB_ST_NON_REVERSIBLE = 3, /* Mici reversibility check failed                  */
  MLB_ST_H_DIVERGED = 4,     /* |h - h_init| > max_delta_h   (notebook cell 44)  */
  MLB_ST_LINALG = 5          /* singular / non-PD factorisation                  */
};

/* ---- models */
enum {
  MLB_MODEL_TOY = 0,  /* toy_2d_model.py:43-87 / notebook; additive noise, dense Gram */
  MLB_MODEL_HIER = 1, /* eight_schools.py:28-53; hierarchical, Woodbury Gram          */
  MLB_MODEL_FN = 2,   /* fitzhugh_nagumo.py:20-64; additive noise, RK4                */
  MLB_MODEL_HH = 3    /* hh_current_stimulus.py:19-203; additive noise, exp-Euler     */
};

enum { MLB_PARAM_UNBOUNDED = 0, MLB_PARAM_NORMAL = 1 }; /* utils.py:37-46 */

/* ---- projection solvers */
enum {
  MLB_SOLVER_QUASI_NEWTON = 0,     /* systems.py:190-228 / solvers.py:16-95        */
  MLB_SOLVER_NEWTON = 1,           /* systems.py:230-270 / solvers.py:98-169       */
  MLB_SOLVER_NEWTON_LS = 2,        /* systems.py:272-326 / solvers.py:172-256      */
  MLB_SOLVER_MICI_NEWTON = 3,      /* mici.solvers.solve_projection_onto_manifold_newton
                                      (toy configs, toy_2d_model.py:170-174)       */
  MLB_SOLVER_MICI_QUASI_NEWTON = 4 /* mici...quasi_newton (same call site)         */
};

enum { MLB_NORM_MAX = 0, MLB_NORM_EUCLIDEAN = 1 }; /* solvers.py:6-13 */

/* Model description.  `data` is a HOST array, copied at creation:
 *   TOY : {sigma, y, coeff}                        (n_data = 3)
 *   HIER: {y_obs[dim_y], sigma[dim_y]}             (n_data = 2*dim_y)
 *   FN  : {n_steps, y_obs[dim_y], dt_seq[n_steps], obs_step[dim_y]}: the static RK4
 *         schedule of mlift/ode.py:45-53 and, per observation, the index of the step
 *         that lands on it (n_data = 1 + 2 dim_y + n_steps)
 *   HH  : {n_steps, y_obs[dim_y], dt_seq[n_steps], obs_step[dim_y], stim[n_steps],
 *         consts[27]}: time grid and per-step stimulus current of
 *         hh_current_stimulus.py:117-122,175-178, then the constants of the reference's
 *         data file in the order k_alpha_n_1..3, k_beta_n_1..3, k_alpha_m_1..3,
 *         k_beta_m_1..3, k_alpha_h_1..3, k_beta_h_1..3, k_p_infinity_1..2, k_tau_p_2..4,
 *         E_Na, E_K, E_leak, c_m                                                     */
typedef struct mlb_model_desc {
  int32_t model;           /* MLB_MODEL_*                                   */
  int32_t dim_y;           /* number of observations                        */
  int32_t parametrization; /* MLB_PARAM_*                                   */
  int32_t n_data;
  const double* data;
} mlb_model_desc;

typedef struct mlb_model mlb_model; /* opaque */

/* Solver + integrator options; mirrors the keyword arguments of the reference solvers
 * (solvers.py:16-26, 172-183) and of Mici's ConstrainedLeapfrogIntegrator
 * (utils.py:322-327).  Passed by pointer per call, so adapters that mutate tolerances
 * between calls (adapters.py:45-47,56-58) keep working. */
typedef struct mlb_solver_opts {
  int32_t solver;                /* MLB_SOLVER_*                 */
  int32_t max_iters;             /* default 50                   */
  int32_t max_line_search_iters; /* default 10                   */
  int32_t norm;                  /* MLB_NORM_*                   */
  double constraint_tol;         /* default 1e-9                 */
  double position_tol;           /* default 1e-8                 */
  double divergence_tol;         /* default 1e10                 */
  double reverse_check_tol;      /* default 2 * position_tol     */
  int32_t reverse_check_norm;    /* MLB_NORM_*                   */
  int32_t flags;                 /* MLB_FLAG_* (0 = defaults)    */
} mlb_solver_opts;

/* mlb_solver_opts.flags.  The streaming (ODE) family runs mlb_leapfrog_steps as a
 * host-driven sequence of sub-phase kernels whose sweeps are split over sensitivity
 * directions / pair groups (several warps per batch of 32 chains); the call then blocks
 * the calling thread while it polls per-iteration convergence counts.  This flag selects
 * the fully asynchronous single-kernel path instead (one thread per chain).  The same holds
 * for mlb_hmc_sample on those models; mlb_nuts_sample on them is always host-driven. */
#define MLB_FLAG_SINGLE_KERNEL 1
/* mlb_nuts_sample follows Mici 0.2.1's DynamicMultinomialHMC defaults (the sampler built
 * at example_models/utils.py:286-288): the no-U-turn criterion is also tested on the two
 * overlapping sub-trajectories of every merge (`do_extra_subtree_checks=True`) and a
 * trajectory is divergent when h - h_init > max_delta_h (one-sided).  These flags switch
 * either off / to the two-sided |h - h_init| test. */
#define MLB_FLAG_NUTS_NO_EXTRA_SUBTREE_CHECKS 2
#define MLB_FLAG_NUTS_TWO_SIDED_DELTA_H 4

/* ---- lifecycle */
int mlb_version(void);
const char* mlb_last_error(void);
int mlb_model_create(const mlb_model_desc* desc, mlb_model** out);
void mlb_model_destroy(mlb_model* m);
/* Scratch of the drivers (tree storage of mlb_nuts_sample, sweep scratch of the ODE and
 * state-space kernels) is stream-ordered memory from a pool owned by the library (one per
 * device; the host application's default pool is not touched) that keeps freed blocks
 * cached between calls.  Hands everything above `keep_bytes` back to the driver. */
int mlb_scratch_trim(size_t keep_bytes);
int mlb_model_dim_u(const mlb_model* m);
int mlb_model_dim_y(const mlb_model* m);
int mlb_model_dim_q(const mlb_model* m); /* dim_u + dim_y (additive) / dim_u + 2 dim_y (hier) */
int mlb_model_jac_rows(const mlb_model* m);  /* rows of the packed Jacobian-block array  */
int mlb_model_gram_rows(const mlb_model* m); /* rows of the packed Gram-component array  */
int mlb_model_gram_dim(const mlb_model* m);  /* K: dim_y (dense) or dim_u (Woodbury)     */

/* Packed block layouts (rows of a [rows][n_chain] SoA array):
 *   jac : dy_du[i][a] at row i*dim_u + a (i < dim_y), then dy_dv[i] (hier only), then
 *         dy_dn[i]                               (blocks of systems.py:568-580, 698-710)
 *   gram: chol[r][c] at row r*K + c (lower triangular, K as above), then d[i] =
 *         dy_dv^2 + dy_dn^2                      (systems.py:591-594, 613-617, 751-757) */

/* ---- the 13 array-level callables of systems.py:328-346 (one entry per row) */

/* _constr: c(q) = generate_y(...) - y_obs.  q [Q][n] -> c [Y][n].  systems.py:564-566, 694-696 */
int mlb_constr(const mlb_model* m, int64_t n, int64_t ld, const double* q, double* c,
               void* stream);
/* _jacob_constr_blocks: -> packed jac, c.  systems.py:568-580, 698-710 */
int mlb_jacob_constr_blocks(const mlb_model* m, int64_t n, int64_t ld, const double* q,
                            double* jac, double* c, void* stream);
/* _decompose_gram: jac -> packed gram.  systems.py:591-594 / 613-617 / 751-757 */
int mlb_decompose_gram(const mlb_model* m, int64_t n, int64_t ld, const double* jac,
                       double* gram, void* stream);
/* _log_det_sqrt_gram: q -> val [n] (+ aux c, jac, gram if non-NULL).  systems.py:603-609/635-641/785-794 */
int mlb_log_det_sqrt_gram(const mlb_model* m, int64_t n, int64_t ld, const double* q,
                          double* val, double* c, double* jac, double* gram, void* stream);
/* _val_and_grad_log_det_sqrt_gram: q -> val [n], grad [Q][n] (+ aux).  systems.py:332-334 */
int mlb_val_and_grad_log_det_sqrt_gram(const mlb_model* m, int64_t n, int64_t ld,
                                       const double* q, double* val, double* grad,
                                       double* c, double* jac, double* gram, void* stream);
/* _lmult_by_jacob_constr: J v.  systems.py:582-584, 712-718 */
int mlb_lmult_by_jacob_constr(const mlb_model* m, int64_t n, int64_t ld, const double* jac,
                              const double* v, double* out, void* stream);
/* _rmult_by_jacob_constr: J^T w.  systems.py:586-587, 720-721 */
int mlb_rmult_by_jacob_constr(const mlb_model* m, int64_t n, int64_t ld, const double* jac,
                              const double* w, double* out, void* stream);
/* _lmult_by_pinv_jacob_constr: J^T (J J^T)^-1 w.  systems.py:173-179 */
int mlb_lmult_by_pinv_jacob_constr(const mlb_model* m, int64_t n, int64_t ld,
                                   const double* jac, const double* gram, const double* w,
                                   double* out, void* stream);
/* _lmult_by_inv_jacob_product: (J_l J_r^T)^-1 w (non-symmetric, LU).  systems.py:599-601/626-633/770-783 */
int mlb_lmult_by_inv_jacob_product(const mlb_model* m, int64_t n, int64_t ld,
                                   const double* jac_l, const double* jac_r,
                                   const double* w, double* out, void* stream);
/* _normal_space_component: J^T (J J^T)^-1 J v.  systems.py:181-188 */
int mlb_normal_space_component(const mlb_model* m, int64_t n, int64_t ld, const double* jac,
                               const double* gram, const double* v, double* out,
                               void* stream);
/* _quasi_newton_projection / _newton_projection / _newton_projection_with_line_search
 * (systems.py:190-326), selected by opts->solver; the two Mici solvers likewise.  The
 * Jacobian blocks (and Gram components for quasi-Newton) of the previous point are
 * passed packed, as the reference solvers fetch them from the cached previous state
 * (solvers.py:60-61,131,209).  Outputs: q_out [Q][n] (may alias q), mu_over_dt [Q][n]
 * (= what the reference subtracts from `state.mom`, solvers.py:84), iters, norm_dq, err,
 * n_constr_calls (line search; may be NULL), status. */
int mlb_solve_projection(const mlb_model* m, int64_t n, int64_t ld, const double* q,
                         const double* jac_prev, const double* gram_prev, double dt,
                         const mlb_solver_opts* opts, double* q_out, double* mu_over_dt,
                         int32_t* iters, double* norm_dq, double* err,
                         int32_t* n_constr_calls, int32_t* status, void* stream);

/* neg_log_dens / grad_neg_log_dens of the extended prior (systems.py:52-87 wrappers;
 * eight_schools.py:49-53, toy_2d_model.py:72-74).  grad may be NULL. */
int mlb_neg_log_dens(const mlb_model* m, int64_t n, int64_t ld, const double* q, double* val,
                     double* grad, void* stream);
/* h1 = nld + log_det_sqrt_gram, dh1_dpos (systems.py:407-411); h1/dh1 may be NULL. */
int mlb_h1(const mlb_model* m, int64_t n, int64_t ld, const double* q, double* h1,
           double* dh1_dpos, void* stream);
/* project_onto_cotangent_space: p -= J^T (J J^T)^-1 J p at position q (systems.py:464-466);
 * in place on p. */
int mlb_project_onto_cotangent_space(const mlb_model* m, int64_t n, int64_t ld,
                                     const double* q, double* p, void* stream);

/* ---- fused drivers (replace the Mici loop around the callables) */

/* n_step constrained leapfrog steps A(dt/2) B(dt) A(dt/2) with reversibility check per
 * chain, state kept on chip between steps.  q, p updated in place to the last
 * successfully completed step.  dt_per_chain may be NULL (use dt).  Optional outputs
 * (NULL to skip): iters_fwd / iters_rev (summed projection iterations), h_init, h_final. */
int mlb_leapfrog_steps(const mlb_model* m, int64_t n, int64_t ld, double* q, double* p,
                       double dt, const double* dt_per_chain, int32_t n_step,
                       const mlb_solver_opts* opts, int32_t* status, int32_t* n_done,
                       int32_t* iters_fwd, int32_t* iters_rev, double* h_init,
                       double* h_final, void* stream);

/* Dual-averaging step-size adaptation state, one per chain, SoA rows:
 *   0 iter, 1 smoothed_log_step_size, 2 adapt_stat_error, 3 log_step_size_reg_target,
 *   4 current step size (in/out).  (Mici DualAveragingStepSizeAdapter; utils.py:281-285) */
#define MLB_ADAPT_ROWS 5
typedef struct mlb_adapt_opts {
  double adapt_stat_target;             /* 0.8  (utils.py:82-87)  */
  double log_step_size_reg_coefficient; /* 0.05 (utils.py:88-93)  */
  double iter_decay_coeff;              /* 0.75 */
  double iter_offset;                   /* 10   */
} mlb_adapt_opts;

/* Per-chain sampling statistics accumulated by mlb_hmc_sample, SoA rows:
 *   0 sum accept_stat, 1 n_accepted, 2 n_convergence_error, 3 n_non_reversible_step,
 *   4 n_hamiltonian_diverged, 5 leapfrog steps completed, 6 projection iterations,
 *   7 sum of tree depths (mlb_nuts_sample only; row 1 there counts transitions whose
 *   state changed). */
#define MLB_STAT_ROWS 8

/* n_iter static-HMC transitions per chain in ONE launch (momentum refresh, n_step
 * constrained leapfrog steps with |dh| > max_delta_h test, Metropolis accept; notebook
 * cell 44 `simulate_proposal` / Mici StaticMetropolisHMC).  Random numbers: Philox4x32-10
 * keyed by (seed, chain0 + chain, iter0 + iteration), so results do not depend on how
 * chains are sharded over GPUs; or supplied (`mom_noise` [n_iter][Q][n], `unif`
 * [n_iter][n]) for decision-by-decision parity tests.  `adapt` ([MLB_ADAPT_ROWS][n],
 * NULL = no adaptation) is updated per transition and its row 4 used as the chain's step
 * size.  `trace_q` ([n_iter / trace_every][trace_dim][n], NULL to skip) receives the first
 * `trace_dim` position components (0 = all Q; the latent parameters u come first),
 * `trace_accept` ([n_iter][n], NULL to skip) per-transition accept_stat,
 * `stats` [MLB_STAT_ROWS][n] is accumulated into (not zeroed).  For the ODE models the
 * transitions are driven from the host over the sub-phase kernels (the call blocks the
 * calling thread) unless opts->flags has MLB_FLAG_SINGLE_KERNEL. */
int mlb_hmc_sample(const mlb_model* m, int64_t n, int64_t ld, double* q, double dt,
                   int32_t n_step, int32_t n_iter, const mlb_solver_opts* opts,
                   double max_delta_h, uint64_t seed, uint64_t chain0, uint32_t iter0,
                   const double* mom_noise, const double* unif, double* adapt,
                   const mlb_adapt_opts* adapt_opts, double* trace_q, int32_t trace_every,
                   int32_t trace_dim, double* trace_accept, double* stats,
                   int32_t* last_status, void* stream);

/* Philox streams used above, exposed for tests (host function, host output). */
void mlb_philox_normals_host(uint64_t seed, uint64_t chain, uint32_t iter, int32_t n,
                             double* out);
double mlb_philox_uniform_host(uint64_t seed, uint64_t chain, uint32_t iter);

/* n_iter dynamic multinomial HMC (NUTS) transitions per chain in ONE launch: Mici's
 * `DynamicMultinomialHMC`, the sampler the reference builds at example_models/utils.py:
 * 286-288 (max_tree_depth :70-73): momentum refresh, binary tree doubling in a uniformly
 * drawn direction with constrained leapfrog steps, multinomial proposal with weights
 * exp(h_init - h), no-U-turn termination per subtree and for the tree, termination on an
 * integrator error or |h - h_init| > max_delta_h.  accept_stat (the adaptation statistic)
 * is the mean of min(1, exp(h_init - h)) over the tree's steps.  Random numbers: momentum
 * normals as in mlb_hmc_sample; the k-th uniform of a transition is Philox counter
 * (k, iter, chain | 1 << 63), k = 1, 2, ... in order of use.  Arguments as mlb_hmc_sample;
 * trace_n_step / trace_depth ([n_iter][n], NULL to skip) receive the tree size / depth
 * of every transition.  Toy / hierarchical models: every chain builds its tree inside one
 * kernel (asynchronous).  ODE models: the trees of all chains grow in lock-step rounds
 * driven from the host over the sub-phase kernels (quasi-Newton / Newton solvers; the
 * call blocks the calling thread). */
int mlb_nuts_sample(const mlb_model* m, int64_t n, int64_t ld, double* q, double dt,
                    int32_t max_tree_depth, int32_t n_iter, const mlb_solver_opts* opts,
                    double max_delta_h, uint64_t seed, uint64_t chain0, uint32_t iter0,
                    const double* mom_noise, double* adapt, const mlb_adapt_opts* adapt_opts,
                    double* trace_q, int32_t trace_every, int32_t trace_dim,
                    double* trace_accept, int32_t* trace_n_step, int32_t* trace_depth,
                    double* stats, int32_t* last_status, void* stream);

/* ---- mlift.linalg tridiagonal primitives (reference mlift/linalg.py:46-114; used by the
 * state-space systems' "tridiagonal + low rank" Gram algebra, systems.py:1189-1238),
 * batched over chains.  All arrays component-major with leading dimension ld >= n:
 * a, c [dim-1][ld] (lower / upper diagonal), b [dim][ld] (main diagonal),
 * d, x [n_rhs][dim][ld] (the reference vmaps the solve over right-hand sides,
 * systems.py:1193-1195, 1221-1223), out [n]. */
int mlb_tridiagonal_solve(int64_t n, int32_t dim, int32_t n_rhs, int64_t ld, const double* a,
                          const double* b, const double* c, const double* d, double* x,
                          void* stream);
int mlb_tridiagonal_pos_def_log_det(int64_t n, int32_t dim, int64_t ld, const double* a,
                                    const double* b, double* out, void* stream);

/* ---- dense helpers of mlift.linalg (reference mlift/linalg.py:6-43; used by the
 * Gaussian-process systems, systems.py:871-899), batched over chains.  Arrays are the
 * reference's with a leading chain axis, chain-major and row-major: l, u, chol_mtx
 * [n][dim][dim] (l / chol_mtx lower-, u upper-triangular; the other triangle is not read),
 * b, x [n][dim][n_rhs], vct [n][dim], tangents [n][n_dir][dim][dim] (symmetric; the
 * reference vmaps over directions, systems.py:868-873), out [n][n_dir][dim]. */
#define MLB_DENSE_MAX_DIM 2048
/* solve l u x = b (linalg.py:6-9) */
int mlb_lu_triangular_solve(int64_t n, int32_t dim, int32_t n_rhs, const double* l,
                            const double* u, const double* b, double* x, void* stream);
/* Jacobian-vector product of mtx -> cholesky(mtx) vct in the direction `tangents`
 * (linalg.py:12-43): chol Phi(chol^-1 t chol^-T) vct, Phi = lower triangle, diagonal halved */
int mlb_jvp_cholesky_mtx_mult_by_vct(int64_t n, int32_t n_dir, int32_t dim,
                                     const double* tangents, const double* chol_mtx,
                                     const double* vct, double* out, void* stream);

/* ---- block algebra of GaussianProcessModelSystem (reference mlift/systems.py:859-918:
 * y = cholesky(covar(u)) n), batched over chains, given the Jacobian blocks dy_du
 * [n][dim_y][dim_u] and dy_dn = cholesky(covar) [n][dim_y][dim_y] (lower; row-major, as the
 * reference's arrays with a leading chain axis).  Gram = dy_du dy_du^T + dy_dn dy_dn^T =
 * "Cholesky-factored covariance + low rank": solves go through the dim_u x dim_u
 * capacitance matrix (Cholesky factor chol_cap_mtx [n][dim_u][dim_u] for the symmetric
 * Gram, pivoted LU for the non-symmetric product of the Newton iteration). */
#define MLB_GP_MAX_DIM_U 16
/* cholesky(covar) per chain (systems.py:869); a matrix that is not positive definite gives
 * a NaN factor for its chain */
int mlb_dense_cholesky(int64_t n, int32_t dim, const double* a, double* l, void* stream);
/* systems.py:885-888 */
int mlb_gp_decompose_gram(int64_t n, int32_t dim_y, int32_t dim_u, const double* dy_du,
                          const double* dy_dn, double* chol_cap_mtx, void* stream);
/* systems.py:890-897: Gram^-1 vct, vct / out [n][dim_y] */
int mlb_gp_lmult_by_inv_gram(int64_t n, int32_t dim_y, int32_t dim_u, const double* dy_du,
                             const double* dy_dn, const double* chol_cap_mtx, const double* vct,
                             double* out, void* stream);
/* systems.py:899-911: (J_l J_r^T)^-1 vct */
int mlb_gp_lmult_by_inv_jacob_product(int64_t n, int32_t dim_y, int32_t dim_u,
                                      const double* dy_du_l, const double* dy_dn_l,
                                      const double* dy_du_r, const double* dy_dn_r,
                                      const double* vct, double* out, void* stream);
/* systems.py:878-883: transpose = 0: J vct (vct [n][dim_u + dim_y] -> out [n][dim_y]);
 * transpose = 1: J^T vct (vct [n][dim_y] -> out [n][dim_u + dim_y]) */
int mlb_gp_jacob_constr_mult(int64_t n, int32_t dim_y, int32_t dim_u, int32_t transpose,
                             const double* dy_du, const double* dy_dn, const double* vct,
                             double* out, void* stream);
/* systems.py:913-918 from the factors: sum log diag chol_cap_mtx + sum log diag dy_dn */
int mlb_gp_log_det_sqrt_gram(int64_t n, int32_t dim_y, int32_t dim_u, const double* dy_dn,
                             const double* chol_cap_mtx, double* out, void* stream);

/* ---- "tridiagonal + low rank" Gram algebra of the state-space systems (reference
 * mlift/systems.py:1161-1243, the closures of PartiallyInvertibleStateSpaceModelSystem),
 * batched over chains, on top of the Thomas recurrence above.  The Jacobian blocks are the
 * reference's tuple (dc_du, dc_dv, (dc_dn[0], dc_dn[1]), dx_dy) (systems.py:1136-1150), one
 * component-major device array each with the same leading dimension ld >= n:
 * dc_du [dim_y * dim_u][ld] (row k * dim_u + a), dc_dv, dc_dn0, dx_dy [dim_y][ld],
 * dc_dn1 [dim_y - 1][ld].  dx_dy is read by mlb_ssm_log_det_sqrt_gram only (may be NULL
 * elsewhere).  The decomposition is the reference's (a [dim_y - 1][ld], b [dim_y][ld],
 * chol_cap_mtx [dim_u * dim_u][ld], row r * dim_u + c, lower triangular, NaN where the
 * capacitance matrix is not positive definite).  1 <= dim_u <= MLB_SSM_MAX_DIM_U.  */
#define MLB_SSM_MAX_DIM_U 8
typedef struct mlb_ssm_blocks {
  const double* dc_du;
  const double* dc_dv;
  const double* dc_dn0;
  const double* dc_dn1;
  const double* dx_dy;
} mlb_ssm_blocks;

/* systems.py:1189-1197 */
int mlb_ssm_decompose_gram(int64_t n, int32_t dim_y, int32_t dim_u, int64_t ld,
                           const mlb_ssm_blocks* J, double* a, double* b, double* chol_cap,
                           void* stream);
/* systems.py:1199-1210: out [dim_y][ld] = (J J^T)^{-1} vct, vct [dim_y][ld], out != vct */
int mlb_ssm_lmult_by_inv_gram(int64_t n, int32_t dim_y, int32_t dim_u, int64_t ld,
                              const mlb_ssm_blocks* J, const double* a, const double* b,
                              const double* chol_cap, const double* vct, double* out,
                              void* stream);
/* systems.py:1212-1232: out = (J_l J_r^T)^{-1} vct (non-symmetric tridiagonal part, pivoted
 * LU on the capacitance matrix as np.linalg.solve) */
int mlb_ssm_lmult_by_inv_jacob_product(int64_t n, int32_t dim_y, int32_t dim_u, int64_t ld,
                                       const mlb_ssm_blocks* J_l, const mlb_ssm_blocks* J_r,
                                       const double* vct, double* out, void* stream);
/* systems.py:1234-1243, given the blocks and their decomposition: out [n] =
 * sum log diag(chol_cap) + tridiagonal_pos_def_log_det(a, b) / 2 - sum log|dx_dy| */
int mlb_ssm_log_det_sqrt_gram(int64_t n, int32_t dim_y, int32_t dim_u, int64_t ld,
                              const mlb_ssm_blocks* J, const double* a, const double* b,
                              const double* chol_cap, double* out, void* stream);
/* systems.py:1161-1175: out [dim_y][ld] = J vct, vct [dim_u + 2 dim_y][ld] = (u, v, n) parts */
int mlb_ssm_lmult_by_jacob_constr(int64_t n, int32_t dim_y, int32_t dim_u, int64_t ld,
                                  const mlb_ssm_blocks* J, const double* vct, double* out,
                                  void* stream);
/* systems.py:1177-1187: out [dim_u + 2 dim_y][ld] = J^T vct, vct [dim_y][ld] */
int mlb_ssm_rmult_by_jacob_constr(int64_t n, int32_t dim_y, int32_t dim_u, int64_t ld,
                                  const mlb_ssm_blocks* J, const double* vct, double* out,
                                  void* stream);

/* systems.py:181-188 and :464-466: mom [dim_u + 2 dim_y][ld] -= J^T (J J^T)^{-1} J mom, in
 * place (normal-space component removed: the momentum re-projection of the leapfrog step) */
int mlb_ssm_project_onto_cotangent_space(int64_t n, int32_t dim_y, int32_t dim_u, int64_t ld,
                                         const mlb_ssm_blocks* J, const double* a,
                                         const double* b, const double* chol_cap, double* mom,
                                         void* stream);
/* Stochastic-volatility model (reference mlift/example_models/stochastic_volatility.py:71-141:
 * constr_split and jacob_constr_split_blocks), `unbounded` prior parametrisation: q
 * [3 + 2 dim_y][ld] = (u = (mu, log sigma, logit((phi + 1) / 2)), v, n), y_obs [dim_y] device
 * (shared by all chains) -> c [dim_y][ld] and, if J_out is not NULL, the five block arrays it
 * points to (WRITTEN; same shapes as above). */
int mlb_ssm_sv_constr_blocks(int64_t n, int32_t dim_y, int64_t ld, const double* q,
                             const double* y_obs, double* c, const mlb_ssm_blocks* J_out,
                             void* stream);
/* What the gradient of log_det_sqrt_gram needs from the Gram matrix G = J J^T (the reference
 * takes that gradient by reverse-mode autodiff of systems.py:1233-1243, jit table
 * systems.py:332-334; in closed form d(1/2 log det G) = sum over the non-zero pattern of J of
 * (G^{-1} J)_{tj} dJ_{tj}): inv_gram_dc_du [dim_y * dim_u][ld] = G^{-1} dc_du (row t * dim_u +
 * u), inv_gram_diag [dim_y][ld] = diag(G^{-1}), inv_gram_offdiag [dim_y - 1][ld] =
 * (G^{-1})_{t,t+1}; a, b, chol_cap from mlb_ssm_decompose_gram. */
int mlb_ssm_band_inv_gram(int64_t n, int32_t dim_y, int32_t dim_u, int64_t ld,
                          const mlb_ssm_blocks* J, const double* a, const double* b,
                          const double* chol_cap, double* inv_gram_dc_du, double* inv_gram_diag,
                          double* inv_gram_offdiag, void* stream);
/* Stochastic-volatility model: grad [3 + 2 dim_y][ld] = d log_det_sqrt_gram / dq, the
 * contraction of the weights above with the derivatives of the entries of
 * jacob_constr_split_blocks (stochastic_volatility.py:92-141); J = the blocks at q. */
int mlb_ssm_sv_grad_log_det_sqrt_gram(int64_t n, int32_t dim_y, int64_t ld, const double* q,
                                      const double* y_obs, const mlb_ssm_blocks* J,
                                      const double* inv_gram_dc_du, const double* inv_gram_diag,
                                      const double* inv_gram_offdiag, double* grad,
                                      void* stream);
/* Bookkeeping of one iteration of the projection loops (systems.py:203-227, 243-269) for the
 * chains with active[i] != 0: error = |c|_max, mu += delta_q, q -= delta_q, n_iters += 1,
 * norm_delta_q = |delta_q|_max, then active[i] = !(n_iters >= max_iters || error >
 * divergence_tol || isnan(error) || (error < constraint_tol && norm_delta_q < position_tol));
 * *n_active (device) is incremented once per chain that stays active.  c [dim_y][ld],
 * delta_q, q, mu [dim_q][ld]; per-chain arrays [n]. */
int mlb_ssm_projection_update(int64_t n, int32_t dim_q, int32_t dim_y, int64_t ld,
                              const double* c, const double* delta_q, double* q, double* mu,
                              int32_t* n_iters, double* norm_delta_q, double* error,
                              int32_t* active, double constraint_tol, double position_tol,
                              double divergence_tol, int32_t max_iters, int32_t* n_active,
                              void* stream);

/* ---- host-buffer entry point (what a reference-side binding calls; see
 * INTEGRATION.md).  q, p are HOST arrays in the reference's natural layout
 * [n_chain][Q] row-major (one NumPy `state.pos` per row); the call stages them through
 * pinned memory, transposes on device, runs mlb_leapfrog_steps and copies results back.
 * Output arrays are host arrays of length n (may be NULL except status). */
int mlb_leapfrog_steps_host(const mlb_model* m, int64_t n, double* q, double* p, double dt,
                            int32_t n_step, const mlb_solver_opts* opts, int32_t* status,
                            int32_t* n_done, int32_t* iters_fwd, int32_t* iters_rev,
                            double* h_init, double* h_final);

/* Transposes between the reference's [n][dim] row-major layout and the SoA layout. */
int mlb_aos_to_soa(int64_t n, int32_t dim, const double* aos, double* soa, int64_t ld,
                   void* stream);
int mlb_soa_to_aos(int64_t n, int32_t dim, const double* soa, int64_t ld, double* aos,
                   void* stream);

/* Launch counter: number of kernels this library has launched in this process (used by
 * bench.py's `gpu_launches`). */
int64_t mlb_launch_count(void);

/* FP64 FMA peak microbenchmark (roofline denominator; MEASURED_PEAKS.json has no fp64
 * entry).  Returns achieved TFLOP/s of a dependent-chain DFMA loop over the full chip. */
double mlb_measure_fp64_tflops(int32_t repeats);

/* Self-test of the lean fp64 primitives used by the kernels (Newton-refined MUFU
 * reciprocal / reciprocal square root, csrc/mlb_common.cuh): x [n] device in; rcp, sqrt,
 * rsqrt [n] device out.  Tests compare them with the IEEE operations. */
int mlb_selftest_fast_math(int64_t n, const double* x, double* rcp, double* sqrt_out,
                           double* rsqrt_out, void* stream);

/* Same for the branch-free exp / expm1 of the Hodgkin-Huxley recurrence: x [n] device in;
 * exp_out, expm1_out [n] device out (arguments are clamped to [-700, 700]). */
int mlb_selftest_exp(int64_t n, const double* x, double* exp_out, double* expm1_out,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MLB_H_ */
