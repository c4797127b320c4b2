// C ABI of libmlb.so (see include/mlb.h): argument checking, model registry, dispatch to
// the per-model launcher tables compiled in mlb_inst_*.cu, layout helpers and the
// host-buffer entry point.
#include "mlb_stream_launch.cuh"

namespace mlb {

std::string& last_error() {
  static thread_local std::string e;
  return e;
}
std::atomic<int64_t>& launch_count() {
  static std::atomic<int64_t> c{0};
  return c;
}

extern const ModelOps ops_toy, ops_hier4u, ops_hier4n, ops_hier8u, ops_hier8n, ops_hier16u,
    ops_hier16n, ops_fn_u, ops_fn_n, ops_hh_u, ops_hh_n;

static const ModelOps* find_ops(int model, int dim_y, int param) {
  const ModelOps* all[] = {&ops_toy,    &ops_hier4u,  &ops_hier4n, &ops_hier8u, &ops_hier8n,
                           &ops_hier16u, &ops_hier16n, &ops_fn_u,   &ops_fn_n,
                           &ops_hh_u,   &ops_hh_n};
  for (const ModelOps* o : all)
    if (o->model == model && (o->dim_y == dim_y || o->dim_y == 0) &&
        (model == MLB_MODEL_TOY || o->param == param))
      return o;
  return nullptr;
}

// Device copy of an ODE model's arrays: {y_obs[Y], dt_seq[n_steps], aux[n_steps]?} as
// doubles followed by obs_step[Y] as int32, in one allocation made on first use on the
// calling thread's current device (one process drives one GPU).  Host data layout
// (mlb_model_create): {n_steps, y_obs[Y], dt_seq[n_steps], obs_step[Y] [, aux[n_steps],
// consts[27]]}.
int ode_data(const mlb_model* m, OdeData* out) {
  std::lock_guard<std::mutex> lock(m->dev_mutex);
  int dev = 0;
  MLB_CUDA_OK(cudaGetDevice(&dev));
  const int Y = m->Y, ns = m->n_steps;
  const bool has_aux = m->model == MLB_MODEL_HH;
  if (m->dev_blob && m->dev_id != dev) {
    g_err = "an ODE model handle is bound to the device of its first use";
    return MLB_ERR_ARG;
  }
  const double* h = m->data.data() + 1;  // y_obs, dt_seq, obs_step, [aux, consts]
  const size_t nd = (size_t)Y + ns + (has_aux ? ns : 0);
  if (!m->dev_blob) {
    std::vector<char> host(nd * sizeof(double) + (size_t)Y * sizeof(int));
    double* hd = reinterpret_cast<double*>(host.data());
    std::memcpy(hd, h, ((size_t)Y + ns) * sizeof(double));
    if (has_aux) std::memcpy(hd + Y + ns, h + 2 * (size_t)Y + ns, (size_t)ns * sizeof(double));
    int* steps = reinterpret_cast<int*>(host.data() + nd * sizeof(double));
    for (int i = 0; i < Y; ++i) steps[i] = (int)h[(size_t)Y + ns + i];
    void* p = nullptr;
    MLB_CUDA_OK(cudaMalloc(&p, host.size()));
    cudaError_t e = cudaMemcpy(p, host.data(), host.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      cudaFree(p);
      g_err = std::string("model constants upload: ") + cudaGetErrorString(e);
      return MLB_ERR_CUDA;
    }
    m->dev_blob = p, m->dev_id = dev;
  }
  const double* d = static_cast<const double*>(m->dev_blob);
  out->y_obs = d, out->dt_seq = d + Y;
  out->aux = has_aux ? d + Y + ns : nullptr;
  out->obs_step = reinterpret_cast<const int*>(d + nd);
  out->Y = Y, out->n_steps = ns;
  for (int k = 0; k < kOdeConsts; ++k) out->consts[k] = 0.0;
  if (m->model == MLB_MODEL_HH) {
    const double* K = h + 2 * (size_t)Y + 2 * (size_t)ns;
    for (int k = 0; k < 27; ++k) out->consts[k] = K[k];
    const int recips[9][2] = {{HH_r_alpha_n_3, HH_k_alpha_n_3}, {HH_r_beta_n_3, HH_k_beta_n_3},
                              {HH_r_alpha_m_3, HH_k_alpha_m_3}, {HH_r_beta_m_3, HH_k_beta_m_3},
                              {HH_r_alpha_h_3, HH_k_alpha_h_3}, {HH_r_beta_h_3, HH_k_beta_h_3},
                              {HH_r_p_infinity_2, HH_k_p_infinity_2},
                              {HH_r_tau_p_4, HH_k_tau_p_4}, {HH_r_c_m, HH_c_m}};
    for (auto& r : recips) out->consts[r[0]] = 1.0 / K[r[1]];
  }
  return MLB_OK;
}

static Opts to_opts(const mlb_solver_opts* s) {
  Opts o;
  o.ctol = s->constraint_tol, o.ptol = s->position_tol, o.dtol = s->divergence_tol;
  o.rev_tol = s->reverse_check_tol;
  o.max_iters = s->max_iters, o.max_ls = s->max_line_search_iters;
  o.norm = s->norm, o.rev_norm = s->reverse_check_norm, o.flags = s->flags;
  return o;
}

// ------------------------------------------------------------------ layout helpers

// [n][dim] row-major <-> [dim][ld]; 32x32 shared-memory tile transpose, both sides coalesced
__global__ void k_transpose(int64_t rows, int64_t cols, const double* in, int64_t ld_in,
                            double* out, int64_t ld_out) {
  __shared__ double tile[32][33];
  int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = in[r * ld_in + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[c * ld_out + r] = tile[threadIdx.x][j];
  }
}

__global__ void k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double x = 1.0000001, y = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, x, y), a1 = fma(a1, x, y), a2 = fma(a2, x, y), a3 = fma(a3, x, y);
    a4 = fma(a4, x, y), a5 = fma(a5, x, y), a6 = fma(a6, x, y), a7 = fma(a7, x, y);
  }
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 12345.678) out[0] = s;
}

__global__ void k_selftest_fast_math(int64_t n, const double* x, double* r, double* s,
                                     double* rs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  r[i] = rcp_fast(x[i]);
  double t;
  s[i] = sqrt_rsqrt_fast(x[i], t);
  rs[i] = t;
}

__global__ void k_selftest_exp(int64_t n, const double* x, double* e, double* em1) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  e[i] = exp_bl(x[i]);
  em1[i] = expm1_bl(x[i]);
}

}  // namespace mlb

using namespace mlb;

extern "C" {

int mlb_version(void) { return MLB_VERSION; }
const char* mlb_last_error(void) { return g_err.c_str(); }
int64_t mlb_launch_count(void) { return g_launches.load(); }

int mlb_model_create(const mlb_model_desc* d, mlb_model** out) {
  MLB_ARG(d && out && d->data);
  bool hier = false, dense = false;
  int U = 0, Y = 0, n_steps = 0;
  if (d->model == MLB_MODEL_TOY) {
    MLB_ARG(d->n_data >= 3 && d->dim_y == 1);
    U = 2, Y = 1, dense = true;
  } else if (d->model == MLB_MODEL_HIER) {
    MLB_ARG(d->n_data >= 2 * d->dim_y && d->dim_y >= 3);
    U = 2, Y = d->dim_y, hier = true;
  } else if (d->model == MLB_MODEL_FN) {
    // data = {n_steps, y_obs[Y], dt_seq[n_steps], obs_step[Y]}
    MLB_ARG(d->dim_y >= 1 && d->n_data >= 1 + d->dim_y);
    n_steps = (int)d->data[0];
    MLB_ARG(n_steps >= 1 && d->n_data >= 1 + 2 * d->dim_y + n_steps);
    for (int i = 0; i < d->dim_y; ++i) {
      const double st = d->data[1 + d->dim_y + n_steps + i];
      MLB_ARG(st >= 0 && st < n_steps && (i == 0 || st >= d->data[d->dim_y + n_steps + i]));
    }
    U = 9, Y = d->dim_y;
  } else if (d->model == MLB_MODEL_HH) {
    // data = {n_steps, y_obs[Y], dt_seq[n_steps], obs_step[Y], stim[n_steps], consts[27]}
    MLB_ARG(d->dim_y >= 1 && d->n_data >= 1);
    n_steps = (int)d->data[0];
    MLB_ARG(n_steps >= 1 && d->n_data >= 1 + 2 * d->dim_y + 2 * n_steps + 27);
    for (int i = 0; i < d->dim_y; ++i) {
      const double st = d->data[1 + d->dim_y + n_steps + i];
      MLB_ARG(st >= 0 && st < n_steps && (i == 0 || st >= d->data[d->dim_y + n_steps + i]));
    }
    U = 7, Y = d->dim_y;
  } else {
    g_err = "model not available in this library build";
    return MLB_ERR_UNSUPPORTED;
  }
  const ModelOps* ops = find_ops(d->model, d->dim_y, d->parametrization);
  if (!ops) {
    g_err = "model / dim_y / parametrization combination not compiled in";
    return MLB_ERR_UNSUPPORTED;
  }
  mlb_model* m = new mlb_model();
  m->model = d->model, m->dim_y = d->dim_y, m->param = d->parametrization;
  m->data.assign(d->data, d->data + d->n_data);
  m->U = U, m->Y = Y, m->ops = ops, m->n_steps = n_steps;
  m->Q = U + (hier ? 2 : 1) * Y;
  m->K = dense ? Y : U;
  m->jac_rows = Y * U + (hier ? 2 : 1) * Y;
  m->gram_rows = m->K * m->K + Y;
  *out = m;
  return MLB_OK;
}
void mlb_model_destroy(mlb_model* m) {
  if (m && m->dev_blob) cudaFree(m->dev_blob);
  if (m && m->host_ws) cudaFree(m->host_ws);
  if (m && m->pinned_word) cudaFreeHost(m->pinned_word);
  if (m)
    for (void* s : m->host_streams)
      if (s) cudaStreamDestroy((cudaStream_t)s);
  delete m;
}
int mlb_scratch_trim(size_t keep_bytes) { return scratch_trim(keep_bytes); }
int mlb_model_dim_u(const mlb_model* m) { return m->U; }
int mlb_model_dim_y(const mlb_model* m) { return m->Y; }
int mlb_model_dim_q(const mlb_model* m) { return m->Q; }
int mlb_model_jac_rows(const mlb_model* m) { return m->jac_rows; }
int mlb_model_gram_rows(const mlb_model* m) { return m->gram_rows; }
int mlb_model_gram_dim(const mlb_model* m) { return m->K; }

#define MLB_COMMON_ARGS()              \
  MLB_ARG(m && n >= 0 && ld >= n);     \
  if (n == 0) return MLB_OK;           \
  cudaStream_t s = (cudaStream_t)stream

int mlb_constr(const mlb_model* m, int64_t n, int64_t ld, const double* q, double* c,
               void* stream) {
  MLB_COMMON_ARGS();
  MLB_ARG(q && c);
  return m->ops->constr(m, n, ld, q, c, s);
}

int mlb_jacob_constr_blocks(const mlb_model* m, int64_t n, int64_t ld, const double* q,
                            double* jac, double* c, void* stream) {
  MLB_COMMON_ARGS();
  MLB_ARG(q && jac);
  return m->ops->jacob(m, n, ld, q, jac, c, s);
}

int mlb_decompose_gram(const mlb_model* m, int64_t n, int64_t ld, const double* jac,
                       double* gram, void* stream) {
  MLB_COMMON_ARGS();
  MLB_ARG(jac && gram);
  return m->ops->decompose(m, n, ld, jac, gram, s);
}

int mlb_log_det_sqrt_gram(const mlb_model* m, int64_t n, int64_t ld, const double* q,
                          double* val, double* c, double* jac, double* gram, void* stream) {
  MLB_COMMON_ARGS();
  MLB_ARG(q);
  return m->ops->logdet(m, n, ld, q, val, nullptr, c, jac, gram, nullptr, nullptr, nullptr,
                        nullptr, s);
}

int mlb_val_and_grad_log_det_sqrt_gram(const mlb_model* m, int64_t n, int64_t ld,
                                       const double* q, double* val, double* grad,
                                       double* c, double* jac, double* gram, void* stream) {
  MLB_COMMON_ARGS();
  MLB_ARG(q && grad);
  return m->ops->logdet(m, n, ld, q, val, grad, c, jac, gram, nullptr, nullptr, nullptr,
                        nullptr, s);
}

int mlb_neg_log_dens(const mlb_model* m, int64_t n, int64_t ld, const double* q, double* val,
                     double* grad, void* stream) {
  MLB_COMMON_ARGS();
  MLB_ARG(q);
  return m->ops->nld(m, n, ld, q, val, grad, s);
}

int mlb_h1(const mlb_model* m, int64_t n, int64_t ld, const double* q, double* h1,
           double* dh1_dpos, void* stream) {
  MLB_COMMON_ARGS();
  MLB_ARG(q && (h1 || dh1_dpos));
  return m->ops->logdet(m, n, ld, q, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                        nullptr, h1, dh1_dpos, s);
}

int mlb_lmult_by_jacob_constr(const mlb_model* m, int64_t n, int64_t ld, const double* jac,
                              const double* vec, double* out, void* stream) {
  MLB_COMMON_ARGS();
  MLB_ARG(jac && vec && out);
  return m->ops->block_op(m, OP_LMULT, n, ld, jac, nullptr, nullptr, vec, out, s);
}
int mlb_rmult_by_jacob_constr(const mlb_model* m, int64_t n, int64_t ld, const double* jac,
                              const double* vec, double* out, void* stream) {
  MLB_COMMON_ARGS();
  MLB_ARG(jac && vec && out);
  return m->ops->block_op(m, OP_RMULT, n, ld, jac, nullptr, nullptr, vec, out, s);
}
int mlb_lmult_by_pinv_jacob_constr(const mlb_model* m, int64_t n, int64_t ld,
                                   const double* jac, const double* gram, const double* vec,
                                   double* out, void* stream) {
  MLB_COMMON_ARGS();
  MLB_ARG(jac && gram && vec && out);
  return m->ops->block_op(m, OP_PINV, n, ld, jac, nullptr, gram, vec, out, s);
}
int mlb_normal_space_component(const mlb_model* m, int64_t n, int64_t ld, const double* jac,
                               const double* gram, const double* vec, double* out,
                               void* stream) {
  MLB_COMMON_ARGS();
  MLB_ARG(jac && gram && vec && out);
  return m->ops->block_op(m, OP_NSC, n, ld, jac, nullptr, gram, vec, out, s);
}
int mlb_lmult_by_inv_jacob_product(const mlb_model* m, int64_t n, int64_t ld,
                                   const double* jac_l, const double* jac_r,
                                   const double* vec, double* out, void* stream) {
  MLB_COMMON_ARGS();
  MLB_ARG(jac_l && jac_r && vec && out);
  return m->ops->block_op(m, OP_INVJP, n, ld, jac_l, jac_r, nullptr, vec, out, s);
}

int mlb_project_onto_cotangent_space(const mlb_model* m, int64_t n, int64_t ld,
                                     const double* q, double* p, void* stream) {
  MLB_COMMON_ARGS();
  MLB_ARG(q && p);
  return m->ops->project_cotangent(m, n, ld, q, p, s);
}

int mlb_solve_projection(const mlb_model* m, int64_t n, int64_t ld, const double* q,
                         const double* jac_prev, const double* gram_prev, double dt,
                         const mlb_solver_opts* opts, double* q_out, double* mu_over_dt,
                         int32_t* iters, double* norm_dq, double* err,
                         int32_t* n_constr_calls, int32_t* status, void* stream) {
  MLB_COMMON_ARGS();
  MLB_ARG(q && jac_prev && opts && q_out && status);
  SolveIO io{q, jac_prev, gram_prev, dt, q_out, mu_over_dt, iters, norm_dq, err,
             n_constr_calls, status};
  return m->ops->solve(m, opts->solver, n, ld, to_opts(opts), io, s);
}

int mlb_leapfrog_steps(const mlb_model* m, int64_t n, int64_t ld, double* q, double* p,
                       double dt, const double* dt_per_chain, int32_t n_step,
                       const mlb_solver_opts* opts, int32_t* status, int32_t* n_done,
                       int32_t* iters_fwd, int32_t* iters_rev, double* h_init,
                       double* h_final, void* stream) {
  MLB_COMMON_ARGS();
  MLB_ARG(q && p && opts && status && n_step >= 0);
  StepsIO io{q, p, dt, dt_per_chain, n_step, status, n_done, iters_fwd, iters_rev, h_init,
             h_final};
  return m->ops->steps(m, opts->solver, n, ld, to_opts(opts), io, s);
}

int mlb_hmc_sample(const mlb_model* m, int64_t n, int64_t ld, double* q, double dt,
                   int32_t n_step, int32_t n_iter, const mlb_solver_opts* opts,
                   double max_delta_h, uint64_t seed, uint64_t chain0, uint32_t iter0,
                   const double* mom_noise, const double* unif, double* adapt,
                   const mlb_adapt_opts* adapt_opts, double* trace_q, int32_t trace_every,
                   int32_t trace_dim, double* trace_accept, double* stats,
                   int32_t* last_status, void* stream) {
  MLB_COMMON_ARGS();
  MLB_ARG(q && opts && n_step >= 0 && n_iter >= 0);
  MLB_ARG(!adapt || adapt_opts);
  SampleArgs a;
  a.dt = dt, a.max_delta_h = max_delta_h, a.n_step = n_step, a.n_iter = n_iter;
  a.trace_every = trace_every > 0 ? trace_every : 1;
  MLB_ARG(trace_dim >= 0 && trace_dim <= m->Q);
  a.trace_dim = trace_dim > 0 ? trace_dim : m->Q;
  a.seed = seed, a.chain0 = chain0, a.iter0 = iter0;
  a.mom_noise = mom_noise, a.unif = unif, a.adapt = adapt;
  a.a_target = adapt_opts ? adapt_opts->adapt_stat_target : 0.8;
  a.a_reg = adapt_opts ? adapt_opts->log_step_size_reg_coefficient : 0.05;
  a.a_decay = adapt_opts ? adapt_opts->iter_decay_coeff : 0.75;
  a.a_offset = adapt_opts ? adapt_opts->iter_offset : 10.0;
  a.trace_q = trace_q, a.trace_accept = trace_accept, a.stats = stats;
  a.last_status = last_status;
  return m->ops->sample(m, opts->solver, n, ld, q, to_opts(opts), a, s);
}

int mlb_nuts_sample(const mlb_model* m, int64_t n, int64_t ld, double* q, double dt,
                    int32_t max_tree_depth, int32_t n_iter, const mlb_solver_opts* opts,
                    double max_delta_h, uint64_t seed, uint64_t chain0, uint32_t iter0,
                    const double* mom_noise, double* adapt, const mlb_adapt_opts* adapt_opts,
                    double* trace_q, int32_t trace_every, int32_t trace_dim,
                    double* trace_accept, int32_t* trace_n_step, int32_t* trace_depth,
                    double* stats, int32_t* last_status, void* stream) {
  MLB_COMMON_ARGS();
  MLB_ARG(q && opts && n_iter >= 0 && max_tree_depth >= 0 && max_tree_depth <= 20);
  MLB_ARG(!adapt || adapt_opts);
  if (!m->ops->nuts) {
    g_err = "the dynamic sampler is not built for this model";
    return MLB_ERR_UNSUPPORTED;
  }
  SampleArgs a;
  a.dt = dt, a.max_delta_h = max_delta_h, a.n_step = 0, a.n_iter = n_iter;
  a.trace_every = trace_every > 0 ? trace_every : 1;
  MLB_ARG(trace_dim >= 0 && trace_dim <= m->Q);
  a.trace_dim = trace_dim > 0 ? trace_dim : m->Q;
  a.seed = seed, a.chain0 = chain0, a.iter0 = iter0;
  a.mom_noise = mom_noise, a.unif = nullptr, a.adapt = adapt;
  a.a_target = adapt_opts ? adapt_opts->adapt_stat_target : 0.8;
  a.a_reg = adapt_opts ? adapt_opts->log_step_size_reg_coefficient : 0.05;
  a.a_decay = adapt_opts ? adapt_opts->iter_decay_coeff : 0.75;
  a.a_offset = adapt_opts ? adapt_opts->iter_offset : 10.0;
  a.trace_q = trace_q, a.trace_accept = trace_accept, a.stats = stats;
  a.last_status = last_status;
  NutsArgs na;
  na.max_depth = max_tree_depth, na.scratch = nullptr, na.lds = 0;
  na.trace_n_step = trace_n_step, na.trace_depth = trace_depth;
  if (n == 0 || n_iter == 0) return MLB_OK;
  return m->ops->nuts(m, opts->solver, n, ld, q, to_opts(opts), a, na, s);
}

void mlb_philox_normals_host(uint64_t seed, uint64_t chain, uint32_t iter, int32_t n,
                             double* out) {
  for (int j = 0; 2 * j < n; ++j) {
    double a, b;
    philox_normal_pair(seed, chain, iter, (uint32_t)j, a, b);
    out[2 * j] = a;
    if (2 * j + 1 < n) out[2 * j + 1] = b;
  }
}
double mlb_philox_uniform_host(uint64_t seed, uint64_t chain, uint32_t iter) {
  return philox_uniform(seed, chain, iter);
}

int mlb_aos_to_soa(int64_t n, int32_t dim, const double* aos, double* soa, int64_t ld,
                   void* stream) {
  MLB_ARG(aos && soa && ld >= n && dim > 0);
  if (n == 0) return MLB_OK;
  dim3 grid((unsigned)((dim + 31) / 32), (unsigned)((n + 31) / 32)), block(32, 8);
  k_transpose<<<grid, block, 0, (cudaStream_t)stream>>>(n, dim, aos, dim, soa, ld);
  return finish_launch();
}
int mlb_soa_to_aos(int64_t n, int32_t dim, const double* soa, int64_t ld, double* aos,
                   void* stream) {
  MLB_ARG(aos && soa && ld >= n && dim > 0);
  if (n == 0) return MLB_OK;
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((dim + 31) / 32)), block(32, 8);
  k_transpose<<<grid, block, 0, (cudaStream_t)stream>>>(dim, n, soa, ld, aos, dim);
  return finish_launch();
}

int mlb_leapfrog_steps_host(const mlb_model* m, int64_t n, double* q, double* p, double dt,
                            int32_t n_step, const mlb_solver_opts* opts, int32_t* status,
                            int32_t* n_done, int32_t* iters_fwd, int32_t* iters_rev,
                            double* h_init, double* h_final) {
  MLB_ARG(m && q && p && opts && status && n >= 0);
  if (n == 0) return MLB_OK;
  const int Q = m->Q;
  // Chains are independent, so the batch is cut into chunks that flow through
  // H2D copy -> fused kernel -> D2H copy on a few streams: the copies of one chunk overlap
  // the kernel of another (PCIe is full duplex), and the call costs about
  // max(copy time, kernel time) instead of their sum.  The register-resident family reads
  // and writes the reference's [n][Q] rows directly (it touches the state once per launch),
  // so a chunk is 2 + 1 + 2 API calls plus one per requested per-chain output; the
  // streaming (ODE) family keeps one chunk and transposes on the device: its batches are
  // small and its kernels long.
  const bool small_state = m->model == MLB_MODEL_TOY || m->model == MLB_MODEL_HIER;
  // chunk size: eight chunks in flight (one per stream), between 2,048 and 16,384 chains.
  // Below ~16,384 chains the fused kernel's duration no longer depends on the chunk size
  // (one warp per scheduler: the latency of a single chain's trajectory, 0.25 ms for 32
  // eight-schools steps), so the call costs about H2D of everything + one kernel + the last
  // chunk's D2H; with eight streams 8,192-chain chunks measured best for 65,536 chains
  // (2.23e9 steps/s against 2.11e9 with four 16,384-chain chunks, profiles/r6n_e2e_sweep.txt;
  // with four streams a fifth chunk waited for the first one's D2H and smaller chunks lost)
  int64_t chunk = n;
  if (small_state) {
    static const int64_t forced = [] {  // MLB_HOST_CHUNK: chunk size for experiments
      const char* e = std::getenv("MLB_HOST_CHUNK");
      return e ? (int64_t)std::atoll(e) : (int64_t)0;
    }();
    static const int64_t parts = [] {  // MLB_HOST_PARTS: chunks a batch is cut into
      const char* e = std::getenv("MLB_HOST_PARTS");
      const int64_t v = e ? (int64_t)std::atoll(e) : 0;
      return v >= 1 ? v : (int64_t)8;
    }();
    chunk = ((n + parts - 1) / parts + 63) / 64 * 64;
    chunk = chunk < 2048 ? 2048 : chunk > 16384 ? 16384 : chunk;
    if (forced >= 64) chunk = forced / 64 * 64;
  }
  const int n_chunk = (int)((n + chunk - 1) / chunk);
  // per chunk: row-major state (2 vec), component-major state (2 vec, ODE family only),
  // per-chain outputs; the staging area lives in the model handle and only grows: no
  // cudaMalloc / cudaFree (both synchronise the device) on the steady-state path.
  const size_t per_chain = (size_t)(small_state ? 2 : 4) * Q * sizeof(double) +
                           2 * sizeof(double) + 4 * sizeof(int32_t);
  const size_t need = per_chain * (size_t)n + 256 * (size_t)n_chunk;
  std::lock_guard<std::mutex> lock(m->host_mutex);
  if (m->host_ws_bytes < need) {
    if (m->host_ws) cudaFree(m->host_ws);
    m->host_ws = nullptr, m->host_ws_bytes = 0;
    MLB_CUDA_OK(cudaMalloc(&m->host_ws, need));
    m->host_ws_bytes = need;
  }
  for (int k = 0; k < mlb_model::kHostStreams; ++k)
    if (!m->host_streams[k])
      MLB_CUDA_OK(cudaStreamCreateWithFlags((cudaStream_t*)&m->host_streams[k],
                                            cudaStreamNonBlocking));
#define MLB_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { g_err = std::string(#expr) + ": " + cudaGetErrorString(e_); rc = MLB_ERR_CUDA; goto done; } } while (0)
  int rc = MLB_OK;
  {
    const Opts o = to_opts(opts);
    char* base = static_cast<char*>(m->host_ws);
    for (int k = 0; k < n_chunk && rc == MLB_OK; ++k) {
      const int64_t i0 = k * chunk, nc = std::min(chunk, n - i0);
      cudaStream_t s = (cudaStream_t)m->host_streams[k % mlb_model::kHostStreams];
      char* dev = base + (per_chain * (size_t)i0 + 256 * (size_t)k) / 256 * 256;
      const size_t vec_bytes = (size_t)nc * Q * sizeof(double);
      double* d_aos_q = (double*)dev;
      double* d_aos_p = d_aos_q + (size_t)nc * Q;
      double* d_q = small_state ? d_aos_q : d_aos_p + (size_t)nc * Q;
      double* d_p = small_state ? d_aos_p : d_q + (size_t)nc * Q;
      double* d_h0 = d_p + (size_t)nc * Q;
      double* d_h1 = d_h0 + nc;
      int32_t* d_status = (int32_t*)(d_h1 + nc);
      int32_t *d_done = d_status + nc, *d_itf = d_done + nc, *d_itr = d_itf + nc;
      MLB_TRY(cudaMemcpyAsync(d_aos_q, q + i0 * Q, vec_bytes, cudaMemcpyHostToDevice, s));
      MLB_TRY(cudaMemcpyAsync(d_aos_p, p + i0 * Q, vec_bytes, cudaMemcpyHostToDevice, s));
      if (!small_state) {
        if ((rc = mlb_aos_to_soa(nc, Q, d_aos_q, d_q, nc, s)) != MLB_OK) break;
        if ((rc = mlb_aos_to_soa(nc, Q, d_aos_p, d_p, nc, s)) != MLB_OK) break;
      }
      // energies only if asked for: their logarithms are not free (mlb_chmc.cuh::evaluate_at)
      StepsIO io{d_q, d_p, dt, nullptr, n_step, d_status, d_done, d_itf, d_itr,
                 h_init ? d_h0 : nullptr, h_final ? d_h1 : nullptr};
      io.row_major = small_state;
      io.batch_total = n;
      if ((rc = m->ops->steps(m, opts->solver, nc, nc, o, io, s)) != MLB_OK) break;
      if (!small_state) {
        if ((rc = mlb_soa_to_aos(nc, Q, d_q, nc, d_aos_q, s)) != MLB_OK) break;
        if ((rc = mlb_soa_to_aos(nc, Q, d_p, nc, d_aos_p, s)) != MLB_OK) break;
      }
      MLB_TRY(cudaMemcpyAsync(q + i0 * Q, d_aos_q, vec_bytes, cudaMemcpyDeviceToHost, s));
      MLB_TRY(cudaMemcpyAsync(p + i0 * Q, d_aos_p, vec_bytes, cudaMemcpyDeviceToHost, s));
      MLB_TRY(cudaMemcpyAsync(status + i0, d_status, nc * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
      if (n_done) MLB_TRY(cudaMemcpyAsync(n_done + i0, d_done, nc * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
      if (iters_fwd) MLB_TRY(cudaMemcpyAsync(iters_fwd + i0, d_itf, nc * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
      if (iters_rev) MLB_TRY(cudaMemcpyAsync(iters_rev + i0, d_itr, nc * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
      if (h_init) MLB_TRY(cudaMemcpyAsync(h_init + i0, d_h0, nc * sizeof(double), cudaMemcpyDeviceToHost, s));
      if (h_final) MLB_TRY(cudaMemcpyAsync(h_final + i0, d_h1, nc * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
  }
done:
  for (int k = 0; k < mlb_model::kHostStreams; ++k) {
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)m->host_streams[k]);
    if (e != cudaSuccess && rc == MLB_OK) {
      g_err = std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e);
      rc = MLB_ERR_CUDA;
    }
  }
#undef MLB_TRY
  return rc;
}

int mlb_selftest_fast_math(int64_t n, const double* x, double* rcp, double* sqrt_out,
                           double* rsqrt_out, void* stream) {
  MLB_ARG(n >= 0 && x && rcp && sqrt_out && rsqrt_out);
  if (n == 0) return MLB_OK;
  k_selftest_fast_math<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      n, x, rcp, sqrt_out, rsqrt_out);
  return finish_launch();
}

int mlb_selftest_exp(int64_t n, const double* x, double* exp_out, double* expm1_out,
                     void* stream) {
  MLB_ARG(n >= 0 && x && exp_out && expm1_out);
  if (n == 0) return MLB_OK;
  k_selftest_exp<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, x, exp_out,
                                                                                 expm1_out);
  return finish_launch();
}

double mlb_measure_fp64_tflops(int32_t repeats) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1.0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* out = nullptr;
  if (cudaMalloc((void**)&out, sizeof(double)) != cudaSuccess) return -1.0;
  const int iters = 1 << 14, threads = 256, blocks = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k_fp64_peak<<<blocks, threads>>>(out, iters);  // warm-up
  double best = 0.0;
  for (int r = 0; r < (repeats > 0 ? repeats : 3); ++r) {
    cudaEventRecord(e0);
    k_fp64_peak<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 8.0 * (double)iters * threads * blocks;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  g_launches.fetch_add(1 + (repeats > 0 ? repeats : 3));
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

}  // extern "C"
