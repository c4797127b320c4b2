// sm_100a kernels + C ABI for the one-chain-per-thread models (toy, eight-schools).
//
// Execution model: chains are embarrassingly parallel (the reference runs one OS process
// per chain, toy_2d_model.py:450); here one CUDA thread owns one chain, the whole chain
// state (q, p, Jacobian blocks, Gram factor, gradient) lives in registers, and chain
// state in HBM is component-major so that a warp's 32 loads of component k are one
// coalesced 256 B transaction.  Model constants travel as __grid_constant__ kernel
// parameters (constant bank operands).  The fused drivers keep state on chip across
// leapfrog steps / HMC transitions so HBM is touched once per launch.
#pragma once
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "mlb_chmc.cuh"
#include "mlb_hier_lean.cuh"
#include "mlb_tma.cuh"

namespace mlb { struct ModelOps; }

struct mlb_model {
  int model, dim_y, param, U, Y, Q, jac_rows, gram_rows, K;
  std::vector<double> data;
  const struct mlb::ModelOps* ops;
  // ODE models: device copy of the constants, uploaded on first use (mlb_api.cu)
  mutable std::mutex dev_mutex;
  mutable void* dev_blob = nullptr;
  mutable int dev_id = -1;
  int n_steps = 0;
  // mlb_leapfrog_steps_host: grow-only device staging area (no malloc / free per call)
  mutable std::mutex host_mutex;  // serialises mlb_leapfrog_steps_host calls on the handle
  mutable void* host_ws = nullptr;
  mutable size_t host_ws_bytes = 0;
  static constexpr int kHostStreams = 8;  // chunk pipeline of the host entry point
  mutable void* host_streams[kHostStreams] = {};
  mutable std::mutex pin_mutex;
  mutable int* pinned_word = nullptr;  // polled by the host-driven streaming path
};

namespace mlb {

std::string& last_error();           // thread-local, defined in mlb_api.cu
std::atomic<int64_t>& launch_count();  // defined in mlb_api.cu
#define g_err (::mlb::last_error())

// Stream-ordered scratch (tree storage, sweep scratch of the ODE / state-space kernels)
// comes from a memory pool OWNED BY THE LIBRARY, one per device, whose release threshold is
// unlimited so that freed blocks stay cached between calls: with the default threshold a
// host loop that synchronises once per round paid for a fresh physical allocation of
// several hundred MB every round.  The process-wide default pool of the host application
// (torch's allocator sits next to it) is left untouched; mlb_scratch_trim() hands the
// cached memory back.  Defined in mlb_linalg.cu.
cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t s);
int scratch_trim(size_t keep_bytes);
#define g_launches (::mlb::launch_count())

#define MLB_CUDA_OK(expr)                                                       \
  do {                                                                          \
    cudaError_t e_ = (expr);                                                    \
    if (e_ != cudaSuccess) {                                                    \
      g_err = std::string(#expr) + ": " + cudaGetErrorString(e_);               \
      return MLB_ERR_CUDA;                                                      \
    }                                                                           \
  } while (0)

#define MLB_ARG(cond)                                   \
  do {                                                  \
    if (!(cond)) {                                      \
      g_err = std::string("bad argument: ") + #cond;    \
      return MLB_ERR_ARG;                               \
    }                                                   \
  } while (0)

template <int D>
MLB_DEV void load_vec(double (&v)[D], const double* a, int64_t ld, int64_t i) {
#pragma unroll
  for (int k = 0; k < D; ++k) v[k] = a[(int64_t)k * ld + i];
}
template <int D>
MLB_DEV void store_vec(const double (&v)[D], double* a, int64_t ld, int64_t i) {
#pragma unroll
  for (int k = 0; k < D; ++k) a[(int64_t)k * ld + i] = v[k];
}

// element (k, chain i) at a[k * es + i * cs]: component-major (es = ld, cs = 1) or the
// reference's chain-major rows (es = 1, cs = D)
template <int D>
MLB_DEV void load_vec(double (&v)[D], const double* a, int64_t es, int64_t cs, int64_t i) {
#pragma unroll
  for (int k = 0; k < D; ++k) v[k] = a[(int64_t)k * es + i * cs];
}
template <int D>
MLB_DEV void store_vec(const double (&v)[D], double* a, int64_t es, int64_t cs, int64_t i) {
#pragma unroll
  for (int k = 0; k < D; ++k) a[(int64_t)k * es + i * cs] = v[k];
}

constexpr int kBlock = 64;  // threads per CTA for the heavy per-chain kernels
#ifndef MLB_LF_MIN_BLOCKS
#define MLB_LF_MIN_BLOCKS 1  // min resident CTAs per SM asked of the fused kernels
#endif

#define MLB_CHAIN_INDEX()                                          \
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; \
  if (i >= n) return;

// ------------------------------------------------------------ individual operations

template <class M>
__global__ void __launch_bounds__(kBlock) k_constr(const __grid_constant__
                                                   typename M::Params P,
                                                   int64_t n, int64_t ld, const double* q,
                                                   double* c) {
  MLB_CHAIN_INDEX();
  double qv[M::Q], cv[M::Y];
  load_vec<M::Q>(qv, q, ld, i);
  M::constr(P, qv, cv);
  store_vec<M::Y>(cv, c, ld, i);
}

template <class M>
__global__ void __launch_bounds__(kBlock)
    k_jacob(const __grid_constant__ typename M::Params P, int64_t n, int64_t ld,
            const double* q, double* jac, double* c) {
  MLB_CHAIN_INDEX();
  double qv[M::Q], cv[M::Y];
  typename M::B::Jac J;
  load_vec<M::Q>(qv, q, ld, i);
  M::jacob(P, qv, J, cv);
  M::B::store_jac(J, jac, ld, i);
  if (c) store_vec<M::Y>(cv, c, ld, i);
}

template <class M>
__global__ void __launch_bounds__(kBlock)
    k_decompose(int64_t n, int64_t ld, const double* jac, double* gram) {
  MLB_CHAIN_INDEX();
  typename M::B::Jac J;
  typename M::B::Gram G;
  M::B::load_jac(J, jac, ld, i);
  M::B::decompose_gram(J, G);
  M::B::store_gram(G, gram, ld, i);
}

// log_det_sqrt_gram, optionally with gradient, aux outputs and h1 / dh1_dpos
template <class M>
__global__ void __launch_bounds__(kBlock)
    k_logdet(const __grid_constant__ typename M::Params P, int64_t n, int64_t ld,
             const double* q, double* val, double* grad, double* c, double* jac,
             double* gram, double* nld_val, double* nld_grad, double* h1, double* dh1) {
  MLB_CHAIN_INDEX();
  constexpr int Q = M::Q;
  double qv[Q], cv[M::Y];
  typename M::B::Jac J;
  typename M::B::Gram G;
  load_vec<Q>(qv, q, ld, i);
  M::jacob(P, qv, J, cv);
  M::B::decompose_gram(J, G);
  const double ld_ = M::B::log_det_sqrt_gram(G);
  if (val) val[i] = ld_;
  if (c) store_vec<M::Y>(cv, c, ld, i);
  if (jac) M::B::store_jac(J, jac, ld, i);
  if (gram) M::B::store_gram(G, gram, ld, i);
  if (grad || nld_val || nld_grad || h1 || dh1) {
    double gn[Q], gl[Q];
    const typename M::Par par = M::params(qv[0], qv[1]);
    const double nv = M::nld(P, qv, par, gn);
#pragma unroll
    for (int k = 0; k < Q; ++k) gl[k] = 0.0;
    M::add_grad_logdet(P, qv, par, J, G, gl);
    if (grad) store_vec<Q>(gl, grad, ld, i);
    if (nld_val) nld_val[i] = nv;
    if (nld_grad) store_vec<Q>(gn, nld_grad, ld, i);
    if (h1) h1[i] = nv + ld_;
    if (dh1) {
#pragma unroll
      for (int k = 0; k < Q; ++k) gn[k] += gl[k];
      store_vec<Q>(gn, dh1, ld, i);
    }
  }
}

template <class M>
__global__ void __launch_bounds__(kBlock)
    k_nld(const __grid_constant__ typename M::Params P, int64_t n, int64_t ld,
          const double* q, double* val, double* grad) {
  MLB_CHAIN_INDEX();
  double qv[M::Q], g[M::Q];
  load_vec<M::Q>(qv, q, ld, i);
  const double v = M::nld(P, qv, M::params(qv[0], qv[1]), g);
  if (val) val[i] = v;
  if (grad) store_vec<M::Q>(g, grad, ld, i);
}

enum { OP_LMULT = 0, OP_RMULT = 1, OP_PINV = 2, OP_NSC = 3, OP_INVJP = 4 };

template <class M, int OP>
__global__ void __launch_bounds__(kBlock)
    k_block_op(int64_t n, int64_t ld, const double* jac, const double* jac_r,
               const double* gram, const double* vec, double* out) {
  MLB_CHAIN_INDEX();
  using B = typename M::B;
  constexpr int Q = M::Q, Y = M::Y;
  typename B::Jac J;
  B::load_jac(J, jac, ld, i);
  if (OP == OP_LMULT) {
    double v[Q], o[Y];
    load_vec<Q>(v, vec, ld, i);
    B::lmult_jacob(J, v, o);
    store_vec<Y>(o, out, ld, i);
  } else if (OP == OP_RMULT) {
    double w[Y], o[Q];
    load_vec<Y>(w, vec, ld, i);
    B::rmult_jacob(J, w, o);
    store_vec<Q>(o, out, ld, i);
  } else if (OP == OP_PINV) {
    typename B::Gram G;
    B::load_gram(G, gram, ld, i);
    double w[Y], o[Q];
    load_vec<Y>(w, vec, ld, i);
    B::lmult_pinv(J, G, w, o);
    store_vec<Q>(o, out, ld, i);
  } else if (OP == OP_NSC) {
    typename B::Gram G;
    B::load_gram(G, gram, ld, i);
    double v[Q], w[Y], o[Q];
    load_vec<Q>(v, vec, ld, i);
    B::lmult_jacob(J, v, w);
    B::lmult_pinv(J, G, w, o);
    store_vec<Q>(o, out, ld, i);
  } else {
    typename B::Jac Jr;
    B::load_jac(Jr, jac_r, ld, i);
    double w[Y], o[Y];
    load_vec<Y>(w, vec, ld, i);
    bool ok = B::lmult_inv_jacob_product(J, Jr, w, o);
    if (!ok) {
#pragma unroll
      for (int k = 0; k < Y; ++k) o[k] = nan("");
    }
    store_vec<Y>(o, out, ld, i);
  }
}

template <class M>
__global__ void __launch_bounds__(kBlock)
    k_project_cotangent(const __grid_constant__ typename M::Params P, int64_t n, int64_t ld,
                        const double* q, double* p) {
  MLB_CHAIN_INDEX();
  double qv[M::Q], pv[M::Q], cv[M::Y];
  typename M::B::Jac J;
  typename M::B::Gram G;
  load_vec<M::Q>(qv, q, ld, i);
  load_vec<M::Q>(pv, p, ld, i);
  M::jacob(P, qv, J, cv);
  M::B::decompose_gram(J, G);
  M::B::project_cotangent(J, G, pv);
  store_vec<M::Q>(pv, p, ld, i);
}

template <class M, int SOLVER>
__global__ void __launch_bounds__(kBlock)
    k_solve(const __grid_constant__ typename M::Params P, int64_t n, int64_t ld,
            const double* q, const double* jac_prev, const double* gram_prev, double dt,
            Opts o, double* q_out, double* mu_out, int32_t* iters, double* norm_dq,
            double* err, int32_t* n_constr, int32_t* status) {
  MLB_CHAIN_INDEX();
  using B = typename M::B;
  double qv[M::Q], mu[M::Q];
  typename B::Jac Jp;
  typename B::Gram Gp;
  load_vec<M::Q>(qv, q, ld, i);
  B::load_jac(Jp, jac_prev, ld, i);
  if (gram_prev) {
    B::load_gram(Gp, gram_prev, ld, i);
  } else {
    B::decompose_gram(Jp, Gp);
  }
  SolveOut r = solve_projection<M, SOLVER>(P, qv, Jp, Gp, dt, o, mu);
  store_vec<M::Q>(qv, q_out, ld, i);
  if (mu_out) store_vec<M::Q>(mu, mu_out, ld, i);
  if (iters) iters[i] = r.iters;
  if (norm_dq) norm_dq[i] = r.ndq;
  if (err) err[i] = r.err;
  if (n_constr) n_constr[i] = r.n_constr;
  status[i] = r.status;
}

// ----------------------------------------------------------------- fused drivers

// State staging (north-star: "chain state staged through shared memory via TMA"): when the
// CTA's [2 Q][kBlock] tile of a component-major state is whole and 16-byte aligned, ONE
// thread fetches it with 2 Q bulk asynchronous copies (cp.async.bulk, mbarrier completion)
// straight into the shared-memory tile that holds the last committed state anyway, and the
// final state leaves the same way; ragged / row-major / unaligned tiles take per-thread
// loads and stores.  MLB_TMA_STATE=0 compiles the copies out (A/B: profiles/r5e_*).
#ifndef MLB_TMA_STATE
#define MLB_TMA_STATE 1
#endif

}  // namespace mlb
#include "mlb_duo.cuh"
namespace mlb {

template <class M, int SOLVER>
__global__ void __launch_bounds__(kBlock, MLB_LF_MIN_BLOCKS)
    k_leapfrog_steps(const __grid_constant__ typename M::Params P, int64_t n, int64_t es,
                     int64_t cs, double* q, double* p, double dt, const double* dt_pc, int n_step,
                     Opts o, int32_t* status, int32_t* n_done, int32_t* iters_fwd,
                     int32_t* iters_rev, double* h_init, double* h_final) {
  constexpr int Q = M::Q;
  __shared__ __align__(128) double commit[2 * Q][kBlock];  // last committed (q, p) of each chain
  __shared__ uint64_t bar;
  const int64_t i0 = (int64_t)blockIdx.x * kBlock;
  const bool tile = MLB_TMA_STATE && i0 + kBlock <= n && bulk_tile_ok(q, es, cs) &&
                    bulk_tile_ok(p, es, cs);  // uniform over the CTA
  if (tile) {
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar, 2 * Q * kBlock * (int)sizeof(double));
#pragma unroll 1
      for (int k = 0; k < Q; ++k) {
        bulk_g2s(&commit[k][0], q + k * es + i0, kBlock * sizeof(double), &bar);
        bulk_g2s(&commit[Q + k][0], p + k * es + i0, kBlock * sizeof(double), &bar);
      }
    }
  }
  MLB_CHAIN_INDEX();
  double qv[Q], pv[Q];
  if (tile) {
    mbar_wait(&bar, 0);
#pragma unroll
    for (int k = 0; k < Q; ++k) qv[k] = commit[k][threadIdx.x], pv[k] = commit[Q + k][threadIdx.x];
  } else {
    load_vec<Q>(qv, q, es, cs, i);
    load_vec<Q>(pv, p, es, cs, i);
  }
  const double dti = dt_pc ? dt_pc[i] : dt;
  const TrajOut t = run_trajectory<M, SOLVER, false, false, kBlock>(
      P, qv, pv, &commit[0][threadIdx.x], &commit[Q][threadIdx.x], n_step, dti, o, 0.0,
      h_init != nullptr || h_final != nullptr);
  if (tile) {
#pragma unroll
    for (int k = 0; k < Q; ++k) commit[k][threadIdx.x] = qv[k], commit[Q + k][threadIdx.x] = pv[k];
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll 1
      for (int k = 0; k < Q; ++k) {
        bulk_s2g(q + k * es + i0, &commit[k][0], kBlock * sizeof(double));
        bulk_s2g(p + k * es + i0, &commit[Q + k][0], kBlock * sizeof(double));
      }
      bulk_commit();
    }
  } else {
    store_vec<Q>(qv, q, es, cs, i);
    store_vec<Q>(pv, p, es, cs, i);
  }
  status[i] = t.status;
  if (n_done) n_done[i] = t.done;
  if (iters_fwd) iters_fwd[i] = t.it_fwd;
  if (iters_rev) iters_rev[i] = t.it_rev;
  if (h_init) h_init[i] = t.h_init;
  if (h_final) h_final[i] = t.h_last;
  if (tile && threadIdx.x == 0) bulk_wait_read();  // the tile must outlive the copies' reads
}

// Shared-memory-parked form of k_leapfrog_steps (mlb_hier_lean.cuh): same device functions,
// 128 registers, 7 CTAs per SM.  Used when it turns two waves of the register-resident
// kernel into one (see Launch::steps).  The initial state goes into the parked columns by
// bulk asynchronous copies as above (the committed state of this form is the global arrays
// themselves, written step by step, so nothing is left to store at the end).
template <class M, int SOLVER>
__global__ void __maxnreg__(128)
    k_leapfrog_steps_lean(const __grid_constant__ typename M::Params P, int64_t n, int64_t es,
                          int64_t cs, double* q, double* p, double dt, const double* dt_pc,
                          int n_step, Opts o, int32_t* status, int32_t* n_done,
                          int32_t* iters_fwd, int32_t* iters_rev, double* h_init,
                          double* h_final) {
  extern __shared__ __align__(128) double lean_columns[];  // [LeanCol::ROWS][kBlock]
  __shared__ uint64_t bar;
  constexpr int Q = M::Q;
  using LC = LeanCol<M, kBlock>;
  static_assert(LC::SQ == 0 && LC::SP == Q, "state rows lead the column");
  const int64_t i0 = (int64_t)blockIdx.x * kBlock;
  const bool tile = MLB_TMA_STATE && i0 + kBlock <= n && bulk_tile_ok(q, es, cs) &&
                    bulk_tile_ok(p, es, cs);
  if (tile) {
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar, 2 * Q * kBlock * (int)sizeof(double));
#pragma unroll 1
      for (int k = 0; k < Q; ++k) {
        bulk_g2s(lean_columns + (LC::SQ + k) * kBlock, q + k * es + i0, kBlock * sizeof(double), &bar);
        bulk_g2s(lean_columns + (LC::SP + k) * kBlock, p + k * es + i0, kBlock * sizeof(double), &bar);
      }
    }
  }
  MLB_CHAIN_INDEX();
  if (tile) mbar_wait(&bar, 0);
  const LC S{lean_columns + threadIdx.x};
  const TrajOut t = run_trajectory_lean<M, SOLVER, kBlock>(
      P, S, q + i * cs, p + i * cs, es, n_step, dt_pc ? dt_pc[i] : dt, o,
      h_init != nullptr || h_final != nullptr, tile);
  status[i] = t.status;
  if (n_done) n_done[i] = t.done;
  if (iters_fwd) iters_fwd[i] = t.it_fwd;
  if (iters_rev) iters_rev[i] = t.it_rev;
  if (h_init) h_init[i] = t.h_init;
  if (h_final) h_final[i] = t.h_last;
}

struct SampleArgs {
  double dt, max_delta_h;
  int n_step, n_iter, trace_every, trace_dim;
  uint64_t seed, chain0;
  uint32_t iter0;
  const double* mom_noise;
  const double* unif;
  double* adapt;
  double a_target, a_reg, a_decay, a_offset;
  double* trace_q;
  double* trace_accept;
  double* stats;
  int32_t* last_status;
};

}  // namespace mlb
#include "mlb_nuts.cuh"
namespace mlb {

template <class M, int SOLVER>
__global__ void __launch_bounds__(kBlock, MLB_LF_MIN_BLOCKS)
    k_hmc_sample(const __grid_constant__ typename M::Params P, int64_t n, int64_t ld,
                 double* q, Opts o, SampleArgs a) {
  constexpr int Q = M::Q;
  __shared__ double commit[2 * Q][kBlock];
  MLB_CHAIN_INDEX();
  double ad_iter = 0, ad_smooth = 0, ad_err = 0, ad_target = 0, dti = a.dt;
  if (a.adapt) {
    ad_iter = a.adapt[0 * ld + i], ad_smooth = a.adapt[1 * ld + i];
    ad_err = a.adapt[2 * ld + i], ad_target = a.adapt[3 * ld + i];
    dti = a.adapt[4 * ld + i];
  }
  double s_acc = 0, s_nacc = 0, s_conv = 0, s_nonrev = 0, s_hdiv = 0, s_steps = 0, s_its = 0;
  int last = MLB_ST_OK;
#pragma unroll 1
  for (int it = 0; it < a.n_iter; ++it) {
    // the chain's current position stays in q (global); the proposal lives in registers
    double pc[Q], qc[Q];
    load_vec<Q>(qc, q, ld, i);
    if (a.mom_noise) {
      load_vec<Q>(pc, a.mom_noise + (int64_t)it * Q * ld, ld, i);
    } else {
      // rolled loop (the Philox + Box-Muller body is ~400 instructions): normals go
      // through this thread's shared-memory column, then into registers
      double* col = &commit[Q][threadIdx.x];
#pragma unroll 1
      for (int j = 0; 2 * j < Q; ++j) {
        double n0, n1;
        philox_normal_pair(a.seed, a.chain0 + (uint64_t)i, a.iter0 + (uint32_t)it,
                           (uint32_t)j, n0, n1);
        col[(2 * j) * kBlock] = n0;
        if (2 * j + 1 < Q) col[(2 * j + 1) * kBlock] = n1;
      }
#pragma unroll
      for (int k = 0; k < Q; ++k) pc[k] = col[k * kBlock];
    }
    const double u = a.unif ? a.unif[(int64_t)it * ld + i]
                            : philox_uniform(a.seed, a.chain0 + (uint64_t)i,
                                             a.iter0 + (uint32_t)it);
    const TrajOut t = run_trajectory<M, SOLVER, true, true, kBlock>(
        P, qc, pc, &commit[0][threadIdx.x], &commit[Q][threadIdx.x], a.n_step, dti, o,
        a.max_delta_h);
    const int st = t.status;
    s_its += t.it_fwd + t.it_rev;
    s_steps += t.done;
    double a_stat = 0.0;
    bool accept = false;
    if (st == MLB_ST_OK) {
      double d = t.h_init - t.h_last;
      a_stat = (d != d) ? 0.0 : exp(d < 0.0 ? d : 0.0);
      accept = u < a_stat;
      if (accept) {
        store_vec<Q>(qc, q, ld, i);
        s_nacc += 1;
      }
    }
    s_acc += a_stat;
    s_conv += (st == MLB_ST_DIVERGED || st == MLB_ST_NOT_CONVERGED || st == MLB_ST_LINALG);
    s_nonrev += (st == MLB_ST_NON_REVERSIBLE);
    s_hdiv += (st == MLB_ST_H_DIVERGED);
    last = st;
    if (a.trace_accept) a.trace_accept[(int64_t)it * ld + i] = a_stat;
    if (a.trace_q && (it + 1) % a.trace_every == 0) {
      if (!accept) load_vec<Q>(qc, q, ld, i);
      double* tq = a.trace_q + (int64_t)((it + 1) / a.trace_every - 1) * a.trace_dim * ld + i;
#pragma unroll
      for (int k = 0; k < Q; ++k)
        if (k < a.trace_dim) tq[(int64_t)k * ld] = qc[k];
    }
    if (a.adapt) {  // Mici DualAveragingStepSizeAdapter.update
      ad_iter += 1;
      const double w = 1.0 / (a.a_offset + ad_iter);
      ad_err = (1.0 - w) * ad_err + w * (a.a_target - a_stat);
      const double sw = pow(1.0 / ad_iter, a.a_decay);
      const double log_eps = ad_target - ad_err * sqrt(ad_iter) / a.a_reg;
      ad_smooth = (1.0 - sw) * ad_smooth + sw * log_eps;
      dti = exp(log_eps);
    }
  }
  if (a.adapt) {
    a.adapt[0 * ld + i] = ad_iter, a.adapt[1 * ld + i] = ad_smooth;
    a.adapt[2 * ld + i] = ad_err, a.adapt[4 * ld + i] = dti;
  }
  if (a.stats) {
    a.stats[0 * ld + i] += s_acc, a.stats[1 * ld + i] += s_nacc;
    a.stats[2 * ld + i] += s_conv, a.stats[3 * ld + i] += s_nonrev;
    a.stats[4 * ld + i] += s_hdiv, a.stats[5 * ld + i] += s_steps;
    a.stats[6 * ld + i] += s_its;
  }
  if (a.last_status) a.last_status[i] = last;
}


// ---------------------------------------------------------------- per-model launchers

struct SolveIO {
  const double *q, *jac_prev, *gram_prev;
  double dt;
  double *q_out, *mu_out;
  int32_t* iters;
  double *norm_dq, *err;
  int32_t *n_constr, *status;
};
struct StepsIO {
  double *q, *p;
  double dt;
  const double* dt_pc;
  int n_step;
  int32_t *status, *n_done, *iters_fwd, *iters_rev;
  double *h_init, *h_final;
  // q, p given as chain-major rows [n][Q] (the reference's layout) instead of [Q][ld];
  // only the register-resident family reads it (it touches q, p once per launch)
  bool row_major = false;
  // chains of the whole batch this launch is a chunk of (host entry point: its chunks run
  // concurrently, so the small-launch form is chosen by the batch, not by the chunk); 0 = n
  int64_t batch_total = 0;
  // sampler use of the ODE family's host-driven path: project p onto the cotangent space
  // first (h_init is taken after that), and stop a chain with MLB_ST_H_DIVERGED when
  // |h - h_init| > max_delta_h after a step (<= 0: no check)
  bool project_first = false;
  double max_delta_h = 0.0;
  // chains with alive_mask[i] == 0 are left untouched (status MLB_ST_OK, nothing done);
  // NULL = all chains (the host-driven dynamic sampler steps only the chains whose tree
  // is still growing)
  const int32_t* alive_mask = nullptr;
};

// Table of launchers for one compiled-in model; one translation unit per model so the
// (model x solver) kernel instantiations compile in parallel.
struct ModelOps {
  int model, dim_y, param;
  int (*constr)(const mlb_model*, int64_t, int64_t, const double*, double*, cudaStream_t);
  int (*jacob)(const mlb_model*, int64_t, int64_t, const double*, double*, double*, cudaStream_t);
  int (*decompose)(const mlb_model*, int64_t, int64_t, const double*, double*, cudaStream_t);
  int (*logdet)(const mlb_model*, int64_t, int64_t, const double*, double*, double*, double*,
                double*, double*, double*, double*, double*, double*, cudaStream_t);
  int (*nld)(const mlb_model*, int64_t, int64_t, const double*, double*, double*, cudaStream_t);
  int (*block_op)(const mlb_model*, int, int64_t, int64_t, const double*, const double*,
                  const double*, const double*, double*, cudaStream_t);
  int (*project_cotangent)(const mlb_model*, int64_t, int64_t, const double*, double*,
                           cudaStream_t);
  int (*solve)(const mlb_model*, int, int64_t, int64_t, const Opts&, const SolveIO&,
               cudaStream_t);
  int (*steps)(const mlb_model*, int, int64_t, int64_t, const Opts&, const StepsIO&,
               cudaStream_t);
  int (*sample)(const mlb_model*, int, int64_t, int64_t, double*, const Opts&,
                const SampleArgs&, cudaStream_t);
  // dynamic multinomial HMC (mlb_nuts.cuh); NULL where it is not built
  int (*nuts)(const mlb_model*, int, int64_t, int64_t, double*, const Opts&,
              const SampleArgs&, const NutsArgs&, cudaStream_t);
};

inline unsigned grid_for(int64_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

// Runs `configure` once per device of this process (kernel function attributes such as the
// dynamic shared-memory limit belong to a device's context) and remembers a success in one
// bit per device; a failure is retried at the next call.
template <class F>
inline bool once_per_device(std::atomic<uint64_t>& ready, F configure) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
  const uint64_t bit = 1ull << dev;
  if (ready.load(std::memory_order_acquire) & bit) return true;
  if (!configure()) return false;
  ready.fetch_or(bit, std::memory_order_release);
  return true;
}

inline int finish_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_err = std::string("kernel launch: ") + cudaGetErrorString(e);
    return MLB_ERR_CUDA;
  }
  return MLB_OK;
}

template <class M>
struct ParamsOf;  // specialised per model: build the kernel-parameter struct from the handle

template <>
struct ParamsOf<ToyModel> {
  static ToyModel::Params make(const mlb_model* m) {
    return ToyModel::Params{m->data[0], m->data[1], m->data[2]};
  }
};
template <int Y, int PARAM>
struct ParamsOf<HierModel<Y, PARAM>> {
  static typename HierModel<Y, PARAM>::Params make(const mlb_model* m) {
    typename HierModel<Y, PARAM>::Params P;
    for (int i = 0; i < Y; ++i) P.y_obs[i] = m->data[i], P.sigma[i] = m->data[Y + i];
    return P;
  }
};

template <class F>
int dispatch_solver(int solver, F&& f) {
  switch (solver) {
    case MLB_SOLVER_QUASI_NEWTON: return f.template operator()<MLB_SOLVER_QUASI_NEWTON>();
    case MLB_SOLVER_NEWTON: return f.template operator()<MLB_SOLVER_NEWTON>();
    case MLB_SOLVER_NEWTON_LS: return f.template operator()<MLB_SOLVER_NEWTON_LS>();
    case MLB_SOLVER_MICI_NEWTON: return f.template operator()<MLB_SOLVER_MICI_NEWTON>();
    case MLB_SOLVER_MICI_QUASI_NEWTON:
      return f.template operator()<MLB_SOLVER_MICI_QUASI_NEWTON>();
  }
  g_err = "unknown solver";
  return MLB_ERR_ARG;
}

template <class M>
struct Launch {
  static auto P(const mlb_model* m) { return ParamsOf<M>::make(m); }
  // hierarchical Woodbury models whose parked column leaves room for 7 CTAs per SM
  static constexpr bool kHasLean = M::B::FAM == FAM_HIER && !M::B::DENSE && M::Y <= 8;
  static int64_t sm_count() {
    static const int v = [] {
      int dev = 0, n = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
      return n > 0 ? n : 148;
    }();
    return v;
  }
  static int constr(const mlb_model* m, int64_t n, int64_t ld, const double* q, double* c,
                    cudaStream_t s) {
    k_constr<M><<<grid_for(n), kBlock, 0, s>>>(P(m), n, ld, q, c);
    return finish_launch();
  }
  static int jacob(const mlb_model* m, int64_t n, int64_t ld, const double* q, double* jac,
                   double* c, cudaStream_t s) {
    k_jacob<M><<<grid_for(n), kBlock, 0, s>>>(P(m), n, ld, q, jac, c);
    return finish_launch();
  }
  static int decompose(const mlb_model*, int64_t n, int64_t ld, const double* jac,
                       double* gram, cudaStream_t s) {
    k_decompose<M><<<grid_for(n), kBlock, 0, s>>>(n, ld, jac, gram);
    return finish_launch();
  }
  static int logdet(const mlb_model* m, int64_t n, int64_t ld, const double* q, double* val,
                    double* grad, double* c, double* jac, double* gram, double* nld_val,
                    double* nld_grad, double* h1, double* dh1, cudaStream_t s) {
    k_logdet<M><<<grid_for(n), kBlock, 0, s>>>(P(m), n, ld, q, val, grad, c, jac, gram,
                                              nld_val, nld_grad, h1, dh1);
    return finish_launch();
  }
  static int nld(const mlb_model* m, int64_t n, int64_t ld, const double* q, double* val,
                 double* grad, cudaStream_t s) {
    k_nld<M><<<grid_for(n), kBlock, 0, s>>>(P(m), n, ld, q, val, grad);
    return finish_launch();
  }
  static int block_op(const mlb_model*, int op, int64_t n, int64_t ld, const double* jac,
                      const double* jac_r, const double* gram, const double* vec,
                      double* out, cudaStream_t s) {
    const unsigned g = grid_for(n);
    switch (op) {
      case OP_LMULT: k_block_op<M, OP_LMULT><<<g, kBlock, 0, s>>>(n, ld, jac, jac_r, gram, vec, out); break;
      case OP_RMULT: k_block_op<M, OP_RMULT><<<g, kBlock, 0, s>>>(n, ld, jac, jac_r, gram, vec, out); break;
      case OP_PINV: k_block_op<M, OP_PINV><<<g, kBlock, 0, s>>>(n, ld, jac, jac_r, gram, vec, out); break;
      case OP_NSC: k_block_op<M, OP_NSC><<<g, kBlock, 0, s>>>(n, ld, jac, jac_r, gram, vec, out); break;
      default: k_block_op<M, OP_INVJP><<<g, kBlock, 0, s>>>(n, ld, jac, jac_r, gram, vec, out); break;
    }
    return finish_launch();
  }
  static int project_cotangent(const mlb_model* m, int64_t n, int64_t ld, const double* q,
                               double* p, cudaStream_t s) {
    k_project_cotangent<M><<<grid_for(n), kBlock, 0, s>>>(P(m), n, ld, q, p);
    return finish_launch();
  }
  static int solve(const mlb_model* m, int solver, int64_t n, int64_t ld, const Opts& o,
                   const SolveIO& io, cudaStream_t s) {
    return dispatch_solver(solver, [&]<int S>() {
      k_solve<M, S><<<grid_for(n), kBlock, 0, s>>>(P(m), n, ld, io.q, io.jac_prev,
                                                  io.gram_prev, io.dt, o, io.q_out, io.mu_out,
                                                  io.iters, io.norm_dq, io.err, io.n_constr,
                                                  io.status);
      return finish_launch();
    });
  }
  static int steps(const mlb_model* m, int solver, int64_t n, int64_t ld, const Opts& o,
                   const StepsIO& io, cudaStream_t s) {
    return dispatch_solver(solver, [&]<int S>() {
      const int64_t es = io.row_major ? 1 : ld, cs = io.row_major ? (int64_t)M::Q : 1;
      {
        // small launches (at most one warp per scheduler): two warps per 32 chains, the
        // reversibility check one step behind the trajectory (mlb_duo.cuh);
        // MLB_DUO_MAX_CHAINS overrides the threshold (0 switches the form off)
        static const int64_t duo_max = [] {
          const char* e = std::getenv("MLB_DUO_MAX_CHAINS");
          return e ? (int64_t)std::atoll(e) : (int64_t)-1;
        }();
        // (8,192 chains x 32 eight-schools steps: 0.254 -> 0.161 ms; 16,384: 0.256 -> 0.227;
        // 32,768: 0.327 -> 0.429: profiles/r6r_duo_sweep.txt)
        // toy model (3 state components, 100 registers): still ahead at 32,768 chains
        // (0.193 -> 0.163 ms), behind at 65,536 (profiles/r6zz_toy_duo.txt)
        const int64_t lim = duo_max >= 0 ? duo_max : sm_count() * (M::Q <= 4 ? 256 : 112);
        if ((io.batch_total > n ? io.batch_total : n) <= lim) {
          constexpr int smem = duo_smem_bytes<M>();
          static std::atomic<uint64_t> ready{0};
          const bool configured = once_per_device(ready, [] {
            return cudaFuncSetAttribute(k_leapfrog_steps_duo<M, S>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        smem) == cudaSuccess;
          });
          if (configured) {
            k_leapfrog_steps_duo<M, S><<<(unsigned)((n + 31) / 32), 64, smem, s>>>(
                P(m), n, es, cs, io.q, io.p, io.dt, io.dt_pc, io.n_step, o, io.status, io.n_done,
                io.iters_fwd, io.iters_rev, io.h_init, io.h_final);
            return finish_launch();
          }
        }
      }
      if constexpr (kHasLean) {
        // the parked form holds 7 CTAs per SM against 4: taken when the batch is more than
        // one wave of the register-resident kernel but fits one wave of the parked one
        // (65,536 chains on 148 SMs: 0.600 -> 0.568 ms; at other sizes it is slower,
        // profiles/r4k_exp_lean.txt)
        const int64_t sms = sm_count();
        if (n > sms * 4 * kBlock && n <= sms * 7 * kBlock && S <= MLB_SOLVER_NEWTON_LS) {
          constexpr int smem = (int)sizeof(double) * LeanCol<M, kBlock>::ROWS * kBlock;
          // (function attributes are per device: one bit per device of this process)
          static std::atomic<uint64_t> ready{0};
          const bool configured = once_per_device(ready, [] {
            cudaFuncSetAttribute(k_leapfrog_steps_lean<M, S>,
                                 cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            return cudaFuncSetAttribute(k_leapfrog_steps_lean<M, S>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        smem) == cudaSuccess;
          });
          if (configured) {
            k_leapfrog_steps_lean<M, S><<<grid_for(n), kBlock, smem, s>>>(
                P(m), n, es, cs, io.q, io.p, io.dt, io.dt_pc, io.n_step, o, io.status, io.n_done,
                io.iters_fwd, io.iters_rev, io.h_init, io.h_final);
            return finish_launch();
          }
        }
      }
      k_leapfrog_steps<M, S><<<grid_for(n), kBlock, 0, s>>>(
          P(m), n, es, cs, io.q, io.p, io.dt, io.dt_pc, io.n_step, o, io.status, io.n_done,
          io.iters_fwd, io.iters_rev, io.h_init, io.h_final);
      return finish_launch();
    });
  }
  static int sample(const mlb_model* m, int solver, int64_t n, int64_t ld, double* q,
                    const Opts& o, const SampleArgs& a, cudaStream_t s) {
    return dispatch_solver(solver, [&]<int S>() {
      k_hmc_sample<M, S><<<grid_for(n), kBlock, 0, s>>>(P(m), n, ld, q, o, a);
      return finish_launch();
    });
  }
  static int nuts(const mlb_model* m, int solver, int64_t n, int64_t ld, double* q,
                  const Opts& o, const SampleArgs& a, const NutsArgs& na0, cudaStream_t s) {
    // per-chain tree storage (edges, summed momentum, proposal, pending siblings per level)
    NutsArgs na = na0;
    na.lds = (n + 31) / 32 * 32;
    // start batching of the per-lane transitions (mlb_nuts.cuh); MLB_NUTS_BATCH in the
    // environment overrides the default for experiments (1 = start at once, 32 = lock step)
    static const int batch = [] {
      const char* e = std::getenv("MLB_NUTS_BATCH");
      const int b = e ? std::atoi(e) : 0;
      return b >= 1 && b <= 32 ? b : 8;
    }();
    na.batch = batch;
    const size_t bytes = sizeof(double) * (size_t)NutsRows<M::Q>::total(na.max_depth) * (size_t)na.lds;
    if (scratch_alloc((void**)&na.scratch, bytes, s) != cudaSuccess) {
      g_err = "tree scratch allocation failed";
      return (int)MLB_ERR_CUDA;
    }
    const int rc = dispatch_solver(solver, [&]<int S>() {
      constexpr int smem = nuts_smem_bytes<M>();
      static std::atomic<uint64_t> ready{0};
      const bool configured = once_per_device(ready, [] {
        return cudaFuncSetAttribute(k_nuts_sample<M, S>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    smem) == cudaSuccess;
      });
      if (!configured) {
        g_err = "shared-memory configuration of the dynamic sampler failed";
        return (int)MLB_ERR_CUDA;
      }
      k_nuts_sample<M, S><<<grid_for(n), kBlock, smem, s>>>(P(m), n, ld, q, o, a, na);
      return finish_launch();
    });
    cudaFreeAsync(na.scratch, s);
    return rc;
  }
};

template <class M>
ModelOps make_ops(int model, int dim_y, int param) {
  using L = Launch<M>;
  return ModelOps{model, dim_y, param, &L::constr, &L::jacob, &L::decompose, &L::logdet,
                  &L::nld, &L::block_op, &L::project_cotangent, &L::solve, &L::steps,
                  &L::sample, &L::nuts};
}

}  // namespace mlb
