// Dynamic multinomial HMC (NUTS) for the register-resident models, all chains in one launch.
//
// B200-native counterpart of Mici's `DynamicMultinomialHMC`, the sampler the reference
// builds at mlift/example_models/utils.py:286-288 (`max_tree_depth` :70-73); SURVEY.md 8f
// rank 1.  Mici's source is not in the reference tree: the algorithm (binary tree doubling
// in a uniformly drawn direction, multinomial proposal with weights exp(h_init - h),
// uniform sampling inside a new subtree and biased progressive sampling between the old
// tree and the new subtree, no-U-turn termination on the summed momentum for every subtree
// and the whole tree, termination on integrator errors and |h - h_init| > max_delta_h) is
// restated from its published form, recursively in the CPU test checker and ITERATIVELY
// here: a thread cannot recurse over register state, so the
// subtree under construction is kept as one pending left sibling per level (first-edge
// momentum, summed momentum, proposal, log weight) in a per-chain column of global scratch;
// after leaf number i the levels whose bit in i is set are merged bottom-up, which visits
// the merges - and consumes the uniform draws - in the recursion's post-order.  Every
// lane runs ONE flat loop whose body is a single constrained leapfrog step from the current
// tree edge (the fused step of mlb_chmc.cuh) plus bookkeeping, so the lanes of a warp stay
// converged on the expensive part however their tree depths differ.
#pragma once
#include "mlb_chmc.cuh"
// (included by mlb_launch.cuh after SampleArgs / kBlock / load_vec are defined)

namespace mlb {

// k-th uniform of one transition: k = 0 is philox_uniform (the static sampler's
// Metropolis draw); the dynamic sampler numbers its draws 1, 2, ... in order of use
MLB_HD double philox_uniform_k(uint64_t seed, uint64_t chain, uint32_t iter, uint32_t k) {
  uint32_t c[4] = {k, iter, (uint32_t)chain,
                   ((uint32_t)(chain >> 32) & 0x7fffffffu) | 0x80000000u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return u01(c[0], c[1]);
}

MLB_DEV double log_add_exp(double a, double b) {
  const double m = a > b ? a : b;
  return m + log(exp(a - m) + exp(b - m));
}

// rows of the per-chain scratch column
template <int Q>
struct NutsRows {
  static constexpr int EN_Q = 0, EN_P = Q, EP_Q = 2 * Q, EP_P = 3 * Q, RHO = 4 * Q,
                       NEXT = 5 * Q, CUR = 6 * Q, SLOT = 9 * Q;  // CUR / slot: PF, RH, PR
  static __host__ __device__ constexpr int slot(int lvl) { return SLOT + 3 * Q * lvl; }
  static __host__ __device__ constexpr int lw(int max_depth) { return SLOT + 3 * Q * max_depth; }
  static __host__ __device__ constexpr int total(int max_depth) { return lw(max_depth) + max_depth; }
};

struct NutsArgs {
  int max_depth;
  double* scratch;  // [NutsRows::total][lds]
  int64_t lds;
  int32_t* trace_n_step;  // [n_iter][n] or NULL
  int32_t* trace_depth;   // [n_iter][n] or NULL
};

template <class M, int SOLVER>
__global__ void __launch_bounds__(kBlock, MLB_LF_MIN_BLOCKS)
    k_nuts_sample(const __grid_constant__ typename M::Params P, int64_t n, int64_t ld, double* q,
                  Opts o, SampleArgs a, NutsArgs na) {
  constexpr int Q = M::Q;
  using R = NutsRows<Q>;
  __shared__ double commit[2 * Q][kBlock];
  MLB_CHAIN_INDEX();
  double* const sc = na.scratch + i;
  const int64_t lds = na.lds;
#define SC(row) sc[(int64_t)(row) * lds]
  double ad_iter = 0, ad_smooth = 0, ad_err = 0, ad_target = 0, dti = a.dt;
  if (a.adapt) {
    ad_iter = a.adapt[0 * ld + i], ad_smooth = a.adapt[1 * ld + i];
    ad_err = a.adapt[2 * ld + i], ad_target = a.adapt[3 * ld + i];
    dti = a.adapt[4 * ld + i];
  }
  double s_acc = 0, s_moved = 0, s_conv = 0, s_nonrev = 0, s_hdiv = 0, s_steps = 0, s_its = 0,
         s_depth = 0;
  int last = MLB_ST_OK;
#pragma unroll 1
  for (int it = 0; it < a.n_iter; ++it) {
    const uint32_t iter = a.iter0 + (uint32_t)it;
    const uint64_t chain = a.chain0 + (uint64_t)i;
    uint32_t kdraw = 1;
    auto draw = [&]() { return philox_uniform_k(a.seed, chain, iter, kdraw++); };
    double qw[Q], pw[Q];
    load_vec<Q>(qw, q, ld, i);
    if (a.mom_noise) {
      load_vec<Q>(pw, a.mom_noise + (int64_t)it * Q * ld, ld, i);
    } else {
      double* col = &commit[Q][threadIdx.x];
#pragma unroll 1
      for (int j = 0; 2 * j < Q; ++j) {
        double n0, n1;
        philox_normal_pair(a.seed, chain, iter, (uint32_t)j, n0, n1);
        col[(2 * j) * kBlock] = n0;
        if (2 * j + 1 < Q) col[(2 * j + 1) * kBlock] = n1;
      }
#pragma unroll
      for (int k = 0; k < Q; ++k) pw[k] = col[k * kBlock];
    }
    // momentum refresh: project onto the cotangent space, h_init (no step); C keeps the
    // evaluation at the point the next step starts from (see run_trajectory_ev)
    Eval<M> C;
    const TrajOut t0 = run_trajectory_ev<M, SOLVER, true, false, kBlock>(
        P, qw, pw, &commit[0][threadIdx.x], &commit[Q][threadIdx.x], 0, dti, o, 0.0, C, false);
    bool have_eval = t0.status == MLB_ST_OK;  // both edges of the one-state tree are here
    int last_dir = 0;
    int st = t0.status;
    const double h0 = t0.h_init;
#pragma unroll 1
    for (int k = 0; k < Q; ++k) {
      SC(R::EN_Q + k) = qw[k], SC(R::EP_Q + k) = qw[k], SC(R::NEXT + k) = qw[k];
      SC(R::EN_P + k) = pw[k], SC(R::EP_P + k) = pw[k], SC(R::RHO + k) = pw[k];
    }
    double lw_tree = 0.0, sum_acc = 0.0;
    int depth = 0, leaf = 0, dir = 1, n_leaf = 0, its = 0;
    bool moved = false;
    bool active = st == MLB_ST_OK && na.max_depth > 0;
#pragma unroll 1
    while (active) {
      if (leaf == 0) {  // new doubling: direction, start from that edge of the tree
        dir = draw() < 0.5 ? 1 : -1;
        if (last_dir != 0 && dir != last_dir) have_eval = false;  // C sits at the other edge
        const int eq = dir > 0 ? R::EP_Q : R::EN_Q, ep = dir > 0 ? R::EP_P : R::EN_P;
#pragma unroll 1
        for (int k = 0; k < Q; ++k) qw[k] = SC(eq + k), pw[k] = SC(ep + k);
      }
      const TrajOut t = run_trajectory_ev<M, SOLVER, false, false, kBlock>(
          P, qw, pw, &commit[0][threadIdx.x], &commit[Q][threadIdx.x], 1, dir * dti, o, 0.0, C,
          have_eval);
      have_eval = t.status == MLB_ST_OK, last_dir = dir;
      n_leaf += 1;
      its += t.it_fwd + t.it_rev;
      bool terminate = false;
      const double h = t.h_last;
      if (t.status != MLB_ST_OK) {
        st = t.status, terminate = true;
      } else if (h != h || fabs(h - h0) > a.max_delta_h) {
        st = MLB_ST_H_DIVERGED, terminate = true;
      }
      if (!terminate) {
        double lw_cur = h0 - h;
        sum_acc += exp(lw_cur < 0.0 ? lw_cur : 0.0);
        // the leaf as a one-state subtree, in registers: it is born after the step and
        // parked before the next one, so it never competes with the step for registers
        // (the first version kept it in global scratch: 160 GB of DRAM writes per launch
        // and 5.3 long-scoreboard stall cycles per issue)
        double c_pf[Q], c_rh[Q], c_pr[Q];
#pragma unroll
        for (int k = 0; k < Q; ++k) c_pf[k] = pw[k], c_rh[k] = pw[k], c_pr[k] = qw[k];
        int lvl = 0;
#pragma unroll 1
        while ((leaf >> lvl) & 1) {  // a left sibling waits at this level: merge
          const int sl = R::slot(lvl);
          const double lw_left = SC(R::lw(na.max_depth) + lvl);
          const double lw_new = log_add_exp(lw_left, lw_cur);
          const bool take_outer = draw() < exp(lw_cur - lw_new);
          double d_first = 0.0, d_last = 0.0;
#pragma unroll
          for (int k = 0; k < Q; ++k) {
            c_rh[k] += SC(sl + Q + k), c_pf[k] = SC(sl + k);
            if (!take_outer) c_pr[k] = SC(sl + 2 * Q + k);
            d_first = fma(c_pf[k], c_rh[k], d_first), d_last = fma(pw[k], c_rh[k], d_last);
          }
          lw_cur = lw_new;
          if (d_first < 0.0 || d_last < 0.0) {
            terminate = true;
            break;
          }
          ++lvl;
        }
        if (!terminate) {  // park as the pending sibling of level lvl (= depth: complete)
          const int sl = R::slot(lvl);
#pragma unroll
          for (int k = 0; k < Q; ++k)
            SC(sl + k) = c_pf[k], SC(sl + Q + k) = c_rh[k], SC(sl + 2 * Q + k) = c_pr[k];
          SC(R::lw(na.max_depth) + lvl) = lw_cur;
        }
      }
      ++leaf;
      if (terminate) {
        active = false;
      } else if (leaf == (1 << depth)) {  // new subtree complete: join it to the tree
        const int sl = R::slot(depth);
        const double lw_sub = SC(R::lw(na.max_depth) + depth);
        const bool take = draw() < exp(lw_sub - lw_tree);
        lw_tree = log_add_exp(lw_tree, lw_sub);
        const int eq = dir > 0 ? R::EP_Q : R::EN_Q, ep = dir > 0 ? R::EP_P : R::EN_P;
        moved |= take;
#pragma unroll 1
        for (int k = 0; k < Q; ++k) {
          if (take) SC(R::NEXT + k) = SC(sl + 2 * Q + k);
          SC(R::RHO + k) += SC(sl + Q + k);
          SC(eq + k) = qw[k], SC(ep + k) = pw[k];
        }
        double d_neg = 0.0, d_pos = 0.0;
#pragma unroll 1
        for (int k = 0; k < Q; ++k) {
          const double rho = SC(R::RHO + k);
          d_neg = fma(SC(R::EN_P + k), rho, d_neg), d_pos = fma(SC(R::EP_P + k), rho, d_pos);
        }
        ++depth, leaf = 0;
        if (d_neg < 0.0 || d_pos < 0.0 || depth >= na.max_depth) active = false;
      }
    }
    const double a_stat = n_leaf > 0 ? sum_acc / n_leaf : 0.0;
#pragma unroll 1
    for (int k = 0; k < Q; ++k) q[(int64_t)k * ld + i] = SC(R::NEXT + k);
    s_acc += a_stat, s_moved += moved, s_steps += n_leaf, s_its += its, s_depth += depth;
    s_conv += (st == MLB_ST_DIVERGED || st == MLB_ST_NOT_CONVERGED || st == MLB_ST_LINALG);
    s_nonrev += (st == MLB_ST_NON_REVERSIBLE);
    s_hdiv += (st == MLB_ST_H_DIVERGED);
    last = st;
    if (a.trace_accept) a.trace_accept[(int64_t)it * ld + i] = a_stat;
    if (na.trace_n_step) na.trace_n_step[(int64_t)it * ld + i] = n_leaf;
    if (na.trace_depth) na.trace_depth[(int64_t)it * ld + i] = depth;
    if (a.trace_q && (it + 1) % a.trace_every == 0) {
      double* tq = a.trace_q + (int64_t)((it + 1) / a.trace_every - 1) * a.trace_dim * ld + i;
#pragma unroll 1
      for (int k = 0; k < a.trace_dim; ++k) tq[(int64_t)k * ld] = SC(R::NEXT + k);
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
#undef SC
  if (a.adapt) {
    a.adapt[0 * ld + i] = ad_iter, a.adapt[1 * ld + i] = ad_smooth;
    a.adapt[2 * ld + i] = ad_err, a.adapt[4 * ld + i] = dti;
  }
  if (a.stats) {
    a.stats[0 * ld + i] += s_acc, a.stats[1 * ld + i] += s_moved;
    a.stats[2 * ld + i] += s_conv, a.stats[3 * ld + i] += s_nonrev;
    a.stats[4 * ld + i] += s_hdiv, a.stats[5 * ld + i] += s_steps;
    a.stats[6 * ld + i] += s_its, a.stats[7 * ld + i] += s_depth;
  }
  if (a.last_status) a.last_status[i] = last;
}

}  // namespace mlb
