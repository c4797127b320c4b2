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

// log(exp(a) + exp(b)) and the share of b in it, exp(b) / (exp(a) + exp(b)): the weight of a
// merged tree and the probability that its proposal comes from the part with log weight b
MLB_DEV double log_add_exp_share(double a, double b, double& share_b) {
  const double m = a > b ? a : b;
  const double ea = exp(a - m), eb = exp(b - m), s = ea + eb;
  share_b = eb / s;
  return m + log(s);
}

// rows of the per-chain scratch column: tree edges (negative / positive time direction),
// summed momentum, proposal, then per level the pending left sibling - first-built and
// last-built momentum, summed momentum, proposal
template <int Q>
struct NutsRows {
  static constexpr int EN_Q = 0, EN_P = Q, EP_Q = 2 * Q, EP_P = 3 * Q, RHO = 4 * Q,
                       NEXT = 5 * Q, SLOT = 6 * Q;
  static constexpr int PF = 0, PL = Q, RH = 2 * Q, PR = 3 * Q, SLOT_ROWS = 4 * Q;
  static __host__ __device__ constexpr int slot(int lvl) { return SLOT + SLOT_ROWS * lvl; }
  static __host__ __device__ constexpr int lw(int max_depth) { return SLOT + SLOT_ROWS * max_depth; }
  static __host__ __device__ constexpr int total(int max_depth) { return lw(max_depth) + max_depth; }
};

struct NutsArgs {
  int max_depth;
  double* scratch;  // [NutsRows::total][lds]
  int64_t lds;
  int32_t* trace_n_step;  // [n_iter][n] or NULL
  int32_t* trace_depth;   // [n_iter][n] or NULL
  int batch;  // lanes of a warp that start their next transition together (see the kernel)
};

// No-U-turn tests when the subtree built first (inner: first-built momentum pf_i, last-built
// pl_i, summed rh_i) is joined by the one built after it (outer: pf_o, pl_o, rh_o), both in
// build order, so the same expressions serve either time direction: the criterion on the
// merged tree (its two edges against the total), and - Mici's `do_extra_subtree_checks`,
// default on - on the two overlapping sub-trajectories "inner + first state of outer" and
// "last state of inner + outer".
struct UTurn {
  double d1 = 0.0, d2 = 0.0, e1a = 0.0, e1b = 0.0, e2a = 0.0, e2b = 0.0;
  MLB_DEV void add(double pf_i, double pl_i, double rh_i, double pf_o, double pl_o,
                   double rh_o, double rho) {
    d1 = fma(pf_i, rho, d1), d2 = fma(pl_o, rho, d2);
    const double r1 = rh_i + pf_o, r2 = rh_o + pl_i;
    e1a = fma(pf_i, r1, e1a), e1b = fma(pf_o, r1, e1b);
    e2a = fma(pl_i, r2, e2a), e2b = fma(pl_o, r2, e2b);
  }
  MLB_DEV bool turned(bool extra) const {
    return d1 < 0.0 || d2 < 0.0 ||
           (extra && (e1a < 0.0 || e1b < 0.0 || e2a < 0.0 || e2b < 0.0));
  }
};

// Shared memory (dynamic, 5 Q doubles per thread, element k of a column at [k * kBlock]):
// the last committed (q, p) of the step code and the subtree under construction.  All tree
// bookkeeping runs as ROLLED loops over these columns and the global scratch column - no
// register arrays, a few hundred instructions of code.
template <class M>
constexpr int nuts_smem_bytes() {
  return (int)sizeof(double) * 5 * M::Q * kBlock;
}

template <class M, int SOLVER>
__global__ void __launch_bounds__(kBlock, MLB_LF_MIN_BLOCKS)
    k_nuts_sample(const __grid_constant__ typename M::Params P, int64_t n, int64_t ld, double* q,
                  Opts o, SampleArgs a, NutsArgs na) {
  constexpr int Q = M::Q;
  using R = NutsRows<Q>;
  extern __shared__ double nuts_smem[];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;  // no early return: the warp votes below need every lane
  double* const cq = nuts_smem + threadIdx.x;  // committed position
  double* const cp = cq + Q * kBlock;          // committed momentum
  double* const cur = cq + 2 * Q * kBlock;     // subtree under construction: PF, RH, PR
  double* const sc = na.scratch + i;
  const int64_t lds = na.lds;
  const bool extra = !(o.flags & MLB_FLAG_NUTS_NO_EXTRA_SUBTREE_CHECKS);
  const bool two_sided = (o.flags & MLB_FLAG_NUTS_TWO_SIDED_DELTA_H) != 0;
#define SC(row) sc[(int64_t)(row) * lds]
  double dti = (valid && a.adapt) ? a.adapt[4 * ld + i] : a.dt;
  int last = MLB_ST_OK;
  const uint64_t chain = a.chain0 + (uint64_t)i;
  // Every lane runs its OWN sequence of transitions: a lane whose tree is complete does not
  // wait for the largest tree of its warp (72 % of the eight-schools trees have 7 leaves,
  // 21 % have 15: in lock step a warp ran ~16 passes per transition for a mean of 8.75
  // leaves).  Starting a transition (momentum draw, evaluation at the start point) runs
  // with only the starting lanes active, so starts are batched: a lane waits until
  // `batch` lanes of its warp are ready, or until no lane is stepping.  The refresh
  // itself shares the one call site of the step code with the other lanes' leaves.
  int it = 0;
  bool in_tr = false, finished = !valid || a.n_iter <= 0;
  // ONE call site of the step code serves the momentum refresh (leading pass: projection
  // onto the cotangent space and h_init, no step) and every tree leaf: with two inlined
  // copies the kernel was 135 KB of SASS and ncu showed 4.3 stall cycles per issue waiting
  // for instructions (profiles/r6a_full_k_nuts_sample_hier.md).  C keeps the evaluation
  // at the point the next step starts from (see run_trajectory_ev).
  Eval<M> C;
  bool refresh = false, have_eval = false, moved = false;
  int last_dir = 0, st = MLB_ST_OK;
  double h0 = 0.0, lw_tree = 0.0, sum_acc = 0.0, lw0 = 0.0;
  int depth = 0, leaf = 0, dir = 1, n_leaf = 0, its = 0;
  uint32_t kdraw = 1;
#pragma unroll 1
  for (;;) {
    const unsigned m_step = __ballot_sync(0xffffffffu, in_tr);
    const unsigned m_wait = __ballot_sync(0xffffffffu, !in_tr && !finished);
    if ((m_step | m_wait) == 0u) break;
    const uint32_t iter = a.iter0 + (uint32_t)it;
    auto draw = [&]() { return philox_uniform_k(a.seed, chain, iter, kdraw++); };
    if (!in_tr && !finished && (m_step == 0u || __popc(m_wait) >= na.batch)) {
      // start of a transition in the committed columns: position, momentum noise
      if (a.mom_noise) {
        const double* mn = a.mom_noise + (int64_t)it * Q * ld + i;
#pragma unroll 1
        for (int k = 0; k < Q; ++k) cp[k * kBlock] = mn[(int64_t)k * ld];
      } else {
#pragma unroll 1
        for (int j = 0; 2 * j < Q; ++j) {
          double n0, n1;
          philox_normal_pair(a.seed, chain, iter, (uint32_t)j, n0, n1);
          cp[(2 * j) * kBlock] = n0;
          if (2 * j + 1 < Q) cp[(2 * j + 1) * kBlock] = n1;
        }
      }
      if (it == 0) {  // (later transitions: the previous one left its result in cq)
#pragma unroll 6
        for (int k = 0; k < Q; ++k) cq[k * kBlock] = q[(int64_t)k * ld + i];
      }
      in_tr = true, refresh = true, have_eval = false, moved = false;
      last_dir = 0, st = MLB_ST_OK, kdraw = 1;
      lw_tree = 0.0, sum_acc = 0.0;
      depth = 0, leaf = 0, dir = 1, n_leaf = 0, its = 0;
    }
    if (in_tr) {
      bool active = true;
      double qw[Q], pw[Q];
      if (!refresh && leaf == 0) {  // new doubling: direction, start from that edge of the tree
        dir = draw() < 0.5 ? 1 : -1;
        if (last_dir != 0 && dir != last_dir) have_eval = false;  // C sits at the other edge
        const double* eq = &SC(dir > 0 ? R::EP_Q : R::EN_Q);
        const double* ep = &SC(dir > 0 ? R::EP_P : R::EN_P);
#pragma unroll
        for (int k = 0; k < Q; ++k) qw[k] = eq[(int64_t)k * lds], pw[k] = ep[(int64_t)k * lds];
      } else {  // continue from the last committed state
#pragma unroll
        for (int k = 0; k < Q; ++k) qw[k] = cq[k * kBlock], pw[k] = cp[k * kBlock];
      }
      const TrajOut t = run_trajectory_ev<M, SOLVER, true, false, kBlock>(
          P, qw, pw, cq, cp, refresh ? 0 : 1, refresh ? dti : dir * dti, o, 0.0, C,
          !refresh && have_eval, true, refresh);
      have_eval = t.status == MLB_ST_OK;
      if (refresh) {  // the tree is the single state (q, projected p)
        refresh = false;
        st = t.status, h0 = t.h_init;
#pragma unroll 1
        for (int k = 0; k < Q; ++k) {
          const double qk = cq[k * kBlock], pk = cp[k * kBlock];
          SC(R::EN_Q + k) = qk, SC(R::EP_Q + k) = qk, SC(R::NEXT + k) = qk;
          SC(R::EN_P + k) = pk, SC(R::EP_P + k) = pk, SC(R::RHO + k) = pk;
        }
        active = st == MLB_ST_OK && na.max_depth > 0;
      } else {
        last_dir = dir;
        n_leaf += 1;
        its += t.it_fwd + t.it_rev;
        bool terminate = false;
        const double h = t.h_last;
        double lw_cur = h0 - h;  // log weight of the subtree under construction
        if (t.status != MLB_ST_OK) {
          st = t.status, terminate = true;
        } else if (h != h || h - h0 > a.max_delta_h || (two_sided && h0 - h > a.max_delta_h)) {
          st = MLB_ST_H_DIVERGED, terminate = true;
        }
        bool joined = false;
        if (!terminate) {
          sum_acc += exp(lw_cur < 0.0 ? lw_cur : 0.0);
          // The leaf is the committed state (cq, cp).  Level 0 never leaves shared memory: an
          // even leaf waits in `cur` as the pending sibling of the odd leaf that follows it
          // (one state: first = last = summed momentum, proposal = its position), and their
          // merge reads both from shared memory.  Only the pending siblings of levels >= 1
          // live in the global scratch column, and a subtree that is complete is joined to
          // the tree straight from `cur`.
          if (!(leaf & 1)) {
            lw0 = lw_cur;
#pragma unroll 2
            for (int k = 0; k < Q; ++k) {
              const double pk = cp[k * kBlock];
              cur[k * kBlock] = pk, cur[(Q + k) * kBlock] = pk;
              cur[(2 * Q + k) * kBlock] = cq[k * kBlock];
            }
          }
          // "Combine" events of this leaf, bottom-up: merges with the pending left siblings of
          // the levels whose bit in `leaf` is set, then - if that completes the subtree of
          // this doubling - its join with the tree.  ONE code site evaluates the weights and
          // draws the uniform for all of them (instruction-cache footprint).
          int lvl = 0;
#pragma unroll 1
          for (;;) {
            const bool merge = (leaf >> lvl) & 1;
            if (!merge && lvl != depth) break;  // park below (levels >= 1)
            const double lw_in = !merge ? lw_tree : lvl == 0 ? lw0 : SC(R::lw(na.max_depth) + lvl);
            double share;  // of the part built last (the subtree under construction) in the sum
            const double lw_new = log_add_exp_share(lw_in, lw_cur, share);
            const double u = draw();
            UTurn ut;
            if (!merge) {
              // join: inner = the tree so far (first-built = its far edge, last-built = the edge
              // the subtree grew from), outer = the subtree, whose last state is the committed
              // one; biased progressive sampling takes the subtree's proposal with probability
              // min(1, w_sub / w_tree) = share / (1 - share)
              const bool take = u * (1.0 - share) < share;
              moved |= take;
              double* eq = &SC(dir > 0 ? R::EP_Q : R::EN_Q);
              double* ep = &SC(dir > 0 ? R::EP_P : R::EN_P);
              const double* fp = &SC(dir > 0 ? R::EN_P : R::EP_P);
#pragma unroll 6
              for (int k = 0; k < Q; ++k) {
                const double rh_i = SC(R::RHO + k), rh_o = cur[(Q + k) * kBlock];
                const double pl_o = cp[k * kBlock], rho = rh_i + rh_o;
                ut.add(fp[(int64_t)k * lds], ep[(int64_t)k * lds], rh_i, cur[k * kBlock], pl_o, rh_o,
                       rho);
                if (take) SC(R::NEXT + k) = cur[(2 * Q + k) * kBlock];
                SC(R::RHO + k) = rho;
                eq[(int64_t)k * lds] = cq[k * kBlock], ep[(int64_t)k * lds] = pl_o;
              }
              lw_tree = lw_new;
              joined = true;
              if (ut.turned(extra)) active = false;
              break;
            }
            const bool take_outer = u < share;
            if (lvl == 0) {  // the previous leaf, still in `cur`
#pragma unroll 2
              for (int k = 0; k < Q; ++k) {
                const double p_i = cur[k * kBlock], p_o = cp[k * kBlock], rho = p_i + p_o;
                ut.add(p_i, p_i, p_i, p_o, p_o, p_o, rho);
                cur[(Q + k) * kBlock] = rho;
                if (take_outer) cur[(2 * Q + k) * kBlock] = cq[k * kBlock];
              }
            } else {
              // (unrolled by 6: the loads of six components are in flight together - rolled,
              // every component paid its own L2 / DRAM round trip, 40 % of the memory stalls)
              const double* sl = &SC(R::slot(lvl));
#pragma unroll 6
              for (int k = 0; k < Q; ++k) {
                const double pf_i = sl[(int64_t)(R::PF + k) * lds];
                const double pl_i = sl[(int64_t)(R::PL + k) * lds];
                const double rh_i = sl[(int64_t)(R::RH + k) * lds];
                const double pr_i = sl[(int64_t)(R::PR + k) * lds];
                const double pf_o = cur[k * kBlock], rh_o = cur[(Q + k) * kBlock];
                const double rho = rh_i + rh_o;
                ut.add(pf_i, pl_i, rh_i, pf_o, cp[k * kBlock], rh_o, rho);
                cur[k * kBlock] = pf_i, cur[(Q + k) * kBlock] = rho;
                if (!take_outer) cur[(2 * Q + k) * kBlock] = pr_i;
              }
            }
            lw_cur = lw_new;
            if (ut.turned(extra)) {
              terminate = true;
              break;
            }
            ++lvl;
          }
          if (!terminate && !joined && lvl >= 1) {  // park as the pending sibling of level lvl
            double* sl = &SC(R::slot(lvl));
#pragma unroll 2
            for (int k = 0; k < Q; ++k) {
              sl[(int64_t)(R::PF + k) * lds] = cur[k * kBlock];
              sl[(int64_t)(R::PL + k) * lds] = cp[k * kBlock];
              sl[(int64_t)(R::RH + k) * lds] = cur[(Q + k) * kBlock];
              sl[(int64_t)(R::PR + k) * lds] = cur[(2 * Q + k) * kBlock];
            }
            SC(R::lw(na.max_depth) + lvl) = lw_cur;
          }
        }
        ++leaf;
        if (terminate) {
          active = false;
        } else if (joined) {
          ++depth, leaf = 0;
          if (depth >= na.max_depth) active = false;
        }
      }
      if (!active) {  // end of this lane's transition: new state, statistics, adaptation
        const double a_stat = n_leaf > 0 ? sum_acc / n_leaf : 0.0;
        const bool tr = a.trace_q && (it + 1) % a.trace_every == 0;
        double* tq =
            tr ? a.trace_q + (int64_t)((it + 1) / a.trace_every - 1) * a.trace_dim * ld + i
               : nullptr;
#pragma unroll 6
        for (int k = 0; k < Q; ++k) {
          const double v = SC(R::NEXT + k);
          q[(int64_t)k * ld + i] = v, cq[k * kBlock] = v;
          if (tr && k < a.trace_dim) tq[(int64_t)k * ld] = v;
        }
        last = st;
        if (a.stats) {
          // accumulated in global memory (eight fewer live doubles in the loop) by
          // reductions without a return value: nothing waits for them
          atomicAdd(&a.stats[0 * ld + i], a_stat), atomicAdd(&a.stats[1 * ld + i], (double)moved);
          const bool conv =
              st == MLB_ST_DIVERGED || st == MLB_ST_NOT_CONVERGED || st == MLB_ST_LINALG;
          if (conv) atomicAdd(&a.stats[2 * ld + i], 1.0);
          if (st == MLB_ST_NON_REVERSIBLE) atomicAdd(&a.stats[3 * ld + i], 1.0);
          if (st == MLB_ST_H_DIVERGED) atomicAdd(&a.stats[4 * ld + i], 1.0);
          atomicAdd(&a.stats[5 * ld + i], (double)n_leaf);
          atomicAdd(&a.stats[6 * ld + i], (double)its);
          atomicAdd(&a.stats[7 * ld + i], (double)depth);
        }
        if (a.trace_accept) a.trace_accept[(int64_t)it * ld + i] = a_stat;
        if (na.trace_n_step) na.trace_n_step[(int64_t)it * ld + i] = n_leaf;
        if (na.trace_depth) na.trace_depth[(int64_t)it * ld + i] = depth;
        if (a.adapt) {  // Mici DualAveragingStepSizeAdapter.update
          double ad_iter = a.adapt[0 * ld + i], ad_smooth = a.adapt[1 * ld + i];
          double ad_err = a.adapt[2 * ld + i];
          const double ad_target = a.adapt[3 * ld + i];
          ad_iter += 1;
          const double w = 1.0 / (a.a_offset + ad_iter);
          ad_err = (1.0 - w) * ad_err + w * (a.a_target - a_stat);
          const double sw = exp(-a.a_decay * log(ad_iter));  // (1 / iter)^decay
          const double log_eps = ad_target - ad_err * sqrt(ad_iter) / a.a_reg;
          ad_smooth = (1.0 - sw) * ad_smooth + sw * log_eps;
          dti = exp(log_eps);
          a.adapt[0 * ld + i] = ad_iter, a.adapt[1 * ld + i] = ad_smooth;
          a.adapt[2 * ld + i] = ad_err, a.adapt[4 * ld + i] = dti;
        }
        ++it;
        in_tr = false, finished = it >= a.n_iter;
      }
    }
  }
#undef SC
  if (valid && a.last_status) a.last_status[i] = last;
}

}  // namespace mlb
