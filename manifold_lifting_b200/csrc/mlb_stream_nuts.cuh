// Dynamic multinomial HMC (NUTS) for the ODE-constrained models, driven from the host.
//
// Same algorithm and same Philox draw order as the in-kernel sampler of the
// register-resident models (mlb_nuts.cuh) and as the recursive restatement in the CPU test
// checker, but a "leaf" here is one constrained leapfrog step of ALL chains whose tree is
// still growing, run through the sub-phase kernels (mlb_stream_phased.cuh, one step from
// each chain's current tree edge, per-chain signed step size, alive mask); the tree
// bookkeeping of a round is one small kernel per chain batch and the host reads one
// integer per round (chains still growing).  Per-chain tree storage is a column of global
// scratch: edges, summed momentum, proposal, working state and one pending sibling per
// level (first-built and last-built momentum, summed momentum, proposal) - (8 + 4 D) Q
// doubles.
#pragma once
#include "mlb_stream_phased.cuh"  // (brings in mlb_launch.cuh and, through it, mlb_nuts.cuh)

namespace mlb {

struct SNutsLayout {
  double* base;
  int64_t lds;
  int Q, D;
  // Q-sized blocks
  enum { EN_Q, EN_P, EP_Q, EP_P, RHO, NEXT, WQ, WP, N_FIXED };
  __host__ __device__ double* blk(int b) const { return base + (int64_t)b * Q * lds; }
  __host__ __device__ double* slot(int lvl, int part) const {
    return base + ((int64_t)(N_FIXED + 4 * lvl + part) * Q) * lds;  // part: PF, PL, RH, PR
  }
  // per-chain doubles: h0, lw_tree, sum_acc, dt (signed, this round), dt0, h_init, h_final, lw[D]
  enum { H0, LW_TREE, SUM_ACC, DT, DT0, H_INIT, H_FINAL, N_DBL };
  __host__ __device__ double* dbl(int r) const {
    return base + ((int64_t)(N_FIXED + 4 * D) * Q + r) * lds;
  }
  __host__ __device__ double* lw(int lvl) const { return dbl(N_DBL + lvl); }
  // per-chain int32: depth, leaf, dir, n_leaf, its, st, active, moved, kdraw, + step outputs
  enum { DEPTH, LEAF, DIR, N_LEAF, ITS, ST, ACTIVE, MOVED, KDRAW, S_STATUS, S_DONE, S_ITF,
         S_ITR, N_INT };
  __host__ __device__ int32_t* i32(int r) const {
    return reinterpret_cast<int32_t*>(dbl(N_DBL + D)) + (int64_t)r * lds;
  }
  int64_t rows() const { return (int64_t)(N_FIXED + 4 * D) * Q + N_DBL + D + (N_INT + 1) / 2 + 1; }
};

// new transition: working state = (q, fresh momentum noise), tree state reset
template <int NCH>
__global__ void __launch_bounds__(32 * NCH)
    kn_begin(SNutsLayout T, int64_t n, int64_t ld, const double* q, SampleArgs a, int it) {
  const int64_t i = (int64_t)blockIdx.x * 32 + threadIdx.x;
  if (i >= n) return;
  int k0, k1;
  row_chunk<NCH>(T.Q, k0, k1);
  const double* mn = a.mom_noise ? a.mom_noise + (int64_t)it * T.Q * ld + i : nullptr;
#pragma unroll 1
  for (int k = k0; k < k1; ++k) {
    T.blk(T.WQ)[(int64_t)k * T.lds + i] = q[(int64_t)k * ld + i];
    double v;
    if (mn) {
      v = mn[(int64_t)k * ld];
    } else {
      double n0, n1;
      philox_normal_pair(a.seed, a.chain0 + (uint64_t)i, a.iter0 + (uint32_t)it,
                         (uint32_t)(k >> 1), n0, n1);
      v = (k & 1) ? n1 : n0;
    }
    T.blk(T.WP)[(int64_t)k * T.lds + i] = v;
  }
  if (threadIdx.y != 0) return;
  const double dt = a.adapt ? a.adapt[4 * ld + i] : a.dt;
  T.dbl(T.DT0)[i] = dt, T.dbl(T.DT)[i] = dt;
  T.i32(T.ACTIVE)[i] = 1;
}

// after the refresh projection (a zero-step call of the sub-phase path): the tree is the
// single state (q, projected p)
template <int NCH>
__global__ void __launch_bounds__(32 * NCH)
    kn_init_tree(SNutsLayout T, int64_t n, int max_depth, int* counter) {
  const int64_t i = (int64_t)blockIdx.x * 32 + threadIdx.x;
  if (i >= n) return;
  int k0, k1;
  row_chunk<NCH>(T.Q, k0, k1);
#pragma unroll 4
  for (int k = k0; k < k1; ++k) {
    const int64_t at = (int64_t)k * T.lds + i;
    const double qk = T.blk(T.WQ)[at], pk = T.blk(T.WP)[at];
    T.blk(T.EN_Q)[at] = qk, T.blk(T.EP_Q)[at] = qk, T.blk(T.NEXT)[at] = qk;
    T.blk(T.EN_P)[at] = pk, T.blk(T.EP_P)[at] = pk, T.blk(T.RHO)[at] = pk;
  }
  if (threadIdx.y != 0) return;
  const int st = T.i32(T.S_STATUS)[i];
  T.dbl(T.H0)[i] = T.dbl(T.H_INIT)[i], T.dbl(T.LW_TREE)[i] = 0.0, T.dbl(T.SUM_ACC)[i] = 0.0;
  T.i32(T.DEPTH)[i] = 0, T.i32(T.LEAF)[i] = 0, T.i32(T.DIR)[i] = 1, T.i32(T.N_LEAF)[i] = 0;
  T.i32(T.ITS)[i] = 0, T.i32(T.ST)[i] = st, T.i32(T.MOVED)[i] = 0, T.i32(T.KDRAW)[i] = 1;
  const int active = st == MLB_ST_OK && max_depth > 0;
  T.i32(T.ACTIVE)[i] = active;
  if (active) atomicAdd(counter, 1);
}

// before a round: chains starting a doubling draw its direction and load that edge
static __global__ void __launch_bounds__(32)
    kn_pre(SNutsLayout T, int64_t n, SampleArgs a, int it) {
  const int64_t i = (int64_t)blockIdx.x * 32 + threadIdx.x;
  if (i >= n || !T.i32(T.ACTIVE)[i]) return;
  if (T.i32(T.LEAF)[i] != 0) return;
  const uint32_t k = (uint32_t)T.i32(T.KDRAW)[i];
  const double u = philox_uniform_k(a.seed, a.chain0 + (uint64_t)i, a.iter0 + (uint32_t)it, k);
  T.i32(T.KDRAW)[i] = (int)k + 1;
  const int dir = u < 0.5 ? 1 : -1;
  T.i32(T.DIR)[i] = dir;
  T.dbl(T.DT)[i] = dir * T.dbl(T.DT0)[i];
  const double* eq = T.blk(dir > 0 ? T.EP_Q : T.EN_Q) + i;
  const double* ep = T.blk(dir > 0 ? T.EP_P : T.EN_P) + i;
  double* wq = T.blk(T.WQ) + i;
  double* wp = T.blk(T.WP) + i;
#pragma unroll 4
  for (int r = 0; r < T.Q; ++r) {
    const int64_t at = (int64_t)r * T.lds;
    wq[at] = eq[at], wp[at] = ep[at];
  }
}

// after a round's step: the leaf joins the subtree under construction (merges bottom-up,
// uniform draws in the recursion's post-order), a completed subtree joins the tree
// (mlb_nuts.cuh has the same bookkeeping with the chain state in registers)
static __global__ void __launch_bounds__(32)
    kn_post(SNutsLayout T, int64_t n, SampleArgs a, int it, int max_depth, int flags,
            int* counter) {
  const int64_t i = (int64_t)blockIdx.x * 32 + threadIdx.x;
  if (i >= n || !T.i32(T.ACTIVE)[i]) return;
  const int Q = T.Q;
  const int64_t lds = T.lds;
  uint32_t kdraw = (uint32_t)T.i32(T.KDRAW)[i];
  auto draw = [&]() {
    return philox_uniform_k(a.seed, a.chain0 + (uint64_t)i, a.iter0 + (uint32_t)it, kdraw++);
  };
  const double* wq = T.blk(T.WQ) + i;
  const double* wp = T.blk(T.WP) + i;
  int depth = T.i32(T.DEPTH)[i], leaf = T.i32(T.LEAF)[i];
  const int dir = T.i32(T.DIR)[i];
  const double h0 = T.dbl(T.H0)[i], h = T.dbl(T.H_FINAL)[i];
  const int sst = T.i32(T.S_STATUS)[i];
  T.i32(T.N_LEAF)[i] += 1;
  T.i32(T.ITS)[i] += T.i32(T.S_ITF)[i] + T.i32(T.S_ITR)[i];
  bool terminate = false;
  if (sst != MLB_ST_OK) {
    T.i32(T.ST)[i] = sst, terminate = true;
  } else if (h != h || h - h0 > a.max_delta_h ||
             ((flags & MLB_FLAG_NUTS_TWO_SIDED_DELTA_H) && h0 - h > a.max_delta_h)) {
    T.i32(T.ST)[i] = MLB_ST_H_DIVERGED, terminate = true;
  }
  const bool extra = !(flags & MLB_FLAG_NUTS_NO_EXTRA_SUBTREE_CHECKS);
  double lw_cur = h0 - h;
  if (!terminate) {
    T.dbl(T.SUM_ACC)[i] += exp(lw_cur < 0.0 ? lw_cur : 0.0);
    // the subtree under construction is assembled in slot `lvl` (the level it will be
    // parked at): start from the leaf, fold in the pending siblings below
    int top = 0;
    while ((leaf >> top) & 1) ++top;  // park level = number of trailing ones of leaf
    double* pf = T.slot(top, 0) + i;
    double* pl = T.slot(top, 1) + i;
    double* rh = T.slot(top, 2) + i;
    double* pr = T.slot(top, 3) + i;
#pragma unroll 4
    for (int r = 0; r < Q; ++r) {
      const int64_t at = (int64_t)r * lds;
      pf[at] = wp[at], pl[at] = wp[at], rh[at] = wp[at], pr[at] = wq[at];
    }
    for (int lvl = 0; lvl < top; ++lvl) {  // merge with the left sibling of level lvl
      double share;
      lw_cur = log_add_exp_share(T.lw(lvl)[i], lw_cur, share);
      const bool take_outer = draw() < share;
      const double* spf = T.slot(lvl, 0) + i;
      const double* spl = T.slot(lvl, 1) + i;
      const double* srh = T.slot(lvl, 2) + i;
      const double* spr = T.slot(lvl, 3) + i;
      UTurn ut;
#pragma unroll 2
      for (int r = 0; r < Q; ++r) {
        const int64_t at = (int64_t)r * lds;
        const double rh_i = srh[at], rh_o = rh[at], p1 = spf[at], rho = rh_i + rh_o;
        ut.add(p1, spl[at], rh_i, pf[at], wp[at], rh_o, rho);
        rh[at] = rho, pf[at] = p1;
        if (!take_outer) pr[at] = spr[at];
      }
      if (ut.turned(extra)) {
        terminate = true;
        break;
      }
    }
    if (!terminate) T.lw(top)[i] = lw_cur;
  }
  ++leaf;
  bool done = terminate;
  if (!terminate && leaf == (1 << depth)) {  // new subtree complete: join it to the tree
    // (inner = the tree so far: first-built = its far edge, last-built = the edge the
    // subtree grew from; outer = the subtree, whose last state is the working one)
    const double lw_sub = T.lw(depth)[i], lw_tree = T.dbl(T.LW_TREE)[i];
    const bool take = draw() < exp(lw_sub - lw_tree);
    double share;
    T.dbl(T.LW_TREE)[i] = log_add_exp_share(lw_tree, lw_sub, share);
    if (take) T.i32(T.MOVED)[i] = 1;
    const double* spf = T.slot(depth, 0) + i;
    const double* srh = T.slot(depth, 2) + i;
    const double* spr = T.slot(depth, 3) + i;
    double* eq = T.blk(dir > 0 ? T.EP_Q : T.EN_Q) + i;
    double* ep = T.blk(dir > 0 ? T.EP_P : T.EN_P) + i;
    const double* fp = T.blk(dir > 0 ? T.EN_P : T.EP_P) + i;
    double* rho = T.blk(T.RHO) + i;
    double* nx = T.blk(T.NEXT) + i;
    UTurn ut;
#pragma unroll 2
    for (int r = 0; r < Q; ++r) {
      const int64_t at = (int64_t)r * lds;
      const double rh_i = rho[at], rh_o = srh[at], pl_o = wp[at], rr = rh_i + rh_o;
      ut.add(fp[at], ep[at], rh_i, spf[at], pl_o, rh_o, rr);
      if (take) nx[at] = spr[at];
      rho[at] = rr;
      eq[at] = wq[at], ep[at] = pl_o;
    }
    ++depth, leaf = 0;
    done = ut.turned(extra) || depth >= max_depth;
  }
  T.i32(T.DEPTH)[i] = depth, T.i32(T.LEAF)[i] = leaf, T.i32(T.KDRAW)[i] = (int)kdraw;
  T.i32(T.ACTIVE)[i] = !done;
  if (!done) atomicAdd(counter, 1);
}

// end of the transition: new state, statistics, traces, dual averaging
template <int NCH>
__global__ void __launch_bounds__(32 * NCH)
    kn_end(SNutsLayout T, int64_t n, int64_t ld, double* q, SampleArgs a, NutsArgs na, int it) {
  const int64_t i = (int64_t)blockIdx.x * 32 + threadIdx.x;
  if (i >= n) return;
  int k0, k1;
  row_chunk<NCH>(T.Q, k0, k1);
#pragma unroll 4
  for (int k = k0; k < k1; ++k) q[(int64_t)k * ld + i] = T.blk(T.NEXT)[(int64_t)k * T.lds + i];
  if (a.trace_q && (it + 1) % a.trace_every == 0) {
    double* tq = a.trace_q + (int64_t)((it + 1) / a.trace_every - 1) * a.trace_dim * ld + i;
    for (int k = k0; k < k1 && k < a.trace_dim; ++k)
      tq[(int64_t)k * ld] = T.blk(T.NEXT)[(int64_t)k * T.lds + i];
  }
  if (threadIdx.y != 0) return;
  const int n_leaf = T.i32(T.N_LEAF)[i], st = T.i32(T.ST)[i], depth = T.i32(T.DEPTH)[i];
  const double a_stat = n_leaf > 0 ? T.dbl(T.SUM_ACC)[i] / n_leaf : 0.0;
  if (a.trace_accept) a.trace_accept[(int64_t)it * ld + i] = a_stat;
  if (na.trace_n_step) na.trace_n_step[(int64_t)it * ld + i] = n_leaf;
  if (na.trace_depth) na.trace_depth[(int64_t)it * ld + i] = depth;
  if (a.stats) {
    a.stats[0 * ld + i] += a_stat, a.stats[1 * ld + i] += T.i32(T.MOVED)[i];
    a.stats[2 * ld + i] +=
        (st == MLB_ST_DIVERGED || st == MLB_ST_NOT_CONVERGED || st == MLB_ST_LINALG);
    a.stats[3 * ld + i] += (st == MLB_ST_NON_REVERSIBLE);
    a.stats[4 * ld + i] += (st == MLB_ST_H_DIVERGED);
    a.stats[5 * ld + i] += n_leaf, a.stats[6 * ld + i] += T.i32(T.ITS)[i];
    a.stats[7 * ld + i] += depth;
  }
  if (a.last_status) a.last_status[i] = st;
  if (a.adapt) {  // Mici DualAveragingStepSizeAdapter.update
    double ad_iter = a.adapt[0 * ld + i], ad_smooth = a.adapt[1 * ld + i];
    double ad_err = a.adapt[2 * ld + i];
    const double ad_target = a.adapt[3 * ld + i];
    ad_iter += 1;
    const double w = 1.0 / (a.a_offset + ad_iter);
    ad_err = (1.0 - w) * ad_err + w * (a.a_target - a_stat);
    const double sw = pow(1.0 / ad_iter, a.a_decay);
    const double log_eps = ad_target - ad_err * sqrt(ad_iter) / a.a_reg;
    ad_smooth = (1.0 - sw) * ad_smooth + sw * log_eps;
    a.adapt[0 * ld + i] = ad_iter, a.adapt[1 * ld + i] = ad_smooth;
    a.adapt[2 * ld + i] = ad_err, a.adapt[4 * ld + i] = exp(log_eps);
  }
}

template <class M, int SOLVER>
int phased_nuts_sample(const mlb_model* m, const OdeData& D, int64_t n, int64_t ld, double* q,
                       const Opts& o, const SampleArgs& a, const NutsArgs& na, cudaStream_t s) {
  constexpr int NCH = kChunkWarps;
  SNutsLayout T;
  T.lds = pad32(n), T.Q = M::U + D.Y, T.D = na.max_depth > 0 ? na.max_depth : 1;
  ScratchBuf sb;
  int rc = sb.alloc(T.rows(), T.lds, s);
  if (rc != MLB_OK) return rc;
  T.base = sb.p;
  int* counter = reinterpret_cast<int*>(sb.p + (T.rows() - 1) * T.lds);
  int* host_count = nullptr;
  {
    std::lock_guard<std::mutex> lock(m->pin_mutex);
    if (!m->pinned_word && cudaMallocHost((void**)&m->pinned_word, sizeof(int)) != cudaSuccess) {
      g_err = "pinned counter allocation failed";
      return MLB_ERR_CUDA;
    }
  }
  // a second pinned word: the step driver polls m->pinned_word while we are between rounds
  if (cudaMallocHost((void**)&host_count, sizeof(int)) != cudaSuccess) {
    g_err = "pinned counter allocation failed";
    return MLB_ERR_CUDA;
  }
  auto read_counter = [&]() -> int {
    cudaMemcpyAsync(host_count, counter, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaMemsetAsync(counter, 0, sizeof(int), s);
    if (cudaStreamSynchronize(s) != cudaSuccess) return -1;
    return *host_count;
  };
  const dim3 g1(sgrid(n)), cblk(32, NCH);
  auto step_io = [&](int n_step, bool pf, const int32_t* mask) {
    StepsIO io{T.blk(T.WQ), T.blk(T.WP), 0.0, T.dbl(T.DT), n_step, T.i32(T.S_STATUS),
               T.i32(T.S_DONE), T.i32(T.S_ITF), T.i32(T.S_ITR), T.dbl(T.H_INIT),
               T.dbl(T.H_FINAL)};
    io.project_first = pf, io.alive_mask = mask;
    return io;
  };
#define MLB_KN(call)                                      \
  do {                                                    \
    call;                                                 \
    if ((rc = finish_launch()) != MLB_OK) goto fail;      \
  } while (0)
  cudaMemsetAsync(counter, 0, sizeof(int), s);
  for (int it = 0; it < a.n_iter; ++it) {
    MLB_KN((kn_begin<NCH><<<g1, cblk, 0, s>>>(T, n, ld, q, a, it)));
    // momentum refresh: projection + h_init through a zero-step call
    if ((rc = phased_leapfrog_steps<M, SOLVER>(m, D, n, T.lds, o, step_io(0, true, nullptr), s)) !=
        MLB_OK)
      goto fail;
    MLB_KN((kn_init_tree<NCH><<<g1, cblk, 0, s>>>(T, n, na.max_depth, counter)));
    int active = read_counter();
    while (active > 0) {
      MLB_KN((kn_pre<<<g1, 32, 0, s>>>(T, n, a, it)));
      if ((rc = phased_leapfrog_steps<M, SOLVER>(m, D, n, T.lds, o,
                                                 step_io(1, false, T.i32(T.ACTIVE)), s)) != MLB_OK)
        goto fail;
      MLB_KN((kn_post<<<g1, 32, 0, s>>>(T, n, a, it, na.max_depth, o.flags, counter)));
      active = read_counter();
    }
    if (active < 0) {
      g_err = "stream synchronisation failed in the tree loop";
      rc = MLB_ERR_CUDA;
      goto fail;
    }
    MLB_KN((kn_end<NCH><<<g1, cblk, 0, s>>>(T, n, ld, q, a, na, it)));
  }
#undef MLB_KN
  cudaStreamSynchronize(s);
  cudaFreeHost(host_count);
  return MLB_OK;
fail:
  cudaStreamSynchronize(s);
  cudaFreeHost(host_count);
  return rc;
}

}  // namespace mlb
