// Phase-structured constrained leapfrog for the streaming (ODE-constrained) family.
//
// The monolithic per-chain kernel (mlb_stream_launch.cuh) leaves a B200 almost empty: a
// FitzHugh-Nagumo batch of 16,384 chains is 512 warps on 592 sub-partitions, and each warp
// walks its sweeps one sensitivity direction / pair group after another.  All chains of a
// launch go through the SAME sequence of sub-phases (only solver iteration counts differ
// per chain), and every Y-sized intermediate already lives in global scratch, so the
// trajectory is re-cut into kernels at sub-phase boundaries and driven from the host:
//
//   * the expensive sweeps run on a 2-D grid (chain block x ROLE): a Jacobian sweep is
//     split by sensitivity direction (FN 4 roles x 2 directions, HH 6 x 1), the
//     second-order sweep by pair group (FN 6, HH 7) - 4-7x more resident warps, no
//     redundant arithmetic beyond the state itself, small per-kernel code;
//   * per-chain control (iteration counters, norms, status, alive / active masks) sits in
//     small device arrays; the solver loop is a host loop that reads ONE int (chains still
//     iterating) per iteration - microseconds against millisecond sweeps; warps whose 32
//     chains have all converged exit at once;
//   * arithmetic is that of the monolithic kernel (same device functions), so the two
//     paths agree bitwise and share the parity tests.
#pragma once
#include <cstdlib>

#include "mlb_stream_launch.cuh"

namespace mlb {

// per-chain control block rows in the scratch allocation (after StreamScratch::rows)
enum { CT_ALIVE, CT_ACTIVE, CT_ITERS, CT_STATUS, CT_DONE, CT_TF, CT_TR, CT_NCONSTR, CT_INT_ROWS };
enum { CD_NDQ, CD_ERR, CD_NLD, CD_LOGDET, CD_HINIT, CD_HLAST, CD_DBL_ROWS };

struct Phased {
  StreamScratch S;
  Col hpart;     // [NG * ND] partial second-order gradients
  double* cd;    // [CD_DBL_ROWS][lds]
  int32_t* ct;   // [CT_INT_ROWS][lds]
  int64_t lds;
  MLB_DEV int32_t& I(int row, int64_t i) const { return ct[(int64_t)row * lds + i]; }
  MLB_DEV double& Dv(int row, int64_t i) const { return cd[(int64_t)row * lds + i]; }
};

struct PhasedLayout {
  double* base;
  int64_t lds;
  int U, Y, NGND;
  // first column of this VIEW of the allocation (chain i of the view lives in column
  // col0 + i; the list of still-iterating chains and the counter belong to the allocation)
  int64_t col0 = 0;
  __host__ __device__ int64_t rows_main() const { return StreamScratch::rows(U, Y); }
  // + one row for the list of still-iterating chains (lds int32 used) + one for the counter
  __host__ __device__ int64_t rows_total() const {
    return rows_main() + NGND + CD_DBL_ROWS + (CT_INT_ROWS + 1) / 2 + 2;
  }
  __host__ __device__ int32_t* list() const {
    return reinterpret_cast<int32_t*>(base + (rows_total() - 2) * lds);
  }
  __host__ __device__ Phased at(int64_t i) const {
    Phased p;
    p.lds = lds;
    double* after = base + StreamScratch::rows(U, Y) * lds;
    p.hpart = Col{after + col0 + i, lds};
    double* cd0 = after + (int64_t)NGND * lds;
    p.cd = cd0 + col0;
    p.ct = reinterpret_cast<int32_t*>(cd0 + (int64_t)CD_DBL_ROWS * lds) + col0;
#ifdef __CUDA_ARCH__
    p.S = StreamScratch::make(base, lds, col0 + i, U, Y);
#endif
    return p;
  }
};

// ---- observation-chunked kernels --------------------------------------------------------
// The Y-sized streaming phases (capacitance accumulation, Woodbury apply, cotangent
// projection, B = Gram^-1 A) are sums over observations followed by a U x U solve and a
// second pass over the observations.  A CTA of NCH warps works on one batch of 32 chains:
// warp j streams observations [j CH, (j+1) CH), the partial sums are combined in shared
// memory IN WARP ORDER (deterministic), warp 0 does the U x U factorisation / solve and
// hands the result back through shared memory.  NCH times more loads in flight per chain
// than the one-thread-per-chain form, which ran at ~5 % of the HBM rate.
constexpr int kChunkWarps = 8;

struct ChunkCtx {
  int lane, wj, k0, k1;
  int64_t i;
  bool in_range;
};
template <int NCH>
MLB_DEV ChunkCtx chunk_ctx(int64_t n, int Y) {
  ChunkCtx c;
  c.lane = threadIdx.x, c.wj = threadIdx.y;
  c.i = (int64_t)blockIdx.x * 32 + c.lane;
  c.in_range = c.i < n;
  const int CH = (Y + NCH - 1) / NCH;
  c.k0 = c.wj * CH;
  c.k1 = c.k0 + CH < Y ? c.k0 + CH : Y;
  if (c.k0 > Y) c.k0 = Y;
  return c;
}
// the same for slot blockIdx.x * 32 + lane of the list of still-iterating chains
template <int NCH>
MLB_DEV ChunkCtx chunk_ctx_listed(const int32_t* list, int64_t n_act, int Y) {
  ChunkCtx c = chunk_ctx<NCH>(n_act, Y);
  c.i = c.in_range ? list[c.i] : 0;
  return c;
}
// rows [k0, k1) of the Q position / momentum rows for warp wj of a (32, NCH) CTA
template <int NCH>
MLB_DEV void row_chunk(int Q, int& k0, int& k1) {
  const int CH = (Q + NCH - 1) / NCH;
  k0 = (int)threadIdx.y * CH;
  k0 = k0 < Q ? k0 : Q;
  k1 = k0 + CH < Q ? k0 + CH : Q;
}

// v <- sum over the CTA's warps of v (added in warp order); the total is returned to warp 0
// only.  sh: [R][32] doubles.  Every thread of the CTA must call it.
template <int R, int NCH>
MLB_DEV void cta_sum_to_warp0(double (&v)[R], double* sh) {
  const int lane = threadIdx.x, wj = threadIdx.y;
#pragma unroll 1
  for (int j = 1; j < NCH; ++j) {
    if (wj == j) {
#pragma unroll
      for (int r = 0; r < R; ++r) sh[r * 32 + lane] = v[r];
    }
    __syncthreads();
    if (wj == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] += sh[r * 32 + lane];
    }
    __syncthreads();
  }
}
// same for a running norm (max for the maximum norm, sum of squares for the Euclidean norm)
template <int NCH>
MLB_DEV double cta_norm_to_warp0(double acc, int kind, double* sh) {
  const int lane = threadIdx.x, wj = threadIdx.y;
  if (wj > 0) sh[(wj - 1) * 32 + lane] = acc;
  __syncthreads();
  if (wj == 0) {
#pragma unroll 1
    for (int j = 1; j < NCH; ++j) {
      const double o = sh[(j - 1) * 32 + lane];
      acc = kind == MLB_NORM_MAX ? (o > acc ? o : acc) : acc + o;
    }
  }
  __syncthreads();
  return acc;
}
// warp 0 publishes R values per chain to every warp of the CTA
template <int R>
MLB_DEV void cta_bcast_from_warp0(double (&v)[R], double* sh) {
  const int lane = threadIdx.x;
  if (threadIdx.y == 0) {
#pragma unroll
    for (int r = 0; r < R; ++r) sh[r * 32 + lane] = v[r];
  }
  __syncthreads();
  if (threadIdx.y != 0) {
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = sh[r * 32 + lane];
  }
  __syncthreads();
}

#define MLB_PH_INDEX()                                              \
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; \
  if (i >= n) return;                                               \
  const Phased P = L.at(i)

// thread -> chain through the list of still-iterating chains (solver-loop kernels)
#define MLB_PH_INDEX_LISTED()                                        \
  const int64_t t_ = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; \
  if (t_ >= n_act) return;                                           \
  const int64_t i = L.list()[t_];                                    \
  const Phased P = L.at(i)

// Compaction of the still-iterating chains (CT_ACTIVE != 0) into L.list(), ascending, and
// their number into *counter: the sweeps of the solver loop are launched over that list
// only, so their cost follows the number of chains that still iterate instead of the
// slowest chain of the batch (Newton iteration counts of a batch spread from 2 to 12+).
// One CTA walks the n flags in tiles of 1,024 (ballot + warp totals in shared memory):
// microseconds against the millisecond sweeps, and the order is deterministic.
template <class M>
__global__ void __launch_bounds__(1024) kp_compact(PhasedLayout L, int64_t n, int* counter) {
  __shared__ int warp_tot[32];
  __shared__ int base_sh;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int32_t* act = L.at(0).ct + (int64_t)CT_ACTIVE * L.lds;
  int32_t* list = L.list();
  if (threadIdx.x == 0) base_sh = 0;
  __syncthreads();
  for (int64_t t0 = 0; t0 < n; t0 += 1024) {
    const int64_t i = t0 + threadIdx.x;
    const bool f = i < n && act[i] != 0;
    const unsigned b = __ballot_sync(0xffffffffu, f);
    if (lane == 0) warp_tot[w] = __popc(b);
    __syncthreads();
    int before = 0;
    for (int k = 0; k < w; ++k) before += warp_tot[k];
    const int base = base_sh;
    if (f) list[base + before + __popc(b & ((1u << lane) - 1u))] = (int32_t)i;
    __syncthreads();
    if (threadIdx.x == 1023) base_sh = base + before + __popc(b);
    __syncthreads();
  }
  if (threadIdx.x == 0) *counter = base_sh;
}

template <class M, int NCH>
__global__ void __launch_bounds__(32 * NCH)
    kp_begin(PhasedLayout L, int64_t n, int64_t ld, const double* q, const double* p,
             const int32_t* alive_mask) {
  MLB_PH_INDEX();
  const int Q = L.U + L.Y;
  const Col qw = col(q, ld, i), pw = col(p, ld, i);
  const int alive = alive_mask ? alive_mask[i] != 0 : 1;
  int k0, k1;
  row_chunk<NCH>(Q, k0, k1);
  if (alive) {
#pragma unroll 4
    for (int k = k0; k < k1; ++k) P.S.cq[k] = qw[k], P.S.cp[k] = pw[k];
  }
  if (threadIdx.y != 0) return;
  P.I(CT_ALIVE, i) = alive, P.I(CT_ACTIVE, i) = 0, P.I(CT_ITERS, i) = 0;
  P.I(CT_STATUS, i) = MLB_ST_OK, P.I(CT_DONE, i) = 0, P.I(CT_TF, i) = 0, P.I(CT_TR, i) = 0;
  P.I(CT_NCONSTR, i) = 0;
  P.Dv(CD_HINIT, i) = nan(""), P.Dv(CD_HLAST, i) = nan("");
}

// Jacobian sweep, one role per blockIdx.y.  src / dst select the position (working qw or
// trial qs) and the destination blocks (Jc at an evaluation, Jt inside Newton).
// (blockDim.y = M::GATE_WARPS warps share the chain batch of a cooperative sweep; the
// lanes that leave early are the same in each of them, so the barriers stay consistent)
template <class M, bool EVAL, int R>
__global__ void __launch_bounds__(kStreamBlock * M::GATE_WARPS)
    kp_jacob(const __grid_constant__ OdeData D, PhasedLayout L, int64_t n, int64_t ld,
             const double* q) {
  extern __shared__ double xch[];
  // n: chains of the batch (EVAL) / length of the list of still-iterating chains
  const int64_t t_ = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t_ >= n) return;
  const int64_t i = EVAL ? t_ : (int64_t)L.list()[t_];
  const Phased P = L.at(i);
  if (!P.I(EVAL ? CT_ALIVE : CT_ACTIVE, i)) return;
  const Col src = EVAL ? col(q, ld, i) : P.S.qs;
  const Col dst = EVAL ? P.S.Jc : P.S.Jt;
  const Col c = EVAL ? Col{nullptr, 0} : P.S.c;
  // R sensitivity directions per role (gridDim.y = ND / R roles)
  [&]<int... K>(std::integer_sequence<int, K...>) {
    ((blockIdx.y == K
          ? Stream<M>::template jacob_role<K * R, R, K == 0, M::GATE_WARPS>(D, src, dst, c, xch)
          : (void)0),
     ...);
  }(std::make_integer_sequence<int, M::ND / R>{});
}

// Launch of a Jacobian sweep over `cnt` chains.  The split of the ND sensitivity directions
// over roles trades latency for redundant work (every role re-integrates the state itself):
// M::ROLE_DIRS directions per role while the launch is small (latency-bound: more, shorter
// threads), M::ROLE_DIRS_WIDE once it fills the GPU (throughput-bound: fewer copies of the
// primal recurrence).  Measured on HH (profiles/r4j_hh_variants*.txt): 8,192 chains 107 ->
// 77 ms per step with 3 directions per role, 1,024 chains 33.5 -> 41 ms.
// launch size above which the sweeps take their wide form (MLB_WIDE_MIN_CHAINS overrides
// the model's default: tuning / tests)
template <class M>
inline int64_t wide_min_chains() {
  static const int64_t v = [] {
    const char* e = std::getenv("MLB_WIDE_MIN_CHAINS");
    return e ? (int64_t)std::atoll(e) : (int64_t)M::WIDE_MIN_CHAINS;
  }();
  return v;
}

template <class M>
inline int64_t wide_min_chains_second_order() {
  static const int64_t v = [] {
    const char* e = std::getenv("MLB_WIDE_MIN_CHAINS_SO");
    return e ? (int64_t)std::atoll(e) : (int64_t)M::WIDE_MIN_CHAINS_SO;
  }();
  return v;
}

template <class M, bool EVAL>
int launch_jacob(const OdeData& D, const PhasedLayout& L, int64_t cnt, int64_t ld,
                 const double* q, cudaStream_t s) {
  const dim3 sblk(kStreamBlock, M::GATE_WARPS);
  if (M::ROLE_DIRS_WIDE != M::ROLE_DIRS && cnt > wide_min_chains<M>()) {
    constexpr int R = M::ROLE_DIRS_WIDE;
    constexpr size_t sh = sizeof(double) * M::template xch_doubles<R, 0>();
    kp_jacob<M, EVAL, R><<<dim3(sgrid(cnt), M::ND / R), sblk, sh, s>>>(D, L, cnt, ld, q);
  } else {
    constexpr int R = M::ROLE_DIRS;
    constexpr size_t sh = sizeof(double) * M::template xch_doubles<R, 0>();
    kp_jacob<M, EVAL, R><<<dim3(sgrid(cnt), M::ND / R), sblk, sh, s>>>(D, L, cnt, ld, q);
  }
  return finish_launch();
}

// residual sweep at the trial position (quasi-Newton)
template <class M>
__global__ void __launch_bounds__(kStreamBlock)
    kp_constr(const __grid_constant__ OdeData D, PhasedLayout L, int64_t n_act) {
  MLB_PH_INDEX_LISTED();
  if (!P.I(CT_ACTIVE, i)) return;
  double u[M::U];
  Stream<M>::load_u(P.S.qs, u);
  const Col qs = P.S.qs;
  Stream<M>::constr(D, u, [&](int k) { return qs[M::U + k]; }, P.S.c, MLB_NORM_MAX);
}

// after the Jacobian sweep of an evaluation: Gram factor, log det, prior, and everything
// of grad 0.5 log det(J J^T) that does not need second-order sensitivities (the arithmetic
// of Stream::decompose / logdet / grad_logdet<false>, observation-chunked)
template <class M, int NCH>
__global__ void __launch_bounds__(32 * NCH)
    kp_eval_mid(const __grid_constant__ OdeData D, PhasedLayout L, int64_t n, int64_t ld,
                double* q, double* p) {
  using S = Stream<M>;
  constexpr int U = M::U, NT = U * (U + 1) / 2;
  extern __shared__ double sh[];
  const int Y = L.Y;
  const ChunkCtx cx = chunk_ctx<NCH>(n, Y);
  const int64_t i = cx.in_range ? cx.i : 0;
  const Phased P = L.at(i);
  const bool live = cx.in_range && P.I(CT_ALIVE, i);
  const Col qw = col(q, ld, i), jac = P.S.Jc, gram = P.S.Gc, B = P.S.Jt, g = P.S.g;
  // pass 1: capacitance matrix (lower triangle) and sum log d over the chunk
  double acc[NT + 1];
#pragma unroll
  for (int r = 0; r < NT + 1; ++r) acc[r] = 0.0;
  if (live && cx.k0 < cx.k1) {
    double prev = jac[U * Y + cx.k0], lp;
    prev *= prev, lp = log(prev);
#pragma unroll 4
    for (int k = cx.k0; k < cx.k1; ++k) {
      const double dn = jac[U * Y + k], d = dn * dn, rd = rcp_fast(d);
      gram[U * U + k] = d;
      if (d != prev) prev = d, lp = log(d);
      acc[NT] += lp;
      double row[U];
#pragma unroll
      for (int a = 0; a < U; ++a) row[a] = jac[k * U + a];
      int t = 0;
#pragma unroll
      for (int a = 0; a < U; ++a) {
        const double ra = row[a] * rd;
#pragma unroll
        for (int b = 0; b <= a; ++b, ++t) acc[t] = fma(ra, row[b], acc[t]);
      }
    }
  }
  cta_sum_to_warp0<NT + 1, NCH>(acc, sh);
  // warp 0: Cholesky, log det, prior, M = C^-1
  double Minv[U * U + 1];  // + the factorisation flag
  double u[U], gu[U], nld = 0.0, ldv = 0.0, trM = 0.0;
#pragma unroll
  for (int r = 0; r < U * U + 1; ++r) Minv[r] = 0.0;
  if (live && cx.wj == 0) {
    typename S::Chol C;
    int t = 0;
#pragma unroll
    for (int a = 0; a < U; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b, ++t) C.L[a][b] = acc[t] + (a == b ? 1.0 : 0.0), C.L[b][a] = C.L[a][b];
    C.ok = chol_lower<U>(C.L, C.rdiag);
#pragma unroll
    for (int a = 0; a < U; ++a)
#pragma unroll
      for (int b = 0; b < U; ++b) gram[a * U + b] = C.L[a][b];
    double sl = 0.0;
#pragma unroll
    for (int a = 0; a < U; ++a) sl += log(C.L[a][a]);
    ldv = fma(0.5, acc[NT], sl);
    S::load_u(qw, u);
    nld = M::nld_u(u, gu);
#pragma unroll
    for (int j = 0; j < U; ++j) {
      double cj[U];
#pragma unroll
      for (int a = 0; a < U; ++a) cj[a] = (a == j) ? 1.0 : 0.0;
      chol_solve<U>(C.L, C.rdiag, cj);
#pragma unroll
      for (int a = 0; a < U; ++a) Minv[a * U + j] = cj[a];
    }
#pragma unroll
    for (int a = 0; a < U; ++a) trM += Minv[a * U + a];
    Minv[U * U] = C.ok ? 1.0 : 0.0;
  }
  cta_bcast_from_warp0<U * U + 1>(Minv, sh);
  // pass 2: B = Gram^-1 A rows, noise part of the gradient
  double part[2] = {0.0, 0.0};  // sigma-column sum, 0.5 |n|^2
  if (live) {
#pragma unroll 4
    for (int k = cx.k0; k < cx.k1; ++k) {
      const double dn = jac[U * Y + k], rd = rcp_fast(dn * dn), ni = qw[U + k];
      double row[U];
#pragma unroll
      for (int a = 0; a < U; ++a) row[a] = jac[k * U + a];
#pragma unroll
      for (int a = 0; a < U; ++a) {
        double t = 0.0;
#pragma unroll
        for (int b = 0; b < U; ++b) t = fma(row[b], Minv[b * U + a], t);
        t *= rd;
        B[k * U + a] = t;
        if (a == M::SIGMA) {
          part[0] = fma(t, dn * ni, part[0]);
          g[U + k] = fma(t, dn, ni);
        }
      }
      part[1] = fma(0.5 * ni, ni, part[1]);
    }
  }
  cta_sum_to_warp0<2, NCH>(part, sh);
  if (live && cx.wj == 0) {
    gu[M::SIGMA] += (double)Y - ((double)U - trM) + part[0];
#pragma unroll
    for (int a = 0; a < U; ++a) g[a] = gu[a];
    P.Dv(CD_NLD, i) = nld + part[1], P.Dv(CD_LOGDET, i) = ldv;
    if (Minv[U * U] == 0.0) {  // hand the chain back at its last committed state
      P.I(CT_STATUS, i) = MLB_ST_LINALG, P.I(CT_ALIVE, i) = 0;
      const Col pw = col(p, ld, i);
      const int Q = U + Y;
#pragma unroll 4
      for (int k = 0; k < Q; ++k) qw[k] = P.S.cq[k], pw[k] = P.S.cp[k];
    }
  }
}

// second-order sweep, one pair group per blockIdx.y; partial gradients to hpart
// (not cooperative: with ten doubles per gating variable to exchange and the potential's
// jet arithmetic repeated in every warp, four gate warps measured 4.8 ms against 3.1 ms)
template <class M, int NGX>
__global__ void __launch_bounds__(kStreamBlock)
    kp_second_order(const __grid_constant__ OdeData D, PhasedLayout L, int64_t n, int64_t ld,
                    const double* q) {
  MLB_PH_INDEX();
  if (!P.I(CT_ALIVE, i)) return;
  constexpr int U = M::U;
  static_assert(M::NPAIR % NGX == 0 && NGX <= M::NG, "pair groups must be equal-sized");
  double u[U], h[U];
  Stream<M>::load_u(col(q, ld, i), u);
#pragma unroll
  for (int a = 0; a < U; ++a) h[a] = 0.0;
  // NGX groups of NPAIR / NGX pairs (gridDim.y = NGX)
  [&]<int... G>(std::integer_sequence<int, G...>) {
    ((blockIdx.y == G
          ? Stream<M>::template second_order_group<G, 1, M::NPAIR / NGX>(D, u, P.S.Jt, h)
          : (void)0),
     ...);
  }(std::make_integer_sequence<int, NGX>{});
#pragma unroll
  for (int d = 0; d < M::ND; ++d) P.hpart[(int)blockIdx.y * M::ND + d] = h[M::dir_u(d)];
  if (NGX < M::NG && blockIdx.y == 0) {  // the consumer adds up all M::NG partials
    for (int k = NGX * M::ND; k < M::NG * M::ND; ++k) P.hpart[k] = 0.0;
  }
}

// Second-order sweep over the batch: M::NG pair groups (one thread each) for a small batch,
// M::NG_WIDE for one that fills the GPU (every group re-integrates the state and all first-
// order sensitivities, so fewer groups = less redundant work, more latency per thread).
template <class M>
int launch_second_order(const OdeData& D, const PhasedLayout& L, int64_t n, int64_t ld,
                        const double* q, cudaStream_t s) {
  if (M::NG_WIDE != M::NG && n > wide_min_chains_second_order<M>()) {
    kp_second_order<M, M::NG_WIDE><<<dim3(sgrid(n), M::NG_WIDE), kStreamBlock, 0, s>>>(D, L, n, ld, q);
  } else {
    kp_second_order<M, M::NG><<<dim3(sgrid(n), M::NG), kStreamBlock, 0, s>>>(D, L, n, ld, q);
  }
  return finish_launch();
}

// [finish a fresh evaluation: add the second-order partials] [half kick] cotangent
// projection p -= J^T Gram^-1 J p [commit the step / record h_init]; observation-chunked.
// commit: 0 no, 1 a completed step (with the |h - h_init| > max_dh test if max_dh > 0),
// 2 the momentum refresh of a sampler transition (committed momentum and h_init are those
// AFTER the projection; no step is counted)
template <class M, int NCH>
__global__ void __launch_bounds__(32 * NCH)
    kp_kick_project(PhasedLayout L, int64_t n, int64_t ld, double* p, double* q, double dt,
                    const double* dt_pc, int fresh, int kick, int first, int commit,
                    double max_dh) {
  using S = Stream<M>;
  constexpr int U = M::U;
  extern __shared__ double sh[];
  const int Y = L.Y;
  const ChunkCtx cx = chunk_ctx<NCH>(n, Y);
  const int64_t i = cx.in_range ? cx.i : 0;
  const Phased P = L.at(i);
  const bool live = cx.in_range && P.I(CT_ALIVE, i);
  const Col pw = col(p, ld, i), qw = col(q, ld, i), jac = P.S.Jc, g = P.S.g, w = P.S.w;
  double gfin[U], vu[U], acc[U + 1];  // acc: A^T D^-1 (J p) partial, |p|^2 partial (first)
#pragma unroll
  for (int r = 0; r < U + 1; ++r) acc[r] = 0.0;
  const double hk = kick ? -0.5 * (dt_pc ? dt_pc[i] : dt) : 0.0;
  if (live) {
#pragma unroll
    for (int a = 0; a < U; ++a) {
      double t = g[a];
      if (fresh) {
        const int d = M::u_dir(a);
        if (d >= 0)
          for (int gi = 0; gi < M::NG; ++gi) t += P.hpart[gi * M::ND + d];
      }
      gfin[a] = t;
      const double pa = pw[a];
      if (first && cx.wj == 0) acc[U] = fma(pa, pa, acc[U]);
      vu[a] = kick ? fma(hk, t, pa) : pa;
    }
#pragma unroll 4
    for (int k = cx.k0; k < cx.k1; ++k) {
      double pk = pw[U + k];
      if (first) acc[U] = fma(pk, pk, acc[U]);
      if (kick) pk = fma(hk, g[U + k], pk), pw[U + k] = pk;
      const double dn = jac[U * Y + k];
      double t = dn * pk, row[U];
#pragma unroll
      for (int a = 0; a < U; ++a) row[a] = jac[k * U + a], t = fma(row[a], vu[a], t);
      w[k] = t;
      const double wr = t * rcp_fast(dn * dn);
#pragma unroll
      for (int a = 0; a < U; ++a) acc[a] = fma(row[a], wr, acc[a]);
    }
  }
  cta_sum_to_warp0<U + 1, NCH>(acc, sh);
  double tt[U];
#pragma unroll
  for (int a = 0; a < U; ++a) tt[a] = acc[a];
  if (live && cx.wj == 0) {
    if (first)
      P.Dv(CD_HINIT, i) = P.Dv(CD_HLAST, i) = P.Dv(CD_NLD, i) + P.Dv(CD_LOGDET, i) + 0.5 * acc[U];
    typename S::Chol C;
    S::load_chol(P.S.Gc, C);
    chol_solve<U>(C.L, C.rdiag, tt);
  }
  cta_bcast_from_warp0<U>(tt, sh);
  double su[U + 1];  // J^T x latent part, |p|^2 partial of the projected momentum
#pragma unroll
  for (int r = 0; r < U + 1; ++r) su[r] = 0.0;
  if (live) {
#pragma unroll 4
    for (int k = cx.k0; k < cx.k1; ++k) {
      const double dn = jac[U * Y + k];
      double t = w[k], row[U];
#pragma unroll
      for (int a = 0; a < U; ++a) row[a] = jac[k * U + a], t = fma(-row[a], tt[a], t);
      const double x = t * rcp_fast(dn * dn);
#pragma unroll
      for (int a = 0; a < U; ++a) su[a] = fma(row[a], x, su[a]);
      const double pk = fma(-dn, x, pw[U + k]);
      pw[U + k] = pk;
      if (commit) {
        P.S.cq[U + k] = qw[U + k], P.S.cp[U + k] = pk;
        su[U] = fma(pk, pk, su[U]);
      }
    }
  }
  cta_sum_to_warp0<U + 1, NCH>(su, sh);
  if (live && cx.wj == 0) {
    double hs = su[U];
#pragma unroll
    for (int a = 0; a < U; ++a) {
      const double pa = vu[a] - su[a];
      pw[a] = pa;
      if (fresh) g[a] = gfin[a];
      if (commit) P.S.cq[a] = qw[a], P.S.cp[a] = pa, hs = fma(pa, pa, hs);
    }
    if (commit) {
      const double h = P.Dv(CD_NLD, i) + P.Dv(CD_LOGDET, i) + 0.5 * hs;
      P.Dv(CD_HLAST, i) = h;
      if (commit == 2) {
        P.Dv(CD_HINIT, i) = h;
      } else {
        P.I(CT_DONE, i) += 1;
        if (max_dh > 0.0 && fabs(h - P.Dv(CD_HINIT, i)) > max_dh)
          P.I(CT_STATUS, i) = MLB_ST_H_DIVERGED, P.I(CT_ALIVE, i) = 0;
      }
    }
  }
}

template <class M, int NCH>
__global__ void __launch_bounds__(32 * NCH)
    kp_solve_begin(PhasedLayout L, int64_t n, int64_t ld, const double* q, const double* p,
                   double dt, const double* dt_pc, double sign, int max_iters) {
  MLB_PH_INDEX();
  const int alive = P.I(CT_ALIVE, i);  // (not written by this kernel)
  const bool go = alive && max_iters > 0;
  if (threadIdx.y == 0) {
    P.I(CT_ACTIVE, i) = go, P.I(CT_ITERS, i) = 0, P.I(CT_NCONSTR, i) = 0;
    P.Dv(CD_NDQ, i) = INFINITY, P.Dv(CD_ERR, i) = -1.0;
  }
  if (!alive) return;
  const int Q = L.U + L.Y;
  const double sdt = sign * (dt_pc ? dt_pc[i] : dt);
  const Col qw = col(q, ld, i), pw = col(p, ld, i);
  int k0, k1;
  row_chunk<NCH>(Q, k0, k1);
#pragma unroll 4
  for (int k = k0; k < k1; ++k) P.S.qs[k] = fma(sdt, pw[k], qw[k]), P.S.mu[k] = 0.0;
}

// one solver iteration after its sweep: |c|, direction, update, new |dq|, and whether the
// chain iterates again (the loop condition of systems.py:215-227 for the NEXT pass);
// observation-chunked
template <class M, int SOLVER, int NCH>
__global__ void __launch_bounds__(32 * NCH)
    kp_update(PhasedLayout L, int64_t n_act, Opts o) {
  using S = Stream<M>;
  constexpr int U = M::U;
  constexpr bool QN = SOLVER == MLB_SOLVER_QUASI_NEWTON;
  constexpr int NA = QN ? U : U * U + U;  // [mat U x U,] rhs U
  extern __shared__ double sh[];
  const int Y = L.Y;
  const ChunkCtx cx = chunk_ctx_listed<NCH>(L.list(), n_act, Y);
  const int64_t i = cx.i;
  const Phased P = L.at(i);
  const bool live = cx.in_range && P.I(CT_ACTIVE, i);
  const Col q = P.S.qs, mu = P.S.mu, c = P.S.c, jp = P.S.Jc;
  const Col jl = QN ? P.S.Jc : P.S.Jt;
  // pass 1: |c|, capacitance matrix of (J_l, J_prev) and right-hand side over the chunk
  double acc[NA], eacc = 0.0;
#pragma unroll
  for (int r = 0; r < NA; ++r) acc[r] = 0.0;
  if (live) {
#pragma unroll 2
    for (int k = cx.k0; k < cx.k1; ++k) {
      const double ck = c[k];
      eacc = norm_accum(eacc, ck, o.norm);
      const double rd = rcp_fast(jl[U * Y + k] * jp[U * Y + k]), wr = ck * rd;
      if (QN) {
#pragma unroll
        for (int a = 0; a < U; ++a) acc[a] = fma(jp[k * U + a], wr, acc[a]);
      } else {
        double rl[U];
#pragma unroll
        for (int b = 0; b < U; ++b) rl[b] = jl[k * U + b] * rd;
#pragma unroll
        for (int a = 0; a < U; ++a) {
          const double ra = jp[k * U + a];
          acc[U * U + a] = fma(ra, wr, acc[U * U + a]);
#pragma unroll
          for (int b = 0; b < U; ++b) acc[a * U + b] = fma(ra, rl[b], acc[a * U + b]);
        }
      }
    }
  }
  cta_sum_to_warp0<NA, NCH>(acc, sh);
  eacc = cta_norm_to_warp0<NCH>(eacc, o.norm, sh);
  double s[U];
#pragma unroll
  for (int a = 0; a < U; ++a) s[a] = 0.0;
  if (live && cx.wj == 0) {
    if (QN) {
      typename S::Chol Cp;
      S::load_chol(P.S.Gc, Cp);
#pragma unroll
      for (int a = 0; a < U; ++a) s[a] = acc[a];
      chol_solve<U>(Cp.L, Cp.rdiag, s);
    } else {
      double mat[U][U];
#pragma unroll
      for (int a = 0; a < U; ++a) {
        s[a] = acc[U * U + a];
#pragma unroll
        for (int b = 0; b < U; ++b) mat[a][b] = acc[a * U + b] + (a == b ? 1.0 : 0.0);
      }
      if (!lu_solve<U>(mat, s)) {
#pragma unroll
        for (int a = 0; a < U; ++a) s[a] = nan("");
      }
    }
  }
  cta_bcast_from_warp0<U>(s, sh);
  // pass 2: x = (c - A_l s) / D, update of the noise coordinates, J_prev^T x latent part
  double du[U], nacc = 0.0;
#pragma unroll
  for (int a = 0; a < U; ++a) du[a] = 0.0;
  if (live) {
#pragma unroll 4
    for (int k = cx.k0; k < cx.k1; ++k) {
      const double x = S::inv_jacprod_x(Y, k, jl, jp, c[k], s);
#pragma unroll
      for (int a = 0; a < U; ++a) du[a] = fma(jp[k * U + a], x, du[a]);
      const double dni = jp[U * Y + k] * x;
      q[U + k] -= dni, mu[U + k] += dni;
      nacc = norm_accum(nacc, dni, o.norm);
    }
  }
  cta_sum_to_warp0<U, NCH>(du, sh);
  nacc = cta_norm_to_warp0<NCH>(nacc, o.norm, sh);
  if (live && cx.wj == 0) {
#pragma unroll
    for (int a = 0; a < U; ++a) {
      q[a] -= du[a];
      mu[a] += du[a];
      nacc = norm_accum(nacc, du[a], o.norm);
    }
    const double ndq = norm_finish(nacc, o.norm), err = norm_finish(eacc, o.norm);
    const int iters = P.I(CT_ITERS, i) + 1;
    P.I(CT_ITERS, i) = iters, P.Dv(CD_NDQ, i) = ndq, P.Dv(CD_ERR, i) = err;
    P.I(CT_ACTIVE, i) = keep_iterating(iters, ndq, err, o);
  }
}

// close a solve: mu / dt, classification (solvers.py:81-95), then either accept the
// retraction (forward) or run the reversibility check (reverse); failures restore the
// last committed state
template <class M, int NCH>
__global__ void __launch_bounds__(32 * NCH)
    kp_solve_end(PhasedLayout L, int64_t n, int64_t ld, double* q, double* p, double dt,
                 const double* dt_pc, double sign, Opts o, int forward) {
  __shared__ double sh[NCH * 32];
  const int Y = L.Y, Q = L.U + Y;
  const ChunkCtx cx = chunk_ctx<NCH>(n, Y);
  const int64_t i = cx.in_range ? cx.i : 0;
  const Phased P = L.at(i);
  const bool live = cx.in_range && P.I(CT_ALIVE, i);
  const Col qw = col(q, ld, i), pw = col(p, ld, i);
  int k0, k1;
  row_chunk<NCH>(Q, k0, k1);
  int st = live ? classify(P.Dv(CD_ERR, i), P.Dv(CD_NDQ, i), o) : MLB_ST_OK;
  double acc = 0.0;
  if (live && st == MLB_ST_OK && !forward) {
#pragma unroll 4
    for (int k = k0; k < k1; ++k) acc = norm_accum(acc, P.S.qs[k] - P.S.cq[k], o.rev_norm);
  }
  // reversibility check: combine the chunks' norms and publish the verdict to every warp
  acc = cta_norm_to_warp0<NCH>(acc, o.rev_norm, sh);
  if (cx.wj == 0) {
    if (live && st == MLB_ST_OK && !forward && norm_finish(acc, o.rev_norm) > o.rev_tol)
      st = MLB_ST_NON_REVERSIBLE;
    sh[cx.lane] = (double)st;
  }
  __syncthreads();
  st = (int)sh[cx.lane];
  if (!live) return;
  if (st == MLB_ST_OK) {
    if (forward) {
      const double rdt = 1.0 / (sign * (dt_pc ? dt_pc[i] : dt));
#pragma unroll 4
      for (int k = k0; k < k1; ++k) pw[k] = fma(-rdt, P.S.mu[k], pw[k]), qw[k] = P.S.qs[k];
    }
  } else {
#pragma unroll 4
    for (int k = k0; k < k1; ++k) qw[k] = P.S.cq[k], pw[k] = P.S.cp[k];
  }
  // (every warp has read the control block before the barriers above)
  if (cx.wj == 0) {
    P.I(forward ? CT_TF : CT_TR, i) += P.I(CT_ITERS, i);
    if (st != MLB_ST_OK) P.I(CT_STATUS, i) = st, P.I(CT_ALIVE, i) = 0;
  }
}

template <class M>
__global__ void __launch_bounds__(kStreamBlock)
    kp_finish(PhasedLayout L, int64_t n, int32_t* status, int32_t* n_done, int32_t* iters_fwd,
              int32_t* iters_rev, double* h_init, double* h_final) {
  MLB_PH_INDEX();
  status[i] = P.I(CT_STATUS, i);
  if (n_done) n_done[i] = P.I(CT_DONE, i);
  if (iters_fwd) iters_fwd[i] = P.I(CT_TF, i);
  if (iters_rev) iters_rev[i] = P.I(CT_TR, i);
  if (h_init) h_init[i] = P.Dv(CD_HINIT, i);
  if (h_final) h_final[i] = P.Dv(CD_HLAST, i);
}

// ------------------------------------------------------------------------ host driver
template <class M, int SOLVER>
int phased_leapfrog_steps_overlapped(const mlb_model* m, const OdeData& D, int64_t n,
                                     int64_t ld, const Opts& o, const StepsIO& io,
                                     cudaStream_t s);
inline bool ode_overlap_enabled();

template <class M, int SOLVER>
int phased_leapfrog_steps(const mlb_model* m, const OdeData& D, int64_t n, int64_t ld,
                          const Opts& o, const StepsIO& io, cudaStream_t s) {
  // plain trajectories of two steps or more: reversibility checks one step behind (below)
  if (io.n_step >= 2 && !io.project_first && !io.alive_mask && io.max_delta_h <= 0.0 &&
      ode_overlap_enabled())
    return phased_leapfrog_steps_overlapped<M, SOLVER>(m, D, n, ld, o, io, s);
  constexpr int U = M::U;
  PhasedLayout L;
  L.lds = pad32(n), L.U = U, L.Y = D.Y, L.NGND = M::NG * M::ND;
  ScratchBuf sb;
  int rc = sb.alloc(L.rows_total(), L.lds, s);
  if (rc != MLB_OK) return rc;
  L.base = sb.p;
  int* counter = reinterpret_cast<int*>(sb.p + (L.rows_total() - 1) * L.lds);
  // pinned word the solver loop polls; owned by the model handle, so host-driven calls on
  // one handle are serialised
  std::lock_guard<std::mutex> lock(m->pin_mutex);
  if (!m->pinned_word && cudaMallocHost((void**)&m->pinned_word, sizeof(int)) != cudaSuccess) {
    g_err = "pinned counter allocation failed";
    return MLB_ERR_CUDA;
  }
  int* host_count = m->pinned_word;
  const dim3 blk(kStreamBlock), g1(sgrid(n));
  constexpr int NCH = kChunkWarps;
  const dim3 cblk(32, NCH);
  // dynamic shared memory of the chunked kernels: [rows][32] doubles
  constexpr size_t sh_eval = sizeof(double) * 32 * (U * U + 1);
  constexpr size_t sh_proj = sizeof(double) * 32 * (U + 1 > NCH ? U + 1 : NCH);
  constexpr size_t sh_upd = sizeof(double) * 32 * (U * U + U > NCH ? U * U + U : NCH);
  auto launched = [&]() { return finish_launch(); };
#define MLB_KP(call)                    \
  do {                                  \
    call;                               \
    if ((rc = launched()) != MLB_OK) return rc; \
  } while (0)
  auto read_counter = [&]() -> int {
    cudaMemcpyAsync(host_count, counter, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaMemsetAsync(counter, 0, sizeof(int), s);
    if (cudaStreamSynchronize(s) != cudaSuccess) return -1;
    return *host_count;
  };
  cudaMemsetAsync(counter, 0, sizeof(int), s);
  MLB_KP((kp_begin<M, NCH><<<g1, cblk, 0, s>>>(L, n, ld, io.q, io.p, io.alive_mask)));
  bool first = true;
  for (int step = 0; step < io.n_step || first; ++step) {
    for (int sub = 0; sub < 3; ++sub) {
      const bool fresh = first || sub == 1;
      if (fresh) {  // evaluate at the working position
        if ((rc = launch_jacob<M, true>(D, L, n, ld, io.q, s)) != MLB_OK) return rc;
        MLB_KP((kp_eval_mid<M, NCH><<<g1, cblk, sh_eval, s>>>(D, L, n, ld, io.q, io.p)));
        if ((rc = launch_second_order<M>(D, L, n, ld, io.q, s)) != MLB_OK) return rc;
      }
      bool fresh_g = fresh;  // the second-order partials are added to the gradient once
      if (first && io.project_first) {  // momentum refresh: project, then h_init
        MLB_KP((kp_kick_project<M, NCH><<<g1, cblk, sh_proj, s>>>(
            L, n, ld, io.p, io.q, io.dt, io.dt_pc, 1, 0, 0, 2, 0.0)));
        first = false, fresh_g = false;
        if (io.n_step <= 0) break;
      }
      if (io.n_step <= 0) {  // only h_init is wanted
        MLB_KP((kp_kick_project<M, NCH><<<g1, cblk, sh_proj, s>>>(L, n, ld, io.p, io.q, io.dt, io.dt_pc,
                                                      1, 0, 1, 0, 0.0)));
        first = false;
        break;
      }
      // note: with first == true the Hamiltonian is recorded BEFORE the kick
      MLB_KP((kp_kick_project<M, NCH><<<g1, cblk, sh_proj, s>>>(
          L, n, ld, io.p, io.q, io.dt, io.dt_pc, fresh_g, sub != 1, first, sub == 2,
          io.max_delta_h)));
      first = false;
      if (sub == 2) break;
      const double sign = sub == 0 ? 1.0 : -1.0;
      MLB_KP((kp_solve_begin<M, NCH><<<g1, cblk, 0, s>>>(L, n, ld, io.q, io.p, io.dt, io.dt_pc, sign,
                                                   o.max_iters)));
      MLB_KP((kp_compact<M><<<1, 1024, 0, s>>>(L, n, counter)));
      int active = read_counter();
      while (active > 0) {  // sweeps over the still-iterating chains only
        const dim3 ga(sgrid(active));
        if (SOLVER == MLB_SOLVER_QUASI_NEWTON)
          MLB_KP((kp_constr<M><<<ga, blk, 0, s>>>(D, L, active)));
        else
          { if ((rc = launch_jacob<M, false>(D, L, active, ld, io.q, s)) != MLB_OK) return rc; }
        MLB_KP((kp_update<M, SOLVER, NCH><<<ga, cblk, sh_upd, s>>>(L, active, o)));
        MLB_KP((kp_compact<M><<<1, 1024, 0, s>>>(L, n, counter)));
        active = read_counter();
      }
      if (active < 0) {
        g_err = "stream synchronisation failed in the solver loop";
        return MLB_ERR_CUDA;
      }
      MLB_KP((kp_solve_end<M, NCH><<<g1, cblk, 0, s>>>(L, n, ld, io.q, io.p, io.dt, io.dt_pc, sign, o,
                                                 sub == 0)));
    }
    if (io.n_step <= 0) break;
  }
  MLB_KP((kp_finish<M><<<g1, blk, 0, s>>>(L, n, io.status, io.n_done, io.iters_fwd,
                                          io.iters_rev, io.h_init, io.h_final)));
#undef MLB_KP
  return MLB_OK;
}

// ---------------------------------------------- reversibility check one step behind
// The reverse solve of step s only yields a verdict; nothing of step s + 1 needs its
// result (mlb_duo.cuh does the same inside one kernel for the register-resident family).
// Here every chain gets a TWIN column (columns [n, 2 n) of a scratch allocation twice as
// wide): at the end of step s the twin receives a copy of what the check needs - working
// (q, p) projected at the new point, Jacobian blocks / Gram factor there, the committed
// state of the start of the step - and its reverse solve then runs in the SAME solver
// rounds as the chain's forward solve of step s + 1: one compaction, one Jacobian sweep, one
// update kernel over the virtual chains of both.  The sweeps of small batches cost the same
// whatever they carry (1.75 ms for 64 ... 3,232 Hodgkin-Huxley chains), so a step costs
// max(forward, reverse) rounds instead of their sum.  The chain commits step s
// optimistically; a failed verdict rolls it back to the state its twin saved (and stops its
// speculative forward solve before its iterations are counted), so states, statuses,
// completed steps and iteration counts are those of the sequential integrator.  All
// kernels are the ones above, launched on three views of the same layout (chains, twins,
// both); two small kernels copy a chain to its twin and apply a verdict.
template <int NCH>
__global__ void __launch_bounds__(32 * NCH)
    kp_twin_copy(PhasedLayout Lm, PhasedLayout Lt, int64_t n, int64_t ld, const double* q,
                 const double* p, double* tq, double* tp, int64_t tld, int gram_rows,
                 int jac_rows) {
  const int64_t i = (int64_t)blockIdx.x * 32 + threadIdx.x;
  if (i >= n) return;
  const Phased P = Lm.at(i), T = Lt.at(i);
  const int alive = P.I(CT_ALIVE, i);
  if (threadIdx.y == 0) {
    T.I(CT_ALIVE, i) = alive, T.I(CT_STATUS, i) = MLB_ST_OK;
    T.I(CT_DONE, i) = P.I(CT_DONE, i), T.Dv(CD_HLAST, i) = P.Dv(CD_HLAST, i);
  }
  if (!alive) return;
  const int Q = Lm.U + Lm.Y;
  const Col qw = col(q, ld, i), pw = col(p, ld, i), tqw = col(tq, tld, i), tpw = col(tp, tld, i);
  for (int k = threadIdx.y; k < Q; k += NCH) {
    tqw[k] = qw[k], tpw[k] = pw[k];
    T.S.cq[k] = P.S.cq[k], T.S.cp[k] = P.S.cp[k];
  }
  for (int k = threadIdx.y; k < jac_rows; k += NCH) T.S.Jc[k] = P.S.Jc[k];
  for (int k = threadIdx.y; k < gram_rows; k += NCH) T.S.Gc[k] = P.S.Gc[k];
}

// verdict of the twin's reverse solve + check (kp_solve_end on the twin view has run):
// failure -> the chain dies at the state the twin saved; the twin is retired either way
template <int NCH>
__global__ void __launch_bounds__(32 * NCH)
    kp_apply_verdict(PhasedLayout Lm, PhasedLayout Lt, int64_t n, int64_t ld, double* q,
                     double* p) {
  const int64_t i = (int64_t)blockIdx.x * 32 + threadIdx.x;
  if (i >= n) return;
  const Phased P = Lm.at(i), T = Lt.at(i);
  const int tst = T.I(CT_STATUS, i);
  const bool failed = P.I(CT_ALIVE, i) && tst != MLB_ST_OK;
  if (failed) {
    const int Q = Lm.U + Lm.Y;
    const Col qw = col(q, ld, i), pw = col(p, ld, i);
    for (int k = threadIdx.y; k < Q; k += NCH) {
      const double a = T.S.cq[k], b = T.S.cp[k];
      qw[k] = a, pw[k] = b, P.S.cq[k] = a, P.S.cp[k] = b;
    }
  }
  __syncthreads();  // every warp has read the chain's flags
  if (threadIdx.y != 0) return;
  if (failed) {
    P.I(CT_STATUS, i) = tst, P.I(CT_ALIVE, i) = 0, P.I(CT_ACTIVE, i) = 0;
    P.I(CT_DONE, i) = T.I(CT_DONE, i), P.Dv(CD_HLAST, i) = T.Dv(CD_HLAST, i);
  }
  T.I(CT_ALIVE, i) = 0, T.I(CT_ACTIVE, i) = 0;
}

template <class M>
__global__ void __launch_bounds__(kStreamBlock)
    kp_merge_twin_counts(PhasedLayout Lm, PhasedLayout Lt, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Lm.at(i).I(CT_TR, i) += Lt.at(i).I(CT_TR, i);
}

inline bool ode_overlap_enabled() {
  static const bool v = [] {
    const char* e = std::getenv("MLB_ODE_OVERLAP");
    return !e || std::atoi(e) != 0;
  }();
  return v;
}

template <class M, int SOLVER>
int phased_leapfrog_steps_overlapped(const mlb_model* m, const OdeData& D, int64_t n,
                                     int64_t ld, const Opts& o, const StepsIO& io,
                                     cudaStream_t s) {
  constexpr int U = M::U;
  constexpr int NCH = kChunkWarps;
  const int Q = U + D.Y;
  PhasedLayout Lf;  // both: chains in columns [0, n), twins in [n, 2 n)
  Lf.lds = pad32(2 * n), Lf.U = U, Lf.Y = D.Y, Lf.NGND = M::NG * M::ND;
  ScratchBuf sb;
  int rc = sb.alloc(Lf.rows_total() + 2 * (int64_t)Q, Lf.lds, s);
  if (rc != MLB_OK) return rc;
  Lf.base = sb.p;
  PhasedLayout Lm = Lf, Lt = Lf;
  Lt.col0 = n;
  double* const tq = sb.p + Lf.rows_total() * Lf.lds + n;  // twins' working (q, p)
  double* const tp = tq + (int64_t)Q * Lf.lds;
  int* counter = reinterpret_cast<int*>(sb.p + (Lf.rows_total() - 1) * Lf.lds);
  std::lock_guard<std::mutex> lock(m->pin_mutex);
  if (!m->pinned_word && cudaMallocHost((void**)&m->pinned_word, sizeof(int)) != cudaSuccess) {
    g_err = "pinned counter allocation failed";
    return MLB_ERR_CUDA;
  }
  int* host_count = m->pinned_word;
  const dim3 blk(kStreamBlock), g1(sgrid(n));
  const dim3 cblk(32, NCH);
  constexpr size_t sh_eval = sizeof(double) * 32 * (U * U + 1);
  constexpr size_t sh_proj = sizeof(double) * 32 * (U + 1 > NCH ? U + 1 : NCH);
  constexpr size_t sh_upd = sizeof(double) * 32 * (U * U + U > NCH ? U * U + U : NCH);
#define MLB_KP(call)                                  \
  do {                                                \
    call;                                             \
    if ((rc = finish_launch()) != MLB_OK) return rc;  \
  } while (0)
  auto read_counter = [&]() -> int {
    cudaMemcpyAsync(host_count, counter, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaMemsetAsync(counter, 0, sizeof(int), s);
    if (cudaStreamSynchronize(s) != cudaSuccess) return -1;
    return *host_count;
  };
  // solver rounds over the still-iterating virtual chains of BOTH views
  auto rounds = [&]() -> int {
    MLB_KP((kp_compact<M><<<1, 1024, 0, s>>>(Lf, 2 * n, counter)));
    int active = read_counter();
    while (active > 0) {
      const dim3 ga(sgrid(active));
      if (SOLVER == MLB_SOLVER_QUASI_NEWTON)
        MLB_KP((kp_constr<M><<<ga, blk, 0, s>>>(D, Lf, active)));
      else if ((rc = launch_jacob<M, false>(D, Lf, active, ld, io.q, s)) != MLB_OK)
        return rc;
      MLB_KP((kp_update<M, SOLVER, NCH><<<ga, cblk, sh_upd, s>>>(Lf, active, o)));
      MLB_KP((kp_compact<M><<<1, 1024, 0, s>>>(Lf, 2 * n, counter)));
      active = read_counter();
    }
    if (active < 0) {
      g_err = "stream synchronisation failed in the solver loop";
      return MLB_ERR_CUDA;
    }
    return MLB_OK;
  };
  auto verdict = [&]() -> int {  // close the twins' reverse solve and apply it
    MLB_KP((kp_solve_end<M, NCH><<<g1, cblk, 0, s>>>(Lt, n, Lf.lds, tq, tp, io.dt, io.dt_pc, -1.0,
                                                     o, 0)));
    MLB_KP((kp_apply_verdict<NCH><<<g1, cblk, 0, s>>>(Lm, Lt, n, ld, io.q, io.p)));
    return MLB_OK;
  };
  cudaMemsetAsync(counter, 0, sizeof(int), s);
  MLB_KP((kp_begin<M, NCH><<<g1, cblk, 0, s>>>(Lm, n, ld, io.q, io.p, nullptr)));
  // twins start retired (control block zero: not alive, not active, status OK, counts 0)
  {
    const Phased T0 = Lt.at(0);
    cudaMemset2DAsync(T0.ct, sizeof(int32_t) * Lf.lds, 0, sizeof(int32_t) * n, CT_INT_ROWS, s);
  }
  if ((rc = launch_jacob<M, true>(D, Lm, n, ld, io.q, s)) != MLB_OK) return rc;
  MLB_KP((kp_eval_mid<M, NCH><<<g1, cblk, sh_eval, s>>>(D, Lm, n, ld, io.q, io.p)));
  if ((rc = launch_second_order<M>(D, Lm, n, ld, io.q, s)) != MLB_OK) return rc;
  for (int step = 0; step < io.n_step; ++step) {
    const int first = step == 0;
    // opening half kick + projection, forward retraction start
    MLB_KP((kp_kick_project<M, NCH><<<g1, cblk, sh_proj, s>>>(Lm, n, ld, io.p, io.q, io.dt,
                                                             io.dt_pc, first, 1, first, 0, 0.0)));
    MLB_KP((kp_solve_begin<M, NCH><<<g1, cblk, 0, s>>>(Lm, n, ld, io.q, io.p, io.dt, io.dt_pc, 1.0,
                                                       o.max_iters)));
    // forward solve of this step together with the reverse solve of the previous one
    if ((rc = rounds()) != MLB_OK) return rc;
    if (step > 0 && (rc = verdict()) != MLB_OK) return rc;
    MLB_KP((kp_solve_end<M, NCH><<<g1, cblk, 0, s>>>(Lm, n, ld, io.q, io.p, io.dt, io.dt_pc, 1.0, o,
                                                     1)));
    // evaluation at the new position, projection of the momentum there
    if ((rc = launch_jacob<M, true>(D, Lm, n, ld, io.q, s)) != MLB_OK) return rc;
    MLB_KP((kp_eval_mid<M, NCH><<<g1, cblk, sh_eval, s>>>(D, Lm, n, ld, io.q, io.p)));
    if ((rc = launch_second_order<M>(D, Lm, n, ld, io.q, s)) != MLB_OK) return rc;
    MLB_KP((kp_kick_project<M, NCH><<<g1, cblk, sh_proj, s>>>(Lm, n, ld, io.p, io.q, io.dt,
                                                             io.dt_pc, 1, 0, 0, 0, 0.0)));
    // hand the check of this step to the twins, then close the step optimistically
    MLB_KP((kp_twin_copy<NCH><<<g1, cblk, 0, s>>>(Lm, Lt, n, ld, io.q, io.p, tq, tp, Lf.lds,
                                                  U * U + D.Y, U * D.Y + D.Y)));
    MLB_KP((kp_solve_begin<M, NCH><<<g1, cblk, 0, s>>>(Lt, n, Lf.lds, tq, tp, io.dt, io.dt_pc, -1.0,
                                                       o.max_iters)));
    MLB_KP((kp_kick_project<M, NCH><<<g1, cblk, sh_proj, s>>>(Lm, n, ld, io.p, io.q, io.dt,
                                                             io.dt_pc, 0, 1, 0, 1, 0.0)));
  }
  if ((rc = rounds()) != MLB_OK) return rc;  // the last step's check runs alone
  if ((rc = verdict()) != MLB_OK) return rc;
  MLB_KP((kp_merge_twin_counts<M><<<g1, blk, 0, s>>>(Lm, Lt, n)));
  MLB_KP((kp_finish<M><<<g1, blk, 0, s>>>(Lm, n, io.status, io.n_done, io.iters_fwd,
                                          io.iters_rev, io.h_init, io.h_final)));
#undef MLB_KP
  return MLB_OK;
}

// ------------------------------------------------------ host-driven static HMC sampler
// One transition = kp_smp_begin (momentum draw, proposal copy, per-chain step size) ->
// phased_leapfrog_steps on the proposal (momentum refresh projected first, |dh| test per
// step) -> kp_smp_end (Metropolis accept, statistics, trace, dual averaging): the
// semantics of ks_hmc_sample (mlb_stream_launch.cuh), with the trajectory on the
// sub-phase kernels.  Per-chain rows after the proposal: dt, h_init, h_final (doubles),
// status, n_done, iters_fwd, iters_rev (int32).
struct SmpLayout {
  double* base;  // [2 Q + 3 + 2][lds]
  int64_t lds;
  int Q;
  __host__ __device__ double* qc() const { return base; }
  __host__ __device__ double* pc() const { return base + (int64_t)Q * lds; }
  __host__ __device__ double* dt() const { return base + (int64_t)2 * Q * lds; }
  __host__ __device__ double* h0() const { return dt() + lds; }
  __host__ __device__ double* h1() const { return dt() + 2 * lds; }
  __host__ __device__ int32_t* ints() const { return reinterpret_cast<int32_t*>(dt() + 3 * lds); }
  __host__ __device__ int32_t* status() const { return ints(); }
  __host__ __device__ int32_t* n_done() const { return ints() + lds; }
  __host__ __device__ int32_t* itf() const { return ints() + 2 * lds; }
  __host__ __device__ int32_t* itr() const { return ints() + 3 * lds; }
  static int64_t rows(int Q) { return 2 * (int64_t)Q + 3 + 2; }
};

template <int NCH>
__global__ void __launch_bounds__(32 * NCH)
    kp_smp_begin(SmpLayout W, int64_t n, int64_t ld, const double* q, SampleArgs a, int it) {
  const int64_t i = (int64_t)blockIdx.x * 32 + threadIdx.x;
  if (i >= n) return;
  const int Q = W.Q;
  int k0, k1;
  row_chunk<NCH>(Q, k0, k1);
  double* qc = W.qc() + i;
  double* pc = W.pc() + i;
  const double* mn = a.mom_noise ? a.mom_noise + (int64_t)it * Q * ld + i : nullptr;
#pragma unroll 1
  for (int k = k0; k < k1; ++k) {
    qc[(int64_t)k * W.lds] = q[(int64_t)k * ld + i];
    double v;
    if (mn) {
      v = mn[(int64_t)k * ld];
    } else {
      double n0, n1;
      philox_normal_pair(a.seed, a.chain0 + (uint64_t)i, a.iter0 + (uint32_t)it,
                         (uint32_t)(k >> 1), n0, n1);
      v = (k & 1) ? n1 : n0;
    }
    pc[(int64_t)k * W.lds] = v;
  }
  if (threadIdx.y == 0) W.dt()[i] = a.adapt ? a.adapt[4 * ld + i] : a.dt;
}

template <int NCH>
__global__ void __launch_bounds__(32 * NCH)
    kp_smp_end(SmpLayout W, int64_t n, int64_t ld, double* q, SampleArgs a, int it) {
  const int64_t i = (int64_t)blockIdx.x * 32 + threadIdx.x;
  if (i >= n) return;
  const int Q = W.Q;
  const int st = W.status()[i];
  double a_stat = 0.0;
  bool accept = false;
  if (st == MLB_ST_OK) {
    const double d = W.h0()[i] - W.h1()[i];
    a_stat = (d != d) ? 0.0 : exp(d < 0.0 ? d : 0.0);
    const double u = a.unif ? a.unif[(int64_t)it * ld + i]
                            : philox_uniform(a.seed, a.chain0 + (uint64_t)i,
                                             a.iter0 + (uint32_t)it);
    accept = u < a_stat;
  }
  int k0, k1;
  row_chunk<NCH>(Q, k0, k1);
  if (accept) {
#pragma unroll 4
    for (int k = k0; k < k1; ++k) q[(int64_t)k * ld + i] = W.qc()[(int64_t)k * W.lds + i];
  }
  if (a.trace_q && (it + 1) % a.trace_every == 0) {
    double* tq = a.trace_q + (int64_t)((it + 1) / a.trace_every - 1) * a.trace_dim * ld + i;
    for (int k = k0; k < k1 && k < a.trace_dim; ++k)
      tq[(int64_t)k * ld] = accept ? W.qc()[(int64_t)k * W.lds + i] : q[(int64_t)k * ld + i];
  }
  if (threadIdx.y != 0) return;
  if (a.trace_accept) a.trace_accept[(int64_t)it * ld + i] = a_stat;
  if (a.stats) {
    a.stats[0 * ld + i] += a_stat, a.stats[1 * ld + i] += accept;
    a.stats[2 * ld + i] +=
        (st == MLB_ST_DIVERGED || st == MLB_ST_NOT_CONVERGED || st == MLB_ST_LINALG);
    a.stats[3 * ld + i] += (st == MLB_ST_NON_REVERSIBLE);
    a.stats[4 * ld + i] += (st == MLB_ST_H_DIVERGED);
    a.stats[5 * ld + i] += W.n_done()[i];
    a.stats[6 * ld + i] += W.itf()[i] + W.itr()[i];
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
int phased_hmc_sample(const mlb_model* m, const OdeData& D, int64_t n, int64_t ld, double* q,
                      const Opts& o, const SampleArgs& a, cudaStream_t s) {
  constexpr int NCH = kChunkWarps;
  SmpLayout W;
  W.lds = pad32(n), W.Q = M::U + D.Y;
  ScratchBuf sb;
  int rc = sb.alloc(SmpLayout::rows(W.Q), W.lds, s);
  if (rc != MLB_OK) return rc;
  W.base = sb.p;
  const dim3 g1(sgrid(n)), cblk(32, NCH);
  for (int it = 0; it < a.n_iter; ++it) {
    kp_smp_begin<NCH><<<g1, cblk, 0, s>>>(W, n, ld, q, a, it);
    if ((rc = finish_launch()) != MLB_OK) return rc;
    StepsIO io{W.qc(), W.pc(), 0.0, W.dt(), a.n_step, W.status(), W.n_done(), W.itf(),
               W.itr(), W.h0(), W.h1()};
    io.project_first = true, io.max_delta_h = a.max_delta_h;
    if ((rc = phased_leapfrog_steps<M, SOLVER>(m, D, n, W.lds, o, io, s)) != MLB_OK) return rc;
    kp_smp_end<NCH><<<g1, cblk, 0, s>>>(W, n, ld, q, a, it);
    if ((rc = finish_launch()) != MLB_OK) return rc;
  }
  return MLB_OK;
}

}  // namespace mlb
