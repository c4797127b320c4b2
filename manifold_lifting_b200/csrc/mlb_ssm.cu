// "Tridiagonal + low rank" Gram algebra of the state-space systems (reference
// mlift/systems.py:1161-1243, PartiallyInvertibleStateSpaceModelSystem), batched over chains.
//
// The constraint Jacobian of a scalar state-space model has blocks
//   dc_du [Y x U] dense, dc_dv [Y] diagonal, dc_dn = (dc_dn0 [Y] diagonal, dc_dn1 [Y-1] first
//   sub-diagonal), dx_dy [Y] (enters the log-determinant only),
// so the Gram matrix is T + dc_du dc_du^T with T symmetric tridiagonal, and every product
// with its inverse is two Thomas solves around a U x U capacitance system (Woodbury).
//
// Execution model: one chain per thread, every array component-major ([rows][ld], element
// (k, chain i) at p[k * ld + i]), so a warp's access to one row of 32 chains is one
// coalesced 256-byte transaction.  Everything that is a SUM over observations (capacitance
// matrix, dc_du^T T^{-1} vct) is accumulated in the forward elimination sweep alone (see
// k_ssm_decompose); only the final solve needs the backward sweep, whose inputs (1 / b', the
// multipliers, the eliminated right-hand side) go through a per-call scratch area / the
// output array in the same layout.  The U x U capacitance matrix, its factor and the
// U-vectors live in registers.  The sweeps'
// loads do not depend on computed values, so each thread fetches kP rows ahead into
// registers and then runs kP dependent steps (one DRAM round trip per kP steps).
// One IEEE division per observation (1 / b'); the reference's other divisions by b' are
// multiplications by that reciprocal (results agree with a NumPy restatement of the
// reference to ~1e-13 relative, tests/test_gpu_state_space.py).
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/mlb.h"
#include "mlb_launch.cuh"

namespace mlb {


struct SsmBlocks {
  const double *du, *dv, *n0, *n1, *dxdy;
};

MLB_DEV int clampi(int k, int lo, int hi) { return k < lo ? lo : (k > hi ? hi : k); }

template <int U>
struct PrefetchDepth {
  static constexpr int value = U <= 2 ? 8 : (U <= 4 ? 6 : 4);
};

// ---- systems.py:1189-1197  decompose_gram:
// a = dc_dn0[:-1] * dc_dn1, b = dc_dv^2 + dc_dn0^2 + pad(dc_dn1^2, (1, 0)),
// cap = I + dc_du^T T^{-1} dc_du (T = tridiag(a, b, a); the reference vmaps tridiagonal_solve
// over the columns of dc_du), chol(cap).
// The Thomas elimination is T = L D L^T (L unit lower bidiagonal with the multipliers w,
// D = diag(b')), so dc_du^T T^{-1} dc_du = sum_k d'_k d'_k^T / b'_k with d' = L^{-1} dc_du
// the forward-eliminated right-hand sides: ONE forward sweep, no backward substitution and
// no scratch -- the kernel reads the blocks once and writes a, b and the factor, i.e. it
// moves exactly its algorithmic bytes.
template <int U>
__global__ void __launch_bounds__(128)
    k_ssm_decompose(int64_t n, int Y, int64_t ld, SsmBlocks J, double* a_out, double* b_out,
                    double* chol_out) {
  constexpr int P = PrefetchDepth<U>::value;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double rbp = 0.0, a_prev = 0.0, n1_prev = 0.0, dp[U], cap[U][U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    dp[u] = 0.0;
#pragma unroll
    for (int v = 0; v < U; ++v) cap[u][v] = u == v ? 1.0 : 0.0;
  }
  for (int k0 = 0; k0 < Y; k0 += P) {
    double du[P][U], dv[P], n0[P], n1[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int k = clampi(k0 + j, 0, Y - 1);
#pragma unroll
      for (int u = 0; u < U; ++u) du[j][u] = J.du[(int64_t)(k * U + u) * ld + i];
      dv[j] = J.dv[(int64_t)k * ld + i], n0[j] = J.n0[(int64_t)k * ld + i];
      n1[j] = Y > 1 ? J.n1[(int64_t)clampi(k, 0, Y - 2) * ld + i] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int k = k0 + j;
      if (k < Y) {
        const double bk = dv[j] * dv[j] + n0[j] * n0[j] + n1_prev * n1_prev;
        const double ak = n0[j] * n1[j];  // meaningful for k < Y - 1
        b_out[(int64_t)k * ld + i] = bk;
        if (k < Y - 1) a_out[(int64_t)k * ld + i] = ak;
        const double w = a_prev * rbp;  // 0 at k = 0
        rbp = 1.0 / (bk - w * a_prev);
#pragma unroll
        for (int u = 0; u < U; ++u) dp[u] = du[j][u] - w * dp[u];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const double z = dp[u] * rbp;
#pragma unroll
          for (int v = 0; v <= u; ++v) cap[u][v] = fma(z, dp[v], cap[u][v]);
        }
        a_prev = ak, n1_prev = n1[j];
      }
    }
  }
  // the factorisation reads the lower triangle only
  double rdiag[U];
  const bool ok = chol_lower<U>(cap, rdiag);
#pragma unroll
  for (int r = 0; r < U; ++r)
#pragma unroll
    for (int c = 0; c < U; ++c)
      chol_out[(int64_t)(r * U + c) * ld + i] =
          ok ? cap[r][c] : __longlong_as_double(0x7ff8000000000000LL);
}

// shared tail of the two inverse products: out = T^{-1} (vct - du s) given the stored
// elimination (w, 1 / b') and the upper diagonal cu.
// Instruction count matters here: at 16,384 chains these sweeps run with one warp per
// scheduler and ncu shows them issue-bound (~95 instructions per row in the first version,
// DESIGN 3.3b).  Hence running pointers with loop-invariant row offsets instead of
// recomputed 64-bit indices, full blocks of P rows without clamps or predicates, and a
// scalar loop for the < P remaining rows.
template <int U, int P>
MLB_DEV void thomas_apply_lowrank(int Y, int64_t ld, int64_t i, const double* du_p,
                                  const double (&s)[U], const double* vct, const double* wv,
                                  const double* cu, const double* rb, double* out) {
  const int Yf = Y / P * P;
  double dp = 0.0;
  {
    const double *pdu = du_p + i, *pv = vct + i, *pw = wv + i;
    double* po = out + i;
    for (int k0 = 0; k0 < Yf; k0 += P) {
      double du[P][U], vv[P], ww[P];
#pragma unroll
      for (int j = 0; j < P; ++j) {
#pragma unroll
        for (int u = 0; u < U; ++u) du[j][u] = pdu[(int64_t)(j * U + u) * ld];
        vv[j] = pv[(int64_t)j * ld], ww[j] = pw[(int64_t)j * ld];
      }
#pragma unroll
      for (int j = 0; j < P; ++j) {
        double d = vv[j];
#pragma unroll
        for (int u = 0; u < U; ++u) d = fma(-du[j][u], s[u], d);
        dp = d - ww[j] * dp;
        po[(int64_t)j * ld] = dp;
      }
      pdu += (int64_t)P * U * ld, pv += (int64_t)P * ld, pw += (int64_t)P * ld;
      po += (int64_t)P * ld;
    }
    for (int k = Yf; k < Y; ++k) {
      double d = *pv;
#pragma unroll
      for (int u = 0; u < U; ++u) d = fma(-pdu[(int64_t)u * ld], s[u], d);
      dp = d - *pw * dp;
      *po = dp;
      pdu += (int64_t)U * ld, pv += ld, pw += ld, po += ld;
    }
  }
  // backward: rows Y-1 ... 0; the rows above the last full block first, one at a time
  double x = 0.0;
  double* po = out + (int64_t)(Y - 1) * ld + i;
  const double* pr = rb + (int64_t)(Y - 1) * ld + i;
  const double* pc = cu + (int64_t)(Y - 1) * ld + i;  // row Y-1 of cu is never read
  for (int k = Y - 1; k >= Yf; --k) {
    const double cv = k < Y - 1 ? *pc : 0.0;
    x = (*po - cv * x) * *pr;
    *po = x;
    po -= ld, pr -= ld, pc -= ld;
  }
  for (int k0 = Yf - 1; k0 >= 0; k0 -= P) {
    double dd[P], cv[P], rv[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      dd[j] = po[-(int64_t)j * ld], rv[j] = pr[-(int64_t)j * ld];
      // only the very first row handled overall (k = Y - 1) has no upper neighbour
      cv[j] = (Yf == Y && k0 == Y - 1 && j == 0) ? 0.0 : pc[-(int64_t)j * ld];
    }
#pragma unroll
    for (int j = 0; j < P; ++j) {
      x = (dd[j] - cv[j] * x) * rv[j];
      po[-(int64_t)j * ld] = x;
    }
    po -= (int64_t)P * ld, pr -= (int64_t)P * ld, pc -= (int64_t)P * ld;
  }
}

// ---- systems.py:1199-1210  lmult_by_inv_gram:
// T^{-1} (vct - dc_du cho_solve(chol_cap, dc_du^T T^{-1} vct)); with T = L D L^T the inner
// product is dc_du^T T^{-1} vct = sum_k d'_k v'_k / b'_k (d' = L^{-1} dc_du, v' = L^{-1} vct):
// one forward sweep, then the solve of the corrected right-hand side.
// scratch: rb [Y][ld], wv [Y][ld]
template <int U>
__global__ void __launch_bounds__(128)
    k_ssm_inv_gram(int64_t n, int Y, int64_t ld, const double* du_p, const double* a,
                   const double* b, const double* chol, const double* vct, double* out,
                   double* rb, double* wv) {
  constexpr int P = PrefetchDepth<U>::value;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double rbp = 0.0, vp = 0.0, dp[U], s[U];
#pragma unroll
  for (int u = 0; u < U; ++u) dp[u] = 0.0, s[u] = 0.0;
  for (int k0 = 0; k0 < Y; k0 += P) {
    double du[P][U], av[P], bv[P], vv[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int k = clampi(k0 + j, 0, Y - 1);
#pragma unroll
      for (int u = 0; u < U; ++u) du[j][u] = du_p[(int64_t)(k * U + u) * ld + i];
      av[j] = k > 0 ? a[(int64_t)(k - 1) * ld + i] : 0.0;
      bv[j] = b[(int64_t)k * ld + i], vv[j] = vct[(int64_t)k * ld + i];
    }
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int k = k0 + j;
      if (k < Y) {
        const double w = av[j] * rbp;
        rbp = 1.0 / (bv[j] - w * av[j]);
        vp = vv[j] - w * vp;
        rb[(int64_t)k * ld + i] = rbp, wv[(int64_t)k * ld + i] = w;
        const double z = vp * rbp;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          dp[u] = du[j][u] - w * dp[u];
          s[u] = fma(dp[u], z, s[u]);
        }
      }
    }
  }
  double L[U][U], rdiag[U];
#pragma unroll
  for (int r = 0; r < U; ++r) {
#pragma unroll
    for (int c = 0; c < U; ++c) L[r][c] = chol[(int64_t)(r * U + c) * ld + i];
    rdiag[r] = 1.0 / L[r][r];
  }
  chol_solve<U>(L, rdiag, s);
  thomas_apply_lowrank<U, P>(Y, ld, i, du_p, s, vct, wv, a, rb, out);
}

// ---- systems.py:1212-1232  lmult_by_inv_jacob_product: (J_l J_r^T)^{-1} vct with
// a = dn1_l * dn0_r[:-1], b = dv_l dv_r + dn0_l dn0_r + pad(dn1_l dn1_r, (1, 0)),
// c = dn0_l[:-1] * dn1_r, cap = I + du_r^T T^{-1} du_l (np.linalg.solve: pivoted LU),
// T^{-1} (vct - du_l cap^{-1} du_r^T T^{-1} vct).
// The elimination is T = L Ub (L unit lower bidiagonal, Ub upper bidiagonal with diagonal b'
// and super-diagonal c), so du_r^T T^{-1} X = (Ub^{-T} du_r)^T (L^{-1} X): with
// z_k = (du_r[k] - c_{k-1} z_{k-1}) / b'_k both the capacitance matrix and the inner product
// with vct are sums over ONE forward sweep.
// scratch: rb, wv, cu [Y][ld] each
// 65,536 chains are 512 CTAs: four resident CTAs per SM (<= 128 registers) keep them in ONE
// wave on 148 SMs; at 160 registers (three per SM) the launch ran 1.15 waves
template <int U>
__global__ void __launch_bounds__(128, U <= 3 ? 4 : 1)
    k_ssm_inv_jacob_product(int64_t n, int Y, int64_t ld, SsmBlocks Jl, SsmBlocks Jr,
                            const double* vct, double* out, double* rb, double* wv,
                            double* cu) {
  constexpr int P = U <= 2 ? 4 : (U <= 4 ? 3 : 2);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double rbp = 0.0, a_prev = 0.0, c_prev = 0.0, n1n1_prev = 0.0, tp = 0.0;
  double dp[U], z[U], cap[U][U], s[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    dp[u] = 0.0, z[u] = 0.0, s[u] = 0.0;
#pragma unroll
    for (int v = 0; v < U; ++v) cap[u][v] = u == v ? 1.0 : 0.0;
  }
  for (int k0 = 0; k0 < Y; k0 += P) {
    double dul[P][U], dur[P][U], dvl[P], dvr[P], n0l[P], n0r[P], n1l[P], n1r[P], vv[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int k = clampi(k0 + j, 0, Y - 1);
      const int k1 = clampi(k, 0, Y - 2 > 0 ? Y - 2 : 0);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        dul[j][u] = Jl.du[(int64_t)(k * U + u) * ld + i];
        dur[j][u] = Jr.du[(int64_t)(k * U + u) * ld + i];
      }
      dvl[j] = Jl.dv[(int64_t)k * ld + i], dvr[j] = Jr.dv[(int64_t)k * ld + i];
      n0l[j] = Jl.n0[(int64_t)k * ld + i], n0r[j] = Jr.n0[(int64_t)k * ld + i];
      n1l[j] = Y > 1 ? Jl.n1[(int64_t)k1 * ld + i] : 0.0;
      n1r[j] = Y > 1 ? Jr.n1[(int64_t)k1 * ld + i] : 0.0;
      vv[j] = vct[(int64_t)k * ld + i];
    }
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int k = k0 + j;
      if (k < Y) {
        const double bk = dvl[j] * dvr[j] + n0l[j] * n0r[j] + n1n1_prev;
        // a_k = dn1_l[k] dn0_r[k] (row k + 1, column k), c_k = dn0_l[k] dn1_r[k]
        const double ak = n1l[j] * n0r[j], ck = n0l[j] * n1r[j];
        const double w = a_prev * rbp;  // 0 at k = 0
        rbp = 1.0 / (bk - w * c_prev);
        rb[(int64_t)k * ld + i] = rbp, wv[(int64_t)k * ld + i] = w;
        if (k < Y - 1) cu[(int64_t)k * ld + i] = ck;
        tp = vv[j] - w * tp;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          dp[u] = dul[j][u] - w * dp[u];
          z[u] = (dur[j][u] - c_prev * z[u]) * rbp;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          s[u] = fma(z[u], tp, s[u]);
#pragma unroll
          for (int v = 0; v < U; ++v) cap[u][v] = fma(z[u], dp[v], cap[u][v]);
        }
        a_prev = ak, c_prev = ck, n1n1_prev = n1l[j] * n1r[j];
      }
    }
  }
  const bool ok = lu_solve<U>(cap, s);
  if (!ok) {
#pragma unroll
    for (int u = 0; u < U; ++u) s[u] = __longlong_as_double(0x7ff8000000000000LL);
  }
  thomas_apply_lowrank<U, 4>(Y, ld, i, Jl.du, s, vct, wv, cu, rb, out);
}

// ---- systems.py:1234-1243  log_det_sqrt_gram from the blocks and the decomposition:
// sum log diag(chol_cap) + tridiagonal_pos_def_log_det(a, b) / 2 - sum log|dx_dy|.
// The reference's log-continuant recursion (linalg.py:104-114) carries l_k = log theta_k of
// the leading minors through an exp and a log1p per observation -- a ~200-cycle dependent
// chain.  Its increments are the Thomas pivots, l_k - l_{k-1} = log b'_k, so here the
// dependent chain is the pivot recurrence alone (one reciprocal) and the logarithms are
// taken off it, of products of two pivots; the two forms agree to rounding
// (mlb_tridiagonal_pos_def_log_det keeps the reference's recursion term by term).
__global__ void __launch_bounds__(128)
    k_ssm_log_det(int64_t n, int Y, int U, int64_t ld, const double* a, const double* b,
                  const double* chol, const double* dxdy, double* out) {
  constexpr int P = 8;  // even: pivots are paired inside one prefetch block
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double rbp = 0.0, st = 0.0, sx = 0.0;
  for (int k0 = 0; k0 < Y; k0 += P) {
    double av[P], bv[P], xv[P], piv[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int k = clampi(k0 + j, 0, Y - 1);
      av[j] = k > 0 ? a[(int64_t)(k - 1) * ld + i] : 0.0;
      bv[j] = b[(int64_t)k * ld + i];
      xv[j] = dxdy[(int64_t)k * ld + i];
    }
#pragma unroll
    for (int j = 0; j < P; ++j) {
      // rows past the end count as pivot 1 and |dx_dy| 1
      const bool in = k0 + j < Y;
      const double w = av[j] * rbp;
      const double bp = bv[j] - w * av[j];
      if (in) rbp = 1.0 / bp;
      piv[j] = in ? bp : 1.0;
      xv[j] = in ? fabs(xv[j]) : 1.0;
    }
#pragma unroll
    for (int j = 0; j < P; j += 2) {
      // a non-positive pivot (matrix not positive definite) must give NaN, also in a pair
      const bool bad = piv[j] <= 0.0 || piv[j + 1] <= 0.0;
      // one logarithm per pair; a product that leaves the normal range (pivots or |dx_dy|
      // around 1e+-160) falls back to the sum of the two logarithms, which stays finite as
      // the reference's per-term recursion does (linalg.py:104-114)
      const double pp = piv[j] * piv[j + 1], xx = xv[j] * xv[j + 1];
      const bool pn = pp >= 2.3e-308 && pp <= 1.7e308, xn = xx >= 2.3e-308 && xx <= 1.7e308;
      st += bad ? log(-1.0) : (pn ? log(pp) : log(piv[j]) + log(piv[j + 1]));
      sx += xn ? log(xx) : log(xv[j]) + log(xv[j + 1]);
    }
  }
  double sc = 0.0;
  for (int u = 0; u < U; ++u) sc += log(chol[(int64_t)(u * U + u) * ld + i]);
  out[i] = sc + 0.5 * st - sx;
}

// ---- systems.py:1161-1175  lmult_by_jacob_constr: J vct, vct = (vct_u, vct_v, vct_n) [Q]
__global__ void __launch_bounds__(128)
    k_ssm_lmult_jacob(int64_t n, int Y, int U, int64_t ld, SsmBlocks J, const double* vct,
                      double* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (i >= n) return;
  const double* vv = vct + (int64_t)U * ld;
  const double* vn = vct + (int64_t)(U + Y) * ld;
  double s = 0.0;
  for (int u = 0; u < U; ++u) s = fma(J.du[(int64_t)(k * U + u) * ld + i], vct[(int64_t)u * ld + i], s);
  s += J.dv[(int64_t)k * ld + i] * vv[(int64_t)k * ld + i];
  s += J.n0[(int64_t)k * ld + i] * vn[(int64_t)k * ld + i];
  if (k > 0) s += J.n1[(int64_t)(k - 1) * ld + i] * vn[(int64_t)(k - 1) * ld + i];
  out[(int64_t)k * ld + i] = s;
}

// ---- systems.py:1177-1187  rmult_by_jacob_constr: J^T vct, vct [Y] -> [Q]
// one thread per chain walks the observations (the U-sized part is a sum over them, added
// in observation order: deterministic), rows prefetched P at a time like the sweeps above
// SUB: out = base - J^T vct (the cotangent projection p - J^T G^{-1} J p, systems.py:464-466;
// base may alias out)
template <int U, bool SUB>
__global__ void __launch_bounds__(128)
    k_ssm_rmult_jacob(int64_t n, int Y, int64_t ld, SsmBlocks J, const double* vct,
                      const double* base, double* out) {
  constexpr int P = PrefetchDepth<U>::value;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double* ov = out + (int64_t)U * ld;
  double* on = out + (int64_t)(U + Y) * ld;
  double acc[U];
#pragma unroll
  for (int u = 0; u < U; ++u) acc[u] = 0.0;
  for (int k0 = 0; k0 < Y; k0 += P) {
    double du[P][U], dv[P], n0[P], n1[P], w[P + 1];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int k = clampi(k0 + j, 0, Y - 1);
#pragma unroll
      for (int u = 0; u < U; ++u) du[j][u] = J.du[(int64_t)(k * U + u) * ld + i];
      dv[j] = J.dv[(int64_t)k * ld + i], n0[j] = J.n0[(int64_t)k * ld + i];
      n1[j] = k < Y - 1 ? J.n1[(int64_t)k * ld + i] : 0.0;
      w[j] = vct[(int64_t)k * ld + i];
    }
    w[P] = k0 + P < Y ? vct[(int64_t)(k0 + P) * ld + i] : 0.0;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int k = k0 + j;
      if (k < Y) {
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] = fma(w[j], du[j][u], acc[u]);
        double tv = w[j] * dv[j];
        double t = w[j] * n0[j];
        if (k < Y - 1) t += n1[j] * w[j + 1];
        if (SUB) {
          tv = base[(int64_t)(U + k) * ld + i] - tv;
          t = base[(int64_t)(U + Y + k) * ld + i] - t;
        }
        ov[(int64_t)k * ld + i] = tv;
        on[(int64_t)k * ld + i] = t;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u)
    out[(int64_t)u * ld + i] = SUB ? base[(int64_t)u * ld + i] - acc[u] : acc[u];
}

// ---- stochastic-volatility model, mlift/example_models/stochastic_volatility.py:71-141
// (constr_split / jacob_constr_split_blocks), `unbounded` prior parametrisation
// (prior.py:17-42): mu = u0, sigma = exp(u1), phi = -1 + 2 expit(u2).
// q rows: u [3], v [Y], n [Y].  x_t = 2 log(y_t / n_t);
// c_0 = mu + sigma v_0 / sqrt(1 - phi^2) - x_0, c_t = mu + phi (x_{t-1} - mu) + sigma v_t - x_t.
// y_obs is shared by all chains (broadcast loads).  JAC = false: residual only (the
// quasi-Newton loop, systems.py:204).
struct SsmBlocksOut {
  double *du, *dv, *n0, *n1, *dxdy;
};

// Every row depends on the inputs only (x_{t-1} = 2 log(y_{t-1} / n_{t-1}) is recomputed, not
// carried), so the grid is (chains, chunks of kSvRows observations): at 16,384 chains one
// thread per chain walking all rows took 0.34 ms (residual) / 0.54 ms (blocks) per call.
constexpr int kSvRows = 32;

template <bool JAC>
__global__ void __launch_bounds__(128)
    k_ssm_sv_blocks(int64_t n, int Y, int64_t ld, const double* q, const double* y_obs,
                    double* c, SsmBlocksOut J) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r0 = blockIdx.y * kSvRows;
  const double mu = q[i], u1 = q[ld + i], u2 = q[2 * ld + i];
  const double sigma = exp(u1), e = 1.0 / (1.0 + exp(-u2));
  const double phi = -1.0 + 2.0 * e;
  const double dmu = 1.0, dsigma = sigma, dphi = 2.0 * e * (1.0 - e);
  const double omp = 1.0 - phi * phi, somp = sqrt(omp);
  const double* v = q + 3 * ld;
  const double* nn = q + (int64_t)(3 + Y) * ld;
  double x_prev = 0.0, n_prev = 1.0;
  if (r0 > 0) {
    n_prev = nn[(int64_t)(r0 - 1) * ld + i];
    x_prev = 2.0 * log(y_obs[r0 - 1] / n_prev);
  }
#pragma unroll 8
  for (int k = r0; k < min(r0 + kSvRows, Y); ++k) {
    const double vk = v[(int64_t)k * ld + i], nk = nn[(int64_t)k * ld + i], yk = y_obs[k];
    const double x = 2.0 * log(yk / nk);
    const double v0s = vk / somp;  // used at k = 0 only
    const double xm = x_prev - mu;
    c[(int64_t)k * ld + i] = k == 0 ? mu + sigma * v0s - x : mu + phi * xm + sigma * vk - x;
    if (JAC) {
      J.du[(int64_t)(k * 3 + 0) * ld + i] = k == 0 ? dmu : (1.0 - phi) * dmu;
      J.du[(int64_t)(k * 3 + 1) * ld + i] = dsigma * (k == 0 ? v0s : vk);
      J.du[(int64_t)(k * 3 + 2) * ld + i] = k == 0 ? dphi * phi * sigma * v0s / omp : xm * dphi;
      J.dv[(int64_t)k * ld + i] = k == 0 ? sigma / somp : sigma;
      J.n0[(int64_t)k * ld + i] = 2.0 / nk;
      if (k > 0) J.n1[(int64_t)(k - 1) * ld + i] = -2.0 * phi / n_prev;
      J.dxdy[(int64_t)k * ld + i] = 2.0 / yk;
    }
    x_prev = x, n_prev = nk;
  }
}

// ---- bookkeeping of one iteration of the projection loops (systems.py:203-227 /
// 243-269) for every chain still iterating: error = |c|, mu += delta, q -= delta, i += 1,
// |delta|, then the reference's keep-iterating test; chains that stay active are counted
// in *n_active (the host loop reads that one integer).  Norm: maximum norm (solvers.py:6-8).
// Two kernels: the element-wise part runs over (chain, chunk of kUpdRows rows) -- one thread
// per chain walking all dim_q rows took 2.2 ms per round at 16,384 chains x dim_q 2,003, 58 %
// of a projection -- and folds its partial maxima into per-chain words with atomicMax on the
// bit patterns (non-negative doubles order like unsigned integers, and |NaN| sits above
// +inf, so a NaN survives the maximum as it does in np.max(abs(x))); the per-chain part then
// applies the test.
constexpr int kUpdRows = 64;

__global__ void __launch_bounds__(128)
    k_ssm_projection_apply(int64_t n, int Q, int Y, int64_t ld, const double* c,
                           const double* delta, double* q, double* mu, const int* active,
                           unsigned long long* max_c, unsigned long long* max_d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !active[i]) return;
  const int r0 = blockIdx.y * kUpdRows;
  unsigned long long md = 0ull, mc = 0ull;
#pragma unroll 8
  for (int k = r0; k < min(r0 + kUpdRows, Q); ++k) {
    const double d = delta[(int64_t)k * ld + i];
    q[(int64_t)k * ld + i] -= d;
    mu[(int64_t)k * ld + i] += d;
    const unsigned long long b = (unsigned long long)__double_as_longlong(fabs(d));
    md = b > md ? b : md;
  }
#pragma unroll 8
  for (int k = r0; k < min(r0 + kUpdRows, Y); ++k) {
    const unsigned long long b =
        (unsigned long long)__double_as_longlong(fabs(c[(int64_t)k * ld + i]));
    mc = b > mc ? b : mc;
  }
  if (md) atomicMax(max_d + i, md);
  if (mc) atomicMax(max_c + i, mc);
}

__global__ void __launch_bounds__(128)
    k_ssm_projection_test(int64_t n, const unsigned long long* max_c,
                          const unsigned long long* max_d, int* iters, double* norm_dq,
                          double* error, int* active, double c_tol, double p_tol,
                          double div_tol, int max_iters, int* n_active) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !active[i]) return;
  const double err = __longlong_as_double((long long)max_c[i]);
  const double nd = __longlong_as_double((long long)max_d[i]);
  const int it = iters[i] + 1;
  iters[i] = it, norm_dq[i] = nd, error[i] = err;
  const bool diverged = err > div_tol || err != err;
  const bool converged = err < c_tol && nd < p_tol;
  const bool keep = !(it >= max_iters || diverged || converged);
  active[i] = keep ? 1 : 0;
  if (keep) atomicAdd(n_active, 1);
}

static SsmBlocks dev_blocks(const mlb_ssm_blocks* J) {
  return SsmBlocks{J->dc_du, J->dc_dv, J->dc_dn0, J->dc_dn1, J->dx_dy};
}

static bool blocks_ok(const mlb_ssm_blocks* J, int Y) {
  return J && J->dc_du && J->dc_dv && J->dc_dn0 && (Y == 1 || J->dc_dn1);
}

// scratch rows in units of [ld] doubles
// ---- band of the inverse Gram matrix and G^{-1} dc_du: what the gradient of
// log_det_sqrt_gram needs.  The reference differentiates systems.py:1233-1243 by reverse mode
// (jit table systems.py:332-334); in closed form d(1/2 log det G) = tr(G^{-1} J dJ^T) = the sum
// over the NON-ZERO PATTERN of J = [dc_du, diag(dc_dv), bidiag(dc_dn0, dc_dn1)] of
// (G^{-1} J)_{tj} dJ_{tj}, which takes G^{-1} dc_du [Y x U] and the diagonal / first
// off-diagonal of G^{-1} only.  With G = T + dc_du dc_du^T, Z = T^{-1} dc_du and the
// capacitance matrix C = I + dc_du^T Z (Woodbury):
//   G^{-1} dc_du = Z C^{-1},  G^{-1} = T^{-1} - Z C^{-1} Z^T,
// and the band of T^{-1} = L^{-T} D^{-1} L^{-1} follows from the backward recurrences
//   S_{t,t} = 1 / b'_t + w_{t+1}^2 S_{t+1,t+1},  S_{t,t+1} = -w_{t+1} S_{t+1,t+1}.
// One forward sweep (elimination; d' = L^{-1} dc_du parked in the ru output, 1 / b' in the gd
// output: no scratch) and one backward sweep that overwrites them row by row:
//   z_t = (d'_t - a_t z_{t+1}) / b'_t,  r_t = C^{-1} z_t  -> ru[t],
//   gd[t] = S_tt - r_t . z_t,  go[t] = S_{t,t+1} - r_t . z_{t+1}.
template <int U>
__global__ void __launch_bounds__(128)
    k_ssm_band_inv_gram(int64_t n, int Y, int64_t ld, const double* du_p, const double* a,
                        const double* b, const double* chol, double* ru, double* gd,
                        double* go) {
  constexpr int P = PrefetchDepth<U>::value;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double rbp = 0.0, dp[U];
#pragma unroll
  for (int u = 0; u < U; ++u) dp[u] = 0.0;
  for (int k0 = 0; k0 < Y; k0 += P) {
    double du[P][U], av[P], bv[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int k = clampi(k0 + j, 0, Y - 1);
#pragma unroll
      for (int u = 0; u < U; ++u) du[j][u] = du_p[(int64_t)(k * U + u) * ld + i];
      av[j] = k > 0 ? a[(int64_t)(k - 1) * ld + i] : 0.0;
      bv[j] = b[(int64_t)k * ld + i];
    }
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int k = k0 + j;
      if (k < Y) {
        const double w = av[j] * rbp;
        rbp = 1.0 / (bv[j] - w * av[j]);
        gd[(int64_t)k * ld + i] = rbp;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          dp[u] = du[j][u] - w * dp[u];
          ru[(int64_t)(k * U + u) * ld + i] = dp[u];
        }
      }
    }
  }
  // C^{-1} = L^{-T} L^{-1} explicitly (U <= 8, registers)
  double L[U][U], rdiag[U], ci[U][U];
#pragma unroll
  for (int r = 0; r < U; ++r) {
#pragma unroll
    for (int c = 0; c < U; ++c) L[r][c] = chol[(int64_t)(r * U + c) * ld + i];
    rdiag[r] = 1.0 / L[r][r];
  }
#pragma unroll
  for (int c = 0; c < U; ++c) {
    double e[U];
#pragma unroll
    for (int r = 0; r < U; ++r) e[r] = r == c ? 1.0 : 0.0;
    chol_solve<U>(L, rdiag, e);
#pragma unroll
    for (int r = 0; r < U; ++r) ci[r][c] = e[r];
  }
  double z_next[U], s_next = 0.0;
#pragma unroll
  for (int u = 0; u < U; ++u) z_next[u] = 0.0;
#pragma unroll 1
  for (int k = Y - 1; k >= 0; --k) {
    const double rb = gd[(int64_t)k * ld + i];
    const double ak = k < Y - 1 ? a[(int64_t)k * ld + i] : 0.0;
    const double w = ak * rb;  // multiplier of row k + 1
    double z[U], r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) z[u] = (ru[(int64_t)(k * U + u) * ld + i] - ak * z_next[u]) * rb;
    const double s_off = -w * s_next, s_diag = rb - w * s_off;
    double qd = 0.0, qo = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < U; ++c) acc = fma(ci[u][c], z[c], acc);
      r[u] = acc;
      qd = fma(acc, z[u], qd);
      qo = fma(acc, z_next[u], qo);
      ru[(int64_t)(k * U + u) * ld + i] = acc;
    }
    gd[(int64_t)k * ld + i] = s_diag - qd;
    if (k < Y - 1) go[(int64_t)k * ld + i] = s_off - qo;
#pragma unroll
    for (int u = 0; u < U; ++u) z_next[u] = z[u];
    s_next = s_diag;
  }
}

// ---- gradient of log_det_sqrt_gram for the stochastic-volatility model: contraction of
// the weights above with the derivatives of the Jacobian entries of
// stochastic_volatility.py:92-141 (mu = u0, sigma = exp(u1), phi = 2 expit(u2) - 1; dx_dy =
// 2 / y does not depend on q).  One chain per thread, one pass over the rows.
__global__ void __launch_bounds__(128)
    k_ssm_sv_grad_log_det(int64_t n, int Y, int64_t ld, const double* q, const double* y_obs,
                          SsmBlocks J, const double* ru, const double* gd, const double* go,
                          double* grad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mu = q[i], u1 = q[ld + i], u2 = q[2 * ld + i];
  const double sigma = exp(u1), e = 1.0 / (1.0 + exp(-u2));
  const double phi = -1.0 + 2.0 * e, dphi = 2.0 * e * (1.0 - e), ddphi = dphi * (1.0 - 2.0 * e);
  const double omp = 1.0 - phi * phi, s1 = sqrt(omp), rs = 1.0 / s1, rs3 = rs / omp,
               rs5 = rs3 / omp;
  const double* v = q + 3 * ld;
  const double* nn = q + (int64_t)(3 + Y) * ld;
  double* gv = grad + 3 * ld;
  double* gn = grad + (int64_t)(3 + Y) * ld;
  double g0 = 0.0, g1 = 0.0, g2 = 0.0;
  double gd_k = gd[i];
  double x_prev = 0.0;
#pragma unroll 1
  for (int k = 0; k < Y; ++k) {
    const double vk = v[(int64_t)k * ld + i], nk = nn[(int64_t)k * ld + i];
    const double rn = 1.0 / nk;
    const double r0 = ru[(int64_t)(k * 3 + 0) * ld + i], r1 = ru[(int64_t)(k * 3 + 1) * ld + i],
                 r2 = ru[(int64_t)(k * 3 + 2) * ld + i];
    const double wdv = gd_k * J.dv[(int64_t)k * ld + i];
    const double dn0 = J.n0[(int64_t)k * ld + i];
    const bool inner = k < Y - 1;
    const double dn1 = inner ? J.n1[(int64_t)k * ld + i] : 0.0;
    const double go_k = inner ? go[(int64_t)k * ld + i] : 0.0;
    const double gd_n = inner ? gd[(int64_t)(k + 1) * ld + i] : 0.0;
    const double r2_next = inner ? ru[(int64_t)((k + 1) * 3 + 2) * ld + i] : 0.0;
    const double w0 = gd_k * dn0 + go_k * dn1, w1 = go_k * dn0 + gd_n * dn1;
    if (k == 0) {
      g1 += r1 * sigma * vk * rs + r2 * dphi * phi * sigma * vk * rs3 + wdv * sigma * rs;
      g2 += r1 * sigma * vk * phi * dphi * rs3 +
            r2 * sigma * vk * (ddphi * phi * rs3 + dphi * dphi * rs3 +
                               3.0 * phi * phi * dphi * dphi * rs5) +
            wdv * sigma * phi * dphi * rs3;
      gv[i] = r1 * sigma * rs + r2 * dphi * phi * sigma * rs3;
    } else {
      g0 -= dphi * r2;
      g1 += sigma * (r1 * vk + wdv);
      g2 += -dphi * r0 + ddphi * r2 * (x_prev - mu);
      gv[(int64_t)k * ld + i] = sigma * r1;
    }
    double gnk = w0 * (-2.0 * rn * rn);
    if (inner) {
      g2 -= 2.0 * dphi * w1 * rn;
      gnk += r2_next * dphi * (-2.0 * rn) + w1 * 2.0 * phi * rn * rn;
    }
    gn[(int64_t)k * ld + i] = gnk;
    x_prev = 2.0 * log(y_obs[k] * rn);
    gd_k = gd_n;
  }
  grad[i] = g0, grad[ld + i] = g1, grad[2 * ld + i] = g2;
}

struct Scratch {
  double* p = nullptr;
  cudaStream_t s;
  int alloc(size_t rows, int64_t ld, cudaStream_t st) {
    s = st;
    MLB_CUDA_OK(scratch_alloc((void**)&p, sizeof(double) * rows * (size_t)ld, st));
    return MLB_OK;
  }
  ~Scratch() {
    if (p) cudaFreeAsync(p, s);
  }
};

#define MLB_SSM_DISPATCH_U(U_, ...)                                    \
  switch (U_) {                                                        \
    case 1: { constexpr int UU = 1; __VA_ARGS__; } break;              \
    case 2: { constexpr int UU = 2; __VA_ARGS__; } break;              \
    case 3: { constexpr int UU = 3; __VA_ARGS__; } break;              \
    case 4: { constexpr int UU = 4; __VA_ARGS__; } break;              \
    case 5: { constexpr int UU = 5; __VA_ARGS__; } break;              \
    case 6: { constexpr int UU = 6; __VA_ARGS__; } break;              \
    case 7: { constexpr int UU = 7; __VA_ARGS__; } break;              \
    case 8: { constexpr int UU = 8; __VA_ARGS__; } break;              \
    default: break;                                                    \
  }

}  // namespace mlb

using namespace mlb;

extern "C" {

int mlb_ssm_decompose_gram(int64_t n, int32_t dim_y, int32_t dim_u, int64_t ld,
                           const mlb_ssm_blocks* J, double* a, double* b, double* chol_cap,
                           void* stream) {
  MLB_ARG(n >= 0 && dim_y >= 1 && dim_u >= 1 && dim_u <= MLB_SSM_MAX_DIM_U && ld >= n);
  MLB_ARG(blocks_ok(J, dim_y) && b && chol_cap && (dim_y == 1 || a));
  if (n == 0) return MLB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = (unsigned)((n + 127) / 128);
  MLB_SSM_DISPATCH_U(dim_u, k_ssm_decompose<UU><<<g, 128, 0, s>>>(n, dim_y, ld, dev_blocks(J), a,
                                                                 b, chol_cap));
  return finish_launch();
}

int mlb_ssm_lmult_by_inv_gram(int64_t n, int32_t dim_y, int32_t dim_u, int64_t ld,
                              const mlb_ssm_blocks* J, const double* a, const double* b,
                              const double* chol_cap, const double* vct, double* out,
                              void* stream) {
  MLB_ARG(n >= 0 && dim_y >= 1 && dim_u >= 1 && dim_u <= MLB_SSM_MAX_DIM_U && ld >= n);
  MLB_ARG(J && J->dc_du && b && chol_cap && vct && out && vct != out && (dim_y == 1 || a));
  if (n == 0) return MLB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  Scratch sc;
  if (int rc = sc.alloc((size_t)dim_y * 2, ld, s)) return rc;
  const unsigned g = (unsigned)((n + 127) / 128);
  MLB_SSM_DISPATCH_U(dim_u, k_ssm_inv_gram<UU><<<g, 128, 0, s>>>(
                                n, dim_y, ld, J->dc_du, a, b, chol_cap, vct, out, sc.p,
                                sc.p + (size_t)dim_y * (size_t)ld));
  return finish_launch();
}

int mlb_ssm_lmult_by_inv_jacob_product(int64_t n, int32_t dim_y, int32_t dim_u, int64_t ld,
                                       const mlb_ssm_blocks* J_l, const mlb_ssm_blocks* J_r,
                                       const double* vct, double* out, void* stream) {
  MLB_ARG(n >= 0 && dim_y >= 1 && dim_u >= 1 && dim_u <= MLB_SSM_MAX_DIM_U && ld >= n);
  MLB_ARG(blocks_ok(J_l, dim_y) && blocks_ok(J_r, dim_y) && vct && out && vct != out);
  if (n == 0) return MLB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  Scratch sc;
  if (int rc = sc.alloc((size_t)dim_y * 3, ld, s)) return rc;
  const size_t row = (size_t)dim_y * (size_t)ld;
  const unsigned g = (unsigned)((n + 127) / 128);
  MLB_SSM_DISPATCH_U(dim_u, k_ssm_inv_jacob_product<UU><<<g, 128, 0, s>>>(
                                n, dim_y, ld, dev_blocks(J_l), dev_blocks(J_r), vct, out, sc.p,
                                sc.p + row, sc.p + 2 * row));
  return finish_launch();
}

int mlb_ssm_log_det_sqrt_gram(int64_t n, int32_t dim_y, int32_t dim_u, int64_t ld,
                              const mlb_ssm_blocks* J, const double* a, const double* b,
                              const double* chol_cap, double* out, void* stream) {
  MLB_ARG(n >= 0 && dim_y >= 1 && dim_u >= 1 && dim_u <= MLB_SSM_MAX_DIM_U && ld >= n);
  MLB_ARG(J && J->dx_dy && b && chol_cap && out && (dim_y == 1 || a));
  if (n == 0) return MLB_OK;
  k_ssm_log_det<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      n, dim_y, dim_u, ld, a, b, chol_cap, J->dx_dy, out);
  return finish_launch();
}

int mlb_ssm_band_inv_gram(int64_t n, int32_t dim_y, int32_t dim_u, int64_t ld,
                          const mlb_ssm_blocks* J, const double* a, const double* b,
                          const double* chol_cap, double* inv_gram_dc_du, double* inv_gram_diag,
                          double* inv_gram_offdiag, void* stream) {
  MLB_ARG(n >= 0 && dim_y >= 1 && dim_u >= 1 && dim_u <= MLB_SSM_MAX_DIM_U && ld >= n);
  MLB_ARG(J && J->dc_du && b && chol_cap && inv_gram_dc_du && inv_gram_diag &&
          (dim_y == 1 || (a && inv_gram_offdiag)));
  if (n == 0) return MLB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = (unsigned)((n + 127) / 128);
  MLB_SSM_DISPATCH_U(dim_u, k_ssm_band_inv_gram<UU><<<g, 128, 0, s>>>(
                                n, dim_y, ld, J->dc_du, a, b, chol_cap, inv_gram_dc_du,
                                inv_gram_diag, inv_gram_offdiag));
  return finish_launch();
}

int mlb_ssm_sv_grad_log_det_sqrt_gram(int64_t n, int32_t dim_y, int64_t ld, const double* q,
                                      const double* y_obs, const mlb_ssm_blocks* J,
                                      const double* inv_gram_dc_du, const double* inv_gram_diag,
                                      const double* inv_gram_offdiag, double* grad,
                                      void* stream) {
  MLB_ARG(n >= 0 && dim_y >= 1 && ld >= n && q && y_obs && blocks_ok(J, dim_y));
  MLB_ARG(inv_gram_dc_du && inv_gram_diag && grad && (dim_y == 1 || inv_gram_offdiag));
  if (n == 0) return MLB_OK;
  k_ssm_sv_grad_log_det<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      n, dim_y, ld, q, y_obs, dev_blocks(J), inv_gram_dc_du, inv_gram_diag, inv_gram_offdiag,
      grad);
  return finish_launch();
}

int mlb_ssm_lmult_by_jacob_constr(int64_t n, int32_t dim_y, int32_t dim_u, int64_t ld,
                                  const mlb_ssm_blocks* J, const double* vct, double* out,
                                  void* stream) {
  MLB_ARG(n >= 0 && dim_y >= 1 && dim_y <= 65535 && dim_u >= 1 &&
          dim_u <= MLB_SSM_MAX_DIM_U && ld >= n);
  MLB_ARG(blocks_ok(J, dim_y) && vct && out);
  if (n == 0) return MLB_OK;
  k_ssm_lmult_jacob<<<dim3((unsigned)((n + 127) / 128), (unsigned)dim_y), 128, 0,
                      (cudaStream_t)stream>>>(n, dim_y, dim_u, ld, dev_blocks(J), vct, out);
  return finish_launch();
}

int mlb_ssm_rmult_by_jacob_constr(int64_t n, int32_t dim_y, int32_t dim_u, int64_t ld,
                                  const mlb_ssm_blocks* J, const double* vct, double* out,
                                  void* stream) {
  MLB_ARG(n >= 0 && dim_y >= 1 && dim_u >= 1 && dim_u <= MLB_SSM_MAX_DIM_U && ld >= n);
  MLB_ARG(blocks_ok(J, dim_y) && vct && out);
  if (n == 0) return MLB_OK;
  const unsigned g = (unsigned)((n + 127) / 128);
  cudaStream_t s = (cudaStream_t)stream;
  MLB_SSM_DISPATCH_U(dim_u, k_ssm_rmult_jacob<UU, false><<<g, 128, 0, s>>>(
                                n, dim_y, ld, dev_blocks(J), vct, nullptr, out));
  return finish_launch();
}

int mlb_ssm_sv_constr_blocks(int64_t n, int32_t dim_y, int64_t ld, const double* q,
                             const double* y_obs, double* c, const mlb_ssm_blocks* J_out,
                             void* stream) {
  MLB_ARG(n >= 0 && dim_y >= 1 && ld >= n && q && y_obs && c);
  MLB_ARG(!J_out || (blocks_ok(J_out, dim_y) && J_out->dx_dy));
  MLB_ARG(dim_y <= 65535 * kSvRows);
  if (n == 0) return MLB_OK;
  const dim3 gy((unsigned)((n + 127) / 128), (unsigned)((dim_y + kSvRows - 1) / kSvRows));
  cudaStream_t s = (cudaStream_t)stream;
  if (J_out) {
    SsmBlocksOut o{const_cast<double*>(J_out->dc_du), const_cast<double*>(J_out->dc_dv),
                   const_cast<double*>(J_out->dc_dn0), const_cast<double*>(J_out->dc_dn1),
                   const_cast<double*>(J_out->dx_dy)};
    k_ssm_sv_blocks<true><<<gy, 128, 0, s>>>(n, dim_y, ld, q, y_obs, c, o);
  } else {
    k_ssm_sv_blocks<false><<<gy, 128, 0, s>>>(n, dim_y, ld, q, y_obs, c, SsmBlocksOut{});
  }
  return finish_launch();
}

int mlb_ssm_projection_update(int64_t n, int32_t dim_q, int32_t dim_y, int64_t ld,
                              const double* c, const double* delta_q, double* q, double* mu,
                              int32_t* n_iters, double* norm_delta_q, double* error,
                              int32_t* active, double constraint_tol, double position_tol,
                              double divergence_tol, int32_t max_iters, int32_t* n_active,
                              void* stream) {
  MLB_ARG(n >= 0 && dim_y >= 1 && dim_q >= dim_y && ld >= n && c && delta_q && q && mu);
  MLB_ARG(n_iters && norm_delta_q && error && active && n_active);
  if (n == 0) return MLB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* mx = nullptr;  // [2][n] partial maxima
  MLB_CUDA_OK(scratch_alloc((void**)&mx, sizeof(unsigned long long) * 2 * (size_t)n, s));
  MLB_CUDA_OK(cudaMemsetAsync(mx, 0, sizeof(unsigned long long) * 2 * (size_t)n, s));
  const unsigned g = (unsigned)((n + 127) / 128);
  k_ssm_projection_apply<<<dim3(g, (unsigned)((dim_q + kUpdRows - 1) / kUpdRows)), 128, 0, s>>>(
      n, dim_q, dim_y, ld, c, delta_q, q, mu, active, mx, mx + n);
  int rc = finish_launch();
  if (rc == MLB_OK) {
    k_ssm_projection_test<<<g, 128, 0, s>>>(n, mx, mx + n, n_iters, norm_delta_q, error, active,
                                           constraint_tol, position_tol, divergence_tol,
                                           max_iters, n_active);
    rc = finish_launch();
  }
  cudaFreeAsync(mx, s);
  return rc;
}

// systems.py:181-188 (normal_space_component) and :464-466 (project_onto_cotangent_space):
// mom -= J^T (J J^T)^{-1} J mom, as three launches: J mom, the Gram solve, and J^T(.)
// subtracted from mom in the same pass
int mlb_ssm_project_onto_cotangent_space(int64_t n, int32_t dim_y, int32_t dim_u, int64_t ld,
                                         const mlb_ssm_blocks* J, const double* a,
                                         const double* b, const double* chol_cap, double* mom,
                                         void* stream) {
  MLB_ARG(n >= 0 && dim_y >= 1 && dim_y <= 65535 && dim_u >= 1 &&
          dim_u <= MLB_SSM_MAX_DIM_U && ld >= n);
  MLB_ARG(blocks_ok(J, dim_y) && b && chol_cap && mom && (dim_y == 1 || a));
  if (n == 0) return MLB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  Scratch sc;  // J mom [Y], G^{-1} J mom [Y], sweep scratch [2 Y]
  if (int rc = sc.alloc((size_t)dim_y * 4, ld, s)) return rc;
  const size_t row = (size_t)dim_y * (size_t)ld;
  double *jp = sc.p, *gjp = sc.p + row, *sw = sc.p + 2 * row;
  const unsigned g = (unsigned)((n + 127) / 128);
  k_ssm_lmult_jacob<<<dim3(g, (unsigned)dim_y), 128, 0, s>>>(n, dim_y, dim_u, ld, dev_blocks(J),
                                                            mom, jp);
  if (int rc = finish_launch()) return rc;
  MLB_SSM_DISPATCH_U(dim_u, k_ssm_inv_gram<UU><<<g, 128, 0, s>>>(n, dim_y, ld, J->dc_du, a, b,
                                                                chol_cap, jp, gjp, sw, sw + row));
  if (int rc = finish_launch()) return rc;
  MLB_SSM_DISPATCH_U(dim_u, k_ssm_rmult_jacob<UU, true><<<g, 128, 0, s>>>(
                                n, dim_y, ld, dev_blocks(J), gjp, mom, mom));
  return finish_launch();
}

}  // extern "C"
