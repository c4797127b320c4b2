// Common device helpers: norms, tiny dense factorizations in registers, Philox.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#include "../../include/mlb.h"

#define MLB_DEV __device__ __forceinline__
#define MLB_HD __host__ __device__ __forceinline__

namespace mlb {

// Solver / integrator options in kernel-parameter form (see mlb_solver_opts).
struct Opts {
  double ctol, ptol, dtol, rev_tol;
  int max_iters, max_ls, norm, rev_norm, flags;
};

// ---- lean fp64 primitives
// The compiler's IEEE double division / sqrt expand to ~15-25 instructions plus an
// out-of-line slow path each; the hot path has dozens of call sites, which made the fused
// kernels larger than the instruction cache.  These versions refine the hardware
// approximation (MUFU.RCP64H / MUFU.RSQ64H, ~2^-22 relative) with two Newton steps:
// error <= ~1 ulp for normal, non-zero finite arguments (not correctly rounded, no
// subnormal handling; tests/test_gpu_parity.py::test_fast_math_primitives checks them
// against the IEEE operations).  0 -> NaN rather than inf: callers test pivots.
MLB_DEV double rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// Reciprocals of N values with one hardware reciprocal per group of four (Montgomery's
// trick: 1/a = (bcd) / (abcd)): 9 multiplications + one refined reciprocal per group instead
// of four refined reciprocals (20 instructions, four MUFU).  Each result carries ~3 roundings
// instead of one; the product of a group must stay inside the fp64 range (callers pass
// values whose fourth power is finite; otherwise the group comes out NaN and the chain is
// flagged like any other non-finite solve).
#ifndef MLB_RCP_BATCH
#define MLB_RCP_BATCH 0
#endif
template <int N>
MLB_DEV void rcp_many(const double (&x)[N], double (&r)[N]) {
  if (MLB_RCP_BATCH && N % 4 == 0) {
#pragma unroll
    for (int g = 0; g < N; g += 4) {
      const double ab = x[g] * x[g + 1], cd = x[g + 2] * x[g + 3];
      const double inv = rcp_fast(ab * cd);
      const double rab = inv * cd, rcd = inv * ab;
      r[g] = rab * x[g + 1], r[g + 1] = rab * x[g];
      r[g + 2] = rcd * x[g + 3], r[g + 3] = rcd * x[g + 2];
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = rcp_fast(x[i]);
  }
}

// s > 0: returns sqrt(s), writes 1 / sqrt(s)
MLB_DEV double sqrt_rsqrt_fast(double s, double& rs) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
  double h = 0.5 * s;
  r = fma(r, fma(-h * r, r, 0.5), r);
  r = fma(r, fma(-h * r, r, 0.5), r);
  double d = s * r;
  d = fma(fma(-d, d, s), 0.5 * r, d);
  rs = fma(r, fma(-d, r, 1.0), r);
  return d;
}

// ---- branch-free exp / expm1
// The Hodgkin-Huxley recurrence evaluates 10 exp and 3 expm1 per step, eight of them
// mutually independent; libdevice's versions carry a range-handling branch each, which
// ends the basic block and makes the thread run them back to back (~200 cycles of
// dependent latency apiece, measured 4300 cycles per ODE step for one warp).  These are
// straight-line (Cody-Waite reduction to |r| <= ln2/2, degree-13 Taylor polynomial
// (remainder < 5e-18), scaling by a constructed power of two) and short (~12 dependent
// fp64 operations).  Arguments are clamped to
// [-700, 700] (no overflow / subnormal handling); NaN propagates, +-inf gives NaN.  Error <= ~1 ulp (exp),
// <= ~2 ulp (expm1), checked against the libdevice functions by
// tests/test_gpu_parity.py::test_fast_math_primitives.
MLB_DEV void exp_reduce(double x, double& r, double& scale) {
  x = fmin(fmax(x, -700.0), 700.0);
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: rounds to nearest integer
  const double t = fma(x, 1.4426950408889634074, magic);
  const double kf = t - magic;
  r = fma(kf, -6.93147180369123816490e-01, x);  // ln2 high part (fdlibm split)
  r = fma(kf, -1.90821492927058770002e-10, r);  // ln2 low part
  const int k = __double2loint(t);
  scale = __hiloint2double((k + 1023) << 20, 0);
}
// (exp(r) - 1 - r) / r^2 = sum_{k<12} r^k / (k+2)!, Estrin's scheme: 5 dependent levels
// instead of Horner's 12 (the evaluation sits on the critical path of the recurrence)
MLB_DEV double exp_poly_tail(double r) {
  const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  const double p01 = fma(r, 1.0 / 6.0, 0.5);
  const double p23 = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  const double p45 = fma(r, 1.0 / 5040.0, 1.0 / 720.0);
  const double p67 = fma(r, 1.0 / 362880.0, 1.0 / 40320.0);
  const double p89 = fma(r, 1.0 / 39916800.0, 1.0 / 3628800.0);
  const double pab = fma(r, 1.0 / 6227020800.0, 1.0 / 479001600.0);
  const double q0 = fma(r2, p23, p01), q1 = fma(r2, p67, p45), q2 = fma(r2, pab, p89);
  return fma(r8, q2, fma(r4, q1, q0));
}
MLB_DEV double exp_bl(double x) {
  double r, s;
  exp_reduce(x, r, s);
  const double em1 = fma(r * r, exp_poly_tail(r), r);
  return fma(s, em1, s) + (x - x);  // x - x: +0 for finite x, NaN for NaN (a select here
                                    // is compiled to a branch around the evaluation)
}
MLB_DEV double expm1_bl(double x) {
  double r, s;
  exp_reduce(x, r, s);
  const double em1 = fma(r * r, exp_poly_tail(r), r);
  return fma(s, em1, s - 1.0) + (x - x);
}

// ---- norms (reference mlift/solvers.py:6-13).  The maximum norm propagates NaN like
// numpy / jax `max`: for non-negative doubles the IEEE bit patterns order like unsigned
// integers and every NaN pattern is above +inf, so an integer max of |v| does both jobs
// off the fp64 pipe.
// sum_{i<N} a(i) * b(i) (+ init) with MLB_NACC independent accumulators: a reduction over
// dim_y as ONE dependent FMA chain leaves the fp64 pipe idle for most of its latency when
// only two warps share a scheduler
#ifndef MLB_NACC
#define MLB_NACC 1
#endif
template <int N, class FA, class FB>
MLB_DEV double dot_acc(double init, FA&& a, FB&& b) {
  constexpr int NA = MLB_NACC < N ? MLB_NACC : N;
  double acc[NA];
#pragma unroll
  for (int k = 0; k < NA; ++k) acc[k] = k == 0 ? init : 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) acc[i % NA] = fma(a(i), b(i), acc[i % NA]);
#pragma unroll
  for (int w = 1; w < NA; w *= 2)
#pragma unroll
    for (int k = 0; k + w < NA; k += 2 * w) acc[k] += acc[k + w];
  return acc[0];
}

template <int N>
MLB_DEV double vec_norm(const double (&v)[N], int kind) {
  if (kind == MLB_NORM_MAX) {
    unsigned long long m = 0ull;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      unsigned long long a =
          (unsigned long long)__double_as_longlong(v[i]) & 0x7fffffffffffffffull;
      m = a > m ? a : m;
    }
    return __longlong_as_double((long long)m);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) s = fma(v[i], v[i], s);
  return sqrt(s);
}

// Maximum norm truncated to its high word: max over the high 32 bits of |v_i| (two integer
// instructions per element instead of a 64-bit compare-and-select), returned as the double
// with that high word and a zero low word.  It compares against a threshold exactly like
// the true norm unless the two share their high word (`norm_hi_ambiguous`), which callers
// resolve with the exact norm; NaN / inf stay above every finite threshold.
template <int N>
MLB_DEV double max_norm_hi(const double (&v)[N]) {
  unsigned m = 0u;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const unsigned a = (unsigned)__double2hiint(v[i]) & 0x7fffffffu;
    m = a > m ? a : m;
  }
  return __hiloint2double((int)m, 0);
}
MLB_DEV bool norm_hi_ambiguous(double truncated, double threshold) {
  return __double2hiint(truncated) == __double2hiint(threshold);
}

// ---- K x K factorizations, fully unrolled, static indexing only (stay in registers)

// in-place lower Cholesky; rdiag <- reciprocal diagonal; false if a pivot is not positive
template <int K>
MLB_DEV bool chol_lower(double (&a)[K][K], double (&rdiag)[K]) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    double s = a[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) s = fma(-a[j][k], a[j][k], s);
    ok &= (s > 0.0);
    double rd;
    a[j][j] = sqrt_rsqrt_fast(s, rd);
    rdiag[j] = rd;
#pragma unroll
    for (int i = j + 1; i < K; ++i) {
      double t = a[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) t = fma(-a[i][k], a[j][k], t);
      a[i][j] = t * rd;
    }
#pragma unroll
    for (int i = 0; i < j; ++i) a[i][j] = 0.0;
  }
  return ok;
}

// solve L L^T x = b in place
template <int K>
MLB_DEV void chol_solve(const double (&L)[K][K], const double (&rdiag)[K], double (&b)[K]) {
#pragma unroll
  for (int i = 0; i < K; ++i) {
    double t = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) t = fma(-L[i][k], b[k], t);
    b[i] = t * rdiag[i];
  }
#pragma unroll
  for (int i = K - 1; i >= 0; --i) {
    double t = b[i];
#pragma unroll
    for (int k = i + 1; k < K; ++k) t = fma(-L[k][i], b[k], t);
    b[i] = t * rdiag[i];
  }
}

// LU with partial pivoting (np.linalg.solve semantics, reference systems.py:601,632);
// a destroyed, b overwritten by the solution.  Pivot search by conditional row swaps.
template <int K>
MLB_DEV bool lu_solve(double (&a)[K][K], double (&b)[K]) {
  bool ok = true;
  double rp[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
#pragma unroll
    for (int i = j + 1; i < K; ++i) {
      bool sw = fabs(a[i][j]) > fabs(a[j][j]);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        double x = a[j][k], y = a[i][k];
        a[j][k] = sw ? y : x;
        a[i][k] = sw ? x : y;
      }
      double x = b[j], y = b[i];
      b[j] = sw ? y : x;
      b[i] = sw ? x : y;
    }
    ok &= (fabs(a[j][j]) > 0.0);
    rp[j] = rcp_fast(a[j][j]);
#pragma unroll
    for (int i = j + 1; i < K; ++i) {
      double f = a[i][j] * rp[j];
#pragma unroll
      for (int k = j + 1; k < K; ++k) a[i][k] = fma(-f, a[j][k], a[i][k]);
      b[i] = fma(-f, b[j], b[i]);
    }
  }
#pragma unroll
  for (int i = K - 1; i >= 0; --i) {
    double t = b[i];
#pragma unroll
    for (int k = i + 1; k < K; ++k) t = fma(-a[i][k], b[k], t);
    b[i] = t * rp[i];
  }
  return ok;
}

// ---- Philox4x32-10 counter-based RNG (Salmon et al. 2011); the stream layout is part of the ABI
MLB_HD void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0, c[1] = n1, c[2] = n2, c[3] = n3;
    k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
  }
}

MLB_HD double u01(uint32_t hi, uint32_t lo) {  // (0,1), 53 bits
  uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}

// stream 0: momentum normals, pair j -> normals (2j, 2j+1); stream 1: accept uniform
MLB_HD void philox_normal_pair(uint64_t seed, uint64_t chain, uint32_t iter, uint32_t j,
                               double& n0, double& n1) {
  uint32_t c[4] = {j, iter, (uint32_t)chain, (uint32_t)(chain >> 32) & 0x7fffffffu};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  double u1 = u01(c[0], c[1]), u2 = u01(c[2], c[3]);
  double r = sqrt(-2.0 * log(u1)), th = 6.283185307179586476925286766559 * u2;
  double s, co;
  sincos(th, &s, &co);
  n0 = r * co, n1 = r * s;
}

MLB_HD double philox_uniform(uint64_t seed, uint64_t chain, uint32_t iter) {
  uint32_t c[4] = {0u, iter, (uint32_t)chain,
                   ((uint32_t)(chain >> 32) & 0x7fffffffu) | 0x80000000u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return u01(c[0], c[1]);
}

}  // namespace mlb
