// Block-structured constraint algebra, one chain per thread, everything in registers.
//
// B200-native counterpart of the per-model closures the reference builds in
//   mlift/systems.py:582-641 (IndependentAdditiveNoiseModelSystem) and
//   mlift/systems.py:712-794 (HierarchicalLatentVariableModelSystem),
// i.e. J = (dy_du dense [Y x U], dy_dv diagonal, dy_dn diagonal) with either a dense
// Y x Y Cholesky of the Gram matrix (dim_u >= dim_y) or the Woodbury / capacitance-matrix
// form (dim_u < dim_y).  Sizes are template parameters so every loop unrolls and every
// array is addressed statically; after inlining the compiler specialises the generic
// block code to each model's structure (e.g. identical dy_dv entries collapse to one
// register).
#pragma once
#include "mlb_common.cuh"

namespace mlb {

enum { FAM_ADDITIVE = 0, FAM_HIER = 1 };

template <int FAM_, int U_, int Y_, bool DENSE_>
struct Blocks {
  static constexpr int FAM = FAM_, U = U_, Y = Y_;
  static constexpr bool DENSE = DENSE_;
  static constexpr int K = DENSE_ ? Y_ : U_;
  static constexpr int Q = U_ + (FAM_ == FAM_HIER ? 2 : 1) * Y_;
  static constexpr int JAC_ROWS = Y_ * U_ + (FAM_ == FAM_HIER ? 2 : 1) * Y_;
  static constexpr int GRAM_ROWS = K * K + Y_;

  struct Jac {
    double du[Y][U];
    double dv[Y];  // unused (zero) for the additive-noise family
    double dn[Y];
  };
  struct Gram {
    double chol[K][K];  // Cholesky factor of the Gram (dense) or capacitance matrix
    double rdiag[K];    // 1 / diag(chol)
    double d[Y];        // dy_dv^2 + dy_dn^2
    double rd[Y];       // 1 / d (hot-path multiplications instead of divisions)
    bool ok;
  };

  // J v   (systems.py:582-584, 712-718)
  static MLB_DEV void lmult_jacob(const Jac& J, const double (&v)[Q], double (&out)[Y]) {
#pragma unroll
    for (int i = 0; i < Y; ++i) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < U; ++a) s = fma(J.du[i][a], v[a], s);
      if (FAM == FAM_HIER) {
        s = fma(J.dv[i], v[U + i], s);
        s = fma(J.dn[i], v[U + Y + i], s);
      } else {
        s = fma(J.dn[i], v[U + i], s);
      }
      out[i] = s;
    }
  }

  // J^T w   (systems.py:586-587, 720-721)
  static MLB_DEV void rmult_jacob(const Jac& J, const double (&w)[Y], double (&out)[Q]) {
#pragma unroll
    for (int a = 0; a < U; ++a) {
      out[a] = dot_acc<Y>(0.0, [&](int i) { return w[i]; }, [&](int i) { return J.du[i][a]; });
    }
#pragma unroll
    for (int i = 0; i < Y; ++i) {
      if (FAM == FAM_HIER) {
        out[U + i] = w[i] * J.dv[i];
        out[U + Y + i] = w[i] * J.dn[i];
      } else {
        out[U + i] = w[i] * J.dn[i];
      }
    }
  }

  // decompose_gram   (systems.py:591-594 dense; 613-617, 751-757 Woodbury)
  static MLB_DEV void decompose_gram(const Jac& J, Gram& G) {
#pragma unroll
    for (int i = 0; i < Y; ++i) {
      double dv = FAM == FAM_HIER ? J.dv[i] : 0.0;
      G.d[i] = fma(dv, dv, J.dn[i] * J.dn[i]);
    }
    rcp_many<Y>(G.d, G.rd);
    if (DENSE) {
#pragma unroll
      for (int i = 0; i < Y; ++i)
#pragma unroll
        for (int j = 0; j < Y; ++j) {
          double s = (i == j) ? G.d[i] : 0.0;
#pragma unroll
          for (int a = 0; a < U; ++a) s = fma(J.du[i][a], J.du[j][a], s);
          G.chol[i][j] = s;
        }
    } else {
#pragma unroll
      for (int a = 0; a < U; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) {
          const double s = dot_acc<Y>((a == b) ? 1.0 : 0.0, [&](int i) { return J.du[i][a]; },
                                      [&](int i) { return J.du[i][b] * G.rd[i]; });
          G.chol[a][b] = s;
          G.chol[b][a] = s;
        }
    }
    G.ok = chol_lower<K>(G.chol, G.rdiag);
  }

  // Gram^-1 w   (systems.py:596-597 dense; 619-624, 759-768 Woodbury)
  static MLB_DEV void lmult_inv_gram(const Jac& J, const Gram& G, const double (&w)[Y],
                                     double (&out)[Y]) {
    if (DENSE) {
      double b[K];
#pragma unroll
      for (int i = 0; i < Y; ++i) b[i] = w[i];
      chol_solve<K>(G.chol, G.rdiag, b);
#pragma unroll
      for (int i = 0; i < Y; ++i) out[i] = b[i];
    } else {
      double t[K];
#pragma unroll
      for (int a = 0; a < U; ++a) {
        t[a] = dot_acc<Y>(0.0, [&](int i) { return J.du[i][a]; },
                          [&](int i) { return w[i] * G.rd[i]; });
      }
      chol_solve<K>(G.chol, G.rdiag, t);
#pragma unroll
      for (int i = 0; i < Y; ++i) {
        double s = w[i];
#pragma unroll
        for (int a = 0; a < U; ++a) s = fma(-J.du[i][a], t[a], s);
        out[i] = s * G.rd[i];
      }
    }
  }

  // (J_l J_r^T)^-1 w, non-symmetric   (systems.py:599-601 dense; 626-633, 770-783)
  static MLB_DEV bool lmult_inv_jacob_product(const Jac& Jl, const Jac& Jr,
                                              const double (&w)[Y], double (&out)[Y]) {
    double rd[Y];
    double dprod[Y];
#pragma unroll
    for (int i = 0; i < Y; ++i) {
      double dv = FAM == FAM_HIER ? Jl.dv[i] * Jr.dv[i] : 0.0;
      dprod[i] = fma(Jl.dn[i], Jr.dn[i], dv);
    }
    rcp_many<Y>(dprod, rd);
    bool ok;
    if (DENSE) {
      double mat[K][K], b[K];
#pragma unroll
      for (int i = 0; i < Y; ++i) {
#pragma unroll
        for (int j = 0; j < Y; ++j) {
          double s = (i == j) ? dprod[i] : 0.0;
#pragma unroll
          for (int a = 0; a < U; ++a) s = fma(Jl.du[i][a], Jr.du[j][a], s);
          mat[i][j] = s;
        }
        b[i] = w[i];
      }
      ok = lu_solve<K>(mat, b);
#pragma unroll
      for (int i = 0; i < Y; ++i) out[i] = b[i];
    } else {
      double mat[K][K], t[K];
#pragma unroll
      for (int a = 0; a < U; ++a) {
#pragma unroll
        for (int b = 0; b < U; ++b) {
          mat[a][b] = dot_acc<Y>((a == b) ? 1.0 : 0.0, [&](int i) { return Jr.du[i][a]; },
                                 [&](int i) { return Jl.du[i][b] * rd[i]; });
        }
        t[a] = dot_acc<Y>(0.0, [&](int i) { return Jr.du[i][a]; },
                          [&](int i) { return w[i] * rd[i]; });
      }
      ok = lu_solve<K>(mat, t);
#pragma unroll
      for (int i = 0; i < Y; ++i) {
        double s = w[i];
#pragma unroll
        for (int a = 0; a < U; ++a) s = fma(-Jl.du[i][a], t[a], s);
        out[i] = s * rd[i];
      }
    }
    return ok;
  }

  // 0.5 log det(J J^T)   (systems.py:603-609 dense; 635-641, 785-794 Woodbury):
  // sum_k log chol_kk + 0.5 sum_i log d_i.  Logs are taken of PAIRWISE products (half
  // as many ~35-instruction log evaluations; a product of two factors overflows only
  // where the factors' own squares would).
  static MLB_DEV double log_det_sqrt_gram(const Gram& G) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k + 1 < K; k += 2) s += log(G.chol[k][k] * G.chol[k + 1][k + 1]);
    if (K & 1) s += log(G.chol[K - 1][K - 1]);
    if (!DENSE) {
      double t = 0.0;
#pragma unroll
      for (int i = 0; i + 1 < Y; i += 2) t += log(G.d[i] * G.d[i + 1]);
      if (Y & 1) t += log(G.d[Y - 1]);
      s = fma(0.5, t, s);
    }
    return s;
  }

  // J^T Gram^-1 w   (default composition, systems.py:173-179)
  static MLB_DEV void lmult_pinv(const Jac& J, const Gram& G, const double (&w)[Y],
                                 double (&out)[Q]) {
    double x[Y];
    lmult_inv_gram(J, G, w, x);
    rmult_jacob(J, x, out);
  }

  // p -= J^T Gram^-1 J p   (systems.py:181-188, 464-466)
  static MLB_DEV void project_cotangent(const Jac& J, const Gram& G, double (&p)[Q]) {
    double w[Y], nsc[Q];
    lmult_jacob(J, p, w);
    lmult_pinv(J, G, w, nsc);
#pragma unroll
    for (int k = 0; k < Q; ++k) p[k] -= nsc[k];
  }

  // ---- packed global-memory layout of the blocks (see mlb.h)
  static MLB_DEV void store_jac(const Jac& J, double* jac, int64_t ld, int64_t i) {
#pragma unroll
    for (int r = 0; r < Y; ++r)
#pragma unroll
      for (int a = 0; a < U; ++a) jac[(int64_t)(r * U + a) * ld + i] = J.du[r][a];
    int row = Y * U;
    if (FAM == FAM_HIER) {
#pragma unroll
      for (int r = 0; r < Y; ++r) jac[(int64_t)(row + r) * ld + i] = J.dv[r];
      row += Y;
    }
#pragma unroll
    for (int r = 0; r < Y; ++r) jac[(int64_t)(row + r) * ld + i] = J.dn[r];
  }
  static MLB_DEV void load_jac(Jac& J, const double* jac, int64_t ld, int64_t i) {
#pragma unroll
    for (int r = 0; r < Y; ++r)
#pragma unroll
      for (int a = 0; a < U; ++a) J.du[r][a] = jac[(int64_t)(r * U + a) * ld + i];
    int row = Y * U;
#pragma unroll
    for (int r = 0; r < Y; ++r) J.dv[r] = 0.0;
    if (FAM == FAM_HIER) {
#pragma unroll
      for (int r = 0; r < Y; ++r) J.dv[r] = jac[(int64_t)(row + r) * ld + i];
      row += Y;
    }
#pragma unroll
    for (int r = 0; r < Y; ++r) J.dn[r] = jac[(int64_t)(row + r) * ld + i];
  }
  static MLB_DEV void store_gram(const Gram& G, double* gram, int64_t ld, int64_t i) {
#pragma unroll
    for (int r = 0; r < K; ++r)
#pragma unroll
      for (int c = 0; c < K; ++c) gram[(int64_t)(r * K + c) * ld + i] = G.chol[r][c];
#pragma unroll
    for (int r = 0; r < Y; ++r) gram[(int64_t)(K * K + r) * ld + i] = G.d[r];
  }
  static MLB_DEV void load_gram(Gram& G, const double* gram, int64_t ld, int64_t i) {
#pragma unroll
    for (int r = 0; r < K; ++r)
#pragma unroll
      for (int c = 0; c < K; ++c) G.chol[r][c] = gram[(int64_t)(r * K + c) * ld + i];
#pragma unroll
    for (int r = 0; r < Y; ++r) {
      G.d[r] = gram[(int64_t)(K * K + r) * ld + i];
      G.rd[r] = rcp_fast(G.d[r]);
    }
#pragma unroll
    for (int r = 0; r < K; ++r) G.rdiag[r] = rcp_fast(G.chol[r][r]);
    G.ok = true;
  }
};

}  // namespace mlb
