// One-chain-per-thread model definitions (closed forms; no autodiff on device).
//
// A model supplies what the reference obtains by tracing `generate_y` with JAX
// (mlift/systems.py:564-580, 694-710) plus the extended-prior density and the gradient
// of 0.5 log det(J J^T) that the reference gets by reverse-mode AD
// (mlift/systems.py:332-334).
#pragma once
#include "mlb_blocks.cuh"

namespace mlb {

// --------------------------------------------------------------------------- toy 2-D
// q = (theta0, theta1, eta); c = F(theta) + sigma eta - y;
// F = theta1^2 + k theta0^2 (theta0^2 - 1/2)   (reference toy_2d_model.py:43-45,80-83;
// k = 2 in the notebook).  nld = |q|^2 / 2 (toy_2d_model.py:72-74).  Additive-noise
// family with dim_u = 2 >= dim_y = 1: dense (scalar) Gram, same algebra as Mici's
// DenseConstrainedEuclideanMetricSystem(dens_wrt_hausdorff=False).
struct ToyModel {
  using B = Blocks<FAM_ADDITIVE, 2, 1, true>;
  static constexpr int Q = B::Q, Y = B::Y, U = B::U;
  struct Params {
    double sigma, y, k;
  };
  struct Par {};  // nothing to share between the evaluations at one position
  static MLB_DEV Par params(const double, const double) { return Par{}; }

  static MLB_DEV void constr(const Params& P, const double (&q)[Q], double (&c)[Y]) {
    double t0 = q[0] * q[0];
    c[0] = fma(q[1], q[1], P.k * t0 * (t0 - 0.5)) + fma(P.sigma, q[2], -P.y);
  }
  static MLB_DEV void jacob(const Params& P, const double (&q)[Q], B::Jac& J,
                            double (&c)[Y]) {
    constr(P, q, c);
    J.du[0][0] = P.k * (4.0 * q[0] * q[0] * q[0] - q[0]);
    J.du[0][1] = 2.0 * q[1];
    J.dv[0] = 0.0;
    J.dn[0] = P.sigma;
  }
  static MLB_DEV void jacob(const Params& P, const double (&q)[Q], const Par&, B::Jac& J,
                            double (&c)[Y]) {
    jacob(P, q, J, c);
  }
  static MLB_DEV double nld(const Params&, const double (&q)[Q], const Par&, double (&g)[Q],
                            bool = true) {
    g[0] = q[0], g[1] = q[1], g[2] = q[2];
    return 0.5 * (q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  }
  // adds d(0.5 log det G)/dq to g
  static MLB_DEV void add_grad_logdet(const Params& P, const double (&q)[Q], const Par&,
                                      const B::Jac& J, const B::Gram& G, double (&g)[Q]) {
    double rg = G.rdiag[0] * G.rdiag[0];
    g[0] += J.du[0][0] * P.k * (12.0 * q[0] * q[0] - 1.0) * rg;
    g[1] += J.du[0][1] * 2.0 * rg;
  }
};

// --------------------------------------------------------------------- eight schools
// y_i = mu(u0) + tau(u1) v_i + sigma_i n_i      (reference eight_schools.py:38-46)
// mu ~ N(0,5), tau ~ half-Cauchy(5)             (eight_schools.py:28-31)
// unbounded: mu = u0, tau = exp(u1)             (prior.py:17-42, transforms.py:43-55)
// normal:    mu = 5 u0, tau = 5 tan(pi((Phi(u1)+1)/2 - 1/2))
//                                               (prior.py:45-55, transforms.py:87-95,137-144)
template <int Y_, int PARAM>
struct HierModel {
  using B = Blocks<FAM_HIER, 2, Y_, (2 >= Y_)>;
  static constexpr int Q = B::Q, Y = B::Y, U = B::U;
  struct Params {
    double y_obs[Y];
    double sigma[Y];
  };
  struct Par {
    double mu, dmu, ddmu, tau, dtau, ddtau;
  };

  static MLB_DEV Par params(const double u0, const double u1) {
    Par p;
    if (PARAM == MLB_PARAM_UNBOUNDED) {
      p.mu = u0, p.dmu = 1.0, p.ddmu = 0.0;
      p.tau = exp(u1), p.dtau = p.tau, p.ddtau = p.tau;
    } else {
      const double pi = 3.14159265358979323846;
      p.mu = 5.0 * u0, p.dmu = 5.0, p.ddmu = 0.0;
      double Phi = 0.5 * erfc(-u1 * 0.70710678118654752440);
      double phi = exp(-0.5 * u1 * u1) * 0.39894228040143267794;
      double arg = pi * ((Phi + 1.0) / 2.0 - 0.5);
      double t = tan(arg), sec2 = fma(t, t, 1.0);
      double darg = 0.5 * pi * phi, ddarg = -u1 * darg;
      p.tau = 5.0 * t;
      p.dtau = 5.0 * sec2 * darg;
      p.ddtau = 5.0 * (2.0 * sec2 * t * darg * darg + sec2 * ddarg);
    }
    return p;
  }

  static MLB_DEV void constr(const Params& P, const double (&q)[Q], double (&c)[Y]) {
    Par p = params(q[0], q[1]);
#pragma unroll
    for (int i = 0; i < Y; ++i)
      c[i] = fma(p.tau, q[2 + i], p.mu) + fma(P.sigma[i], q[2 + Y + i], -P.y_obs[i]);
  }
  static MLB_DEV void jacob(const Params& P, const double (&q)[Q], typename B::Jac& J,
                            double (&c)[Y]) {
    jacob(P, q, params(q[0], q[1]), J, c);
  }
  static MLB_DEV void jacob(const Params& P, const double (&q)[Q], const Par& p,
                            typename B::Jac& J, double (&c)[Y]) {
#pragma unroll
    for (int i = 0; i < Y; ++i) {
      c[i] = fma(p.tau, q[2 + i], p.mu) + fma(P.sigma[i], q[2 + Y + i], -P.y_obs[i]);
      J.du[i][0] = p.dmu;
      J.du[i][1] = p.dtau * q[2 + i];
      J.dv[i] = p.tau;
      J.dn[i] = P.sigma[i];
    }
  }
  // extended prior nld (eight_schools.py:49-53) with gradient
  // want_val = false: gradient only (the value costs a log1p; the fused trajectory needs it
  // at its first and last point only)
  static MLB_DEV double nld(const Params&, const double (&q)[Q], const Par& p,
                            double (&g)[Q], bool want_val = true) {
    double val;
    if (PARAM == MLB_PARAM_UNBOUNDED) {
      // normal(0,5) on u0; half-Cauchy(5) pulled back through tau = exp(u1):
      // log1p((tau/5)^2) - log(d tau/d u1)   (distributions.py:78-80, 449-450), and
      // log(d tau / d u1) = log(exp(u1)) = u1
      double r = p.tau * 0.2, r2 = r * r;
      double s = q[0] * 0.2;
      val = 0.5 * s * s + (want_val ? log1p(r2) : 0.0) - q[1];
      g[0] = q[0] * 0.04;
      g[1] = fma(2.0 * r2, rcp_fast(1.0 + r2), -1.0);
    } else {
      val = 0.5 * q[0] * q[0] + 0.5 * q[1] * q[1];
      g[0] = q[0], g[1] = q[1];
    }
    double s = 0.0, t = 0.0;
#pragma unroll
    for (int i = 0; i < Y; ++i) {
      s = fma(q[2 + i], q[2 + i], s);
      t = fma(q[2 + Y + i], q[2 + Y + i], t);
      g[2 + i] = q[2 + i];
      g[2 + Y + i] = q[2 + Y + i];
    }
    return val + 0.5 * (s + t);
  }
  // d(0.5 log det G) = sum_ia B_ia dA_ia + 0.5 sum_i (G^-1)_ii dD_i with B = G^-1 A
  // = D^-1 A C^-1 (Woodbury).  Only mu(u0), tau(u1) have second derivatives.
  static MLB_DEV void add_grad_logdet(const Params&, const double (&q)[Q], const Par& p,
                                      const typename B::Jac& J, const typename B::Gram& G,
                                      double (&g)[Q]) {
    static_assert(!B::DENSE, "hierarchical model needs dim_y > dim_u");
    double s0 = 0.0, s1 = 0.0, sd = 0.0;
#pragma unroll
    for (int i = 0; i < Y; ++i) {
      double x[2] = {J.du[i][0], J.du[i][1]};
      chol_solve<2>(G.chol, G.rdiag, x);
      double b0 = x[0] * G.rd[i], b1 = x[1] * G.rd[i];
      double gii = G.rd[i] - (J.du[i][0] * x[0] + J.du[i][1] * x[1]) * (G.rd[i] * G.rd[i]);
      s0 += b0;
      s1 = fma(b1, q[2 + i], s1);
      sd += gii;
      g[2 + i] = fma(p.dtau, b1, g[2 + i]);
    }
    g[0] = fma(p.ddmu, s0, g[0]);
    g[1] += p.ddtau * s1 + p.tau * p.dtau * sd;
  }
};

}  // namespace mlb
