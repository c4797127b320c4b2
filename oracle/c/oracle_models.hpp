// Model definitions for the CPU oracle (closed forms).  TEST INFRASTRUCTURE ONLY.
//
// Each model supplies the constraint, its Jacobian blocks, the extended-prior negative
// log density with gradient, and the gradient of 0.5*log det(Gram).  The closed forms
// are derived from the reference's model code (cited per model) and are checked
// against the autodiff restatement in oracle/mlift_oracle by tests/test_oracle_*.py.
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

namespace orc {

enum { FAMILY_ADDITIVE = 0, FAMILY_HIER = 1 };
enum { MODEL_TOY = 0, MODEL_HIER = 1, MODEL_FN = 2, MODEL_HH = 3 };
enum { PARAM_UNBOUNDED = 0, PARAM_NORMAL = 1 };

struct Model {
  int family = FAMILY_ADDITIVE, U = 0, Y = 0, Q = 0;
  bool dense = false;
  virtual ~Model() {}
  virtual void constr(const double* q, double* c) const = 0;
  // dy_du [Y][U] row-major, dy_dv [Y] (hier only), dy_dn [Y], c [Y]
  virtual void jacob(const double* q, double* dy_du, double* dy_dv, double* dy_dn,
                     double* c) const = 0;
  // returns nld; writes gradient if grad != nullptr
  virtual double nld(const double* q, double* grad) const = 0;
  // gradient of 0.5 log det (J J^T) with respect to q
  virtual void grad_logdet(const double* q, const double* dy_du, const double* dy_dv,
                           const double* dy_dn, const double* chol, const double* d,
                           double* g) const = 0;
};

// ---------------------------------------------------------------------------- toy
// q = (theta0, theta1, eta), c = F(theta) + sigma*eta - y,
// F = theta1^2 + k*theta0^2*(theta0^2 - 1/2)      (toy_2d_model.py:43-45, 80-83;
// k = 2 in notebooks/Two-dimensional_example.ipynb), nld = |q|^2/2 (:72-74).
// Treated as an additive-noise system with dim_u = 2 >= dim_y = 1, i.e. the dense
// Gram branch; identical algebra to Mici's DenseConstrainedEuclideanMetricSystem
// with dens_wrt_hausdorff=False (toy_2d_model.py:175-182).
struct ToyModel : Model {
  double sigma, y, k;
  ToyModel(double sigma_, double y_, double k_) : sigma(sigma_), y(y_), k(k_) {
    family = FAMILY_ADDITIVE, U = 2, Y = 1, Q = 3, dense = true;
  }
  void constr(const double* q, double* c) const override {
    c[0] = q[1] * q[1] + k * q[0] * q[0] * (q[0] * q[0] - 0.5) + sigma * q[2] - y;
  }
  void jacob(const double* q, double* dy_du, double*, double* dy_dn,
             double* c) const override {
    constr(q, c);
    dy_du[0] = k * (4.0 * q[0] * q[0] * q[0] - q[0]);
    dy_du[1] = 2.0 * q[1];
    dy_dn[0] = sigma;
  }
  double nld(const double* q, double* grad) const override {
    if (grad) grad[0] = q[0], grad[1] = q[1], grad[2] = q[2];
    return 0.5 * (q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  }
  void grad_logdet(const double* q, const double* dy_du, const double*, const double*,
                   const double* chol, const double*, double* g) const override {
    double gram = chol[0] * chol[0];
    g[0] = dy_du[0] * k * (12.0 * q[0] * q[0] - 1.0) / gram;
    g[1] = dy_du[1] * 2.0 / gram;
    g[2] = 0.0;
  }
};

// ------------------------------------------------------------------ eight schools
// y_i = mu(u0) + tau(u1) v_i + sigma_i n_i          (eight_schools.py:38-46)
// priors mu ~ N(0,5), tau ~ half-Cauchy(5)           (eight_schools.py:28-31)
// unbounded: mu = u0, tau = exp(u1)                  (prior.py:17-42, transforms.py:43-55)
// normal:    mu = 5 u0, tau = 5 tan(pi((Phi(u1)+1)/2 - 1/2))
//                                                    (prior.py:45-55, transforms.py:87-95,137-144)
struct HierModel : Model {
  std::vector<double> y_obs, sigma;
  int param;
  HierModel(int Y_, int param_, const double* y_, const double* s_) : param(param_) {
    family = FAMILY_HIER, U = 2, Y = Y_, Q = 2 + 2 * Y_, dense = (U >= Y);
    y_obs.assign(y_, y_ + Y_), sigma.assign(s_, s_ + Y_);
  }
  struct Par {
    double mu, dmu, ddmu, tau, dtau, ddtau;
  };
  Par params(const double* u) const {
    Par p;
    if (param == PARAM_UNBOUNDED) {
      p.mu = u[0], p.dmu = 1.0, p.ddmu = 0.0;
      p.tau = std::exp(u[1]), p.dtau = p.tau, p.ddtau = p.tau;
    } else {
      const double pi = 3.14159265358979323846;
      p.mu = 0.0 + 5.0 * u[0], p.dmu = 5.0, p.ddmu = 0.0;
      double Phi = 0.5 * std::erfc(-u[1] * 0.70710678118654752440);
      double phi = std::exp(-0.5 * u[1] * u[1]) * 0.39894228040143267794;
      double arg = pi * ((Phi + 1.0) / 2.0 - 0.5);
      double t = std::tan(arg), sec2 = 1.0 + t * t;
      double darg = 0.5 * pi * phi, ddarg = -u[1] * darg;
      p.tau = 5.0 * t;
      p.dtau = 5.0 * sec2 * darg;
      p.ddtau = 5.0 * (2.0 * sec2 * t * darg * darg + sec2 * ddarg);
    }
    return p;
  }
  void constr(const double* q, double* c) const override {
    Par p = params(q);
    const double *v = q + 2, *n = q + 2 + Y;
    for (int i = 0; i < Y; ++i) c[i] = p.mu + p.tau * v[i] + sigma[i] * n[i] - y_obs[i];
  }
  void jacob(const double* q, double* dy_du, double* dy_dv, double* dy_dn,
             double* c) const override {
    Par p = params(q);
    const double *v = q + 2, *n = q + 2 + Y;
    for (int i = 0; i < Y; ++i) {
      c[i] = p.mu + p.tau * v[i] + sigma[i] * n[i] - y_obs[i];
      dy_du[2 * i] = p.dmu, dy_du[2 * i + 1] = p.dtau * v[i];
      dy_dv[i] = p.tau, dy_dn[i] = sigma[i];
    }
  }
  double nld(const double* q, double* grad) const override {
    Par p = params(q);
    double val;
    if (param == PARAM_UNBOUNDED) {
      double r = p.tau / 5.0;
      val = (q[0] / 5.0) * (q[0] / 5.0) / 2.0 + std::log1p(r * r) - std::log(p.tau);
      if (grad) grad[0] = q[0] / 25.0, grad[1] = 2.0 * r * r / (1.0 + r * r) - 1.0;
    } else {
      val = 0.5 * q[0] * q[0] + 0.5 * q[1] * q[1];
      if (grad) grad[0] = q[0], grad[1] = q[1];
    }
    double s = 0.0, t = 0.0;
    for (int i = 0; i < Y; ++i) s += q[2 + i] * q[2 + i], t += q[2 + Y + i] * q[2 + Y + i];
    if (grad)
      for (int i = 0; i < 2 * Y; ++i) grad[2 + i] = q[2 + i];
    return val + s / 2.0 + t / 2.0;
  }
  // d(0.5 log det G) = sum_ia B_ia dA_ia + 0.5 sum_i (G^-1)_ii dD_i,  B = G^-1 A.
  // Uses G^-1 computed column by column through whichever factorisation is stored.
  void grad_logdet(const double* q, const double* dy_du, const double* dy_dv,
                   const double* dy_dn, const double* chol, const double* d,
                   double* g) const override {
    Par p = params(q);
    const double* v = q + 2;
    std::vector<double> Ginv_diag(Y), B((size_t)Y * 2);
    if (!dense) {
      // Woodbury: G^-1 = D^-1 - D^-1 A C^-1 A^T D^-1 ;  B = D^-1 A C^-1
      double L00 = chol[0], L10 = chol[2], L11 = chol[3];
      auto csolve = [&](double b0, double b1, double& x0, double& x1) {
        double z0 = b0 / L00, z1 = (b1 - L10 * z0) / L11;
        x1 = z1 / L11, x0 = (z0 - L10 * x1) / L00;
      };
      for (int i = 0; i < Y; ++i) {
        double a0 = dy_du[2 * i], a1 = dy_du[2 * i + 1], x0, x1;
        csolve(a0, a1, x0, x1);
        B[2 * i] = x0 / d[i], B[2 * i + 1] = x1 / d[i];
        Ginv_diag[i] = 1.0 / d[i] - (a0 * x0 + a1 * x1) / (d[i] * d[i]);
      }
    } else {
      // dense Y x Y Cholesky: solve for each unit vector
      std::vector<double> col(Y);
      std::vector<double> Ginv((size_t)Y * Y);
      for (int j = 0; j < Y; ++j) {
        for (int i = 0; i < Y; ++i) col[i] = (i == j) ? 1.0 : 0.0;
        for (int i = 0; i < Y; ++i) {
          double t = col[i];
          for (int k = 0; k < i; ++k) t -= chol[(size_t)i * Y + k] * col[k];
          col[i] = t / chol[(size_t)i * Y + i];
        }
        for (int i = Y - 1; i >= 0; --i) {
          double t = col[i];
          for (int k = i + 1; k < Y; ++k) t -= chol[(size_t)k * Y + i] * col[k];
          col[i] = t / chol[(size_t)i * Y + i];
        }
        for (int i = 0; i < Y; ++i) Ginv[(size_t)i * Y + j] = col[i];
      }
      for (int i = 0; i < Y; ++i) {
        Ginv_diag[i] = Ginv[(size_t)i * Y + i];
        double b0 = 0.0, b1 = 0.0;
        for (int j = 0; j < Y; ++j)
          b0 += Ginv[(size_t)i * Y + j] * dy_du[2 * j],
              b1 += Ginv[(size_t)i * Y + j] * dy_du[2 * j + 1];
        B[2 * i] = b0, B[2 * i + 1] = b1;
      }
    }
    (void)dy_dv, (void)dy_dn;
    double s0 = 0.0, s1 = 0.0, sd = 0.0;
    for (int i = 0; i < Y; ++i) {
      s0 += B[2 * i];
      s1 += B[2 * i + 1] * v[i];
      sd += Ginv_diag[i];
      g[2 + i] = p.dtau * B[2 * i + 1];
      g[2 + Y + i] = 0.0;
    }
    g[0] = p.ddmu * s0;
    g[1] = p.ddtau * s1 + p.tau * p.dtau * sd;
  }
};

Model* make_fn_model(int dim_y, int param, const double* data, int n_data);
Model* make_hh_model(int dim_y, int param, const double* data, int n_data);

inline Model* make_model(int model_id, int dim_y, int param, const double* data,
                         int n_data) {
  switch (model_id) {
    case MODEL_TOY:
      return n_data >= 3 ? new ToyModel(data[0], data[1], data[2]) : nullptr;
    case MODEL_HIER:
      return n_data >= 2 * dim_y ? new HierModel(dim_y, param, data, data + dim_y)
                                 : nullptr;
    case MODEL_FN:
      return make_fn_model(dim_y, param, data, n_data);
    case MODEL_HH:
      return make_hh_model(dim_y, param, data, n_data);
  }
  return nullptr;
}

}  // namespace orc
