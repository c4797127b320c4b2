// Forward-mode jets on the device: value, N first derivatives and a compile-time SUBSET of
// NP second derivatives (pairs P0 .. P0+NP-1 of the upper triangle), all in registers
// with static indexing.  Used by ODE models whose right-hand side is too intricate to
// differentiate by hand twice (Hodgkin-Huxley); a second-order pass is split into pair
// groups so that one group's jets fit the register file, the first-order part being
// recomputed per group.  (The reference gets the same derivatives from jax.jacfwd and
// reverse-over-forward autodiff of the traced model, mlift/systems.py:332-334, 568-580.)
#pragma once
#include <utility>

#include "mlb_common.cuh"

namespace mlb {

// pair index k <-> (a, b), a <= b, row-major over the upper triangle of an N x N matrix
template <int N>
struct PairMap {
  static __host__ __device__ constexpr int a(int k) {
    int r = 0;
    while (k >= N - r) k -= N - r, ++r;
    return r;
  }
  static __host__ __device__ constexpr int b(int k) {
    int r = 0;
    while (k >= N - r) k -= N - r, ++r;
    return r + k;
  }
};

template <int N, int NP, int P0>
struct DJet {
  double v;
  double g[N];
  double h[NP ? NP : 1];
};

template <int N, int NP, int P0, class F>
MLB_DEV void for_pairs(F&& f) {
  if constexpr (NP > 0) {
    [&]<int... K>(std::integer_sequence<int, K...>) {
      (f.template operator()<K, PairMap<N>::a(P0 + K), PairMap<N>::b(P0 + K)>(), ...);
    }(std::make_integer_sequence<int, NP>{});
  }
}

#define MLB_JET_T template <int N, int NP, int P0>
#define MLB_JET DJet<N, NP, P0>

MLB_JET_T MLB_DEV MLB_JET jconst(double c) {
  MLB_JET r;
  r.v = c;
#pragma unroll
  for (int a = 0; a < N; ++a) r.g[a] = 0.0;
#pragma unroll
  for (int k = 0; k < (NP ? NP : 1); ++k) r.h[k] = 0.0;
  return r;
}
MLB_JET_T MLB_DEV MLB_JET operator+(const MLB_JET& x, const MLB_JET& y) {
  MLB_JET r;
  r.v = x.v + y.v;
#pragma unroll
  for (int a = 0; a < N; ++a) r.g[a] = x.g[a] + y.g[a];
#pragma unroll
  for (int k = 0; k < NP; ++k) r.h[k] = x.h[k] + y.h[k];
  return r;
}
MLB_JET_T MLB_DEV MLB_JET operator-(const MLB_JET& x, const MLB_JET& y) {
  MLB_JET r;
  r.v = x.v - y.v;
#pragma unroll
  for (int a = 0; a < N; ++a) r.g[a] = x.g[a] - y.g[a];
#pragma unroll
  for (int k = 0; k < NP; ++k) r.h[k] = x.h[k] - y.h[k];
  return r;
}
MLB_JET_T MLB_DEV MLB_JET operator-(const MLB_JET& x) {
  MLB_JET r;
  r.v = -x.v;
#pragma unroll
  for (int a = 0; a < N; ++a) r.g[a] = -x.g[a];
#pragma unroll
  for (int k = 0; k < NP; ++k) r.h[k] = -x.h[k];
  return r;
}
MLB_JET_T MLB_DEV MLB_JET operator*(const MLB_JET& x, const MLB_JET& y) {
  MLB_JET r;
  r.v = x.v * y.v;
#pragma unroll
  for (int a = 0; a < N; ++a) r.g[a] = fma(x.g[a], y.v, x.v * y.g[a]);
  for_pairs<N, NP, P0>([&]<int K, int A, int B>() {
    r.h[K] = fma(x.h[K], y.v, fma(x.g[A], y.g[B], fma(x.g[B], y.g[A], x.v * y.h[K])));
  });
  return r;
}
MLB_JET_T MLB_DEV MLB_JET operator*(double s, const MLB_JET& x) {
  MLB_JET r;
  r.v = s * x.v;
#pragma unroll
  for (int a = 0; a < N; ++a) r.g[a] = s * x.g[a];
#pragma unroll
  for (int k = 0; k < NP; ++k) r.h[k] = s * x.h[k];
  return r;
}
MLB_JET_T MLB_DEV MLB_JET operator*(const MLB_JET& x, double s) { return s * x; }
MLB_JET_T MLB_DEV MLB_JET operator+(const MLB_JET& x, double c) {
  MLB_JET r = x;
  r.v += c;
  return r;
}
MLB_JET_T MLB_DEV MLB_JET operator+(double c, const MLB_JET& x) { return x + c; }
MLB_JET_T MLB_DEV MLB_JET operator-(const MLB_JET& x, double c) { return x + (-c); }
MLB_JET_T MLB_DEV MLB_JET operator-(double c, const MLB_JET& x) { return (-x) + c; }

// f(x) given f, f', f'' at x.v
MLB_JET_T MLB_DEV MLB_JET jchain(const MLB_JET& x, double f, double df, double ddf) {
  MLB_JET r;
  r.v = f;
#pragma unroll
  for (int a = 0; a < N; ++a) r.g[a] = df * x.g[a];
  for_pairs<N, NP, P0>([&]<int K, int A, int B>() {
    r.h[K] = fma(df, x.h[K], ddf * x.g[A] * x.g[B]);
  });
  return r;
}
MLB_JET_T MLB_DEV MLB_JET jexp(const MLB_JET& x) {
  const double e = exp_bl(x.v);
  return jchain(x, e, e, e);
}
MLB_JET_T MLB_DEV MLB_JET jexpm1(const MLB_JET& x) {
  const double e = expm1_bl(x.v);
  return jchain(x, e, e + 1.0, e + 1.0);
}
MLB_JET_T MLB_DEV MLB_JET jrecip(const MLB_JET& x) {
  const double r = rcp_fast(x.v), r2 = r * r;
  return jchain(x, r, -r2, 2.0 * r2 * r);
}
MLB_JET_T MLB_DEV MLB_JET operator/(const MLB_JET& x, const MLB_JET& y) { return x * jrecip(y); }
MLB_JET_T MLB_DEV MLB_JET operator/(double c, const MLB_JET& y) { return c * jrecip(y); }

// order-0 versions so that model code can be written once
MLB_DEV double jexp(double x) { return exp_bl(x); }
MLB_DEV double jexpm1(double x) { return expm1_bl(x); }
MLB_DEV double jrecip(double x) { return rcp_fast(x); }

// A scalar / jet as JetSize<T>::value doubles at base[j * stride] (exchange of recurrence
// state between the warps that share a chain batch)
template <class T>
struct JetSize {
  static constexpr int value = 1;
};
MLB_JET_T struct JetSize<MLB_JET> {
  static constexpr int value = 1 + N + NP;
};
MLB_DEV void jet_store(const double& x, double* base, int) { base[0] = x; }
MLB_DEV void jet_load(double& x, const double* base, int) { x = base[0]; }
MLB_JET_T MLB_DEV void jet_store(const MLB_JET& x, double* base, int stride) {
  base[0] = x.v;
#pragma unroll
  for (int a = 0; a < N; ++a) base[(1 + a) * stride] = x.g[a];
#pragma unroll
  for (int k = 0; k < NP; ++k) base[(1 + N + k) * stride] = x.h[k];
}
MLB_JET_T MLB_DEV void jet_load(MLB_JET& x, const double* base, int stride) {
  x.v = base[0];
#pragma unroll
  for (int a = 0; a < N; ++a) x.g[a] = base[(1 + a) * stride];
#pragma unroll
  for (int k = 0; k < NP; ++k) x.h[k] = base[(1 + N + k) * stride];
}

// constants / seeded variables of either scalar type
template <class T>
struct JetOps;
template <>
struct JetOps<double> {
  static MLB_DEV double constant(double c) { return c; }
  static MLB_DEV double seed(double val, int, double) { return val; }
};
MLB_JET_T struct JetOps<MLB_JET> {
  static MLB_DEV MLB_JET constant(double c) { return jconst<N, NP, P0>(c); }
  // variable with value `val` whose derivative in direction d is `slope`
  static MLB_DEV MLB_JET seed(double val, int d, double slope) {
    MLB_JET r = jconst<N, NP, P0>(val);
#pragma unroll
    for (int a = 0; a < N; ++a) r.g[a] = (a == d) ? slope : 0.0;
    return r;
  }
};

#undef MLB_JET_T
#undef MLB_JET

}  // namespace mlb
