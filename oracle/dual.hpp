// oracle/dual.hpp — TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
//
// Forward-mode dual number with N partials, restating the arithmetic rules of ceres::Jet<double,N>
// (Ceres 2.1.0, include/ceres/jet.h — external dependency of the reference, pinned at
// /root/reference/installation/Dockerfile:105; not vendored, restated from its published definition).
// The reference differentiates its functors with Jet<double,26> through
// ceres::AutoDiffCostFunction (src/BundleAdjustment/BundleAdjustment.h:206,212,219,273).
#pragma once
#include <cmath>

namespace lfba_oracle {

template <int N>
struct Dual {
  double a;
  double v[N];

  Dual() : a(0.0) {
    for (int i = 0; i < N; ++i) v[i] = 0.0;
  }
  Dual(double s) : a(s) {  // NOLINT: implicit on purpose, mirrors T(0.5) in the functors
    for (int i = 0; i < N; ++i) v[i] = 0.0;
  }
  Dual(double s, int k) : a(s) {
    for (int i = 0; i < N; ++i) v[i] = 0.0;
    v[k] = 1.0;
  }
};

template <int N>
inline Dual<N> operator+(const Dual<N>& f, const Dual<N>& g) {
  Dual<N> h;
  h.a = f.a + g.a;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i];
  return h;
}
template <int N>
inline Dual<N> operator-(const Dual<N>& f, const Dual<N>& g) {
  Dual<N> h;
  h.a = f.a - g.a;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i];
  return h;
}
template <int N>
inline Dual<N> operator-(const Dual<N>& f) {
  Dual<N> h;
  h.a = -f.a;
  for (int i = 0; i < N; ++i) h.v[i] = -f.v[i];
  return h;
}
// product rule: (f g)' = f.a g' + f' g.a
template <int N>
inline Dual<N> operator*(const Dual<N>& f, const Dual<N>& g) {
  Dual<N> h;
  h.a = f.a * g.a;
  for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a;
  return h;
}
// quotient as Ceres does it: one reciprocal, then (f' - (f/g) g') / g
template <int N>
inline Dual<N> operator/(const Dual<N>& f, const Dual<N>& g) {
  Dual<N> h;
  const double g_inv = 1.0 / g.a;
  const double q = f.a * g_inv;
  h.a = q;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * g_inv;
  return h;
}
template <int N>
inline Dual<N>& operator+=(Dual<N>& f, const Dual<N>& g) {
  f = f + g;
  return f;
}
template <int N>
inline Dual<N>& operator-=(Dual<N>& f, const Dual<N>& g) {
  f = f - g;
  return f;
}
template <int N>
inline Dual<N>& operator*=(Dual<N>& f, const Dual<N>& g) {
  f = f * g;
  return f;
}
template <int N>
inline Dual<N>& operator/=(Dual<N>& f, const Dual<N>& g) {
  f = f / g;
  return f;
}
// comparisons look at the scalar part only (Jet semantics): `fL < T(0.0)`
template <int N>
inline bool operator<(const Dual<N>& f, const Dual<N>& g) {
  return f.a < g.a;
}
template <int N>
inline bool operator>(const Dual<N>& f, const Dual<N>& g) {
  return f.a > g.a;
}

template <int N>
inline Dual<N> sin(const Dual<N>& f) {
  Dual<N> h;
  h.a = std::sin(f.a);
  const double c = std::cos(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i];
  return h;
}
template <int N>
inline Dual<N> cos(const Dual<N>& f) {
  Dual<N> h;
  h.a = std::cos(f.a);
  const double ms = -std::sin(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = ms * f.v[i];
  return h;
}
// pow(f, g) for the general (non-special-cased) branch of ceres::pow(Jet, Jet): value pow(f.a, g.a),
// partials g.a * pow(f.a, g.a - 1) * f' + pow(f.a, g.a) * log(f.a) * g'.  The reference only calls it
// as ceres::pow(x, T(0.5)) (BundleAdjustment.h:264), where g' == 0.
template <int N>
inline Dual<N> pow(const Dual<N>& f, const Dual<N>& g) {
  Dual<N> h;
  const double t1 = std::pow(f.a, g.a);
  const double t2 = g.a * std::pow(f.a, g.a - 1.0);
  const double t3 = t1 * std::log(f.a);
  h.a = t1;
  for (int i = 0; i < N; ++i) h.v[i] = t2 * f.v[i] + t3 * g.v[i];
  return h;
}

// plain-double overloads so one templated functor serves cost-only and Jacobian evaluation
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
inline double pow(double x, double y) { return std::pow(x, y); }

}  // namespace lfba_oracle
