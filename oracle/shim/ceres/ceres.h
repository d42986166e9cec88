// oracle/shim/ceres/ceres.h — TEST INFRASTRUCTURE.
// Minimal stand-in for the pieces of Ceres Solver 2.1.0 that the reference's
// src/BundleAdjustment/BundleAdjustment.h touches at COMPILE time: Jet, CostFunction,
// AutoDiffCostFunction and ceres::pow. Written from scratch following Ceres' documented Jet algebra
// (include/ceres/jet.h): it exists so that the reference header can be compiled unmodified, in place,
// by oracle/ref_bridge.cpp (Ceres itself is absent from this image).
#pragma once
#include <cmath>
#include <cstddef>
#include <type_traits>

namespace ceres {

template <typename T, int N>
struct Jet {
  T a;
  T v[N];
  Jet() : a() {
    for (int i = 0; i < N; ++i) v[i] = T();
  }
  Jet(const T& value) : a(value) {  // NOLINT (implicit, like Ceres)
    for (int i = 0; i < N; ++i) v[i] = T();
  }
  Jet(int value) : a(T(value)) {  // NOLINT: T(2) in templated Eigen code
    for (int i = 0; i < N; ++i) v[i] = T();
  }
  Jet(const T& value, int k) : a(value) {
    for (int i = 0; i < N; ++i) v[i] = T();
    v[k] = T(1.0);
  }
};

template <typename T, int N>
inline Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h;
  h.a = f.a + g.a;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i];
  return h;
}
template <typename T, int N>
inline Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h;
  h.a = f.a - g.a;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i];
  return h;
}
template <typename T, int N>
inline Jet<T, N> operator-(const Jet<T, N>& f) {
  Jet<T, N> h;
  h.a = -f.a;
  for (int i = 0; i < N; ++i) h.v[i] = -f.v[i];
  return h;
}
template <typename T, int N>
inline Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h;
  h.a = f.a * g.a;
  for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a;
  return h;
}
template <typename T, int N>
inline Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h;
  const T g_a_inverse = T(1.0) / g.a;
  const T f_a_by_g_a = f.a * g_a_inverse;
  h.a = f_a_by_g_a;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse;
  return h;
}
#define LFBA_JET_COMPOUND(op)                                        \
  template <typename T, int N>                                       \
  inline Jet<T, N>& operator op##=(Jet<T, N>& f, const Jet<T, N>& g) { \
    f = f op g;                                                      \
    return f;                                                        \
  }
LFBA_JET_COMPOUND(+)
LFBA_JET_COMPOUND(-)
LFBA_JET_COMPOUND(*)
LFBA_JET_COMPOUND(/)
#undef LFBA_JET_COMPOUND
#define LFBA_JET_CMP(op)                                           \
  template <typename T, int N>                                     \
  inline bool operator op(const Jet<T, N>& f, const Jet<T, N>& g) { \
    return f.a op g.a;                                             \
  }
LFBA_JET_CMP(<)
LFBA_JET_CMP(<=)
LFBA_JET_CMP(>)
LFBA_JET_CMP(>=)
LFBA_JET_CMP(==)
LFBA_JET_CMP(!=)
#undef LFBA_JET_CMP

template <typename T, int N>
inline Jet<T, N> sin(const Jet<T, N>& f) {
  Jet<T, N> h;
  h.a = std::sin(f.a);
  const T c = std::cos(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i];
  return h;
}
template <typename T, int N>
inline Jet<T, N> cos(const Jet<T, N>& f) {
  Jet<T, N> h;
  h.a = std::cos(f.a);
  const T ms = -std::sin(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = ms * f.v[i];
  return h;
}
template <typename T, int N>
inline Jet<T, N> sqrt(const Jet<T, N>& f) {
  Jet<T, N> h;
  h.a = std::sqrt(f.a);
  const T two_a_inverse = T(1.0) / (T(2.0) * h.a);
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * two_a_inverse;
  return h;
}
// general branch of ceres::pow(Jet, Jet) (the special cases concern f.a == 0 with g >= 1 and f.a < 0 with
// integer g, neither reachable from the reference's pow(x, T(0.5)) on a squared distance)
template <typename T, int N>
inline Jet<T, N> pow(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h;
  const T tmp1 = std::pow(f.a, g.a);
  const T tmp2 = g.a * std::pow(f.a, g.a - T(1.0));
  const T tmp3 = tmp1 * std::log(f.a);
  h.a = tmp1;
  for (int i = 0; i < N; ++i) h.v[i] = tmp2 * f.v[i] + tmp3 * g.v[i];
  return h;
}
inline double pow(double x, double y) { return std::pow(x, y); }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
inline double sqrt(double x) { return std::sqrt(x); }

class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
};

namespace internal {
template <int... Ns>
struct Sum;
template <>
struct Sum<> {
  static const int value = 0;
};
template <int N0, int... Ns>
struct Sum<N0, Ns...> {
  static const int value = N0 + Sum<Ns...>::value;
};
}  // namespace internal

// Forward-mode autodiff cost function: one Jet<double, sum(Ns)> per parameter, like Ceres' AutoDiff.
template <typename Functor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(Functor* f) : functor_(f) {}
  ~AutoDiffCostFunction() override { delete functor_; }

  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    const int sizes[] = {Ns...};
    const int nblocks = (int)sizeof...(Ns);
    if (jacobians == nullptr) return call(*functor_, parameters, residuals, nblocks);
    const int kTotal = internal::Sum<Ns...>::value;
    typedef Jet<double, internal::Sum<Ns...>::value> J;
    J x[internal::Sum<Ns...>::value];
    J out[kNumResiduals];
    const J* ptrs[sizeof...(Ns)];
    int off = 0;
    for (int b = 0; b < nblocks; ++b) {
      ptrs[b] = x + off;
      for (int j = 0; j < sizes[b]; ++j) x[off + j] = J(parameters[b][j], off + j);
      off += sizes[b];
    }
    if (!call(*functor_, ptrs, out, nblocks)) return false;
    for (int r = 0; r < kNumResiduals; ++r) residuals[r] = out[r].a;
    off = 0;
    for (int b = 0; b < nblocks; ++b) {
      if (jacobians[b] != nullptr)
        for (int r = 0; r < kNumResiduals; ++r)
          for (int j = 0; j < sizes[b]; ++j) jacobians[b][r * sizes[b] + j] = out[r].v[off + j];
      off += sizes[b];
    }
    (void)kTotal;
    return true;
  }

 private:
  template <typename T>
  static bool call(const Functor& f, T const* const* p, T* res, int nblocks) {
    return call_n(f, p, res, std::integral_constant<int, (int)sizeof...(Ns)>());
  }
  template <typename T>
  static bool call_n(const Functor& f, T const* const* p, T* res, std::integral_constant<int, 1>) {
    return f(p[0], res);
  }
  template <typename T>
  static bool call_n(const Functor& f, T const* const* p, T* res, std::integral_constant<int, 2>) {
    return f(p[0], p[1], res);
  }
  template <typename T>
  static bool call_n(const Functor& f, T const* const* p, T* res, std::integral_constant<int, 3>) {
    return f(p[0], p[1], p[2], res);
  }
  Functor* functor_;
};

}  // namespace ceres
