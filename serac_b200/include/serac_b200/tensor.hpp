// serac_b200: host/device tensor, dual and tuple algebra used inside q-functions.
//
// This is the device-ready counterpart of the reference's arithmetic vocabulary
//   tensor / zero           (src/serac/numerics/functional/tensor.hpp)
//   dual                    (src/serac/numerics/functional/dual.hpp)
//   tuple / get             (src/serac/numerics/functional/tuple.hpp)
//   get_value / dual maths  (src/serac/numerics/functional/tuple_tensor_dual_functions.hpp)
// written from scratch for nvcc (sm_100a).  The user-visible names (serac::tensor, serac::tuple,
// serac::get<i>, dot, inv, det, tr, sym, dev, transpose, detApIm1, Identity, ...) are kept so that the
// reference's material models and q-function functors compile unchanged; the representation differs:
//
//  * forward-mode derivatives are carried by `dual<N>`: a value plus a FLAT array of N partials that the
//    evaluation kernel seeds in chunks (see domain_integral.cuh).  Structurally-zero derivative blocks
//    (what the reference expresses with tuple<zero, tensor<...>> gradient types,
//    tuple_tensor_dual_functions.hpp:124-163) fall out as plain-double outputs of a pass and are never stored.
//  * every loop has compile-time bounds and is `#pragma unroll`ed, so tensors live in registers.
#pragma once

#include <cmath>
#include <cstdint>
#include <type_traits>
#include <utility>

#if defined(__CUDACC__)
#define SERAC_HOST_DEVICE __host__ __device__
#define SERAC_DEVICE __device__
#define SB_UNROLL _Pragma("unroll")
#else
#define SERAC_HOST_DEVICE
#define SERAC_DEVICE
#define SB_UNROLL
#endif
#define SB_HD SERAC_HOST_DEVICE inline constexpr
#define SB_HDF SERAC_HOST_DEVICE inline

namespace serac {

// ------------------------------------------------------------------------------------------------
// zero: a value that is structurally nothing (reference: tensor.hpp:123-258)
// ------------------------------------------------------------------------------------------------
struct zero {
  SB_HD operator double() const { return 0.0; }
  template <typename... I>
  SB_HD zero operator()(I...) const
  {
    return zero{};
  }
  template <typename T>
  SB_HD zero& operator=(const T&)
  {
    return *this;
  }
};

template <typename T>
struct is_zero : std::false_type {
};
template <>
struct is_zero<zero> : std::true_type {
};

SB_HD zero operator+(zero, zero) { return {}; }
SB_HD zero operator-(zero, zero) { return {}; }
SB_HD zero operator-(zero) { return {}; }
SB_HD zero operator*(zero, zero) { return {}; }
template <typename T, typename = std::enable_if_t<!is_zero<T>::value>>
SB_HD T operator+(zero, const T& b)
{
  return b;
}
template <typename T, typename = std::enable_if_t<!is_zero<T>::value>>
SB_HD T operator+(const T& a, zero)
{
  return a;
}
template <typename T, typename = std::enable_if_t<!is_zero<T>::value>>
SB_HD auto operator-(zero, const T& b)
{
  return -b;
}
template <typename T, typename = std::enable_if_t<!is_zero<T>::value>>
SB_HD T operator-(const T& a, zero)
{
  return a;
}
template <typename T, typename = std::enable_if_t<!is_zero<T>::value>>
SB_HD zero operator*(zero, const T&)
{
  return {};
}
template <typename T, typename = std::enable_if_t<!is_zero<T>::value>>
SB_HD zero operator*(const T&, zero)
{
  return {};
}
template <typename T, typename = std::enable_if_t<!is_zero<T>::value>>
SB_HD zero operator/(zero, const T&)
{
  return {};
}
template <typename T>
SB_HD zero dot(zero, const T&)
{
  return {};
}
template <typename T, typename = std::enable_if_t<!is_zero<T>::value>>
SB_HD zero dot(const T&, zero)
{
  return {};
}

// ------------------------------------------------------------------------------------------------
// dual<N>: value + N partial derivatives (forward mode)
// ------------------------------------------------------------------------------------------------
template <int N>
struct dual {
  double value;
  double gradient[N];
};

template <typename T>
struct is_dual_number : std::false_type {
};
template <int N>
struct is_dual_number<dual<N>> : std::true_type {
};

template <typename T>
inline constexpr bool is_scalar_v = std::is_arithmetic<T>::value || is_dual_number<T>::value;

/// number of partials carried by a scalar type (0 for double)
template <typename T>
struct num_partials {
  static constexpr int value = 0;
};
template <int N>
struct num_partials<dual<N>> {
  static constexpr int value = N;
};

SB_HD double get_value(double x) { return x; }
SB_HD zero get_value(zero) { return {}; }
template <int N>
SB_HD double get_value(const dual<N>& x)
{
  return x.value;
}

template <int N>
SB_HD dual<N> make_dual_constant(double v)
{
  dual<N> r{};
  r.value = v;
  return r;
}

// --- dual (+,-) ---
template <int N>
SB_HD dual<N> operator+(const dual<N>& a, const dual<N>& b)
{
  dual<N> r{};
  r.value = a.value + b.value;
  SB_UNROLL
  for (int i = 0; i < N; i++) r.gradient[i] = a.gradient[i] + b.gradient[i];
  return r;
}
template <int N>
SB_HD dual<N> operator+(const dual<N>& a, double b)
{
  dual<N> r = a;
  r.value += b;
  return r;
}
template <int N>
SB_HD dual<N> operator+(double a, const dual<N>& b)
{
  dual<N> r = b;
  r.value += a;
  return r;
}
template <int N>
SB_HD dual<N> operator-(const dual<N>& a)
{
  dual<N> r{};
  r.value = -a.value;
  SB_UNROLL
  for (int i = 0; i < N; i++) r.gradient[i] = -a.gradient[i];
  return r;
}
template <int N>
SB_HD dual<N> operator-(const dual<N>& a, const dual<N>& b)
{
  dual<N> r{};
  r.value = a.value - b.value;
  SB_UNROLL
  for (int i = 0; i < N; i++) r.gradient[i] = a.gradient[i] - b.gradient[i];
  return r;
}
template <int N>
SB_HD dual<N> operator-(const dual<N>& a, double b)
{
  dual<N> r = a;
  r.value -= b;
  return r;
}
template <int N>
SB_HD dual<N> operator-(double a, const dual<N>& b)
{
  dual<N> r{};
  r.value = a - b.value;
  SB_UNROLL
  for (int i = 0; i < N; i++) r.gradient[i] = -b.gradient[i];
  return r;
}
// --- dual (*,/) ---
template <int N>
SB_HD dual<N> operator*(const dual<N>& a, const dual<N>& b)
{
  dual<N> r{};
  r.value = a.value * b.value;
  SB_UNROLL
  for (int i = 0; i < N; i++) r.gradient[i] = a.value * b.gradient[i] + a.gradient[i] * b.value;
  return r;
}
template <int N>
SB_HD dual<N> operator*(const dual<N>& a, double b)
{
  dual<N> r{};
  r.value = a.value * b;
  SB_UNROLL
  for (int i = 0; i < N; i++) r.gradient[i] = a.gradient[i] * b;
  return r;
}
template <int N>
SB_HD dual<N> operator*(double a, const dual<N>& b)
{
  return b * a;
}
template <int N>
SB_HD dual<N> operator/(const dual<N>& a, double b)
{
  const double inv = 1.0 / b;
  dual<N>      r{};
  r.value = a.value * inv;
  SB_UNROLL
  for (int i = 0; i < N; i++) r.gradient[i] = a.gradient[i] * inv;
  return r;
}
template <int N>
SB_HD dual<N> operator/(double a, const dual<N>& b)
{
  const double inv = 1.0 / b.value;
  const double s   = -a * inv * inv;
  dual<N>      r{};
  r.value = a * inv;
  SB_UNROLL
  for (int i = 0; i < N; i++) r.gradient[i] = s * b.gradient[i];
  return r;
}
template <int N>
SB_HD dual<N> operator/(const dual<N>& a, const dual<N>& b)
{
  const double inv = 1.0 / b.value;
  const double q   = a.value * inv;
  dual<N>      r{};
  r.value = q;
  SB_UNROLL
  for (int i = 0; i < N; i++) r.gradient[i] = (a.gradient[i] - q * b.gradient[i]) * inv;
  return r;
}
template <int N, typename T>
SB_HD dual<N>& operator+=(dual<N>& a, const T& b)
{
  a = a + b;
  return a;
}
template <int N, typename T>
SB_HD dual<N>& operator-=(dual<N>& a, const T& b)
{
  a = a - b;
  return a;
}
template <int N, typename T>
SB_HD dual<N>& operator*=(dual<N>& a, const T& b)
{
  a = a * b;
  return a;
}
template <int N, typename T>
SB_HD dual<N>& operator/=(dual<N>& a, const T& b)
{
  a = a / b;
  return a;
}

// comparisons act on the value (reference: dual.hpp:170-215)
#define SB_DUAL_CMP(OP)                                    \
  template <int N>                                         \
  SB_HD bool operator OP(const dual<N>& a, const dual<N>& b) \
  {                                                        \
    return a.value OP b.value;                             \
  }                                                        \
  template <int N>                                         \
  SB_HD bool operator OP(const dual<N>& a, double b)       \
  {                                                        \
    return a.value OP b;                                   \
  }                                                        \
  template <int N>                                         \
  SB_HD bool operator OP(double a, const dual<N>& b)       \
  {                                                        \
    return a OP b.value;                                   \
  }
SB_DUAL_CMP(<)
SB_DUAL_CMP(>)
SB_DUAL_CMP(<=)
SB_DUAL_CMP(>=)
SB_DUAL_CMP(==)
SB_DUAL_CMP(!=)
#undef SB_DUAL_CMP

/// apply f(value) with derivative df(value) to a dual
template <int N>
SB_HD dual<N> chain(const dual<N>& a, double f, double df)
{
  dual<N> r{};
  r.value = f;
  SB_UNROLL
  for (int i = 0; i < N; i++) r.gradient[i] = df * a.gradient[i];
  return r;
}

template <int N>
SB_HDF dual<N> sqrt(const dual<N>& a)
{
  const double s = ::sqrt(a.value);
  return chain(a, s, 0.5 / s);
}
template <int N>
SB_HDF dual<N> exp(const dual<N>& a)
{
  const double e = ::exp(a.value);
  return chain(a, e, e);
}
template <int N>
SB_HDF dual<N> log(const dual<N>& a)
{
  return chain(a, ::log(a.value), 1.0 / a.value);
}
template <int N>
SB_HDF dual<N> log1p(const dual<N>& a)
{
  return chain(a, ::log1p(a.value), 1.0 / (1.0 + a.value));
}
template <int N>
SB_HDF dual<N> sin(const dual<N>& a)
{
  return chain(a, ::sin(a.value), ::cos(a.value));
}
template <int N>
SB_HDF dual<N> cos(const dual<N>& a)
{
  return chain(a, ::cos(a.value), -::sin(a.value));
}
template <int N>
SB_HDF dual<N> tan(const dual<N>& a)
{
  const double t = ::tan(a.value);
  return chain(a, t, 1.0 + t * t);
}
template <int N>
SB_HDF dual<N> atan(const dual<N>& a)
{
  return chain(a, ::atan(a.value), 1.0 / (1.0 + a.value * a.value));
}
template <int N>
SB_HDF dual<N> acos(const dual<N>& a)
{
  return chain(a, ::acos(a.value), -1.0 / ::sqrt(1.0 - a.value * a.value));
}
template <int N>
SB_HDF dual<N> asin(const dual<N>& a)
{
  return chain(a, ::asin(a.value), 1.0 / ::sqrt(1.0 - a.value * a.value));
}
template <int N>
SB_HDF dual<N> tanh(const dual<N>& a)
{
  const double t = ::tanh(a.value);
  return chain(a, t, 1.0 - t * t);
}
template <int N>
SB_HDF dual<N> abs(const dual<N>& a)
{
  return (a.value < 0) ? -a : a;
}
template <int N>
SB_HDF dual<N> pow(const dual<N>& a, double b)
{
  const double v = ::pow(a.value, b);
  return chain(a, v, b * ::pow(a.value, b - 1.0));
}
template <int N>
SB_HDF dual<N> pow(double a, const dual<N>& b)
{
  const double v = ::pow(a, b.value);
  return chain(b, v, v * ::log(a));
}
template <int N>
SB_HDF dual<N> pow(const dual<N>& a, const dual<N>& b)
{
  const double v  = ::pow(a.value, b.value);
  const double da = b.value * ::pow(a.value, b.value - 1.0);
  const double db = v * ::log(a.value);
  dual<N>      r{};
  r.value = v;
  SB_UNROLL
  for (int i = 0; i < N; i++) r.gradient[i] = da * a.gradient[i] + db * b.gradient[i];
  return r;
}
template <int N>
SB_HD dual<N> max(const dual<N>& a, const dual<N>& b)
{
  return (a.value > b.value) ? a : b;
}
template <int N>
SB_HD dual<N> min(const dual<N>& a, const dual<N>& b)
{
  return (a.value < b.value) ? a : b;
}

// make the plain-double maths visible inside serac:: (the dual overloads above would otherwise hide the global ones);
// using-declarations name the same entities as ::f, so a caller's `using namespace serac` stays unambiguous
using ::acos;
using ::asin;
using ::atan;
using ::cos;
using ::exp;
using ::fabs;
using ::log;
using ::log1p;
using ::pow;
using ::sin;
using ::sqrt;
using ::tan;
using ::tanh;
using std::abs;

// ------------------------------------------------------------------------------------------------
// tensor<T, n...>: nested fixed-size arrays (aggregate, so brace-initialisation works like the
// reference's, tensor.hpp:29-118).  A(i,j,...) and A[i][j]... both index it.
// ------------------------------------------------------------------------------------------------
template <typename T, int... n>
struct tensor;

template <typename T, int m>
struct tensor<T, m> {
  using value_type                  = T;
  static constexpr int ndim         = 1;
  static constexpr int first_dim    = m;
  static constexpr int total        = m;
  T                    data[m];

  SB_HD T& operator[](int i) { return data[i]; }
  SB_HD const T& operator[](int i) const { return data[i]; }
  SB_HD T& operator()(int i) { return data[i]; }
  SB_HD const T& operator()(int i) const { return data[i]; }
};

template <typename T, int m, int... n>
struct tensor<T, m, n...> {
  using value_type                  = T;
  static constexpr int ndim         = 1 + int(sizeof...(n));
  static constexpr int first_dim    = m;
  static constexpr int total        = m * tensor<T, n...>::total;
  tensor<T, n...>      data[m];

  SB_HD tensor<T, n...>& operator[](int i) { return data[i]; }
  SB_HD const tensor<T, n...>& operator[](int i) const { return data[i]; }
  SB_HD tensor<T, n...>& operator()(int i) { return data[i]; }
  SB_HD const tensor<T, n...>& operator()(int i) const { return data[i]; }
  template <typename... I>
  SB_HD auto& operator()(int i, I... rest)
  {
    return data[i](rest...);
  }
  template <typename... I>
  SB_HD const auto& operator()(int i, I... rest) const
  {
    return data[i](rest...);
  }
};

template <typename T, int n1>
tensor(const T (&)[n1]) -> tensor<T, n1>;
template <typename T, int n1, int n2>
tensor(const T (&)[n1][n2]) -> tensor<T, n1, n2>;

template <typename T>
struct is_tensor : std::false_type {
};
template <typename T, int... n>
struct is_tensor<tensor<T, n...>> : std::true_type {
};

template <int n>
using vec = tensor<double, n>;
template <int m, int n>
using mat = tensor<double, m, n>;

/// number of scalar entries
SB_HD int size(double) { return 1; }
SB_HD int size(zero) { return 0; }
template <int N>
SB_HD int size(const dual<N>&)
{
  return 1;
}
template <typename T, int... n>
SB_HD int size(const tensor<T, n...>&)
{
  return (n * ...);
}
template <typename T, int m, int... n>
SB_HD int leading_dimension(const tensor<T, m, n...>&)
{
  return m;
}

// ---- element-wise machinery (recursive on the leading dimension) ----
namespace detail {

template <typename T>
struct scalar_of {
  using type = T;
};
template <typename T, int... n>
struct scalar_of<tensor<T, n...>> {
  using type = T;
};

template <typename NewT, typename Tens>
struct rebind;
template <typename NewT, typename T, int... n>
struct rebind<NewT, tensor<T, n...>> {
  using type = tensor<NewT, n...>;
};

template <typename F, typename S, typename T, int m>
SB_HD void zip(const F& f, tensor<decltype(F{}(S{}, T{})), m>& out, const tensor<S, m>& a, const tensor<T, m>& b)
{
  SB_UNROLL
  for (int i = 0; i < m; i++) out.data[i] = f(a.data[i], b.data[i]);
}
template <typename F, typename S, typename T, int m, int n0, int... n>
SB_HD void zip(const F& f, tensor<decltype(F{}(S{}, T{})), m, n0, n...>& out, const tensor<S, m, n0, n...>& a,
               const tensor<T, m, n0, n...>& b)
{
  SB_UNROLL
  for (int i = 0; i < m; i++) zip(f, out.data[i], a.data[i], b.data[i]);
}

template <typename F, typename S, int m>
SB_HD void map(const F& f, tensor<decltype(F{}(S{})), m>& out, const tensor<S, m>& a)
{
  SB_UNROLL
  for (int i = 0; i < m; i++) out.data[i] = f(a.data[i]);
}
template <typename F, typename S, int m, int n0, int... n>
SB_HD void map(const F& f, tensor<decltype(F{}(S{})), m, n0, n...>& out, const tensor<S, m, n0, n...>& a)
{
  SB_UNROLL
  for (int i = 0; i < m; i++) map(f, out.data[i], a.data[i]);
}

struct add_op {
  template <typename S, typename T>
  SB_HD auto operator()(const S& a, const T& b) const
  {
    return a + b;
  }
};
struct sub_op {
  template <typename S, typename T>
  SB_HD auto operator()(const S& a, const T& b) const
  {
    return a - b;
  }
};
struct neg_op {
  template <typename S>
  SB_HD auto operator()(const S& a) const
  {
    return -a;
  }
};
template <typename S>
struct scale_op {
  S s;
  template <typename T>
  SB_HD auto operator()(const T& a) const
  {
    return a * s;
  }
};
template <typename S>
struct lscale_op {
  S s;
  template <typename T>
  SB_HD auto operator()(const T& a) const
  {
    return s * a;
  }
};
template <typename S>
struct div_op {
  S s;
  template <typename T>
  SB_HD auto operator()(const T& a) const
  {
    return a / s;
  }
};
template <typename S>
struct rdiv_op {
  S s;
  template <typename T>
  SB_HD auto operator()(const T& a) const
  {
    return s / a;
  }
};
struct value_op {
  template <typename T>
  SB_HD double operator()(const T& a) const
  {
    return get_value(a);
  }
};
struct value_op_any {
  template <typename T>
  SB_HD auto operator()(const T& a) const
  {
    return get_value(a);
  }
};
}  // namespace detail

template <typename S, typename T, int... n>
SB_HD auto operator+(const tensor<S, n...>& a, const tensor<T, n...>& b)
{
  tensor<decltype(S{} + T{}), n...> out{};
  detail::zip(detail::add_op{}, out, a, b);
  return out;
}
template <typename S, typename T, int... n>
SB_HD auto operator-(const tensor<S, n...>& a, const tensor<T, n...>& b)
{
  tensor<decltype(S{} - T{}), n...> out{};
  detail::zip(detail::sub_op{}, out, a, b);
  return out;
}
template <typename S, int... n>
SB_HD auto operator-(const tensor<S, n...>& a)
{
  tensor<S, n...> out{};
  detail::map(detail::neg_op{}, out, a);
  return out;
}
template <typename S, typename T, int... n, typename = std::enable_if_t<is_scalar_v<S>>>
SB_HD auto operator*(const S& s, const tensor<T, n...>& a)
{
  tensor<decltype(S{} * T{}), n...> out{};
  detail::map(detail::lscale_op<S>{s}, out, a);
  return out;
}
template <typename S, typename T, int... n, typename = std::enable_if_t<is_scalar_v<S>>>
SB_HD auto operator*(const tensor<T, n...>& a, const S& s)
{
  tensor<decltype(T{} * S{}), n...> out{};
  detail::map(detail::scale_op<S>{s}, out, a);
  return out;
}
template <typename S, typename T, int... n, typename = std::enable_if_t<is_scalar_v<S>>>
SB_HD auto operator/(const tensor<T, n...>& a, const S& s)
{
  if constexpr (std::is_same_v<T, double> && std::is_arithmetic_v<S>) {
    // one FP64 division instead of one per component (each costs ~a dozen FP64 instructions on the device);
    // differs from the per-component quotient by at most 1 ulp
    return a * (1.0 / s);
  } else {
    tensor<decltype(T{} / S{}), n...> out{};
    detail::map(detail::div_op<S>{s}, out, a);
    return out;
  }
}
template <typename S, typename T, int... n, typename = std::enable_if_t<is_scalar_v<S>>>
SB_HD auto operator/(const S& s, const tensor<T, n...>& a)
{
  tensor<decltype(S{} / T{}), n...> out{};
  detail::map(detail::rdiv_op<S>{s}, out, a);
  return out;
}
template <typename S, typename T, int... n>
SB_HD auto& operator+=(tensor<S, n...>& a, const tensor<T, n...>& b)
{
  a = a + b;
  return a;
}
template <typename S, int... n>
SB_HD auto& operator+=(tensor<S, n...>& a, zero)
{
  return a;
}
template <typename S, typename T, int... n>
SB_HD auto& operator-=(tensor<S, n...>& a, const tensor<T, n...>& b)
{
  a = a - b;
  return a;
}
template <typename S, int... n>
SB_HD auto& operator-=(tensor<S, n...>& a, zero)
{
  return a;
}

/// strip derivative information (reference: tuple_tensor_dual_functions.hpp:283-303)
template <typename T, int... n>
SB_HD auto get_value(const tensor<T, n...>& a)
{
  tensor<double, n...> out{};
  detail::map(detail::value_op{}, out, a);
  return out;
}

// ---- construction helpers ----
template <int n1, typename F>
SB_HD auto make_tensor(F f)
{
  using T = decltype(f(0));
  tensor<T, n1> A{};
  SB_UNROLL
  for (int i = 0; i < n1; i++) A(i) = f(i);
  return A;
}
template <int n1, int n2, typename F>
SB_HD auto make_tensor(F f)
{
  using T = decltype(f(0, 0));
  tensor<T, n1, n2> A{};
  SB_UNROLL
  for (int i = 0; i < n1; i++) {
    SB_UNROLL
    for (int j = 0; j < n2; j++) A(i, j) = f(i, j);
  }
  return A;
}
template <int n1, int n2, int n3, typename F>
SB_HD auto make_tensor(F f)
{
  using T = decltype(f(0, 0, 0));
  tensor<T, n1, n2, n3> A{};
  SB_UNROLL
  for (int i = 0; i < n1; i++) {
    SB_UNROLL
    for (int j = 0; j < n2; j++) {
      SB_UNROLL
      for (int k = 0; k < n3; k++) A(i, j, k) = f(i, j, k);
    }
  }
  return A;
}
template <int n1, int n2, int n3, int n4, typename F>
SB_HD auto make_tensor(F f)
{
  using T = decltype(f(0, 0, 0, 0));
  tensor<T, n1, n2, n3, n4> A{};
  SB_UNROLL
  for (int i = 0; i < n1; i++) {
    SB_UNROLL
    for (int j = 0; j < n2; j++) {
      SB_UNROLL
      for (int k = 0; k < n3; k++) {
        SB_UNROLL
        for (int l = 0; l < n4; l++) A(i, j, k, l) = f(i, j, k, l);
      }
    }
  }
  return A;
}

/// dense identity (the reference returns an isotropic_tensor from Identity<dim>(), isotropic_tensor.hpp;
/// a dense matrix has the same arithmetic meaning in every expression the material models use)
template <int dim>
SB_HD tensor<double, dim, dim> DenseIdentity()
{
  tensor<double, dim, dim> I{};
  SB_UNROLL
  for (int i = 0; i < dim; i++) {
    SB_UNROLL
    for (int j = 0; j < dim; j++) I(i, j) = (i == j) ? 1.0 : 0.0;
  }
  return I;
}
template <int dim>
SB_HD tensor<double, dim, dim> Identity()
{
  return DenseIdentity<dim>();
}

// ---- contractions ----
/// dot: contract the last index of A with the first index of B (reference: tensor.hpp:708-900)
template <typename S, typename T, int m>
SB_HD auto dot(const tensor<S, m>& a, const tensor<T, m>& b)
{
  decltype(S{} * T{}) s{};
  SB_UNROLL
  for (int i = 0; i < m; i++) s = s + a[i] * b[i];
  return s;
}
template <typename S, typename T, int m, int p0, int... p>
SB_HD auto dot(const tensor<S, m>& a, const tensor<T, m, p0, p...>& B)
{
  tensor<decltype(S{} * T{}), p0, p...> out{};
  SB_UNROLL
  for (int i = 0; i < m; i++) out = out + a[i] * B[i];
  return out;
}
namespace detail {
template <int m, typename T>
struct prepend {
  using type = tensor<T, m>;
};
template <int m, typename T, int... n>
struct prepend<m, tensor<T, n...>> {
  using type = tensor<T, m, n...>;
};
}  // namespace detail
template <typename S, typename T, int m, int n0, int... n, int... p>
SB_HD auto dot(const tensor<S, m, n0, n...>& A, const tensor<T, p...>& B)
{
  using sub_t = decltype(dot(A[0], B));
  typename detail::prepend<m, sub_t>::type out{};
  SB_UNROLL
  for (int i = 0; i < m; i++) out[i] = dot(A[i], B);
  return out;
}

/// tensor * tensor is shorthand for dot(A, B) (reference: tensor.hpp:995-1002)
template <typename S, typename T, int... m, int... n>
SB_HD auto operator*(const tensor<S, m...>& A, const tensor<T, n...>& B)
{
  return dot(A, B);
}

/// vector . matrix . vector
template <typename S, typename T, typename U, int m, int n>
SB_HD auto dot(const tensor<S, m>& u, const tensor<T, m, n>& A, const tensor<U, n>& v)
{
  return dot(u, dot(A, v));
}

template <typename S, typename T, int m, int n>
SB_HD auto outer(const tensor<S, m>& a, const tensor<T, n>& b)
{
  tensor<decltype(S{} * T{}), m, n> out{};
  SB_UNROLL
  for (int i = 0; i < m; i++) {
    SB_UNROLL
    for (int j = 0; j < n; j++) out(i, j) = a[i] * b[j];
  }
  return out;
}
template <typename S, typename T, int m, typename = std::enable_if_t<is_scalar_v<S>>>
SB_HD auto outer(const S& a, const tensor<T, m>& b)
{
  return a * b;
}
template <typename S, typename T, int m, typename = std::enable_if_t<is_scalar_v<S>>>
SB_HD auto outer(const tensor<T, m>& b, const S& a)
{
  return b * a;
}

/// full contraction of two same-shaped tensors (reference: inner / double_dot, tensor.hpp:668-1000)
template <typename S, typename T, int m>
SB_HD auto inner(const tensor<S, m>& a, const tensor<T, m>& b)
{
  return dot(a, b);
}
template <typename S, typename T, int m, int n0, int... n>
SB_HD auto inner(const tensor<S, m, n0, n...>& A, const tensor<T, m, n0, n...>& B)
{
  decltype(S{} * T{}) s{};
  SB_UNROLL
  for (int i = 0; i < m; i++) s = s + inner(A[i], B[i]);
  return s;
}
SB_HD double inner(double a, double b) { return a * b; }

template <typename S, typename T, int m, int n>
SB_HD auto double_dot(const tensor<S, m, n>& A, const tensor<T, m, n>& B)
{
  return inner(A, B);
}
/// A_ijkl B_kl
template <typename S, typename T, int m, int n, int p, int q>
SB_HD auto double_dot(const tensor<S, m, n, p, q>& A, const tensor<T, p, q>& B)
{
  tensor<decltype(S{} * T{}), m, n> out{};
  SB_UNROLL
  for (int i = 0; i < m; i++) {
    SB_UNROLL
    for (int j = 0; j < n; j++) out(i, j) = inner(A(i, j), B);
  }
  return out;
}
/// A_ijk B_jk
template <typename S, typename T, int m, int n, int p>
SB_HD auto double_dot(const tensor<S, m, n, p>& A, const tensor<T, n, p>& B)
{
  tensor<decltype(S{} * T{}), m> out{};
  SB_UNROLL
  for (int i = 0; i < m; i++) out(i) = inner(A(i), B);
  return out;
}

template <typename T, int... n>
SB_HD auto squared_norm(const tensor<T, n...>& A)
{
  return inner(A, A);
}
SB_HD double squared_norm(double a) { return a * a; }
template <typename T, int... n>
SB_HDF auto norm(const tensor<T, n...>& A)
{
  return sqrt(squared_norm(A));
}
SB_HDF double norm(double a) { return fabs(a); }
template <typename T, int... n>
SB_HDF auto normalize(const tensor<T, n...>& A)
{
  return A / norm(A);
}

template <typename T, int m, int n>
SB_HD auto transpose(const tensor<T, m, n>& A)
{
  tensor<T, n, m> At{};
  SB_UNROLL
  for (int i = 0; i < n; i++) {
    SB_UNROLL
    for (int j = 0; j < m; j++) At(i, j) = A(j, i);
  }
  return At;
}
template <typename T, int n>
SB_HD auto tr(const tensor<T, n, n>& A)
{
  T t{};
  SB_UNROLL
  for (int i = 0; i < n; i++) t = t + A(i, i);
  return t;
}
template <typename T, int n>
SB_HD auto sym(const tensor<T, n, n>& A)
{
  tensor<T, n, n> S{};
  SB_UNROLL
  for (int i = 0; i < n; i++) {
    SB_UNROLL
    for (int j = 0; j < n; j++) S(i, j) = 0.5 * (A(i, j) + A(j, i));
  }
  return S;
}
template <typename T, int n>
SB_HD auto antisym(const tensor<T, n, n>& A)
{
  tensor<T, n, n> S{};
  SB_UNROLL
  for (int i = 0; i < n; i++) {
    SB_UNROLL
    for (int j = 0; j < n; j++) S(i, j) = 0.5 * (A(i, j) - A(j, i));
  }
  return S;
}
template <typename T, int n>
SB_HD auto dev(const tensor<T, n, n>& A)
{
  auto            t = tr(A) / double(n);
  tensor<T, n, n> D = A;
  SB_UNROLL
  for (int i = 0; i < n; i++) D(i, i) = D(i, i) - t;
  return D;
}
template <typename T, int n>
SB_HD tensor<T, n, n> diag(const tensor<T, n>& d)
{
  tensor<T, n, n> D{};
  SB_UNROLL
  for (int i = 0; i < n; i++) D(i, i) = d[i];
  return D;
}
template <typename T, int n>
SB_HD tensor<T, n> diag(const tensor<T, n, n>& D)
{
  tensor<T, n> d{};
  SB_UNROLL
  for (int i = 0; i < n; i++) d[i] = D(i, i);
  return d;
}
template <typename T, int n>
SB_HD tensor<T, n, n> diagonal_matrix(const tensor<T, n, n>& A)
{
  return diag(diag(A));
}

template <typename T>
SB_HD auto det(const tensor<T, 2, 2>& A)
{
  return A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0);
}
template <typename T>
SB_HD auto det(const tensor<T, 3, 3>& A)
{
  return A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) - A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0)) +
         A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
}
/// det(A + I) - 1 without cancellation for small A (reference: tensor.hpp:1266-1296): tr A + I2(A) + det A
template <typename T>
SB_HD auto detApIm1(const tensor<T, 2, 2>& A)
{
  return A(0, 0) + A(1, 1) + (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0));
}
template <typename T>
SB_HD auto detApIm1(const tensor<T, 3, 3>& A)
{
  auto m01 = A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0);
  auto m02 = A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0);
  auto m12 = A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1);
  return (A(0, 0) + A(1, 1) + A(2, 2)) + (m01 + m02 + m12) + det(A);
}

template <typename T>
SB_HD auto inv(const tensor<T, 2, 2>& A)
{
  auto            id = 1.0 / det(A);
  tensor<T, 2, 2> B{};
  B(0, 0) = A(1, 1) * id;
  B(0, 1) = -A(0, 1) * id;
  B(1, 0) = -A(1, 0) * id;
  B(1, 1) = A(0, 0) * id;
  return B;
}
template <typename T>
SB_HD auto inv(const tensor<T, 3, 3>& A)
{
  tensor<T, 3, 3> C{};  // cofactors (transposed)
  C(0, 0) = A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1);
  C(0, 1) = A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2);
  C(0, 2) = A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1);
  C(1, 0) = A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2);
  C(1, 1) = A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0);
  C(1, 2) = A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2);
  C(2, 0) = A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0);
  C(2, 1) = A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1);
  C(2, 2) = A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0);
  auto id = 1.0 / (A(0, 0) * C(0, 0) + A(0, 1) * C(1, 0) + A(0, 2) * C(2, 0));
  return C * id;
}

template <typename S, typename T>
SB_HD auto cross(const tensor<S, 3>& a, const tensor<T, 3>& b)
{
  tensor<decltype(S{} * T{}), 3> c{};
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
  return c;
}
/// normal of the surface spanned by the columns of a 3x2 jacobian (reference: tensor.hpp:906-928)
template <typename T>
SB_HD auto cross(const tensor<T, 3, 2>& A)
{
  tensor<T, 3> a{{A(0, 0), A(1, 0), A(2, 0)}};
  tensor<T, 3> b{{A(0, 1), A(1, 1), A(2, 1)}};
  return cross(a, b);
}
template <typename T>
SB_HD auto cross(const tensor<T, 2, 1>& A)
{
  tensor<T, 2> c{};
  c[0] = A(1, 0);
  c[1] = -A(0, 0);
  return c;
}

// ------------------------------------------------------------------------------------------------
// tuple (reference: tuple.hpp).  get<i>, structured bindings, element-wise arithmetic.
// ------------------------------------------------------------------------------------------------
namespace detail {
template <int I, typename T>
struct leaf {
  T value;
};
template <typename Seq, typename... T>
struct tuple_storage;
template <int... I, typename... T>
struct tuple_storage<std::integer_sequence<int, I...>, T...> : leaf<I, T>... {
  SB_HD tuple_storage() : leaf<I, T>{}... {}
  SB_HD tuple_storage(const T&... v) : leaf<I, T>{v}... {}
};
template <int I, typename... T>
struct type_at;
template <typename T0, typename... T>
struct type_at<0, T0, T...> {
  using type = T0;
};
template <int I, typename T0, typename... T>
struct type_at<I, T0, T...> {
  using type = typename type_at<I - 1, T...>::type;
};
}  // namespace detail

template <typename... T>
struct tuple : detail::tuple_storage<std::make_integer_sequence<int, int(sizeof...(T))>, T...> {
  using base = detail::tuple_storage<std::make_integer_sequence<int, int(sizeof...(T))>, T...>;
  SB_HD tuple() : base() {}
  SB_HD tuple(const T&... v) : base(v...) {}
  /// member get so that structured bindings work on device without ADL surprises
  template <std::size_t I>
  SB_HD auto& get()
  {
    using E = typename detail::type_at<int(I), T...>::type;
    return static_cast<detail::leaf<int(I), E>&>(*this).value;
  }
  template <std::size_t I>
  SB_HD const auto& get() const
  {
    using E = typename detail::type_at<int(I), T...>::type;
    return static_cast<const detail::leaf<int(I), E>&>(*this).value;
  }
};
template <>
struct tuple<> {
};
template <typename... T>
tuple(T...) -> tuple<T...>;

template <typename... T>
SB_HD tuple<T...> make_tuple(const T&... v)
{
  return tuple<T...>(v...);
}

template <int I, typename... T>
SB_HD auto& get(tuple<T...>& t)
{
  return t.template get<std::size_t(I)>();
}
template <int I, typename... T>
SB_HD const auto& get(const tuple<T...>& t)
{
  return t.template get<std::size_t(I)>();
}
/// get<i>(zero) is zero (reference: tensor.hpp:239-250)
template <int I>
SB_HD zero get(zero)
{
  return {};
}

template <typename T>
struct is_tuple : std::false_type {
};
template <typename... T>
struct is_tuple<tuple<T...>> : std::true_type {
};

namespace detail {
template <typename F, typename... S, typename... T, int... I>
SB_HD auto tuple_zip(const F& f, const tuple<S...>& a, const tuple<T...>& b, std::integer_sequence<int, I...>)
{
  return serac::make_tuple(f(get<I>(a), get<I>(b))...);
}
template <typename F, typename... S, int... I>
SB_HD auto tuple_map(const F& f, const tuple<S...>& a, std::integer_sequence<int, I...>)
{
  return serac::make_tuple(f(get<I>(a))...);
}
}  // namespace detail

template <typename... S, typename... T>
SB_HD auto operator+(const tuple<S...>& a, const tuple<T...>& b)
{
  return detail::tuple_zip(detail::add_op{}, a, b, std::make_integer_sequence<int, int(sizeof...(S))>{});
}
template <typename... S, typename... T>
SB_HD auto operator-(const tuple<S...>& a, const tuple<T...>& b)
{
  return detail::tuple_zip(detail::sub_op{}, a, b, std::make_integer_sequence<int, int(sizeof...(S))>{});
}
template <typename... S>
SB_HD auto operator-(const tuple<S...>& a)
{
  return detail::tuple_map(detail::neg_op{}, a, std::make_integer_sequence<int, int(sizeof...(S))>{});
}
template <typename... S>
SB_HD auto operator*(double s, const tuple<S...>& a)
{
  return detail::tuple_map(detail::lscale_op<double>{s}, a, std::make_integer_sequence<int, int(sizeof...(S))>{});
}
template <typename... S>
SB_HD auto operator*(const tuple<S...>& a, double s)
{
  return detail::tuple_map(detail::scale_op<double>{s}, a, std::make_integer_sequence<int, int(sizeof...(S))>{});
}
template <typename... S>
SB_HD auto operator/(const tuple<S...>& a, double s)
{
  return detail::tuple_map(detail::div_op<double>{s}, a, std::make_integer_sequence<int, int(sizeof...(S))>{});
}

template <typename... T>
SB_HD auto get_value(const tuple<T...>& t)
{
  return detail::tuple_map(detail::value_op_any{}, t, std::make_integer_sequence<int, int(sizeof...(T))>{});
}

// ------------------------------------------------------------------------------------------------
// isotropic tensor functions of a symmetric 3x3 tensor (tuple_tensor_dual_functions.hpp:791-1045)
// ------------------------------------------------------------------------------------------------
/// eigenvalues (ascending) and eigenvectors (columns of Q) of a symmetric 3x3 matrix; closed form after Scherzinger &
/// Dohrmann, "A robust algorithm for finding the eigenvalues and eigenvectors of 3x3 symmetric matrices" -- the
/// algorithm the reference uses (tuple_tensor_dual_functions.hpp:802-912): most distinct eigenvalue of the deviator
/// from the trigonometric solution, its eigenvector from the best-conditioned pair of rows, the other two from the
/// 2x2 problem in the orthogonal complement.
SB_HDF void eig_symm(const tensor<double, 3, 3>& A, tensor<double, 3>& eigvals, tensor<double, 3, 3>& eigvecs)
{
  const double          pi = 3.14159265358979323846;
  tensor<double, 3>     eta{};
  tensor<double, 3, 3>  Q = DenseIdentity<3>();
  const auto            Ad = dev(A);
  const double          J2 = 0.5 * inner(Ad, Ad);
  const double          J3 = det(Ad);
  if (J2 > 0.0) {
    double       c     = (0.5 * J3) * pow(3.0 / J2, 1.5);
    c                  = fmin(fmax(c, -1.0), 1.0);
    const double alpha = acos(c) / 3.0;
    const double amp   = 2.0 * sqrt(J2 / 3.0);
    eta[0]             = (6.0 * alpha < pi) ? amp * cos(alpha) : amp * cos(alpha + 2.0 * pi / 3.0);

    // rows of (A_dev - eta0 I)^T; the two that span its range best give the eigenvector of eta0 as their normal
    tensor<double, 3, 3> r{};
    int                  imax = 0;
    double               nmax = -1.0;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) r[i][j] = Ad[j][i] - (i == j ? eta[0] : 0.0);
      const double nr = norm(r[i]);
      if (nmax < nr) {
        imax = i;
        nmax = nr;
      }
    }
    const tensor<double, 3> s0 = normalize(r[imax]);
    const tensor<double, 3> ra = r[(imax + 1) % 3], rb = r[(imax + 2) % 3];
    const tensor<double, 3> t1 = ra - dot(ra, s0) * s0;
    const tensor<double, 3> t2 = rb - dot(rb, s0) * s0;
    const tensor<double, 3> s1 = normalize((norm(t1) > norm(t2)) ? t1 : t2);
    const tensor<double, 3> v0 = cross(s0, s1);
    for (int i = 0; i < 3; i++) Q[i][0] = v0[i];

    // the remaining two eigenvalues: the 2x2 problem in span{s0, s1}
    const tensor<double, 3> As0 = dot(Ad, s0), As1 = dot(Ad, s1);
    const double            A11 = dot(s0, As0), A12 = dot(s0, As1), A21 = dot(s1, As0), A22 = dot(s1, As1);
    const double            d   = A11 - A22;
    const double            sg  = double((0.0 < d) - (d < 0.0));
    const double            delta = 0.5 * sg * sqrt(d * d + 4.0 * A12 * A21);
    eta[1]                        = 0.5 * (A11 + A22) - delta;
    eta[2]                        = 0.5 * (A11 + A22) + delta;
    if (fabs(delta) <= 1.0e-15) {  // repeated pair: any basis of the complement
      for (int i = 0; i < 3; i++) {
        Q[i][1] = s0[i];
        Q[i][2] = s1[i];
      }
    } else {
      const tensor<double, 3> u1 = As0 - eta[1] * s0;
      const tensor<double, 3> u2 = As1 - eta[1] * s1;
      const tensor<double, 3> w  = normalize((norm(u1) > norm(u2)) ? u1 : u2);
      const tensor<double, 3> v1 = normalize(cross(w, v0));
      const tensor<double, 3> v2 = normalize(cross(v0, v1));
      for (int i = 0; i < 3; i++) {
        Q[i][1] = v1[i];
        Q[i][2] = v2[i];
      }
    }
  }
  const double shift = tr(A) / 3.0;
  for (int i = 0; i < 3; i++) eta[i] += shift;
  // ascending order
  int o[3] = {0, 1, 2};
  auto swp = [](int& a, int& b) {
    const int t = a;
    a           = b;
    b           = t;
  };
  if (eta[o[0]] > eta[o[1]]) swp(o[0], o[1]);
  if (eta[o[1]] > eta[o[2]]) swp(o[1], o[2]);
  if (eta[o[0]] > eta[o[1]]) swp(o[0], o[1]);
  for (int a = 0; a < 3; a++) {
    eigvals[a] = eta[o[a]];
    for (int i = 0; i < 3; i++) eigvecs[i][a] = Q[i][o[a]];
  }
}

/// f(A) = Q diag(f(lambda)) Q^T; with a dual input the derivative uses the eigenvalue secant function
/// g(l1, l2) = (f(l1) - f(l2)) / (l1 - l2) (-> f'(l1) when equal), written stably by the caller
/// (tuple_tensor_dual_functions.hpp:963-1000)
template <typename T, typename Function, typename EigvalSecantFunction>
SB_HDF auto symmetric_mat3_function(const tensor<T, 3, 3>& A, const Function& f, const EigvalSecantFunction& g)
{
  tensor<double, 3>    lambda{};
  tensor<double, 3, 3> Q{};
  eig_symm(get_value(A), lambda, Q);
  tensor<double, 3, 3> fA{};
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) {
      double v = 0.0;
      for (int a = 0; a < 3; a++) v += Q[i][a] * f(lambda[a]) * Q[j][a];
      fA[i][j] = v;
    }
  }
  if constexpr (!is_dual_number<T>::value) {
    return fA;
  } else {
    constexpr int N = num_partials<T>::value;
    // rotate the input's partials into the eigenbasis, scale by the secant table, rotate back
    double G[3][3];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) G[a][b] = g(lambda[a], lambda[b]);
    tensor<T, 3, 3> out{};
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        out[i][j].value = fA[i][j];
        for (int n = 0; n < N; n++) out[i][j].gradient[n] = 0.0;
      }
    for (int n = 0; n < N; n++) {
      double B[3][3];  // Q^T dA Q
      for (int a = 0; a < 3; a++) {
        for (int b = 0; b < 3; b++) {
          double v = 0.0;
          for (int k = 0; k < 3; k++)
            for (int l = 0; l < 3; l++) v += Q[k][a] * A[k][l].gradient[n] * Q[l][b];
          B[a][b] = G[a][b] * v;
        }
      }
      for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) {
          double v = 0.0;
          for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) v += Q[i][a] * B[a][b] * Q[j][b];
          out[i][j].gradient[n] = v;
        }
      }
    }
    return out;
  }
}

/// logarithm of a symmetric positive definite matrix (tuple_tensor_dual_functions.hpp:1007-1020)
template <typename T>
SB_HDF auto log_symm(const tensor<T, 3, 3>& A)
{
  auto g = [](double l1, double l2) {
    if (l1 == l2) return 1.0 / l1;
    const double y = l1 / l2;
    return (log(y) / (y - 1.0)) / l2;
  };
  return symmetric_mat3_function(A, [](double x) { return log(x); }, g);
}

/// exponential of a symmetric matrix (tuple_tensor_dual_functions.hpp:1028-1041)
template <typename T>
SB_HDF auto exp_symm(const tensor<T, 3, 3>& A)
{
  auto g = [](double l1, double l2) {
    if (l1 == l2) return exp(l1);
    const double arg = l1 - l2;
    return exp(l2) * expm1(arg) / arg;
  };
  return symmetric_mat3_function(A, [](double x) { return exp(x); }, g);
}

/// compile-time spatial dimension tag and DependsOn<> (reference: functional.hpp / geometry.hpp)
template <int d>
struct Dimension {
  static constexpr int value = d;
  SB_HD operator int() const { return d; }
};
template <int... i>
struct DependsOn {
};

/// sentinel quadrature-data types (reference: quadrature_data.hpp:20-37)
struct Nothing {
};
struct Empty {
};

}  // namespace serac

namespace std {
template <typename... T>
struct tuple_size<serac::tuple<T...>> : std::integral_constant<std::size_t, sizeof...(T)> {
};
template <std::size_t I, typename... T>
struct tuple_element<I, serac::tuple<T...>> {
  using type = typename serac::detail::type_at<int(I), T...>::type;
};
}  // namespace std
