// serac_b200: device-ready scalar root finder used by the J2 return map.
//
// Same contract as the reference's solve_scalar_equation
// (src/serac/numerics/functional/tuple_tensor_dual_functions.hpp:519-677): Newton's method safeguarded by bisection on
// a bracket [lower_bound, upper_bound]; when any parameter carries derivatives the returned root carries
// dx/dp = -(df/dp) / (df/dx) (implicit function theorem).  The iteration visits the same iterates as the reference so
// that converged roots agree to rounding, which the <=1e-12 parity budget on J2 residuals needs.  Differences: no slic
// calls (a failed bracket or an exhausted iteration budget is reported through SolverStatus only), usable in device code.
#pragma once

#include "serac_b200/tensor.hpp"

namespace serac {

struct ScalarSolverOptions {
  double       xtol;      ///< absolute tolerance on the Newton correction
  double       rtol;      ///< absolute tolerance on |f(x)|
  unsigned int max_iter;  ///< iteration budget
};

struct SolverStatus {
  bool         converged;
  unsigned int iterations;
  double       residual;
};

namespace detail {
template <typename... T>
struct max_partials;
template <>
struct max_partials<> {
  static constexpr int value = 0;
};
template <typename T0, typename... T>
struct max_partials<T0, T...> {
  static constexpr int here  = num_partials<typename scalar_of<T0>::type>::value;
  static constexpr int rest  = max_partials<T...>::value;
  static constexpr int value = here > rest ? here : rest;
};
}  // namespace detail

template <typename function, typename... ParamTypes>
SERAC_HOST_DEVICE inline auto solve_scalar_equation(const function& f, double x0, double lower_bound,
                                                    double upper_bound, ScalarSolverOptions options,
                                                    ParamTypes... params)
{
  // value and d/dx of f at x, parameters frozen at their values
  auto eval = [&](double xx, double& fv, double& dfv) {
    dual<1> xd{xx, {1.0}};
    auto    R = f(xd, get_value(params)...);
    fv        = R.value;
    dfv       = R.gradient[0];
  };

  double x = x0, df_dx = 0.0, fval = 0.0;
  double fl = f(lower_bound, get_value(params)...);
  double fh = f(upper_bound, get_value(params)...);

  unsigned int iterations = 0;
  bool         converged  = false;
  const bool   bracketed  = !(fl * fh > 0);

  if (fl == 0) {
    x         = lower_bound;
    converged = true;
  } else if (fh == 0) {
    x         = upper_bound;
    converged = true;
  }

  if (converged) {
    eval(x, fval, df_dx);
  } else {
    // orient the search so that f(xl) < 0
    double xl = lower_bound, xh = upper_bound;
    if (fl > 0) {
      xl = upper_bound;
      xh = lower_bound;
    }
    if (x0 < lower_bound || x0 > upper_bound) x0 = 0.5 * (lower_bound + upper_bound);
    x                  = x0;
    double delta_x_old = ::fabs(upper_bound - lower_bound);
    double delta_x     = delta_x_old;
    eval(x, fval, df_dx);

    while (!converged) {
      if (iterations == options.max_iter) break;

      const bool out_of_bracket = ((x - xh) * df_dx - fval > 0) || ((x - xl) * df_dx - fval < 0);
      const bool too_slow       = ::fabs(2. * fval) > ::fabs(delta_x_old * df_dx);
      if (out_of_bracket || too_slow) {  // bisect
        delta_x_old = delta_x;
        delta_x     = 0.5 * (xh - xl);
        x           = xl + delta_x;
        converged   = (x == xl);
      } else {  // Newton
        delta_x_old       = delta_x;
        delta_x           = fval / df_dx;
        const double prev = x;
        x -= delta_x;
        converged = (x == prev);
      }

      eval(x, fval, df_dx);
      converged = converged || (::fabs(delta_x) < options.xtol) || (::fabs(fval) < options.rtol);

      if (fval < 0) {
        xl = x;
      } else {
        xh = x;
      }
      ++iterations;
    }
  }

  constexpr int N = detail::max_partials<ParamTypes...>::value;
  if constexpr (N > 0) {
    auto         R = f(x, params...);  // dual<N>: value and df/dp . dp
    SolverStatus status{converged && bracketed, iterations, R.value};
    dual<N>      xd{};
    xd.value = x;
    SB_UNROLL
    for (int k = 0; k < N; k++) xd.gradient[k] = -R.gradient[k] / df_dx;
    return tuple{xd, status};
  } else {
    double       R = f(x, params...);
    SolverStatus status{converged && bracketed, iterations, R};
    return tuple{x, status};
  }
}

}  // namespace serac
