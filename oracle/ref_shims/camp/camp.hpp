// TEST INFRASTRUCTURE (oracle/): the two camp aliases the reference's kernels use (function_signature.hpp:22).
#pragma once
#include <utility>
namespace camp {
template <class T, T... i>
using int_seq = std::integer_sequence<T, i...>;
template <class T, T n>
using make_int_seq = std::make_integer_sequence<T, n>;
}  // namespace camp
