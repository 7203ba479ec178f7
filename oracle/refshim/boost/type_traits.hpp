// Minimal stand-in for boost/type_traits.hpp (oracle/_ref only).
#pragma once
#include <type_traits>
namespace boost {
template <class T> struct is_integral : std::is_integral<T> {};
}
