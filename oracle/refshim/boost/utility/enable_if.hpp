// Minimal stand-in for boost/utility/enable_if.hpp (oracle/_ref only).
#pragma once
#include <type_traits>
namespace boost {
template <bool B, class T = void> struct enable_if_c : std::enable_if<B, T> {};
}
