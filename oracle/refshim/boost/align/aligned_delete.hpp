// Minimal stand-in for boost/align/aligned_delete.hpp (oracle/_ref only).
#pragma once
#include "aligned_alloc.hpp"
namespace boost { namespace alignment {
struct aligned_delete {
  template <class T> void operator()(T* p) const { if (p) { p->~T(); aligned_free(p); } }
};
}}
