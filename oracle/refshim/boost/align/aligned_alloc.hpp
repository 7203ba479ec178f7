// Minimal stand-in for boost/align/aligned_alloc.hpp (boost is not installed in
// this image). Test infrastructure only: lets the reference's stage headers
// compile unmodified for oracle/_ref. Not part of the shipped product.
#pragma once
#include <cstdlib>
namespace boost { namespace alignment {
inline void* aligned_alloc(std::size_t alignment, std::size_t size) {
  void* p = nullptr;
  if (alignment < sizeof(void*)) alignment = sizeof(void*);
  if (posix_memalign(&p, alignment, size ? size : alignment) != 0) return nullptr;
  return p;
}
inline void aligned_free(void* p) { std::free(p); }
}}
