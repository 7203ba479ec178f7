// Minimal stand-in for boost/align/aligned_allocator.hpp (oracle/_ref only).
#pragma once
#include <cstddef>
#include <new>
#include "aligned_alloc.hpp"
namespace boost { namespace alignment {
template <class T, std::size_t Alignment = alignof(T)>
struct aligned_allocator {
  typedef T value_type;
  typedef T* pointer;
  typedef const T* const_pointer;
  typedef T& reference;
  typedef const T& const_reference;
  typedef std::size_t size_type;
  typedef std::ptrdiff_t difference_type;
  template <class U> struct rebind { typedef aligned_allocator<U, Alignment> other; };
  aligned_allocator() noexcept {}
  template <class U> aligned_allocator(const aligned_allocator<U, Alignment>&) noexcept {}
  T* allocate(std::size_t n) {
    void* p = boost::alignment::aligned_alloc(Alignment, n * sizeof(T));
    if (!p) throw std::bad_alloc();
    return static_cast<T*>(p);
  }
  void deallocate(T* p, std::size_t) noexcept { boost::alignment::aligned_free(p); }
};
template <class T, class U, std::size_t A>
bool operator==(const aligned_allocator<T, A>&, const aligned_allocator<U, A>&) noexcept { return true; }
template <class T, class U, std::size_t A>
bool operator!=(const aligned_allocator<T, A>&, const aligned_allocator<U, A>&) noexcept { return false; }
}}
