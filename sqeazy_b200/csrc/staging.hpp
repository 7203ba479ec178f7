// Host <-> device transfers behind the SQY_* entry points (caller-owned host buffers of any kind).
//
// The reference works in the caller's memory with `nthreads` OpenMP threads (sqeazy_algorithms.hpp:14-22,
// src/sqeazy.cpp:108-142). Here the same knob sizes the host side of the PCIe hop: page-locked caller buffers go to
// the device with one DMA; pageable ones (a plain malloc / std::vector / Java heap buffer: what the sqy CLI and the
// BridJ bindings pass) are moved through a ring of page-locked chunks, `nthreads` host threads filling chunk c+1
// while the DMA engine moves chunk c. With nthreads == 1 the driver's own pageable path is used (it is the same
// single-threaded copy). There is one ring per device; the caller holds that device's lock (api.cu). A slot remembers
// across calls and streams whether a DMA on it is still outstanding.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace sqyb {

// both return a cudaError_t-compatible int (0 = success); work is ordered on `st`.
// h2d: on return the host buffer has been consumed (it may be reused), the device copy may still be in flight on `st`.
// d2h: on return the host buffer is complete when the path was staged; otherwise it is complete after `st` is synchronised.
int staged_h2d(void* d_dst, const void* h_src, size_t bytes, int nthreads, cudaStream_t st);
int staged_d2h(void* h_dst, const void* d_src, size_t bytes, int nthreads, cudaStream_t st);
void staging_release();          // frees the page-locked ring of the current device (sqyx_release_scratch)
bool host_buffer_is_pageable(const void* p);   // not page-locked / registered with CUDA
int staging_threads(int nthreads);   // the reference's rule: <= 0 or more than the machine has => all cores (capped at 16 here)

}  // namespace sqyb
