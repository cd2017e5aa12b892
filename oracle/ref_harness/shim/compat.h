// Force-included (-include) when compiling the reference's translation units with
// nvcc 12.9 / g++ 13 on Linux.  The reference was written against MSVC + CUDA 11.1
// (CMakeLists.txt:58-97) and uses three things that toolchain accepted:
//   * atomicAdd(size_t*, int)            src/dcgrid/dcgrid_utils.cuh:118
//   * atomicCAS(size_t*, size_t, size_t) src/utils/cudamath.cuh:306,316
//   * the integer literal suffix Ui32    src/dcgrid/dcgrid_utils.cuh:86
// On LP64 size_t is `unsigned long`, for which CUDA ships no atomic overloads, so we
// forward to the `unsigned long long` ones (same width).  No reference source is
// copied or modified.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#ifdef __CUDACC__
static __device__ __forceinline__ size_t atomicAdd(size_t *addr, int v) {
  return (size_t)atomicAdd(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)(long long)v);
}
static __device__ __forceinline__ size_t atomicCAS(size_t *addr, size_t cmp, size_t val) {
  return (size_t)atomicCAS(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)cmp,
                           (unsigned long long)val);
}
#endif

#ifdef __CUDACC__
__host__ __device__
#endif
constexpr uint32_t operator""Ui32(unsigned long long v) { return (uint32_t)v; }
