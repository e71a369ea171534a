// Shared device/host helpers for the diaglib_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define DLB_CUDA_CHECK(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      std::fprintf(stderr, "diaglib_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e), \
                   __FILE__, __LINE__, cudaGetErrorString(_e));                                \
      std::abort();                                                                            \
    }                                                                                          \
  } while (0)

namespace dlb {

// ---- FP64 tensor-core MMA (SASS: DMMA).  m8n8k4, A row-major 8x4, B col-major 4x8.
// lane l holds A[l>>2][l&3], B[l&3][l>>2], C[l>>2][2*(l&3) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- cp.async (SASS: LDGSTS) with zero fill: copies src_bytes (<= N) and zero-fills the rest.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace dlb
