// gemm_common.cuh — tile constants and PTX helpers shared by the DMMA contraction kernels (k_density.cu, k_gemm.cu)
#pragma once
#include "xc_internal.h"
#include <cuda_runtime.h>

namespace xcb {


#ifndef XC_GK
#define XC_GK 16
#endif
#ifndef XC_STAGES
#define XC_STAGES 3
#endif
constexpr int GM = 128, GN = 64, GK = XC_GK, GSTAGES = XC_STAGES;
constexpr int NTHREADS = 256;
// density: A stage is [k][m] (m contiguous), ld 132; B stage is [n][k] (k contiguous), ld 20
constexpr int D_LDA = GM + 4, D_LDB = GK + 4;
constexpr int D_ASZ = GK * D_LDA, D_BSZ = GN * D_LDB, D_STAGE = D_ASZ + D_BSZ;

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool pred) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, bool pred) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  const int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void red_add_f64(double *addr, double v) {
  asm volatile("red.global.add.f64 [%0], %1;\n" ::"l"(addr), "d"(v) : "memory");
}


} // namespace xcb
