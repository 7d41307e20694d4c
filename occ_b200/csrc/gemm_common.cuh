// gemm_common.cuh — tile constants and PTX helpers shared by the DMMA contraction kernels (k_density.cu, k_gemm.cu)
#pragma once
#include "xc_internal.h"
#include <atomic>
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

// Opt-in to > 48 KiB of dynamic shared memory.  The attribute belongs to the (function, device) pair, and one process
// may drive handles on several devices from several threads (occ_xc_multi_create), so the
// "already done" state is one bit per device ordinal; cudaFuncSetAttribute is idempotent, a race only repeats it.
struct SmemOptIn {
  std::atomic<unsigned long long> done[4]; // 256 device ordinals
  SmemOptIn() {
    for (auto &d : done) d.store(0ull);
  }
  template <class Kernel> void ensure(Kernel kernel, int bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::atomic<unsigned long long> &word = done[(dev >> 6) & 3];
    const unsigned long long bit = 1ull << (dev & 63);
    if (word.load(std::memory_order_acquire) & bit) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    word.fetch_or(bit, std::memory_order_release);
  }
};

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

// ---- deterministic (bit-repeatable) accumulation: options.deterministic = 1 -------------------------------------
// A cross-CTA sum of doubles depends on the arrival order of the atomics.  Here every addend x is split into
//   hi = fl(fl(x + C) - C),  C = 1.5 * 2^E   -> a multiple of 2^(E-52), |x - hi| <= 2^(E-53)
//   lo = fl(fl((x - hi) + c) - c), c = 1.5 * 2^(E-36) -> a multiple of 2^(E-88)
// and hi / lo are accumulated in two separate arrays.  All partial sums of multiples of one power of two are exactly
// representable while they stay below 2^53 quanta (|sum hi| < 2^(E+1), up to 2^18 addends for lo), so the two sums do
// not depend on the order; what is dropped is below 2^(E-89) per addend.  The kernels that use it are compiled in
// separate translation units (k_*_det.cu include the same source with XC_DETERMINISTIC), the default build is untouched.
__device__ __forceinline__ void red_add_binned(double *addr, long long lo_off, double x, double C, double c) {
  const double hi = __dsub_rn(__dadd_rn(x, C), C);
  const double r = __dsub_rn(x, hi);
  const double lo = __dsub_rn(__dadd_rn(r, c), c);
  red_add_f64(addr, hi);
  red_add_f64(addr + lo_off, lo);
}
#ifdef XC_DETERMINISTIC
#define XC_SYM(name) name##_det
#define XC_ACC_ADD(P, addr, v) red_add_binned(addr, (P).det_acc_off, v, 25165824.0 /* 1.5 * 2^24 */, 3.662109375e-4 /* 1.5 * 2^-12 */)
#define XC_K_ADD(P, addr, v) red_add_binned(addr, (P).det_K_off, v, 1572864.0 /* 1.5 * 2^20 */, 2.288818359375e-5 /* 1.5 * 2^-16 */)
#else
#define XC_SYM(name) name
#define XC_ACC_ADD(P, addr, v) red_add_f64(addr, v)
#define XC_K_ADD(P, addr, v) red_add_f64(addr, v)
#endif

} // namespace xcb
