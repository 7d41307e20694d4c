// k_dgemm.cu — plain (optionally strided-batched) FP64 GEMM on the DMMA pipe, for the dense products around the hot
// kernels that would otherwise go to cuBLAS:
//   D2_s = C_s C_s^T            occ_xc_build_mo with D2 = NULL (the density matrix never crosses PCIe)
//   F_t  = Phi_t D2q            seminumerical exchange, one product per batch of 128 points
//   Psi_t = Phi_t (P C),  F_t = Psi_t C^T      its factorised form (seminumerical_exchange.cpp:660-815)
// Same tile design as k_density (gemm_common.cuh): CTA 128 x 64, K chunks of 16, 8 warps as 4 x 2 with 32 x 32 warp
// tiles, 3-stage cp.async pipeline, shared-memory leading dimensions = 4 (mod 16) doubles.
//   C[M x N] (ldc) = alpha * A[M x K] * op(B),  all column-major
//   A: element (m, k) at A[k * lda + m]                                   (m contiguous)
//   BT = true : op(B) = B^T, B is N x K, element (n, k) at B[k * ldb + n]   (n contiguous)
//   BT = false: op(B) = B,   B is K x N, element (k, n) at B[n * ldb + k]   (k contiguous)
// Batches: blockIdx.z, operand pointers advance by strideA / strideB / strideC doubles (0 = shared operand).
#include "gemm_common.cuh"
#include <algorithm>

namespace xcb {

namespace {
constexpr int G_LDBT = GN + 4; // BT stage is [k][n]
constexpr int G_BTSZ = GK * G_LDBT;
constexpr int G_STAGE = D_ASZ + (D_BSZ > G_BTSZ ? D_BSZ : G_BTSZ);

template <bool BT, bool VEC>
__global__ void __launch_bounds__(NTHREADS, 2)
    k_dgemm(int M, int N, int K, double alpha, const double *__restrict__ A, int lda, long long strideA,
            const double *__restrict__ B, int ldb, long long strideB, double *__restrict__ C, int ldc, long long strideC) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  const int m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
  A += strideA * blockIdx.z;
  B += strideB * blockIdx.z;
  C += strideC * blockIdx.z;
  const int nk = (K + GK - 1) / GK;

  auto load_stage = [&](int kc, int stage) {
    double *As = smem + stage * G_STAGE;
    double *Bs = As + D_ASZ;
    // A chunk: GK columns of 128 rows
    if (VEC) {
#pragma unroll
      for (int r = 0; r < GK * 64 / NTHREADS; ++r) {
        const int seg = tid + NTHREADS * r;
        const int col = seg >> 6, s2 = (seg & 63) * 2;
        const int k = kc * GK + col, m = m0 + s2;
        const bool ok = k < K && m < M; // M is even on this path, so the pair never straddles the end
        cp_async16(As + col * D_LDA + s2, A + static_cast<size_t>(ok ? k : 0) * lda + (ok ? m : 0), ok);
      }
    } else {
#pragma unroll
      for (int r = 0; r < GK * 128 / NTHREADS; ++r) {
        const int e = tid + NTHREADS * r;
        const int col = e >> 7, mm = e & 127;
        const int k = kc * GK + col, m = m0 + mm;
        const bool ok = k < K && m < M;
        cp_async8(As + col * D_LDA + mm, A + static_cast<size_t>(ok ? k : 0) * lda + (ok ? m : 0), ok);
      }
    }
    if (BT) { // stage [k][n], n contiguous
      if (VEC) {
#pragma unroll
        for (int r = 0; r < GK * 32 / NTHREADS; ++r) {
          const int seg = tid + NTHREADS * r;
          const int col = seg >> 5, s2 = (seg & 31) * 2;
          const int k = kc * GK + col, n = n0 + s2;
          const bool ok = k < K && n < N;
          cp_async16(Bs + col * G_LDBT + s2, B + static_cast<size_t>(ok ? k : 0) * ldb + (ok ? n : 0), ok);
        }
      } else {
#pragma unroll
        for (int r = 0; r < GK * 64 / NTHREADS; ++r) {
          const int e = tid + NTHREADS * r;
          const int col = e >> 6, nn = e & 63;
          const int k = kc * GK + col, n = n0 + nn;
          const bool ok = k < K && n < N;
          cp_async8(Bs + col * G_LDBT + nn, B + static_cast<size_t>(ok ? k : 0) * ldb + (ok ? n : 0), ok);
        }
      }
    } else { // stage [n][k], k contiguous (the layout of k_density's gathered operand)
#pragma unroll
      for (int r = 0; r < GK * 64 / NTHREADS; ++r) {
        const int e = tid + NTHREADS * r;
        const int kk = e % GK, nn = e / GK;
        const int k = kc * GK + kk, n = n0 + nn;
        const bool ok = k < K && n < N;
        cp_async8(Bs + nn * D_LDB + kk, B + static_cast<size_t>(ok ? n : 0) * ldb + (ok ? k : 0), ok);
      }
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

#pragma unroll
  for (int s = 0; s < GSTAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }
  for (int kc = 0; kc < nk; ++kc) {
    cp_async_wait<GSTAGES - 2>();
    __syncthreads();
    const double *As = smem + (kc % GSTAGES) * G_STAGE + wm * 32 + gid;
    const double *Bs = smem + (kc % GSTAGES) * G_STAGE + D_ASZ + (BT ? wn * 32 + gid : (wn * 32 + gid) * D_LDB);
#pragma unroll
    for (int k4 = 0; k4 < GK / 4; ++k4) {
      double a[4], b[4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) a[mi] = As[(k4 * 4 + tig) * D_LDA + mi * 8];
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
        b[ni] = BT ? Bs[(k4 * 4 + tig) * G_LDBT + ni * 8] : Bs[ni * 8 * D_LDB + k4 * 4 + tig];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
      if (k4 == GK / 8 - 1) { // refill the stage consumed last iteration from the middle of this chunk's DMMAs
        if (kc + GSTAGES - 1 < nk) load_stage(kc + GSTAGES - 1, (kc + GSTAGES - 1) % GSTAGES);
        cp_async_commit();
      }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int ni = 0; ni < 4; ++ni)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + wn * 32 + ni * 8 + 2 * tig + h;
      if (n >= N) continue;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const int m = m0 + wm * 32 + mi * 8 + gid;
        if (m < M) C[static_cast<size_t>(n) * ldc + m] = alpha * acc[mi][ni][h];
      }
    }
}

template <bool BT, bool VEC>
void launch_one(int M, int N, int K, double alpha, const double *A, int lda, long long sA, const double *B, int ldb,
                long long sB, double *C, int ldc, long long sC, int batches, cudaStream_t st) {
  const int smem_bytes = GSTAGES * G_STAGE * sizeof(double);
  static SmemOptIn optin;
  optin.ensure(k_dgemm<BT, VEC>, smem_bytes);
  dim3 grid((M + GM - 1) / GM, (N + GN - 1) / GN, batches);
  k_dgemm<BT, VEC><<<grid, NTHREADS, smem_bytes, st>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, C, ldc, sC);
}
} // namespace

void launch_dgemm(bool b_transposed, int M, int N, int K, double alpha, const double *A, int lda, long long strideA,
                  const double *B, int ldb, long long strideB, double *C, int ldc, long long strideC, int batches,
                  void *stream) {
  if (M <= 0 || N <= 0 || batches <= 0) return;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // 16-byte copies need even leading dimensions / strides / extents and 16-byte aligned bases
  auto even = [](long long v) { return (v & 1) == 0; };
  const bool a_vec = even(lda) && even(strideA) && even(M) && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
  const bool b_vec = !b_transposed || (even(ldb) && even(strideB) && even(N) && (reinterpret_cast<uintptr_t>(B) & 15) == 0);
  const bool vec = a_vec && b_vec;
  if (b_transposed) {
    if (vec) launch_one<true, true>(M, N, K, alpha, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, batches, st);
    else launch_one<true, false>(M, N, K, alpha, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, batches, st);
  } else {
    if (vec) launch_one<false, true>(M, N, K, alpha, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, batches, st);
    else launch_one<false, false>(M, N, K, alpha, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, batches, st);
  }
}

// out[i] = s * in[i]  (density_scale / orbital_scale of the host entry points)
__global__ void k_scale(const double *__restrict__ in, double *__restrict__ out, size_t n, double s) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = s * in[i];
}
void launch_scale(const double *in, double *out, size_t n, double s, void *stream) {
  if (n == 0) return;
  const unsigned grid = static_cast<unsigned>(std::min<size_t>((n + 255) / 256, 148 * 16));
  k_scale<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, n, s);
}

// a[i] += b[i]  (folds the low bins of the deterministic accumulation into the high bins)
__global__ void k_axpy1(double *__restrict__ a, const double *__restrict__ b, size_t n) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    a[i] += b[i];
}
void launch_add_inplace(double *a, const double *b, size_t n, void *stream) {
  if (n == 0) return;
  const unsigned grid = static_cast<unsigned>(std::min<size_t>((n + 255) / 256, 148 * 16));
  k_axpy1<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, n);
}

// ---- FP64 tensor-pipe ceiling, measured live (the roofline denominator of bench.py) ----------------------------
// Register-resident mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4, the only f64 MMA shape sm_100a executes): 8 warps per CTA,
// 2 CTAs per SM like the contraction kernels, 16 independent accumulator pairs per thread so the pipe never waits on
// a dependency, no memory traffic inside the loop.  Flops per DMMA: 8 x 8 x 4 x 2 = 512 per warp.
__global__ void __launch_bounds__(NTHREADS, 2) k_dmma_peak(double *out, int iters) {
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a = 1e-3 * (threadIdx.x + 1), b = 1.0 + 1e-6 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(acc[i][0], acc[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
  out[static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x] = s;
}

// returns 0 on success; *tflops = sustained DMMA rate over ~`target_ms` of the current device
int measure_dmma_peak(double target_ms, double *tflops, double *measured_ms) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
  const int blocks = 2 * sms;
  double *d_out = nullptr;
  if (cudaMalloc(&d_out, static_cast<size_t>(blocks) * NTHREADS * sizeof(double)) != cudaSuccess) return 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto run = [&](int iters, float *ms) -> bool {
    cudaEventRecord(e0);
    k_dmma_peak<<<blocks, NTHREADS>>>(d_out, iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess || cudaGetLastError() != cudaSuccess) return false;
    cudaEventElapsedTime(ms, e0, e1);
    return true;
  };
  float ms = 0.f;
  bool ok = run(2000, &ms) && run(20000, &ms); // warm-up, then calibrate
  int iters = 20000;
  if (ok && ms > 0.f) iters = static_cast<int>(std::min(4.0e7, std::max(2.0e4, 20000.0 * target_ms / ms)));
  ok = ok && run(iters, &ms);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  if (!ok || !(ms > 0.f)) return 1;
  const double flops = 512.0 * 16.0 * iters * (NTHREADS / 32.0) * blocks;
  *tflops = flops / (ms * 1e-3) / 1e12;
  if (measured_ms) *measured_ms = ms;
  return 0;
}

} // namespace xcb
