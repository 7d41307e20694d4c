// In-loop ceiling of the XC contraction kernels on B200: the exact per-chunk instruction mix of k_vxc_resident
// (per k4 step 2 + 8 LDS.64 -> 16 DMMA.8x8x4, 16 x 64 warp tile) and of k_density (4 + 4 LDS.64 -> 16 DMMA, 32 x 32 warp
// tile) with all operands already in shared memory: no global loads, no cp.async, no barriers, no epilogue.
// 8 warps per CTA, 2 CTAs per SM (shared memory sized like the real kernels).  Whatever these reach is what the
// DMMA + LDS register-feeding pattern allows; the gap between the real kernels and these numbers is pipeline/prologue loss.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_loop dmma_loop.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void mma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 16 x 64 warp tile: A ring slice [16][20] per warp, B resident [64][132]
__global__ void __launch_bounds__(256, 2) k_loop_16x64(double *out, int chunks) {
  extern __shared__ double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
  for (int i = tid; i < 64 * 132 + 8 * 2 * 320; i += 256) smem[i] = 1e-3 * (i % 97);
  __syncthreads();
  const double *Bres = smem, *Aw = smem + 64 * 132 + warp * 640;
  double acc[2][8][2];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  for (int q = 0; q < chunks; ++q) {
    const double *As = Aw + (q & 1) * 320 + gid * 20 + tig;
    const double *Bs = Bres + gid * 132 + tig + (q & 7) * 16;
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      double a[2], b[8];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) a[mi] = As[mi * 8 * 20 + k4 * 4];
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) b[ni] = Bs[ni * 8 * 132 + k4 * 4];
#pragma unroll
      for (int ni = 0; ni < 8; ++ni)
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) mma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
  }
  double s = 0;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) s += acc[mi][ni][0] + acc[mi][ni][1];
  out[blockIdx.x * 256 + tid] = s;
}

// 32 x 32 warp tile: A stage [16 k][132], B stage [64 n][20], 3 stages (k_density)
__global__ void __launch_bounds__(256, 2) k_loop_32x32(double *out, int chunks) {
  extern __shared__ double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  constexpr int STAGE = 16 * 132 + 64 * 20;
  for (int i = tid; i < 3 * STAGE; i += 256) smem[i] = 1e-3 * (i % 97);
  __syncthreads();
  double acc[4][4][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  for (int q = 0; q < chunks; ++q) {
    const double *As = smem + (q % 3) * STAGE + wm * 32 + gid;
    const double *Bs = smem + (q % 3) * STAGE + 16 * 132 + (wn * 32 + gid) * 20;
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      double a[4], b[4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) a[mi] = As[(k4 * 4 + tig) * 132 + mi * 8];
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[ni * 8 * 20 + k4 * 4 + tig];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) mma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
  }
  double s = 0;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) s += acc[mi][ni][0] + acc[mi][ni][1];
  out[blockIdx.x * 256 + tid] = s;
}

template <typename F> float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int nsm = p.multiProcessorCount, chunks = 40000;
  double *out; CK(cudaMalloc(&out, sizeof(double) * nsm * 4 * 256));
  const int smem1 = (64 * 132 + 8 * 2 * 320) * 8, smem2 = 3 * (16 * 132 + 64 * 20) * 8;
  CK(cudaFuncSetAttribute(k_loop_16x64, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
  CK(cudaFuncSetAttribute(k_loop_32x32, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
  for (int cps : {1, 2}) {
    const int blocks = nsm * cps;
    const double flops = (double)blocks * 8 * chunks * 64 * 512.0;
    float ms = time_ms([&] { k_loop_16x64<<<blocks, 256, smem1>>>(out, chunks); });
    printf("{\"bench\": \"loop_16x64_vxc_mix\", \"ctas_per_sm\": %d, \"tflops\": %.2f}\n", cps, flops / ms / 1e9);
    ms = time_ms([&] { k_loop_32x32<<<blocks, 256, smem2>>>(out, chunks); });
    printf("{\"bench\": \"loop_32x32_density_mix\", \"ctas_per_sm\": %d, \"tflops\": %.2f}\n", cps, flops / ms / 1e9);
  }
  return 0;
}
