// Microbenchmarks that set the roofline denominators for the XC build on B200 (sm_100a):
//   * FP64 tensor (DMMA, mma.sync f64) peak for each legal shape
//   * FP64 CUDA-core (DFMA) peak, and DMMA+DFMA issued concurrently from different warps
//   * FP64 exp() throughput (collocation primitive loop)
//   * red.global.add.f64 throughput (Vxc scatter)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void mma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void k_mma884(double *out, int iters) {
  double c[NACC][2];
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) mma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_mma1684(double *out, int iters) {
  double c[NACC][4]; double a[2] = {threadIdx.x * 1e-3, 0.5}; double b = threadIdx.x * 2e-3;
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) mma1684(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_mma1688(double *out, int iters) {
  double c[NACC][4]; double a[4] = {threadIdx.x * 1e-3, 0.5, 0.25, 0.125}; double b[2] = {threadIdx.x * 2e-3, 1.0};
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) mma1688(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_mma16816(double *out, int iters) {
  double c[NACC][4]; double a[8]; double b[4];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = threadIdx.x * 2e-3 - i;
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) mma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double *out, int iters) {
  double c[NACC]; double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// warps with (warp_id % 2 == 0) issue DMMA m16n8k8; others issue DFMA. iters_fma scaled so both take similar time.
__global__ void k_mixed(double *out, int iters_mma, int iters_fma) {
  int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp & 1) {
    double c[16]; double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = i;
    for (int it = 0; it < iters_fma; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
  } else {
    double c[8][4]; double a[4] = {threadIdx.x * 1e-3, 0.5, 0.25, 0.125}; double b[2] = {threadIdx.x * 2e-3, 1.0};
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters_mma; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) mma1688(c[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double *out, int iters) {
  double x = -1e-3 * (threadIdx.x + 1), s = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) { s += exp(x); x -= 1e-4; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// each thread adds to `per` addresses; consecutive lanes -> consecutive doubles (coalesced), tiles spread over buffer
__global__ void k_red(double *buf, size_t n, int per) {
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int i = 0; i < per; i++) {
    size_t idx = (tid + (size_t)i * stride * 7919) % n;
    asm volatile("red.global.add.f64 [%0], %1;" :: "l"(buf + idx), "d"(1.0) : "memory");
  }
}
// scatter pattern like the Vxc tile: each warp adds 8 runs of 4 consecutive doubles with column stride ld
__global__ void k_red_tile(double *buf, size_t ld, size_t ncol, int per) {
  int lane = threadIdx.x & 31;
  size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  for (int i = 0; i < per; i++) {
    size_t col = (warp * 131 + (size_t)i * 17 + (lane >> 2)) % ncol;
    size_t row = (warp * 29 + (size_t)i * 8 + (lane & 3) * 2) % (ld - 2);
    double *p = buf + col * ld + row;
    asm volatile("red.global.add.f64 [%0], %1;" :: "l"(p), "d"(1.0) : "memory");
    asm volatile("red.global.add.f64 [%0], %1;" :: "l"(p + 1), "d"(1.0) : "memory");
  }
}

template <typename F>
float time_ms(F f, int reps = 5) {
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
  int nsm = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\"}\n", p.name, nsm, p.major, p.minor);
  double *out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
  const int iters = 20000;
  for (int wpb : {4, 8, 16, 32}) {
    int threads = wpb * 32; int blocks = nsm * (wpb <= 8 ? 2 : 1);
    double nwarps = (double)blocks * wpb;
    float ms;
    ms = time_ms([&] { k_mma884<8><<<blocks, threads>>>(out, iters); });
    printf("{\"bench\": \"dmma_m8n8k4\", \"warps_per_block\": %d, \"blocks\": %d, \"tflops\": %.2f}\n", wpb, blocks, nwarps * iters * 8 * 2.0 * 8 * 8 * 4 / ms / 1e9);
    ms = time_ms([&] { k_mma1684<8><<<blocks, threads>>>(out, iters); });
    printf("{\"bench\": \"dmma_m16n8k4\", \"warps_per_block\": %d, \"blocks\": %d, \"tflops\": %.2f}\n", wpb, blocks, nwarps * iters * 8 * 2.0 * 16 * 8 * 4 / ms / 1e9);
    ms = time_ms([&] { k_mma1688<8><<<blocks, threads>>>(out, iters); });
    printf("{\"bench\": \"dmma_m16n8k8\", \"warps_per_block\": %d, \"blocks\": %d, \"tflops\": %.2f}\n", wpb, blocks, nwarps * iters * 8 * 2.0 * 16 * 8 * 8 / ms / 1e9);
    ms = time_ms([&] { k_mma16816<8><<<blocks, threads>>>(out, iters / 2); });
    printf("{\"bench\": \"dmma_m16n8k16\", \"warps_per_block\": %d, \"blocks\": %d, \"tflops\": %.2f}\n", wpb, blocks, nwarps * (iters / 2) * 8 * 2.0 * 16 * 8 * 16 / ms / 1e9);
    ms = time_ms([&] { k_dfma<16><<<blocks, threads>>>(out, iters); });
    printf("{\"bench\": \"dfma\", \"warps_per_block\": %d, \"blocks\": %d, \"tflops\": %.2f}\n", wpb, blocks, nwarps * 32 * iters * 16 * 2.0 / ms / 1e9);
  }
  // accumulator-count (ILP) sweep for m16n8k8 at 8 warps/block, 1 block/SM (the XC kernels' shape)
  {
    int threads = 256, blocks = nsm; double nwarps = blocks * 8.0; float ms;
    ms = time_ms([&] { k_mma1688<1><<<blocks, threads>>>(out, iters); });
    printf("{\"bench\": \"dmma_m16n8k8_ilp1\", \"tflops\": %.2f}\n", nwarps * iters * 1 * 2.0 * 16 * 8 * 8 / ms / 1e9);
    ms = time_ms([&] { k_mma1688<2><<<blocks, threads>>>(out, iters); });
    printf("{\"bench\": \"dmma_m16n8k8_ilp2\", \"tflops\": %.2f}\n", nwarps * iters * 2 * 2.0 * 16 * 8 * 8 / ms / 1e9);
    ms = time_ms([&] { k_mma1688<4><<<blocks, threads>>>(out, iters); });
    printf("{\"bench\": \"dmma_m16n8k8_ilp4\", \"tflops\": %.2f}\n", nwarps * iters * 4 * 2.0 * 16 * 8 * 8 / ms / 1e9);
    ms = time_ms([&] { k_mma1688<16><<<blocks, threads>>>(out, iters); });
    printf("{\"bench\": \"dmma_m16n8k8_ilp16\", \"tflops\": %.2f}\n", nwarps * iters * 16 * 2.0 * 16 * 8 * 8 / ms / 1e9);
  }
  // mixed: DMMA on even warps, DFMA on odd warps
  {
    int threads = 512, blocks = nsm; float ms;
    int im = 20000;
    for (int ifm : {0, 20000, 40000, 80000}) {
      ms = time_ms([&] { k_mixed<<<blocks, threads>>>(out, im, ifm); });
      double fl_mma = blocks * 8.0 * im * 8 * 2.0 * 16 * 8 * 8, fl_fma = blocks * 8.0 * 32 * (double)ifm * 16 * 2.0;
      printf("{\"bench\": \"mixed_dmma_dfma\", \"iters_fma\": %d, \"ms\": %.3f, \"tflops_mma\": %.2f, \"tflops_fma\": %.2f, \"tflops_total\": %.2f}\n",
             ifm, ms, fl_mma / ms / 1e9, fl_fma / ms / 1e9, (fl_mma + fl_fma) / ms / 1e9);
    }
    ms = time_ms([&] { k_mixed<<<blocks, threads>>>(out, 0, 80000); });
    printf("{\"bench\": \"mixed_dfma_only_oddwarps\", \"ms\": %.3f, \"tflops_fma\": %.2f}\n", ms, blocks * 8.0 * 32 * 80000.0 * 16 * 2.0 / ms / 1e9);
  }
  {
    int threads = 256, blocks = nsm * 8; float ms = time_ms([&] { k_exp<<<blocks, threads>>>(out, 2000); });
    printf("{\"bench\": \"exp_f64\", \"gexp_per_s\": %.2f}\n", (double)blocks * threads * 2000 * 8 / ms / 1e6);
  }
  {
    size_t n = 3348ull * 3348ull; double *buf; CK(cudaMalloc(&buf, n * 8)); CK(cudaMemset(buf, 0, n * 8));
    int threads = 256, blocks = nsm * 8, per = 256; float ms;
    ms = time_ms([&] { k_red<<<blocks, threads>>>(buf, n, per); });
    printf("{\"bench\": \"red_add_f64_coalesced_spread\", \"gatomics_per_s\": %.2f}\n", (double)blocks * threads * per / ms / 1e6);
    ms = time_ms([&] { k_red_tile<<<blocks, threads>>>(buf, 3348, 3348, per); });
    printf("{\"bench\": \"red_add_f64_tile_scatter\", \"gatomics_per_s\": %.2f}\n", (double)blocks * threads * per * 2 / ms / 1e6);
    size_t n2 = 512ull * 512ull;
    ms = time_ms([&] { k_red<<<blocks, threads>>>(buf, n2, per); });
    printf("{\"bench\": \"red_add_f64_coalesced_2MB\", \"gatomics_per_s\": %.2f}\n", (double)blocks * threads * per / ms / 1e6);
  }
  return 0;
}
