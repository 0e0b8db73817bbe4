// Microbenchmark: legacy mma.sync TF32 m16n8k8 issue rate on sm_100a vs FP32 FFMA, to size the
// 3xTF32 tile-GEMM plan.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC>
__global__ void mma_kernel(float* out, int iters) {
  float c[NACC][4];
  unsigned a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
  b[0] = __float_as_uint(0.5f); b[1] = __float_as_uint(0.25f);
  for (int n = 0; n < NACC; ++n) for (int i = 0; i < 4; ++i) c[n][i] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int n = 0; n < NACC; ++n) mma_tf32(c[n], a, b);
  }
  float s = 0.f;
  for (int n = 0; n < NACC; ++n) for (int i = 0; i < 4; ++i) s += c[n][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ffma_kernel(float* out, int iters) {
  float c[32];
  float a = 1.0f + threadIdx.x * 1e-6f, b = 0.999f;
  for (int n = 0; n < 32; ++n) c[n] = n;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int n = 0; n < 32; ++n) c[n] = fmaf(c[n], a, b);
  }
  float s = 0.f;
  for (int n = 0; n < 32; ++n) s += c[n];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, 148 * 2048 * sizeof(float) * 2);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      mma_kernel<8><<<nsm, warps * 32>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mmas = (double)nsm * warps * iters * 8;
    double macs = mmas * 16 * 8 * 8;
    printf("mma.sync tf32 m16n8k8: %2d warps/SM  %.3f ms  %.1f MMA/us/SM  %.2f TMAC/s dense (%.1f MAC/clk/SM @ %d MHz nominal)\n",
           warps, ms, mmas / nsm / (ms * 1e3), macs / (ms * 1e-3) / 1e12, macs / nsm / (ms * 1e-3) / (khz * 1e3), khz / 1000);
  }
  for (int warps : {8, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      ffma_kernel<<<nsm, warps * 32>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)nsm * warps * 32 * iters * 32;
    printf("ffma: %2d warps/SM  %.3f ms  %.2f TMAC/s (%.1f MAC/clk/SM nominal)\n", warps, ms, fmas / (ms * 1e-3) / 1e12,
           fmas / nsm / (ms * 1e-3) / (khz * 1e3));
  }
  printf("cuda err: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
