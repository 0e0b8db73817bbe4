// tcgen05.mma issue-rate probe: cycles per M = 128, K = 8 tf32 MMA as a function of N, operand source (SS / TS) and the
// number of independent accumulators, with the issue loop fully unrolled and descriptors advanced by constants (the way
// the library's issuer does it).  One CTA, one issuing lane.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/umma_rate_probe tools/umma_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include "../loopy_slam_b200/csrc/lsr_umma.cuh"
using namespace lsr::umma;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

template <int N, int NACC, bool TS, int NMMA>
__global__ void __launch_bounds__(128) rate(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (warp == 0) tmem_alloc(&tslot, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  long long t0 = 0;
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = idesc_tf32(128, N);
      const uint32_t a_lo = ((2048u >> 4) << 16) | ((smem_u32(smem) >> 4) & 0x3fffu);
      const uint32_t b_lo = (((uint32_t)N * 16u >> 4) << 16) | ((smem_u32(smem + 65536) >> 4) & 0x3fffu);
      const uint32_t hiw = (128u >> 4) | (1u << 14);
      t0 = clock64();
#pragma unroll
      for (int i = 0; i < NMMA; ++i) {
        const uint32_t d = tb + (TS ? 256 : 0) * 0 + (uint32_t)((i % NACC) * N) + (TS ? 0 : 0);
        const uint64_t db = ((uint64_t)hiw << 32) | (b_lo + (uint32_t)((i & 3) * ((2 * N * 16) >> 4)));
        if (TS) {
          mma_ts(d, tb + 512 - 64 + (uint32_t)((i & 3) * 8), db, idesc, 1u);
        } else {
          const uint64_t da = ((uint64_t)hiw << 32) | (a_lo + (uint32_t)((i & 3) * 256));
          mma_ss(d, da, db, idesc, 1u);
        }
      }
      mma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (tid == 0) out[0] = clock64() - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

template <int N, int NACC, bool TS>
static void run(long long* d) {
  constexpr int NMMA = 96;
  CK(cudaFuncSetAttribute(rate<N, NACC, TS, NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  long long best = 1ll << 60;
  for (int rep = 0; rep < 3; ++rep) {
    rate<N, NACC, TS, NMMA><<<1, 128, 160 * 1024>>>(d);
    CK(cudaDeviceSynchronize());
    long long c;
    CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
    if (c < best) best = c;
  }
  printf("%s N=%3d nacc=%d: %6lld cycles / %d MMAs = %6.1f cycles per MMA (floor N/2 = %d)\n", TS ? "TS" : "SS", N, NACC, best, NMMA,
         (double)best / NMMA, N / 2);
}

int main() {
  long long* d;
  CK(cudaMalloc(&d, 8));
  run<16, 1, false>(d); run<32, 1, false>(d); run<64, 1, false>(d); run<128, 1, false>(d); run<208, 1, false>(d); run<256, 1, false>(d);
  run<32, 2, false>(d); run<64, 2, false>(d); run<128, 2, false>(d); run<32, 4, false>(d); run<64, 4, false>(d);
  run<32, 1, true>(d); run<64, 1, true>(d); run<128, 1, true>(d); run<208, 1, true>(d);
  run<64, 2, true>(d); run<128, 2, true>(d);
  return 0;
}
