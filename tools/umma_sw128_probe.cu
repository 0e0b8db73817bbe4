// Bring-up probe for the round-2 tcgen05 backward: 128-byte-swizzled operands for kind::tf32.
//
// The backward needs every activation-gradient tile Z (128 rows x F features) in TWO contractions:
//   dW  = Z^T X   contraction over the ROWS      (weight gradient)
//   dX  = Z  W    contraction over the FEATURES  (input gradient)
// One physical layout serves both if the tensor core accepts it under two descriptors:
//   bytes   : atom column q = rows 32q .. 32q+31; inside it feature f owns a 128-byte line holding its 32 rows,
//             16-byte chunks XOR-swizzled with (f % 8)  (Swizzle<3,4,3>, the TMA / UMMA SWIZZLE_128B pattern).
//             A row-owning thread writes Z[r][f] with a 4-byte store: the 32 lanes of a warp (32 rows, same f) fill
//             exactly one 128-byte line -> conflict-free, and the same image in global memory is a coalesced store.
//   view 1  : K-major  A (M = feature, K = row):   SBO = 1024 (8-feature group), K step = +32 B inside the line
//   view 2  : MN-major A (M = row,     K = feature): LBO = atom-column stride, SBO = 1024, K step = +1024 B
// Test 1 checks view 1 (A and B), test 2 checks view 2 against a no-swizzle K-major B (the library's weight format),
// test 3 measures tcgen05.mma issue cost vs N for dependent / interleaved accumulator chains.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/umma_sw128_probe tools/umma_sw128_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../loopy_slam_b200/csrc/lsr_umma.cuh"

using namespace lsr::umma;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)

struct Cfg {
  int test;      // 1 / 2 / 3
  int layout;    // descriptor layout_type of the swizzled operands (2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B)
  int swz;       // 0: Swizzle<3,4,3>, 1: Swizzle<2,5,2>
  int lbo, sbo;  // test 2: fields of the MN-major A descriptor; test 1: lbo only (sbo fixed 1024)
  int bmn;       // test 2: 1 = describe B (the same swizzled bytes of X^T) as MN-major too instead of the no-swizzle weights
  int n, nmma, nacc;   // test 3
};

__device__ __forceinline__ uint64_t desc_l(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return smem_desc(addr, lbo, sbo) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ uint32_t swz_off(uint32_t off, int swz) {
  return swz == 0 ? (off ^ (((off >> 7) & 7u) << 4)) : (off ^ (((off >> 7) & 3u) << 5));
}

constexpr int OFF_ZT = 0;                       // [4 atom columns][128 features][128 B] = 64 KB
constexpr int OFF_XT = OFF_ZT + 4 * 128 * 128;  // [4][64 features][128 B] = 32 KB
constexpr int OFF_WN = OFF_XT + 4 * 64 * 128;   // no-swizzle K-major [N = 128][K = 128]: 32 slabs x 2048 B = 64 KB
constexpr int SMEM_TOTAL = OFF_WN + 32 * 2048 + 1024;

__global__ void __launch_bounds__(128) probe(const float* Z, const float* X, const float* W, float* D, long long* cyc, Cfg c) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tslot, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  const int r = tid, q = r >> 5;
  // Z^T, X^T: thread = row, 4-byte transposing stores into the swizzled image
  for (int f = 0; f < 128; ++f)
    *reinterpret_cast<float*>(smem + OFF_ZT + q * 16384 + swz_off(f * 128 + (r & 31) * 4, c.swz)) = Z[r * 128 + f];
  for (int f = 0; f < 64; ++f)
    *reinterpret_cast<float*>(smem + OFF_XT + q * 8192 + swz_off(f * 128 + (r & 31) * 4, c.swz)) = X[r * 64 + f];
  // W[i][f] (thread = i): canonical no-swizzle K-major, element (i, f) at (f / 4) * 2048 + i * 16 + (f % 4) * 4
  for (int f = 0; f < 128; f += 4)
    *reinterpret_cast<float4*>(smem + OFF_WN + (f >> 2) * 2048 + tid * 16) = *reinterpret_cast<const float4*>(W + tid * 128 + f);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  int ncols = 0;
  long long t0 = 0;
  if (c.test == 1) {
    ncols = 64;
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = idesc_tf32(128, 64);
      uint32_t acc = 0;
      for (int k8 = 0; k8 < 16; ++k8) {
        const int qq = k8 >> 2, j = k8 & 3;
        const uint64_t da = desc_l(smem_u32(smem + OFF_ZT) + qq * 16384 + j * 32, c.lbo, 1024, c.layout);
        const uint64_t db = desc_l(smem_u32(smem + OFF_XT) + qq * 8192 + j * 32, c.lbo, 1024, c.layout);
        mma_ss(tb, da, db, idesc, acc);
        acc = 1u;
      }
      mma_commit(&bar);
    }
  } else if (c.test == 2) {
    ncols = 128;
    if (tid == 0) {
      tc_fence_after();
      if (!c.bmn) {
        // D[r][i] = sum_f Z[r][f] W[i][f]
        const uint32_t idesc = idesc_tf32(128, 128) | (1u << 15);   // A MN-major
        uint32_t acc = 0;
        for (int k8 = 0; k8 < 16; ++k8) {
          const uint64_t da = desc_l(smem_u32(smem + OFF_ZT) + k8 * 1024, c.lbo, c.sbo, c.layout);
          const uint64_t db = smem_desc(smem_u32(smem + OFF_WN) + k8 * 2 * 2048, 2048, 128);
          mma_ss(tb, da, db, idesc, acc);
          acc = 1u;
        }
      } else {
        // D[f][j] = sum_r Z^T[f][r] X^T[j][r] with A K-major (view 1) and B = X^T described MN-major?  No: B MN-major means
        // B[K][N] with N contiguous -- here K = feature of X (64), N = rows (128): D[r'][r] = sum_j Z? not a needed product.
        // Instead: D[r][j'] over K = feature f < 64 with A = Z (MN-major, first 64 features) and B = X^T read as
        // B[N = row][K = feature]: that is X itself, K-major would need features contiguous -> MN-major B view of X^T bytes
        // is B[K = feature][N = row].  D[m][n] = sum_f Z[m][f] X[n][f], N = 128 rows.
        const uint32_t idesc = idesc_tf32(128, 128) | (1u << 15) | (1u << 16);
        uint32_t acc = 0;
        for (int k8 = 0; k8 < 8; ++k8) {
          const uint64_t da = desc_l(smem_u32(smem + OFF_ZT) + k8 * 1024, c.lbo, c.sbo, c.layout);
          const uint64_t db = c.lbo > c.sbo ? desc_l(smem_u32(smem + OFF_XT) + k8 * 1024, c.lbo / 2, c.sbo, c.layout)   // X^T atom column = 8 KB
                                          : desc_l(smem_u32(smem + OFF_XT) + k8 * 1024, c.lbo, c.sbo / 2, c.layout);
          mma_ss(tb, da, db, idesc, acc);
          acc = 1u;
        }
        ncols = 128;
      }
      mma_commit(&bar);
    }
  } else {
    // timing: nmma MMAs of shape 128 x n x 8 round-robin over nacc accumulators (data irrelevant)
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = idesc_tf32(128, c.n);
      t0 = clock64();
      for (int i = 0; i < c.nmma; ++i) {
        const int k8 = i & 15;
        const uint64_t da = smem_desc(smem_u32(smem + OFF_WN) + k8 * 2 * 2048, 2048, 128);
        const uint64_t db = smem_desc(smem_u32(smem + OFF_ZT) + (k8 & 7) * 2 * (uint32_t)c.n * 16, (uint32_t)c.n * 16, 128);
        mma_ss(tb + (uint32_t)((i % c.nacc) * c.n), da, db, idesc, i >= c.nacc ? 1u : 0u);
      }
      mma_commit(&bar);
    }
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (c.test == 3) {
    if (tid == 0) cyc[0] = clock64() - t0;
  } else {
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_addr(tb, 32 * warp, c0), v);
      tmem_wait_ld();
      for (int j = 0; j < 32; ++j) D[tid * ncols + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  std::vector<float> Z(128 * 128), X(128 * 64), W(128 * 128);
  for (int r = 0; r < 128; ++r)
    for (int f = 0; f < 128; ++f) Z[r * 128 + f] = (float)(((r * 7 + f * 3 + (r * f) % 5) % 9) - 4);
  for (int r = 0; r < 128; ++r)
    for (int j = 0; j < 64; ++j) X[r * 64 + j] = (float)(((r * 5 + j * 11 + (r * j) % 3) % 7) - 3);
  for (int i = 0; i < 128; ++i)
    for (int f = 0; f < 128; ++f) W[i * 128 + f] = (float)(((i * 3 + f * 5 + (i * f) % 7) % 5) - 2);
  float *dZ, *dX, *dW, *dD;
  long long* dC;
  CK(cudaMalloc(&dZ, Z.size() * 4)); CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4));
  CK(cudaMalloc(&dD, 128 * 128 * 4)); CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dZ, Z.data(), Z.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
  int ok_any[3] = {0, 0, 0};
  auto run = [&](Cfg c, const char* name) {
    CK(cudaMemset(dD, 0xff, 128 * 128 * 4));
    probe<<<1, 128, SMEM_TOTAL>>>(dZ, dX, dW, dD, dC, c);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-58s CUDA error: %s\n", name, cudaGetErrorString(e)); exit(3); }
    if (c.test == 3) {
      long long cy;
      CK(cudaMemcpy(&cy, dC, 8, cudaMemcpyDeviceToHost));
      printf("%-40s N=%3d nacc=%d: %6lld cycles for %d MMAs = %.1f cycles / MMA\n", name, c.n, c.nacc, cy, c.nmma, (double)cy / c.nmma);
      return;
    }
    const int ncols = c.test == 1 ? 64 : 128;
    std::vector<float> D(128 * ncols);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    float first_got = 0, first_exp = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < ncols; ++n) {
        double ref = 0;
        if (c.test == 1) for (int rr = 0; rr < 128; ++rr) ref += (double)Z[rr * 128 + m] * X[rr * 64 + n];
        else if (!c.bmn) for (int f = 0; f < 128; ++f) ref += (double)Z[m * 128 + f] * W[n * 128 + f];
        else for (int f = 0; f < 64; ++f) ref += (double)Z[m * 128 + f] * X[n * 64 + f];
        if (D[m * ncols + n] != (float)ref) { if (!bad) { first_got = D[m * ncols + n]; first_exp = (float)ref; } ++bad; }
      }
    printf("%-58s %s  (%d / %d wrong; first got %g expected %g)\n", name, bad ? "MISMATCH" : "EXACT", bad, 128 * ncols, first_got, first_exp);
    if (!bad) ok_any[c.test] = 1;
  };
  // test 1: K-major SWIZZLE_128B for A and B (row contraction)
  run(Cfg{1, 2, 0, 0, 1024, 0, 0, 0, 0}, "T1 K-major SW128 (layout 2, swz<3,4,3>, LBO 0)");
  run(Cfg{1, 2, 0, 16, 1024, 0, 0, 0, 0}, "T1 K-major SW128 (layout 2, swz<3,4,3>, LBO 16)");
  run(Cfg{1, 1, 1, 0, 1024, 0, 0, 0, 0}, "T1 K-major SW128_BASE32B (layout 1, swz<2,5,2>)");
  // test 2: MN-major A view of the same bytes
  run(Cfg{2, 2, 0, 16384, 1024, 0, 0, 0, 0}, "T2 A MN-major SW128 (LBO 16K, SBO 1K)");
  run(Cfg{2, 2, 0, 1024, 16384, 0, 0, 0, 0}, "T2 A MN-major SW128 (LBO 1K, SBO 16K)");
  run(Cfg{2, 1, 1, 16384, 1024, 0, 0, 0, 0}, "T2 A MN-major SW128_BASE32B swz<2,5,2> (LBO 16K, SBO 1K)");
  run(Cfg{2, 1, 1, 1024, 16384, 0, 0, 0, 0}, "T2 A MN-major SW128_BASE32B swz<2,5,2> (LBO 1K, SBO 16K)");
  run(Cfg{2, 1, 0, 16384, 1024, 0, 0, 0, 0}, "T2 A MN-major SW128_BASE32B swz<3,4,3> (LBO 16K, SBO 1K)");
  run(Cfg{2, 2, 0, 16384, 1024, 1, 0, 0, 0}, "T2 A and B MN-major SW128 (LBO 16K/8K, SBO 1K)");
  run(Cfg{2, 2, 0, 1024, 16384, 1, 0, 0, 0}, "T2 A and B MN-major SW128 (LBO 1K/512, SBO 16K)");
  // test 4: does the tensor core TRUNCATE fp32 operands to tf32 (ignore the low 13 mantissa bits)?  If so the raw fp32 image
  // can serve as the "hi" operand of the 3xTF32 split and only the "lo" image has to be produced.
  {
    std::vector<float> Zr(Z.size()), Xr(X.size()), Zm(Z.size()), Xm(X.size());
    srand(7);
    for (size_t i = 0; i < Z.size(); ++i) { Zr[i] = ((float)rand() / RAND_MAX - 0.5f) * 3.f; uint32_t u; memcpy(&u, &Zr[i], 4); u &= 0xffffe000u; memcpy(&Zm[i], &u, 4); }
    for (size_t i = 0; i < X.size(); ++i) { Xr[i] = ((float)rand() / RAND_MAX - 0.5f) * 3.f; uint32_t u; memcpy(&u, &Xr[i], 4); u &= 0xffffe000u; memcpy(&Xm[i], &u, 4); }
    std::vector<float> D1(128 * 64), D2(128 * 64);
    Cfg c{1, 2, 0, 0, 1024, 0, 0, 0, 0};
    CK(cudaMemcpy(dZ, Zr.data(), Z.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dX, Xr.data(), X.size() * 4, cudaMemcpyHostToDevice));
    probe<<<1, 128, SMEM_TOTAL>>>(dZ, dX, dW, dD, dC, c); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D1.data(), dD, D1.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(dZ, Zm.data(), Z.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dX, Xm.data(), X.size() * 4, cudaMemcpyHostToDevice));
    probe<<<1, 128, SMEM_TOTAL>>>(dZ, dX, dW, dD, dC, c); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D2.data(), dD, D2.size() * 4, cudaMemcpyDeviceToHost));
    int diff = 0;
    for (size_t i = 0; i < D1.size(); ++i) diff += memcmp(&D1[i], &D2[i], 4) != 0;
    printf("T4 raw fp32 operands vs operands masked to 19 bits: %d / %zu results differ  -> %s\n", diff, D1.size(),
           diff ? "the hardware ROUNDS (or uses) the low bits" : "the hardware TRUNCATES: raw image == hi image");
    CK(cudaMemcpy(dZ, Z.data(), Z.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
  }
  // test 3: issue cost
  for (int n : {32, 64, 128, 208, 256})
    for (int nacc : {1, 2}) {
      if (n * nacc > 512) continue;
      run(Cfg{3, 0, 0, 0, 0, 0, n, 96, nacc}, "T3 dependent / interleaved chains");
    }
  printf("summary: T1 %s, T2 %s\n", ok_any[1] ? "OK" : "FAILED", ok_any[2] ? "OK" : "FAILED");
  return 0;
}
