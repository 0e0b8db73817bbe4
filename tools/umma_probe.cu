// Bring-up probe of the tcgen05 building blocks (lsr_umma.cuh / lsr_umma_prog.cuh) on a real B200:
//   1. one 128 x N x K GEMM, operands laid out by hand, A from shared memory (SS) or from TMEM (TS), exact on
//      tf32-representable inputs (mode 4: the other assignment of the descriptor LBO / SBO fields, which faults);
//   2. the streamed-weight engine (producer / issuer / epilogue warps) on a colour-trunk-shaped 5-layer MLP,
//      checked against an fp64 host evaluation, then timed over many tiles on all SMs;
//   3. (open) weight-gradient shaped GEMMs with MN-major operands: kind::tf32 + SWIZZLE_NONE yields zeros so far;
//   5. the same GEMMs on the verified K-major path: operands transposed by the row-owning threads, padded LBO.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/umma_probe tools/umma_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../loopy_slam_b200/csrc/lsr_umma_prog.cuh"

using namespace lsr::umma;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

// ------------------------------------------------------------------------------------------- test 1
// D = A . B^T, A (128 x K) and B (N x K) row-major in global.  swap: exchange the LBO / SBO fields.
__global__ void __launch_bounds__(128) probe_gemm(const float* A, const float* B, float* D, int K, int N, int swap, int ts,
                                                  int three) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* sAhi = smem;
  uint8_t* sAlo = sAhi + 128 * K * 4;
  uint8_t* sBhi = sAlo + 128 * K * 4;
  uint8_t* sBlo = sBhi + N * K * 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tslot, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  // operands -> canonical layout
  for (int e = tid; e < 128 * K; e += 128) {
    const int r = e / K, k = e % K;
    uint32_t hi, lo;
    split_hi_lo(A[e], hi, lo);
    const int off = (k >> 2) * 128 * 16 + r * 16 + (k & 3) * 4;
    *reinterpret_cast<uint32_t*>(sAhi + off) = three ? hi : __float_as_uint(A[e]);
    *reinterpret_cast<uint32_t*>(sAlo + off) = lo;
  }
  for (int e = tid; e < N * K; e += 128) {
    const int r = e / K, k = e % K;
    uint32_t hi, lo;
    split_hi_lo(B[e], hi, lo);
    const int off = (k >> 2) * N * 16 + r * 16 + (k & 3) * 4;
    *reinterpret_cast<uint32_t*>(sBhi + off) = three ? hi : __float_as_uint(B[e]);
    *reinterpret_cast<uint32_t*>(sBlo + off) = lo;
  }
  if (ts) {   // A -> TMEM columns [256, 256 + K) (hi) and [384, 384 + K) (lo); thread = row
    for (int k0 = 0; k0 < K; k0 += 32) {
      uint32_t vh[32], vl[32];
      for (int j = 0; j < 32; ++j) {
        const float x = (k0 + j < K) ? A[tid * K + k0 + j] : 0.f;
        split_hi_lo(x, vh[j], vl[j]);
        if (!three) vh[j] = __float_as_uint(x);
      }
      tmem_st32(tmem_addr(tb, 32 * warp, 256 + k0), vh);
      tmem_st32(tmem_addr(tb, 32 * warp, 384 + k0), vl);
    }
    tmem_wait_st();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = idesc_tf32(128, N);
    const uint32_t a_slab = 128 * 16, b_slab = N * 16;
    uint32_t acc = 0;
    for (int k8 = 0; k8 < K / 8; ++k8) {
      const uint32_t ahi = smem_u32(sAhi) + k8 * 2 * a_slab, alo = smem_u32(sAlo) + k8 * 2 * a_slab;
      const uint32_t bhi = smem_u32(sBhi) + k8 * 2 * b_slab, blo = smem_u32(sBlo) + k8 * 2 * b_slab;
      const uint64_t dah = swap ? smem_desc(ahi, 128, a_slab) : smem_desc(ahi, a_slab, 128);
      const uint64_t dal = swap ? smem_desc(alo, 128, a_slab) : smem_desc(alo, a_slab, 128);
      const uint64_t dbh = swap ? smem_desc(bhi, 128, b_slab) : smem_desc(bhi, b_slab, 128);
      const uint64_t dbl = swap ? smem_desc(blo, 128, b_slab) : smem_desc(blo, b_slab, 128);
      if (ts) {
        const uint32_t th = tb + 256 + k8 * 8, tl = tb + 384 + k8 * 8;
        if (three) { mma_ts(tb, tl, dbh, idesc, acc); mma_ts(tb, th, dbl, idesc, 1u); acc = 1u; }
        mma_ts(tb, th, dbh, idesc, acc);
      } else {
        if (three) { mma_ss(tb, dal, dbh, idesc, acc); mma_ss(tb, dah, dbl, idesc, 1u); acc = 1u; }
        mma_ss(tb, dah, dbh, idesc, acc);
      }
      acc = 1u;
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_addr(tb, 32 * warp, c0), v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) D[tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------------- test 3
// Weight-gradient shaped GEMM: D[m][n] = sum_r A[r][m] * B[r][n] (contraction over the 128 ROWS), A (128 x 128) and
// B (128 x N) stored exactly like the forward's activation tiles (row-thread float4 stores: (f/4) * 2048 + r * 16 +
// (f % 4) * 4), but described to the tensor core as MN-major operands.  swap: exchange the LBO / SBO fields.
__global__ void __launch_bounds__(128) probe_gemm_mn(const float* A, const float* B, float* D, int N, int swap, int three, int which) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* sAhi = smem;
  uint8_t* sAlo = sAhi + 128 * 128 * 4;
  uint8_t* sBhi = sAlo + 128 * 128 * 4;
  uint8_t* sBlo = sBhi + 128 * N * 4;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tslot, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  for (int f = 0; f < 128; f += 4) {   // thread = row
    float4 v = *reinterpret_cast<const float4*>(A + tid * 128 + f);
    if (!three) { store_a_split(sAhi, sAlo, tid, f, v); *reinterpret_cast<float4*>(sAhi + (f >> 2) * 2048 + tid * 16) = v; }
    else store_a_split(sAhi, sAlo, tid, f, v);
  }
  for (int f = 0; f < N; f += 4) {
    float4 v = *reinterpret_cast<const float4*>(B + tid * N + f);
    if (!three) { store_a_split(sBhi, sBlo, tid, f, v); *reinterpret_cast<float4*>(sBhi + (f >> 2) * 2048 + tid * 16) = v; }
    else store_a_split(sBhi, sBlo, tid, f, v);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = idesc_tf32(128, N) | ((which & 1) ? (1u << 15) : 0u) | ((which & 2) ? (1u << 16) : 0u);   // A / B MN-major
    uint32_t acc = 0;
    for (int k8 = 0; k8 < 16; ++k8) {   // 8 rows per MMA: one 128-byte core matrix along K
      const uint32_t off = k8 * 128;
      const uint64_t dah = swap ? smem_desc(smem_u32(sAhi) + off, 2048, 128) : smem_desc(smem_u32(sAhi) + off, 128, 2048);
      const uint64_t dal = swap ? smem_desc(smem_u32(sAlo) + off, 2048, 128) : smem_desc(smem_u32(sAlo) + off, 128, 2048);
      const uint64_t dbh = swap ? smem_desc(smem_u32(sBhi) + off, 2048, 128) : smem_desc(smem_u32(sBhi) + off, 128, 2048);
      const uint64_t dbl = swap ? smem_desc(smem_u32(sBlo) + off, 2048, 128) : smem_desc(smem_u32(sBlo) + off, 128, 2048);
      if (three) { mma_ss(tb, dal, dbh, idesc, acc); mma_ss(tb, dah, dbl, idesc, 1u); acc = 1u; }
      mma_ss(tb, dah, dbh, idesc, acc);
      acc = 1u;
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_addr(tb, 32 * warp, c0), v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) D[tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------------- test 5
// Row-contraction (weight-gradient shaped) GEMM with K-MAJOR operands only: D[m][n] = sum_r X[r][m] * Y[r][n].
// The row-owning threads write X^T and Y^T themselves: element (feature f, row r) goes to slab r / 4 at
// (r / 4) * LBO + f * 16 + (r % 4) * 4 with LBO = F * 16 + 16 -- the 16 bytes of padding make the 32 lanes of a
// warp (32 consecutive rows, same f) hit 32 different banks, so the transposing 4-byte stores are conflict-free.
__global__ void __launch_bounds__(128) probe_gemm_rows(const float* X, const float* Y, float* D, int N, int three) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t lbo_a = 128 * 16 + 16, lbo_b = N * 16 + 16;
  uint8_t* sAhi = smem;
  uint8_t* sAlo = sAhi + 32 * lbo_a;
  uint8_t* sBhi = sAlo + 32 * lbo_a;
  uint8_t* sBlo = sBhi + 32 * lbo_b;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tslot, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  const int r = tid;   // thread = row
  for (int f = 0; f < 128; ++f) {
    uint32_t hi, lo;
    const float x = X[r * 128 + f];
    split_hi_lo(x, hi, lo);
    const uint32_t off = (r >> 2) * lbo_a + f * 16 + (r & 3) * 4;
    *reinterpret_cast<uint32_t*>(sAhi + off) = three ? hi : __float_as_uint(x);
    *reinterpret_cast<uint32_t*>(sAlo + off) = lo;
  }
  for (int f = 0; f < N; ++f) {
    uint32_t hi, lo;
    const float y = Y[r * N + f];
    split_hi_lo(y, hi, lo);
    const uint32_t off = (r >> 2) * lbo_b + f * 16 + (r & 3) * 4;
    *reinterpret_cast<uint32_t*>(sBhi + off) = three ? hi : __float_as_uint(y);
    *reinterpret_cast<uint32_t*>(sBlo + off) = lo;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = idesc_tf32(128, N);
    uint32_t acc = 0;
    for (int k8 = 0; k8 < 16; ++k8) {   // 8 rows per MMA = two slabs
      const uint64_t dah = smem_desc(smem_u32(sAhi) + k8 * 2 * lbo_a, lbo_a, 128), dal = smem_desc(smem_u32(sAlo) + k8 * 2 * lbo_a, lbo_a, 128);
      const uint64_t dbh = smem_desc(smem_u32(sBhi) + k8 * 2 * lbo_b, lbo_b, 128), dbl = smem_desc(smem_u32(sBlo) + k8 * 2 * lbo_b, lbo_b, 128);
      if (three) { mma_ss(tb, dah, dbl, idesc, acc); mma_ss(tb, dal, dbh, idesc, 1u); acc = 1u; }
      mma_ss(tb, dah, dbh, idesc, acc);
      acc = 1u;
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_addr(tb, 32 * warp, c0), v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) D[tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------------- test 2
#ifndef PROBE_NS
#define PROBE_NS 2
#endif
constexpr int NS = PROBE_NS;
constexpr int HC = 128, ECC = 40, CD = 32;
// dynamic smem carve-up (bytes from the 1024-aligned base)
constexpr int OFF_RING = 0;
constexpr int OFF_EHI = OFF_RING + NS * UM_STAGE_BYTES;      // e' 40 -> 10 slabs
constexpr int OFF_ELO = OFF_EHI + 10 * UM_A_SLAB;
constexpr int OFF_CHI = OFF_ELO + 10 * UM_A_SLAB;            // c 32 -> 8 slabs
constexpr int OFF_CLO = OFF_CHI + 8 * UM_A_SLAB;
constexpr int OFF_OPS = OFF_CLO + 8 * UM_A_SLAB;
constexpr int OFF_PIPE = OFF_OPS + UM_MAX_OPS * (int)sizeof(UOp);
constexpr int SMEM_TOTAL = OFF_PIPE + 256;
constexpr int TM_ACC0 = 0, TM_ACC1 = 128, TM_AHI = 256, TM_ALO = 384;

__device__ __forceinline__ float softplus100(float x) {
  const float y = 100.f * x;
  return y > 20.f ? x : log1pf(expf(y)) * 0.01f;
}
__device__ __forceinline__ float softplus100_fast(float x) {
  float e, l;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 144.26950408889634f));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.f + e));
  return x > 0.2f ? x : l * 0.0069314718055994531f;
}

struct TrunkArgs {
  const float* E;       // [tiles][128][40]
  const float* C;       // [tiles][128][32]
  const float* bias;    // [5][128] layer biases, then [5][128] fc_c biases
  const float* wpk;     // packed UMMA weights
  UOp ops[32];          // by value: read through the constant bank
  int n_ops;
  float* S;             // [tiles][5][128][128] softplus outputs
  float* H;             // [tiles][5][128][128] layer outputs
  int ntiles;
  long long* cycles;    // per-CTA cycle counter (tile loop only)
  long long* trace;     // [2][64] timestamps of CTA 0 on its 3rd tile: issuer / epilogue thread 0
  int mode;             // bit 0: skip global stores, bit 1: MUFU softplus, bit 2: no activation math at all
};

constexpr int OFF_BIAS = SMEM_TOTAL;                      // [10][128] floats
constexpr int SMEM_ALL = OFF_BIAS + 10 * 128 * 4;

template <bool STORE>
__global__ void __launch_bounds__(320, 1) trunk_kernel(const __grid_constant__ TrunkArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  UPipe<NS>* pipe = reinterpret_cast<UPipe<NS>*>(smem + OFF_PIPE);
  float* sBias = reinterpret_cast<float*>(smem + OFF_BIAS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 10 * 128; i += blockDim.x) sBias[i] = a.bias[i];
  if (warp == 8) tmem_alloc(&pipe->tmem_base, 512);
  if (tid == 0) pipe_init<NS>(pipe, 8);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = pipe->tmem_base;
  const long long t0 = clock64();
  if (warp == 9) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x)
        producer_tile<NS>(a.ops, a.n_ops, a.wpk, smem + OFF_RING, pipe, it);
    }
  } else if (warp == 8) {
    uint32_t it = 0, a_par = 0;
    int nt = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++nt)
      issuer_tile<NS>(a.ops, a.n_ops, smem_u32(smem), smem_u32(smem + OFF_RING), pipe, tb, it, a_par,
                      (a.trace && blockIdx.x == 0 && nt == 2) ? a.trace : nullptr);
  } else {
    EpiSync es;
    const int row = 32 * (warp & 3) + lane, half = warp >> 2;
    const uint32_t lane_base = 32 * (warp & 3);
    int nt = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++nt) {
      long long* trc = (a.trace && blockIdx.x == 0 && nt == 2 && tid == 0) ? a.trace + 64 : nullptr;
      if (trc) *trc++ = clock64();
      // inputs -> shared-memory operands (e', c)
      const float* E = a.E + (size_t)tile * 128 * ECC;
      const float* C = a.C + (size_t)tile * 128 * CD;
      for (int it = tid; it < 128 * (ECC / 4); it += 256) {
        const int r = it % 128, q = it / 128;
        store_a_split(smem + OFF_EHI, smem + OFF_ELO, r, q * 4, *reinterpret_cast<const float4*>(E + r * ECC + q * 4));
      }
      for (int it = tid; it < 128 * (CD / 4); it += 256) {
        const int r = it % 128, q = it / 128;
        store_a_split(smem + OFF_CHI, smem + OFF_CLO, r, q * 4, *reinterpret_cast<const float4*>(C + r * CD + q * 4));
      }
      es.signal_a(pipe);
      if (trc) *trc++ = clock64();
      for (int li = 0; li < 5; ++li) {
        es.wait_d(pipe, 0);
        if (trc) *trc++ = clock64();
        const float4* b4 = reinterpret_cast<const float4*>(sBias + li * HC);
        const float4* u4 = reinterpret_cast<const float4*>(sBias + (5 + li) * HC);
        float* Sg = a.S + (((size_t)tile * 5 + li) * 128 + row) * HC;
        float* Hg = a.H + (((size_t)tile * 5 + li) * 128 + row) * HC;
        uint32_t v1[2][16], v2[2][16];
        tmem_ld16(tmem_addr(tb, lane_base, TM_ACC0 + 64 * half), v1[0]);
        tmem_ld16(tmem_addr(tb, lane_base, TM_ACC1 + 64 * half), v2[0]);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col0 = 64 * half + 16 * c;
          if (c < 3) {   // next 16 columns fly while these are processed
            tmem_ld16(tmem_addr(tb, lane_base, TM_ACC0 + col0 + 16), v1[(c + 1) & 1]);
            tmem_ld16(tmem_addr(tb, lane_base, TM_ACC1 + col0 + 16), v2[(c + 1) & 1]);
          }
          uint32_t (&x1)[16] = v1[c & 1];
          uint32_t (&x2)[16] = v2[c & 1];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 bb = b4[(col0 + j) >> 2], uu = u4[(col0 + j) >> 2];
            const float bv[4] = {bb.x, bb.y, bb.z, bb.w}, uv[4] = {uu.x, uu.y, uu.z, uu.w};
            float s[4], h[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              s[t] = softplus100_fast(__uint_as_float(x1[j + t]) + bv[t]);
              h[t] = s[t] + (__uint_as_float(x2[j + t]) + uv[t]);
              split_hi_lo(h[t], x1[j + t], x2[j + t]);
            }
            if (STORE) {
              *reinterpret_cast<float4*>(Sg + col0 + j) = make_float4(s[0], s[1], s[2], s[3]);
              *reinterpret_cast<float4*>(Hg + col0 + j) = make_float4(h[0], h[1], h[2], h[3]);
            }
          }
          tmem_st16(tmem_addr(tb, lane_base, TM_AHI + col0), x1);
          tmem_st16(tmem_addr(tb, lane_base, TM_ALO + col0), x2);
          if (c < 3) tmem_wait_ld();
        }
        if (trc) *trc++ = clock64();
        if (li < 4) es.signal_a(pipe);   // the last layer feeds no further GEMM in this probe
        else { tmem_wait_st(); }
        if (trc) *trc++ = clock64();
      }
    }
  }
  const long long t1 = clock64();
  tc_fence_before();
  __syncthreads();
  if (tid == 0 && a.cycles) a.cycles[blockIdx.x] = t1 - t0;
  if (warp == 8) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------------- host
static float frand() { return (float)rand() / (float)RAND_MAX * 2.f - 1.f; }

static double check(const char* name, const std::vector<float>& got, const std::vector<double>& ref) {
  double num = 0, den = 0, mx = 0;
  for (size_t i = 0; i < ref.size(); ++i) {
    const double d = (double)got[i] - ref[i];
    num += d * d; den += ref[i] * ref[i];
    if (fabs(d) > mx) mx = fabs(d);
  }
  const double rel = sqrt(num / (den + 1e-300));
  printf("  %-44s rel-L2 %.3e  max-abs %.3e\n", name, rel, mx);
  return rel;
}

static int test_gemm(int K, int N, int swap, int ts, int three, bool exact_inputs) {
  std::vector<float> A(128 * K), B(N * K), D(128 * N);
  for (auto& v : A) v = exact_inputs ? (float)((rand() % 17) - 8) * 0.125f : frand();
  for (auto& v : B) v = exact_inputs ? (float)((rand() % 17) - 8) * 0.25f : frand();
  std::vector<double> ref(128 * N);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * (double)B[n * K + k];
      ref[m * N + n] = s;
    }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, D.size() * 4));
  const int smem = (2 * 128 * K + 2 * N * K) * 4 + 1024;
  CK(cudaFuncSetAttribute(probe_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_gemm<<<1, 128, smem>>>(dA, dB, dD, K, N, swap, ts, three);
  cudaError_t e = cudaDeviceSynchronize();
  char name[128];
  snprintf(name, sizeof name, "gemm K=%d N=%d %s %s %s", K, N, ts ? "TS" : "SS", swap ? "LBO<->SBO swapped" : "LBO=K SBO=MN", three ? "3xTF32" : "1xTF32");
  if (e != cudaSuccess) { printf("  %-44s CUDA error: %s\n", name, cudaGetErrorString(e)); return -1; }
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  const double rel = check(name, D, ref);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return rel < (three || exact_inputs ? 2e-6 : 2e-3) ? 1 : 0;
}

static int test_gemm_mn(int N, int swap, int three, bool exact_inputs, int pattern = 0, int which = 3) {
  std::vector<float> A(128 * 128), B(128 * N), D(128 * N);
  for (auto& v : A) v = exact_inputs ? (float)((rand() % 17) - 8) * 0.125f : frand();
  for (auto& v : B) v = exact_inputs ? (float)((rand() % 17) - 8) * 0.25f : frand();
  if (pattern == 1) { for (auto& v : A) v = 1.f; for (auto& v : B) v = 1.f; }
  if (pattern == 2) {   // A[r][m] = 1 iff r == 3 and m == 5; B[r][n] = n + 1 iff r == 3  ->  D[5][n] = n + 1
    for (auto& v : A) v = 0.f; for (auto& v : B) v = 0.f;
    A[3 * 128 + 5] = 1.f;
    for (int n = 0; n < N; ++n) B[3 * N + n] = (float)(n + 1);
  }
  std::vector<double> ref(128 * N);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int r = 0; r < 128; ++r) s += (double)A[r * 128 + m] * (double)B[r * N + n];
      ref[m * N + n] = s;
    }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, D.size() * 4));
  const int smem = (2 * 128 * 128 + 2 * 128 * N) * 4 + 1024;
  CK(cudaFuncSetAttribute(probe_gemm_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_gemm_mn<<<1, 128, smem>>>(dA, dB, dD, N, swap, three, which);
  cudaError_t e = cudaDeviceSynchronize();
  char name[128];
  snprintf(name, sizeof name, "A^T.B (MN-major %s%s) N=%d %s %s", (which & 1) ? "A" : "", (which & 2) ? "B" : "", N, swap ? "LBO=MN SBO=K" : "LBO=K(128B) SBO=MN(2048B)", three ? "3xTF32" : "1xTF32");
  if (e != cudaSuccess) { printf("  %-44s CUDA error: %s\n", name, cudaGetErrorString(e)); return -1; }
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  const double rel = check(name, D, ref);
  if (pattern) {
    int nz = 0; double sum = 0;
    for (size_t i = 0; i < D.size(); ++i) { if (D[i] != 0.f) { if (nz < 12) printf("    D[%zu][%zu] = %g\n", i / N, i % N, D[i]); ++nz; } sum += D[i]; }
    printf("    pattern %d: %d non-zeros, sum %g\n", pattern, nz, sum);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return rel < (three || exact_inputs ? 2e-6 : 2e-3) ? 1 : 0;
}

static int test_gemm_rows(int N, int three, bool exact_inputs) {
  std::vector<float> A(128 * 128), B(128 * N), D(128 * N);
  for (auto& v : A) v = exact_inputs ? (float)((rand() % 17) - 8) * 0.125f : frand();
  for (auto& v : B) v = exact_inputs ? (float)((rand() % 17) - 8) * 0.25f : frand();
  std::vector<double> ref(128 * N);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int r = 0; r < 128; ++r) s += (double)A[r * 128 + m] * (double)B[r * N + n];
      ref[m * N + n] = s;
    }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, D.size() * 4));
  const int smem = 2 * 32 * (128 * 16 + 16) + 2 * 32 * (N * 16 + 16) + 1024;
  CK(cudaFuncSetAttribute(probe_gemm_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_gemm_rows<<<1, 128, smem>>>(dA, dB, dD, N, three);
  cudaError_t e = cudaDeviceSynchronize();
  char name[128];
  snprintf(name, sizeof name, "X^T.Y rows-as-K, K-major padded LBO N=%d %s", N, three ? "3xTF32" : "1xTF32");
  if (e != cudaSuccess) { printf("  %-44s CUDA error: %s\n", name, cudaGetErrorString(e)); return -1; }
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  const double rel = check(name, D, ref);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return rel < (three || exact_inputs ? 2e-6 : 2e-3) ? 1 : 0;
}

int main(int argc, char** argv) {
  srand(1219);
  int only = argc > 1 ? atoi(argv[1]) : 0;
  if (only == 0 || only == 1) {
    printf("test 1: single GEMMs (exact-in-tf32 inputs: any layout error shows as O(1) mismatch)\n");
    for (int ts = 0; ts < 2; ++ts)
      if (test_gemm(32, 128, 0, ts, 0, true) < 0) { printf("  (context lost, stopping)\n"); return 1; }
    printf("  3xTF32, random inputs:\n");
    test_gemm(40, 128, 0, 0, 1, false);
    test_gemm(64, 128, 0, 1, 1, false);
    test_gemm(128, 32, 0, 1, 1, false);
    test_gemm(128, 16, 0, 1, 1, false);
    test_gemm(56, 128, 0, 0, 1, false);
    printf("  1xTF32, random inputs (expected ~5e-4):\n");
    test_gemm(64, 128, 0, 0, 0, false);
  }
  if (only == 0 || only == 5) {
    printf("test 5: weight-gradient shaped GEMMs on the verified K-major path (transposed operands, padded LBO)\n");
    if (test_gemm_rows(64, 0, true) < 0) { printf("  (context lost, stopping)\n"); return 1; }
    test_gemm_rows(64, 1, false);
    test_gemm_rows(48, 1, false);
    test_gemm_rows(32, 1, false);
  }
  if (only == 4) {
    printf("test 4: the OTHER assignment of the descriptor LBO / SBO fields (how the convention was found; on B200 this\n"
           "        faults with an illegal memory access, i.e. LBO = K direction, SBO = row direction is the right one)\n");
    for (int ts = 0; ts < 2; ++ts)
      if (test_gemm(32, 128, 1, ts, 0, true) < 0) { printf("  (context lost, stopping)\n"); return 1; }
  }
  if (only == 3) {
    printf("test 3: weight-gradient shaped GEMMs, operands MN-major (contraction over the rows)\n");
    for (int swap = 0; swap < 2; ++swap)
      if (test_gemm_mn(64, swap, 0, true) < 0) { printf("  (context lost, stopping)\n"); return 1; }
    for (int which = 0; which < 4; ++which)
      for (int swap = 0; swap < 2; ++swap) test_gemm_mn(64, swap, 0, true, 1, which);
  }
  if (only == 0 || only == 2) {
    printf("test 2: streamed-weight engine on a colour-trunk-shaped MLP\n");
    // weights in a blob exactly like nn.Linear: W_i (128, K_i) row-major, U_i (128, 32)
    const int Kin[5] = {ECC, HC, HC, ECC + HC, HC};
    std::vector<float> blob;
    int wofs[5], uofs[5];
    for (int i = 0; i < 5; ++i) {
      wofs[i] = (int)blob.size();
      const float sc = 1.0f / sqrtf((float)Kin[i]);
      for (int e = 0; e < HC * Kin[i]; ++e) blob.push_back(frand() * sc * 1.7f);
    }
    for (int i = 0; i < 5; ++i) {
      uofs[i] = (int)blob.size();
      for (int e = 0; e < HC * CD; ++e) blob.push_back(frand() * 0.17f);
    }
    std::vector<float> bias(10 * HC);
    for (auto& v : bias) v = frand() * 0.05f;
    static UProgram P;
    UBuilder Bd(&P);
    int jw[5], ju[5], jw3e = -1;
    for (int i = 0; i < 5; ++i) {
      if (i == 3) {
        jw3e = Bd.weights(wofs[3], ECC + HC, 0, ECC, HC, HC);
        jw[3] = Bd.weights(wofs[3], ECC + HC, ECC, HC, HC, HC);
      } else {
        jw[i] = Bd.weights(wofs[i], Kin[i], 0, Kin[i], HC, HC);
      }
      ju[i] = Bd.weights(uofs[i], CD, 0, CD, HC, HC);
    }
    for (int i = 0; i < 5; ++i) {
      if (i == 0) Bd.gemm(jw[0], false, OFF_EHI, OFF_ELO, TM_ACC0, true, true, 0);
      else if (i == 3) {
        Bd.gemm(jw3e, false, OFF_EHI, OFF_ELO, TM_ACC0, true, true, 0);
        Bd.gemm(jw[3], true, TM_AHI, TM_ALO, TM_ACC0, false, false, 0);
      } else Bd.gemm(jw[i], true, TM_AHI, TM_ALO, TM_ACC0, true, true, 0);
      Bd.gemm(ju[i], false, OFF_CHI, OFF_CLO, TM_ACC1, true, false, 1);
    }
    printf("  program: %d ops, %d weight jobs, %d packed floats\n", P.n_ops, P.n_jobs, P.packed_floats);
    const int tiles_check = 3;
    const int tiles_time = 148 * 8;
    std::vector<float> E((size_t)tiles_time * 128 * ECC), C((size_t)tiles_time * 128 * CD);
    for (auto& v : E) v = frand();
    for (auto& v : C) v = frand() * 0.3f;
    float *dblob, *dwpk, *dE, *dC, *dbias, *dS, *dH;
    UPackJob* djobs; long long* dcyc;
    CK(cudaMalloc(&dblob, blob.size() * 4)); CK(cudaMemcpy(dblob, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dwpk, (size_t)P.packed_floats * 4));
    CK(cudaMalloc(&dE, E.size() * 4)); CK(cudaMemcpy(dE, E.data(), E.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dC, C.size() * 4)); CK(cudaMemcpy(dC, C.data(), C.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dbias, bias.size() * 4)); CK(cudaMemcpy(dbias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
    const size_t plane = (size_t)tiles_time * 5 * 128 * HC;
    CK(cudaMalloc(&dS, plane * 4)); CK(cudaMalloc(&dH, plane * 4));
    CK(cudaMalloc(&djobs, sizeof(UPackJob) * P.n_jobs)); CK(cudaMemcpy(djobs, P.jobs, sizeof(UPackJob) * P.n_jobs, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dcyc, 148 * 8));
    pack_umma_kernel<<<dim3(16, P.n_jobs), 256>>>(dblob, dwpk, djobs);
    CK(cudaDeviceSynchronize());
    TrunkArgs ta;
    ta.E = dE; ta.C = dC; ta.bias = dbias; ta.wpk = dwpk; memcpy(ta.ops, P.ops, sizeof(UOp) * P.n_ops); ta.n_ops = P.n_ops; ta.S = dS; ta.H = dH;
    ta.ntiles = tiles_check; ta.cycles = dcyc; ta.mode = 0; ta.trace = nullptr;
    long long* dtrace; CK(cudaMalloc(&dtrace, 128 * 8));
    CK(cudaFuncSetAttribute(trunk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALL + 1024));
    CK(cudaFuncSetAttribute(trunk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALL + 1024));
    trunk_kernel<true><<<2, 320, SMEM_ALL + 1024>>>(ta);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  trunk kernel: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    // fp64 reference of the same MLP
    std::vector<float> S((size_t)tiles_check * 5 * 128 * HC), H(S.size());
    CK(cudaMemcpy(S.data(), dS, S.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(H.data(), dH, H.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<double> Sr(S.size()), Hr(H.size());
    for (int t = 0; t < tiles_check; ++t)
      for (int r = 0; r < 128; ++r) {
        const float* e_ = &E[((size_t)t * 128 + r) * ECC];
        const float* c_ = &C[((size_t)t * 128 + r) * CD];
        std::vector<double> h(HC, 0.0), hn(HC);
        for (int li = 0; li < 5; ++li) {
          for (int n = 0; n < HC; ++n) {
            double z = bias[li * HC + n];
            const float* w = &blob[wofs[li] + (size_t)n * Kin[li]];
            if (li == 0) for (int k = 0; k < ECC; ++k) z += (double)w[k] * e_[k];
            else if (li == 3) {
              for (int k = 0; k < ECC; ++k) z += (double)w[k] * e_[k];
              for (int k = 0; k < HC; ++k) z += (double)w[ECC + k] * h[k];
            } else for (int k = 0; k < HC; ++k) z += (double)w[k] * h[k];
            const double y = 100.0 * z;
            const double s = y > 20.0 ? z : log1p(exp(y)) * 0.01;
            double uu = bias[(5 + li) * HC + n];
            const float* uw = &blob[uofs[li] + (size_t)n * CD];
            for (int k = 0; k < CD; ++k) uu += (double)uw[k] * c_[k];
            hn[n] = s + uu;
            Sr[(((size_t)t * 5 + li) * 128 + r) * HC + n] = s;
            Hr[(((size_t)t * 5 + li) * 128 + r) * HC + n] = hn[n];
          }
          h = hn;
        }
      }
    check("trunk softplus outputs (5 layers, 3 tiles)", S, Sr);
    check("trunk layer outputs h", H, Hr);
    for (int t = 0; t < tiles_check; ++t) {
      std::vector<float> g(H.begin() + (size_t)t * 5 * 128 * HC, H.begin() + (size_t)(t + 1) * 5 * 128 * HC);
      std::vector<double> rr(Hr.begin() + (size_t)t * 5 * 128 * HC, Hr.begin() + (size_t)(t + 1) * 5 * 128 * HC);
      char nm[64]; snprintf(nm, sizeof nm, "  tile %d h (all layers)", t);
      check(nm, g, rr);
    }
    // per-layer error of the last layer only
    {
      std::vector<float> g; std::vector<double> rr;
      for (int t = 0; t < tiles_check; ++t)
        for (size_t i = 0; i < (size_t)128 * HC; ++i) {
          g.push_back(H[((size_t)t * 5 + 4) * 128 * HC + i]);
          rr.push_back(Hr[((size_t)t * 5 + 4) * 128 * HC + i]);
        }
      check("  last layer h4 only", g, rr);
    }
    // timing: all SMs, 8 tiles per CTA
    ta.ntiles = tiles_time;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 4; ++rep) {
      ta.mode = rep / 2;
      CK(cudaMemset(dtrace, 0, 128 * 8));
      ta.trace = (rep & 1) ? dtrace : nullptr;
      cudaEventRecord(e0);
      if (ta.mode == 0) trunk_kernel<true><<<148, 320, SMEM_ALL + 1024>>>(ta);
      else trunk_kernel<false><<<148, 320, SMEM_ALL + 1024>>>(ta);
      cudaEventRecord(e1);
      e = cudaEventSynchronize(e1);
      if (e != cudaSuccess) { printf("  timing run: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long cyc[148];
      CK(cudaMemcpy(cyc, dcyc, sizeof cyc, cudaMemcpyDeviceToHost));
      long long mx = 0; for (int i = 0; i < 148; ++i) if (cyc[i] > mx) mx = cyc[i];
      if (rep & 1) {
        long long tr[128];
        CK(cudaMemcpy(tr, dtrace, sizeof tr, cudaMemcpyDeviceToHost));
        const long long base = tr[64];
        printf("    issuer  (wait_a done, commit_d) per layer:");
        for (int i = 0; i < 10 && tr[i]; ++i) printf("%s%lld", i % 2 == 0 ? " | " : " ", tr[i] - base);
        printf("\n    epilogue thread 0 (start, inputs signalled, then per layer: d_ready, before signal, after signal):\n     ");
        for (int i = 0; i < 40 && tr[64 + i]; ++i) printf(" %lld", tr[64 + i] - base);
        printf("\n");
      }
      const double rows = (double)tiles_time * 128;
      const double macs = rows * (40.0 * 128 + 3 * 128.0 * 128 + 168.0 * 128 + 5 * 32.0 * 128);
      printf("  mode %d: %d tiles on 148 CTAs: %.3f ms, %.0f cycles/tile, %.1f cycles/row-layer, %.1f algorithmic TFLOP/s (x3 executed)\n",
             ta.mode, tiles_time, ms, (double)mx / 8.0, (double)mx / 8.0 / 128 / 5, 2 * macs / (ms * 1e-3) / 1e12);
    }
  }
  printf("probe done\n");
  return 0;
}
