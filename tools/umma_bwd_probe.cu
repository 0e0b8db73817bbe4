// DRAFT (compiles; NOT yet run on a GPU -- the round's GPU budget was spent when it was written).
// Bring-up probe for the round-2 tcgen05 BACKWARD of the colour trunk: one 128-row tile,
// L uniform layers  h_l = softplus100(W_l h_{l-1} + b_l) + (U_l c + u_l),  run backwards from G = dL/dh_{L-1}:
//
//   per layer l = L-1 .. 0                                           tensor-core work (all K-major, 3xTF32)
//     [dU_l | du_l] += G^T [c | 1]                                    rows contracted, operands transposed by their owners
//     dC            += G U_l                                          A = G in TMEM, B = U_l^T streamed
//     Z = G * (1 - exp(-100 s_l))                                     epilogue (s_l = saved softplus output)
//     [dW_l | db_l] += Z^T [h_{l-1} | 1]                              rows contracted
//     G <- Z W_l                                                      A = Z in TMEM, B = W_l^T streamed
//
// The bias gradients ride along as a ones column of the row-contracted B operand.  Row contraction = the 128 rows are
// the K dimension: both operands are written TRANSPOSED by the row-owning threads (element (feature f, row r) at
// (r/4) * LBO + f * 16 + (r % 4) * 4, LBO = F * 16 + 16: conflict-free, verified by umma_probe mode 5) and, because the
// pair would be 280 KB as hi + lo for 128 rows, in two 64-row halves that accumulate into the same TMEM columns.
//
// TMEM columns: [0,128) G / Z hi, [128,256) lo (row-major A operand); [256,288) dC; [288,472) the row-contracted
// results ([dU|du]: 48, [dW|db]: 144) and, after they have been flushed, G_prev (128).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/umma_bwd_probe tools/umma_bwd_probe.cu
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

constexpr int HC = 128, CD = 32, L = 3;
constexpr int NU = 48;                 // [c | 1 | pad]  (N must be a multiple of 16)
constexpr int NW = 144;                // [h | 1 | pad]
constexpr int NS = 2;
// TMEM
constexpr uint32_t T_AHI = 0, T_ALO = 128, T_DC = 256, T_R = 288;
// shared memory (bytes)
constexpr int LBO_AT = HC * 16 + 16;   // transposed G / Z: 128 features
constexpr int LBO_BU = NU * 16 + 16;
constexpr int LBO_BW = NW * 16 + 16;
constexpr int HALF_SLABS = 16;         // 64 rows = 16 slabs of 4 rows
constexpr int OFF_RING = 0;
constexpr int OFF_AT_HI = OFF_RING + NS * UM_STAGE_BYTES;
constexpr int OFF_AT_LO = OFF_AT_HI + HALF_SLABS * LBO_AT;
constexpr int OFF_BT_HI = OFF_AT_LO + HALF_SLABS * LBO_AT;
constexpr int OFF_BT_LO = OFF_BT_HI + HALF_SLABS * LBO_BW;
constexpr int OFF_PIPE = OFF_BT_LO + HALF_SLABS * LBO_BW;
constexpr int SMEM_TOTAL = OFF_PIPE + 256;
static_assert(SMEM_TOTAL <= 232448, "shared memory");

struct BwdArgs {
  const float* G0;      // [128][128] upstream gradient
  const float* S;       // [L][128][128] softplus outputs
  const float* Hp;      // [L][128][128] h_{l-1} (input of layer l)
  const float* C;       // [128][32]
  const float* wpk;     // packed TRANSPOSED weights: per layer [U_l^T (N=32,K=128) | W_l^T (N=128,K=128)] in chunk format
  uint32_t u_off[L], w_off[L];   // float offsets of the first chunk of U_l^T / W_l^T
  float *dW, *db, *dU, *du;      // [L][128][128], [L][128], [L][128][32], [L][128]  (atomically accumulated)
  float *dC, *Gout;              // [128][32], [128][128]
};

__device__ __forceinline__ float softplus100_grad_from_out(float sp) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(sp * -144.26950408889634f));
  return 1.f - e;
}

// one K = 8 step of a GEMM whose two operands sit in shared memory with their own LBO
__device__ __forceinline__ void mma3_local(uint32_t d, uint32_t ah, uint32_t al, uint32_t a_lbo, uint32_t bh, uint32_t bl,
                                           uint32_t b_lbo, uint32_t idesc, uint32_t acc) {
  const uint64_t dah = smem_desc(ah, a_lbo, 128), dal = smem_desc(al, a_lbo, 128);
  const uint64_t dbh = smem_desc(bh, b_lbo, 128), dbl = smem_desc(bl, b_lbo, 128);
  mma_ss(d, dah, dbl, idesc, acc);
  mma_ss(d, dal, dbh, idesc, 1u);
  mma_ss(d, dah, dbh, idesc, 1u);
}

// transposed store of one value of row r (0..63 inside the half) and feature f
__device__ __forceinline__ void store_t(uint8_t* hi, uint8_t* lo, uint32_t lbo, int r, int f, float v) {
  uint32_t h, l;
  split_hi_lo(v, h, l);
  const uint32_t off = (uint32_t)(r >> 2) * lbo + (uint32_t)f * 16u + (uint32_t)(r & 3) * 4u;
  *reinterpret_cast<uint32_t*>(hi + off) = h;
  *reinterpret_cast<uint32_t*>(lo + off) = l;
}

__device__ __forceinline__ void bar_compute256() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

__global__ void __launch_bounds__(320, 1) trunk_bwd_kernel(const __grid_constant__ BwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  UPipe<NS>* pipe = reinterpret_cast<UPipe<NS>*>(smem + OFF_PIPE);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 8) tmem_alloc(&pipe->tmem_base, 512);
  if (tid == 0) pipe_init<NS>(pipe, 8);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = pipe->tmem_base;
  const uint32_t sbase = smem_u32(smem);

  if (warp == 9) {
    // ---------------------------------------------------------------- producer: per layer U_l^T (1 chunk), W_l^T (4 chunks)
    if (lane == 0) {
      uint32_t it = 0;
      for (int l = L - 1; l >= 0; --l) {
        for (int c = 0; c < 5; ++c, ++it) {
          const uint32_t stage = it % NS, use = it / NS;
          if (use > 0) mbar_wait(&pipe->empty[stage], (use - 1) & 1);
          const uint32_t hb = c == 0 ? CD * HC * 4 : HC * UM_KC * 4;                       // bytes of the hi block
          const float* src = a.wpk + (c == 0 ? a.u_off[l] : a.w_off[l] + (c - 1) * 2 * HC * UM_KC);
          uint8_t* dst = smem + OFF_RING + stage * UM_STAGE_BYTES;
          mbar_arrive_expect_tx(&pipe->full[stage], 2 * hb);
          bulk_g2s(dst, src, hb, &pipe->full[stage]);
          bulk_g2s(dst + UM_STAGE_BYTES / 2, reinterpret_cast<const uint8_t*>(src) + hb, hb, &pipe->full[stage]);
        }
      }
    }
  } else if (warp == 8) {
    // ---------------------------------------------------------------- issuer (same order as the compute warps below)
    uint32_t it = 0, a_par = 0;
    auto wait_a = [&]() { mbar_wait(&pipe->a_ready, a_par); a_par ^= 1; tc_fence_after(); };
    auto ring_gemm = [&](uint32_t d_col, int n, int nk8, uint32_t a_col0, bool first) {   // A from TMEM, B = one ring chunk
      const uint32_t stage = it % NS, use = it / NS;
      mbar_wait(&pipe->full[stage], use & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t lbo_b = (uint32_t)n * 16u, idesc = idesc_tf32(128, n);
        uint32_t bh = ((uint32_t)(lbo_b >> 4) << 16) | (((sbase + OFF_RING + stage * UM_STAGE_BYTES) >> 4) & 0x3fffu);
        uint32_t bl = ((uint32_t)(lbo_b >> 4) << 16) | (((sbase + OFF_RING + stage * UM_STAGE_BYTES + UM_STAGE_BYTES / 2) >> 4) & 0x3fffu);
        uint32_t acc = first ? 0u : 1u;
        for (int k8 = 0; k8 < nk8; ++k8) {
          mma3_ts(tb + d_col, tb + T_AHI + a_col0 + k8 * 8, tb + T_ALO + a_col0 + k8 * 8, bh, bl, UM_DESC_HIWORD, idesc, acc);
          acc = 1u; bh += (2 * lbo_b) >> 4; bl += (2 * lbo_b) >> 4;
        }
        mma_commit(&pipe->empty[stage]);
      }
      __syncwarp();
      ++it;
    };
    auto local_gemm = [&](uint32_t d_col, int n, uint32_t b_lbo, bool first) {   // 64 rows of a row-contracted GEMM
      if (elect_one()) {
        const uint32_t idesc = idesc_tf32(128, n);
        uint32_t acc = first ? 0u : 1u;
        for (int k8 = 0; k8 < HALF_SLABS / 2; ++k8) {
          mma3_local(tb + d_col, sbase + OFF_AT_HI + k8 * 2 * LBO_AT, sbase + OFF_AT_LO + k8 * 2 * LBO_AT, LBO_AT,
                     sbase + OFF_BT_HI + k8 * 2 * b_lbo, sbase + OFF_BT_LO + k8 * 2 * b_lbo, b_lbo, idesc, acc);
          acc = 1u;
        }
      }
      __syncwarp();
    };
    auto commit_d = [&]() { if (elect_one()) mma_commit(&pipe->d_ready[0]); __syncwarp(); };
    for (int l = L - 1; l >= 0; --l) {
      for (int q = 0; q < 2; ++q) {            // [dU | du] += G^T [c | 1], 64 rows at a time
        wait_a();
        local_gemm(T_R, NU, LBO_BU, q == 0);
        if (q == 1) ring_gemm(T_DC, CD, HC / 8, 0, l == L - 1);   // dC += G U_l  (one chunk: N = 32, K = 128)
        commit_d();
      }
      for (int q = 0; q < 2; ++q) {            // [dW | db] += Z^T [h | 1]
        wait_a();
        local_gemm(T_R, NW, LBO_BW, q == 0);
        commit_d();
      }
      wait_a();                                // G_prev = Z W_l  (four chunks of K = 32)
      for (int c = 0; c < 4; ++c) ring_gemm(T_R, HC, UM_KC / 8, c * UM_KC, c == 0);
      commit_d();
    }
  } else {
    // ---------------------------------------------------------------- compute warps: thread = row (= TMEM lane)
    EpiSync es;
    const int row = 32 * (warp & 3) + lane, half = warp >> 2;
    const uint32_t lane_base = 32u * (warp & 3);
    const int qmine = (warp & 3) >> 1;          // which 64-row half this thread's row belongs to
    const int rloc = row & 63;
    // G0 -> TMEM A operand
    for (int c0 = 64 * half; c0 < 64 * half + 64; c0 += 16) {
      uint32_t hi[16], lo[16];
      for (int j = 0; j < 16; ++j) split_hi_lo(a.G0[row * HC + c0 + j], hi[j], lo[j]);
      tmem_st16(tmem_addr(tb, lane_base, T_AHI + c0), hi);
      tmem_st16(tmem_addr(tb, lane_base, T_ALO + c0), lo);
    }
    tmem_wait_st();
    for (int l = L - 1; l >= 0; --l) {
      // current G / Z of this thread's 64 columns, kept in registers across the layer
      float g[64];
      for (int c = 0; c < 4; ++c) {
        uint32_t hi[16], lo[16];
        tmem_ld16(tmem_addr(tb, lane_base, T_AHI + 64 * half + 16 * c), hi);
        tmem_ld16(tmem_addr(tb, lane_base, T_ALO + 64 * half + 16 * c), lo);
        tmem_wait_ld();
        for (int j = 0; j < 16; ++j) g[16 * c + j] = __uint_as_float(hi[j]) + __uint_as_float(lo[j]);   // exact
      }
      // ---- [dU | du]: two 64-row halves
      for (int q = 0; q < 2; ++q) {
        if (q == 1) es.wait_d(pipe, 0);         // the MMAs of half 0 have read AT / BT
        if (qmine == q) {
          for (int j = 0; j < 64; ++j) store_t(smem + OFF_AT_HI, smem + OFF_AT_LO, LBO_AT, rloc, 64 * half + j, g[j]);
          if (half == 0) {
            for (int f = 0; f < CD; ++f) store_t(smem + OFF_BT_HI, smem + OFF_BT_LO, LBO_BU, rloc, f, a.C[row * CD + f]);
            store_t(smem + OFF_BT_HI, smem + OFF_BT_LO, LBO_BU, rloc, CD, 1.f);
            for (int f = CD + 1; f < NU; ++f) store_t(smem + OFF_BT_HI, smem + OFF_BT_LO, LBO_BU, rloc, f, 0.f);
          }
        }
        es.signal_a(pipe);
      }
      es.wait_d(pipe, 0);
      // flush [dU | du]: lane = OUTPUT feature `row`, columns = c index; this half: 16 columns (+ du from column 32)
      {
        uint32_t x[16];
        tmem_ld16(tmem_addr(tb, lane_base, T_R + 16 * half), x);
        tmem_wait_ld();
        for (int j = 0; j < 16; ++j) atomicAdd(a.dU + ((size_t)l * HC + row) * CD + 16 * half + j, __uint_as_float(x[j]));
        if (half == 0) {
          tmem_ld16(tmem_addr(tb, lane_base, T_R + 32), x);
          tmem_wait_ld();
          atomicAdd(a.du + l * HC + row, __uint_as_float(x[0]));
        }
      }
      // ---- Z = G * softplus'(s), back into the A operand, and [dW | db] in two halves
      for (int j = 0; j < 64; ++j) g[j] *= softplus100_grad_from_out(a.S[((size_t)l * 128 + row) * HC + 64 * half + j]);
      for (int c = 0; c < 4; ++c) {
        uint32_t hi[16], lo[16];
        for (int j = 0; j < 16; ++j) split_hi_lo(g[16 * c + j], hi[j], lo[j]);
        tmem_st16(tmem_addr(tb, lane_base, T_AHI + 64 * half + 16 * c), hi);
        tmem_st16(tmem_addr(tb, lane_base, T_ALO + 64 * half + 16 * c), lo);
      }
      for (int q = 0; q < 2; ++q) {
        if (q == 1) es.wait_d(pipe, 0);
        if (qmine == q) {
          for (int j = 0; j < 64; ++j) store_t(smem + OFF_AT_HI, smem + OFF_AT_LO, LBO_AT, rloc, 64 * half + j, g[j]);
          const float* hp = a.Hp + ((size_t)l * 128 + row) * HC;
          for (int j = 0; j < 64; ++j) store_t(smem + OFF_BT_HI, smem + OFF_BT_LO, LBO_BW, rloc, 64 * half + j, hp[64 * half + j]);
          if (half == 0) {
            store_t(smem + OFF_BT_HI, smem + OFF_BT_LO, LBO_BW, rloc, HC, 1.f);
            for (int f = HC + 1; f < NW; ++f) store_t(smem + OFF_BT_HI, smem + OFF_BT_LO, LBO_BW, rloc, f, 0.f);
          }
        }
        es.signal_a(pipe);
      }
      es.wait_d(pipe, 0);
      // flush [dW | db]: lane = output feature, this half: 64 input-feature columns (+ db from column 128)
      for (int c = 0; c < 4; ++c) {
        uint32_t x[16];
        tmem_ld16(tmem_addr(tb, lane_base, T_R + 64 * half + 16 * c), x);
        tmem_wait_ld();
        for (int j = 0; j < 16; ++j)
          atomicAdd(a.dW + ((size_t)l * HC + row) * HC + 64 * half + 16 * c + j, __uint_as_float(x[j]));
      }
      if (half == 0) {
        uint32_t x[16];
        tmem_ld16(tmem_addr(tb, lane_base, T_R + HC), x);
        tmem_wait_ld();
        atomicAdd(a.db + l * HC + row, __uint_as_float(x[0]));
      }
      // ---- G_prev = Z W_l into the columns [dW | db] just left
      es.signal_a_tmem(pipe);
      es.wait_d(pipe, 0);
      for (int c = 0; c < 4; ++c) {
        uint32_t x[16], hi[16], lo[16];
        tmem_ld16(tmem_addr(tb, lane_base, T_R + 64 * half + 16 * c), x);
        tmem_wait_ld();
        for (int j = 0; j < 16; ++j) split_hi_lo(__uint_as_float(x[j]), hi[j], lo[j]);
        if (l == 0)
          for (int j = 0; j < 16; ++j) a.Gout[row * HC + 64 * half + 16 * c + j] = __uint_as_float(x[j]);
        tmem_st16(tmem_addr(tb, lane_base, T_AHI + 64 * half + 16 * c), hi);
        tmem_st16(tmem_addr(tb, lane_base, T_ALO + 64 * half + 16 * c), lo);
      }
      tmem_wait_st();
      tc_fence_before();
      bar_compute256();      // every thread's new G is in TMEM before anyone reloads its own columns (same thread: trivially
      tc_fence_after();      // ordered; kept as the place where a cross-thread consumer would hook in)
    }
    // dC
    {
      uint32_t x[16];
      tmem_ld16(tmem_addr(tb, lane_base, T_DC + 16 * half), x);
      tmem_wait_ld();
      for (int j = 0; j < 16; ++j) a.dC[row * CD + 16 * half + j] = __uint_as_float(x[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------------- host
static float frand() { return (float)rand() / (float)RAND_MAX * 2.f - 1.f; }

// B[n][k] (n_rows x k_cols, k contiguous per row in the SOURCE sense B[n][k] = src(n, k)) -> chunk format of
// lsr_umma_prog.cuh: chunks of um_kc(n) k-values, each [hi | lo], slab-major ((k/4) * n + row) * 4 + k % 4
template <class F>
static void pack_chunks(std::vector<float>& out, int n, int K, F src) {
  const int KC = um_kc(n);
  for (int k0 = 0; k0 < K; k0 += KC) {
    const int kc = std::min(KC, K - k0);
    const size_t base = out.size();
    out.resize(base + 2 * (size_t)n * kc);
    for (int k = 0; k < kc; ++k)
      for (int r = 0; r < n; ++r) {
        const float v = src(r, k0 + k);
        uint32_t u;
        memcpy(&u, &v, 4);
        uint32_t hi = u & 0xffffe000u;
        float hif;
        memcpy(&hif, &hi, 4);
        const float lof = v - hif;
        const size_t e = ((size_t)(k / 4) * n + r) * 4 + k % 4;
        out[base + e] = hif;
        out[base + (size_t)n * kc + e] = lof;
      }
  }
}

static double rel_l2(const std::vector<float>& got, const std::vector<double>& ref, const char* name) {
  double num = 0, den = 0;
  for (size_t i = 0; i < ref.size(); ++i) { const double d = got[i] - ref[i]; num += d * d; den += ref[i] * ref[i]; }
  const double r = sqrt(num / (den + 1e-300));
  printf("  %-28s rel-L2 %.3e\n", name, r);
  return r;
}

int main() {
  srand(1219);
  std::vector<float> W(L * HC * HC), U(L * HC * CD), G0(128 * HC), S(L * 128 * HC), Hp(L * 128 * HC), C(128 * CD);
  for (auto& v : W) v = frand() * 0.15f;
  for (auto& v : U) v = frand() * 0.17f;
  for (auto& v : G0) v = frand();
  for (auto& v : S) v = fabsf(frand()) * 0.05f;      // softplus outputs are >= 0; small values keep softplus' away from 1
  for (auto& v : Hp) v = frand();
  for (auto& v : C) v = frand() * 0.3f;
  // packed transposed weights
  std::vector<float> wpk;
  BwdArgs a;
  for (int l = 0; l < L; ++l) {
    a.u_off[l] = (uint32_t)wpk.size();
    pack_chunks(wpk, CD, HC, [&](int n, int k) { return U[((size_t)l * HC + k) * CD + n]; });      // U_l^T: [n = c idx][k = out]
    a.w_off[l] = (uint32_t)wpk.size();
    pack_chunks(wpk, HC, HC, [&](int n, int k) { return W[((size_t)l * HC + k) * HC + n]; });      // W_l^T: [n = in][k = out]
  }
  // fp64 reference
  std::vector<double> g(128 * HC), z(128 * HC), dWr(L * HC * HC, 0.0), dbr(L * HC, 0.0), dUr(L * HC * CD, 0.0), dur(L * HC, 0.0),
      dCr(128 * CD, 0.0), gout(128 * HC);
  for (size_t i = 0; i < g.size(); ++i) g[i] = G0[i];
  for (int l = L - 1; l >= 0; --l) {
    for (int r = 0; r < 128; ++r)
      for (int o = 0; o < HC; ++o) {
        const double gv = g[r * HC + o];
        dur[l * HC + o] += gv;
        for (int c = 0; c < CD; ++c) { dUr[((size_t)l * HC + o) * CD + c] += gv * C[r * CD + c]; dCr[r * CD + c] += gv * U[((size_t)l * HC + o) * CD + c]; }
        const double zv = gv * (1.0 - exp(-100.0 * (double)S[((size_t)l * 128 + r) * HC + o]));
        z[r * HC + o] = zv;
        dbr[l * HC + o] += zv;
        for (int i = 0; i < HC; ++i) dWr[((size_t)l * HC + o) * HC + i] += zv * Hp[((size_t)l * 128 + r) * HC + i];
      }
    for (int r = 0; r < 128; ++r)
      for (int i = 0; i < HC; ++i) {
        double s = 0;
        for (int o = 0; o < HC; ++o) s += z[r * HC + o] * W[((size_t)l * HC + o) * HC + i];
        gout[r * HC + i] = s;
      }
    g = gout;
  }
  // device
  float *dG0, *dS, *dHp, *dCc, *dwpk, *ddW, *ddb, *ddU, *ddu, *ddC, *dGout;
  auto up = [&](float** p, const std::vector<float>& h) { CK(cudaMalloc(p, h.size() * 4)); CK(cudaMemcpy(*p, h.data(), h.size() * 4, cudaMemcpyHostToDevice)); };
  auto zero = [&](float** p, size_t n) { CK(cudaMalloc(p, n * 4)); CK(cudaMemset(*p, 0, n * 4)); };
  up(&dG0, G0); up(&dS, S); up(&dHp, Hp); up(&dCc, C); up(&dwpk, wpk);
  zero(&ddW, dWr.size()); zero(&ddb, dbr.size()); zero(&ddU, dUr.size()); zero(&ddu, dur.size()); zero(&ddC, dCr.size()); zero(&dGout, gout.size());
  a.G0 = dG0; a.S = dS; a.Hp = dHp; a.C = dCc; a.wpk = dwpk; a.dW = ddW; a.db = ddb; a.dU = ddU; a.du = ddu; a.dC = ddC; a.Gout = dGout;
  CK(cudaFuncSetAttribute(trunk_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
  trunk_bwd_kernel<<<1, 320, SMEM_TOTAL>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("trunk_bwd_kernel: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  auto down = [&](float* p, size_t n) { std::vector<float> h(n); CK(cudaMemcpy(h.data(), p, n * 4, cudaMemcpyDeviceToHost)); return h; };
  printf("trunk backward, %d layers, one 128-row tile, vs fp64:\n", L);
  double worst = 0;
  worst = fmax(worst, rel_l2(down(ddW, dWr.size()), dWr, "dW"));
  worst = fmax(worst, rel_l2(down(ddb, dbr.size()), dbr, "db"));
  worst = fmax(worst, rel_l2(down(ddU, dUr.size()), dUr, "dU"));
  worst = fmax(worst, rel_l2(down(ddu, dur.size()), dur, "du"));
  worst = fmax(worst, rel_l2(down(ddC, dCr.size()), dCr, "dC"));
  worst = fmax(worst, rel_l2(down(dGout, gout.size()), gout, "G after the last layer"));
  printf("%s (worst %.3e)\n", worst < 2e-5 ? "OK" : "MISMATCH", worst);
  return worst < 2e-5 ? 0 : 1;
}
