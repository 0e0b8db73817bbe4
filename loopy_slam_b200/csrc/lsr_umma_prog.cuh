// The "GEMM program" of one tile of the tcgen05 render kernels.
//
// Every dense contraction of a tile is D[128 x N] (+)= A[128 x K] . W[N x K]^T with fp32 operands
// evaluated as an error-compensated 3xTF32 product (A_lo.W_hi + A_hi.W_lo + A_hi.W_hi, fp32 accumulate
// in TMEM).  The weights are pre-split into (hi, lo) and pre-arranged in the canonical UMMA layout
// (lsr_umma.cuh) by pack_umma_kernel, cut into K-chunks of <= 32, so that a chunk is ONE contiguous
// block [hi | lo] in global memory = one cp.async.bulk into a ring stage.  The static list of chunks of
// a tile (UOp) is walked in lockstep by two single-thread roles:
//   producer : wait empty[stage] -> expect_tx -> bulk copy                     (weights L2 -> smem ring)
//   issuer   : (wait a_ready) -> wait full[stage] -> nk8 x 3 tcgen05.mma -> commit empty[stage]
//              (-> commit d_ready[i])
// while the epilogue warps follow the same order in code: wait d_ready[i] -> tcgen05.ld -> bias /
// activation -> write the next A operand (tcgen05.st to TMEM, or st.shared in UMMA layout) -> arrive a_ready.
#pragma once
#include "lsr_umma.cuh"

namespace lsr {
namespace umma {

constexpr int UM_M = 128;                  // rows (samples) per tile = TMEM lanes
constexpr int UM_KC = 32;                  // contraction values per streamed weight chunk of a 128-row matrix
constexpr int UM_STAGE_BYTES = 2 * 128 * UM_KC * 4;   // [hi | lo] of a 128 x 32 chunk = 32 KB
// Narrow matrices get proportionally deeper chunks (N = 32 -> K = 128): the same stage bytes, fewer L2 round trips.
__host__ __device__ constexpr int um_kc(int n) { return (UM_STAGE_BYTES / 8) / n < 128 ? (UM_STAGE_BYTES / 8) / n : 128; }
constexpr int UM_A_SLAB = UM_M * 16;       // bytes of one 4-wide K slab of an A operand in shared memory (= LBO of A)
constexpr int UM_MAX_OPS = 96;
constexpr int UM_MAX_JOBS = 40;

struct UOp {           // 32 bytes, device-ready (read through the constant bank: kernel parameter)
  uint32_t src;        // float offset of the chunk's hi block in the packed-UMMA weight buffer (lo block follows it)
  uint32_t half_bytes; // bytes of each block = n * kc * 4
  uint32_t a_hi, a_lo; // TS: TMEM column of the chunk's first k (hi / lo copies); SS: byte offset in dynamic smem
  uint32_t idesc;      // instruction descriptor (M = 128, N = n)
  uint32_t b_lbo_word; // (LBO >> 4) << 16 of the B descriptors, LBO = n * 16
  uint16_t b_k8_step;  // descriptor start-address increment per K = 8 step = (2 * LBO) >> 4
  uint16_t d_col;      // first accumulator column
  uint8_t nk8;         // K/8 steps in this chunk (1..16)
  uint8_t flags;       // UOP_*
  uint8_t commit_d;    // 0: none, 1 / 2: commit d_ready[0 / 1] after this chunk
  uint8_t pad;
};
static_assert(sizeof(UOp) == 32, "UOp");
enum { UOP_TS = 1, UOP_FIRST = 2, UOP_WAIT_A = 4 };

struct UPackJob {      // one GEMM's weights -> chunked UMMA layout
  int32_t src;         // blob offset of W (n_valid x ld row-major), first used column col0
  int32_t ld, col0;
  int32_t k_valid;     // contraction length (zero-padded to a multiple of 8 per chunk)
  int32_t n, n_valid;  // padded / real row count
  int32_t dst;         // float offset in the packed-UMMA buffer
  int32_t total;       // floats written (hi + lo, all chunks)
};

struct UProgram {
  UOp ops[UM_MAX_OPS];
  int n_ops;
  UPackJob jobs[UM_MAX_JOBS];
  int n_jobs;
  int packed_floats;
};

// ---- host-side builder ---------------------------------------------------------------------------
struct UBuilder {
  UProgram* P;
  explicit UBuilder(UProgram* p) : P(p) { p->n_ops = 0; p->n_jobs = 0; p->packed_floats = 0; }
  // Registers the weight matrix once; returns the job index (the same packed copy can feed several GEMMs).
  int weights(int src, int ld, int col0, int k_valid, int n, int n_valid) {
    UPackJob& j = P->jobs[P->n_jobs];
    j.src = src; j.ld = ld; j.col0 = col0; j.k_valid = k_valid; j.n = n; j.n_valid = n_valid;
    j.dst = P->packed_floats;
    int total = 0;
    const int KC = um_kc(n);
    for (int k0 = 0; k0 < k_valid; k0 += KC) {
      const int kc = (((k_valid - k0 < KC) ? (k_valid - k0) : KC) + 7) / 8 * 8;
      total += 2 * n * kc;
    }
    j.total = total;
    P->packed_floats += total;
    return P->n_jobs++;
  }
  // Appends the chunk ops of D[:, d_col : d_col + n] (+)= A . W^T for a registered weight job.
  // a_hi / a_lo: TMEM columns (ts) or dynamic-smem byte offsets (ss) of the first k.
  void gemm(int job, bool ts, uint32_t a_hi, uint32_t a_lo, int d_col, bool first, bool wait_a, int commit_d) {
    const UPackJob& j = P->jobs[job];
    int off = j.dst;
    int c = 0;
    const int KC = um_kc(j.n);
    const int nchunks = (j.k_valid + KC - 1) / KC;
    for (int k0 = 0; k0 < j.k_valid; k0 += KC, ++c) {
      const int kc = (((j.k_valid - k0 < KC) ? (j.k_valid - k0) : KC) + 7) / 8 * 8;
      UOp& o = P->ops[P->n_ops++];
      o.src = (uint32_t)off;
      o.half_bytes = (uint32_t)(j.n * kc * 4);
      o.a_hi = a_hi + (ts ? (uint32_t)k0 : (uint32_t)(k0 / 4) * UM_A_SLAB);
      o.a_lo = a_lo + (ts ? (uint32_t)k0 : (uint32_t)(k0 / 4) * UM_A_SLAB);
      o.idesc = idesc_tf32(UM_M, j.n);
      o.b_lbo_word = (uint32_t)(j.n * 16 >> 4) << 16;
      o.b_k8_step = (uint16_t)((2 * j.n * 16) >> 4);
      o.d_col = (uint16_t)d_col;
      o.nk8 = (uint8_t)(kc / 8);
      o.flags = (uint8_t)((ts ? UOP_TS : 0) | ((first && c == 0) ? UOP_FIRST : 0) | ((wait_a && c == 0) ? UOP_WAIT_A : 0));
      o.commit_d = (c == nchunks - 1) ? (uint8_t)commit_d : 0;
      o.pad = 0;
      off += 2 * j.n * kc;
    }
  }
};

#ifdef __CUDACC__
// ---- weight re-layout ----------------------------------------------------------------------------
// Element e of job jb -> its (hi, lo) words in the chunked layout.  Destination element order inside a chunk:
// [hi block | lo block], each slab-major: ((k/4) * n + row) * 4 + k % 4; all chunks but the last hold um_kc(n) k-values.
__device__ __forceinline__ void pack_umma_element(const UPackJob& jb, int e, const float* __restrict__ blob,
                                                  float* __restrict__ packed) {
  const int KC = um_kc(jb.n);
  const int full = jb.n * KC;
  const int c = e / full;
  const int k0 = c * KC;
  const int kc = (((jb.k_valid - k0 < KC) ? (jb.k_valid - k0) : KC) + 7) / 8 * 8;
  const int r = e - c * full;            // index inside the chunk's hi block: (slab, row, k%4)
  const int slab = r / (jb.n * 4), row = (r / 4) % jb.n, kq = r % 4;
  const int k = k0 + slab * 4 + kq;
  float v = 0.f;
  if (row < jb.n_valid && k < jb.k_valid) v = blob[jb.src + (size_t)row * jb.ld + jb.col0 + k];
  uint32_t hi, lo;
  split_hi_lo(v, hi, lo);
  float* chunk = packed + jb.dst + (size_t)c * 2 * full;
  chunk[r] = __uint_as_float(hi);
  chunk[jb.n * kc + r] = __uint_as_float(lo);
}
// grid (blocks, n_jobs): job list in global memory (bring-up probe; the library passes its jobs by value)
static __global__ void pack_umma_kernel(const float* __restrict__ blob, float* __restrict__ packed, const UPackJob* __restrict__ jobs) {
  const UPackJob jb = jobs[blockIdx.y];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < jb.total / 2; e += gridDim.x * blockDim.x)
    pack_umma_element(jb, e, blob, packed);
}

// ---- pipeline state (shared memory) ----------------------------------------------------------------
template <int NS>
struct UPipe {
  uint64_t full[NS], empty[NS];
  uint64_t a_ready, d_ready[2];
  uint32_t tmem_base;
  uint32_t pad;
};

template <int NS>
__device__ __forceinline__ void pipe_init(UPipe<NS>* p, uint32_t n_epilogue_warps) {
  for (int s = 0; s < NS; ++s) { mbar_init(&p->full[s], 1); mbar_init(&p->empty[s], 1); }
  mbar_init(&p->a_ready, n_epilogue_warps);
  mbar_init(&p->d_ready[0], 1);
  mbar_init(&p->d_ready[1], 1);
  fence_barrier_init();
}

// producer role (one thread): streams the chunks of one tile; `it` = running chunk counter of this CTA.
// hi block -> stage base, lo block -> stage base + UM_STAGE_BYTES / 2 (fixed, so the B descriptors of a stage
// do not depend on the chunk).
template <int NS>
__device__ __forceinline__ void producer_tile(const UOp* __restrict__ ops, int n_ops, const float* __restrict__ wpk,
                                              uint8_t* ring, UPipe<NS>* p, uint32_t& it) {
  for (int i = 0; i < n_ops; ++i, ++it) {
    const uint32_t stage = it % NS, use = it / NS;
    if (use > 0) mbar_wait(&p->empty[stage], (use - 1) & 1);
    const uint32_t hb = ops[i].half_bytes;
    const float* src = wpk + ops[i].src;
    uint8_t* dst = ring + stage * UM_STAGE_BYTES;
    mbar_arrive_expect_tx(&p->full[stage], 2 * hb);
    bulk_g2s(dst, src, hb, &p->full[stage]);
    bulk_g2s(dst + UM_STAGE_BYTES / 2, reinterpret_cast<const uint8_t*>(src) + hb, hb, &p->full[stage]);
  }
}

// three MMAs of one K = 8 step: D (+)= A_hi.B_lo ; D += A_lo.B_hi ; D += A_hi.B_hi  (small terms first; the two
// that read the B_hi tile back to back)
__device__ __forceinline__ void mma3_ts(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t bh_lo, uint32_t bl_lo,
                                        uint32_t b_hiword, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 bh, bl;\n\t"
      "setp.ne.b32 p, %7, 0;\n\t"
      "setp.eq.b32 q, 0, 0;\n\t"
      "mov.b64 bh, {%3, %5};\n\t"
      "mov.b64 bl, {%4, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bl, %6, p;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], bh, %6, q;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bh, %6, q;\n\t}\n" ::"r"(d),
      "r"(a_hi), "r"(a_lo), "r"(bh_lo), "r"(bl_lo), "r"(b_hiword), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma3_ss(uint32_t d, uint32_t ah_lo, uint32_t al_lo, uint32_t a_hiword, uint32_t bh_lo,
                                        uint32_t bl_lo, uint32_t b_hiword, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 ah, al, bh, bl;\n\t"
      "setp.ne.b32 p, %8, 0;\n\t"
      "setp.eq.b32 q, 0, 0;\n\t"
      "mov.b64 ah, {%1, %3};\n\t"
      "mov.b64 al, {%2, %3};\n\t"
      "mov.b64 bh, {%4, %6};\n\t"
      "mov.b64 bl, {%5, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bl, %7, p;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], al, bh, %7, q;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bh, %7, q;\n\t}\n" ::"r"(d),
      "r"(ah_lo), "r"(al_lo), "r"(a_hiword), "r"(bh_lo), "r"(bl_lo), "r"(b_hiword), "r"(idesc), "r"(acc)
      : "memory");
}

constexpr uint32_t UM_DESC_HIWORD = (128u >> 4) | (1u << 14);   // SBO = 128 B, descriptor version 1, SWIZZLE_NONE
constexpr uint32_t UM_A_LBO_WORD = (uint32_t)(UM_A_SLAB >> 4) << 16;
constexpr uint32_t UM_A_K8_STEP = (2 * UM_A_SLAB) >> 4;

// issuer role: executed by ONE WHOLE WARP (converged; all values warp-uniform and the op table read through
// the constant bank, so descriptor arithmetic stays in the uniform datapath); one elected lane issues.
template <int NS>
__device__ __forceinline__ void issuer_tile(const UOp* __restrict__ ops, int n_ops, uint32_t smem_base, uint32_t ring_addr,
                                            UPipe<NS>* p, uint32_t tmem_base, uint32_t& it, uint32_t& a_par,
                                            long long* trace = nullptr) {   // trace: bring-up probe only
  for (int i = 0; i < n_ops; ++i, ++it) {
    const uint32_t stage = it % NS, use = it / NS;
    const uint32_t flags = ops[i].flags;
    if (flags & UOP_WAIT_A) { mbar_wait(&p->a_ready, a_par); a_par ^= 1; if (trace) { if ((threadIdx.x & 31) == 0) *trace = clock64(); ++trace; } }
    mbar_wait(&p->full[stage], use & 1);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t lbo_word = ops[i].b_lbo_word, step = ops[i].b_k8_step, idesc = ops[i].idesc;
      uint32_t bh = lbo_word | (((ring_addr + stage * UM_STAGE_BYTES) >> 4) & 0x3fffu);
      uint32_t bl = lbo_word | (((ring_addr + stage * UM_STAGE_BYTES + UM_STAGE_BYTES / 2) >> 4) & 0x3fffu);
      const uint32_t d = tmem_base + ops[i].d_col;
      uint32_t acc = (flags & UOP_FIRST) ? 0u : 1u;
      const int nk8 = ops[i].nk8;
      if (flags & UOP_TS) {
        uint32_t ah = tmem_base + ops[i].a_hi, al = tmem_base + ops[i].a_lo;
#pragma unroll 4
        for (int k8 = 0; k8 < nk8; ++k8) {
          mma3_ts(d, ah, al, bh, bl, UM_DESC_HIWORD, idesc, acc);
          acc = 1u; ah += 8; al += 8; bh += step; bl += step;
        }
      } else {
        uint32_t ah = UM_A_LBO_WORD | (((smem_base + ops[i].a_hi) >> 4) & 0x3fffu);
        uint32_t al = UM_A_LBO_WORD | (((smem_base + ops[i].a_lo) >> 4) & 0x3fffu);
#pragma unroll 4
        for (int k8 = 0; k8 < nk8; ++k8) {
          mma3_ss(d, ah, al, UM_DESC_HIWORD, bh, bl, UM_DESC_HIWORD, idesc, acc);
          acc = 1u; ah += UM_A_K8_STEP; al += UM_A_K8_STEP; bh += step; bl += step;
        }
      }
      mma_commit(&p->empty[stage]);
      if (ops[i].commit_d) mma_commit(&p->d_ready[ops[i].commit_d - 1]);
    }
    __syncwarp();
    if (trace && ops[i].commit_d) { if ((threadIdx.x & 31) == 0) *trace = clock64(); ++trace; }
  }
}

// epilogue-side handshakes
struct EpiSync {
  uint32_t d_par[2];
  __device__ EpiSync() { d_par[0] = 0; d_par[1] = 0; }
  template <int NS>
  __device__ __forceinline__ void wait_d(UPipe<NS>* p, int which) {
    mbar_wait(&p->d_ready[which], d_par[which]);
    d_par[which] ^= 1;
    tc_fence_after();
  }
  // this thread's TMEM loads / stores and shared-memory operand writes are done -> the issuer may go on
  // (a_ready counts WARPS: every lane fences its own writes, the warp converges, one lane arrives -- 512 per-thread
  //  arrivals on one shared-memory word cost ~0.5 k cycles per hand-over)
  template <int NS>
  __device__ __forceinline__ void signal_a(UPipe<NS>* p) {
    tmem_wait_st();
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&p->a_ready);
  }
  // same, when the hand-over is TMEM only (no shared-memory operand was written since the last signal): skips
  // the generic->async proxy fence, which is a MEMBAR that also waits for this thread's global stores in flight
  template <int NS>
  __device__ __forceinline__ void signal_a_tmem(UPipe<NS>* p) {
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&p->a_ready);
  }
};

// float4 of 4 consecutive k (k % 4 == 0) of row r -> shared-memory A operand (hi and lo tiles)
__device__ __forceinline__ void store_a_split(uint8_t* a_hi, uint8_t* a_lo, int r, int k, float4 v) {
  uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
  split_hi_lo(v.x, h0, l0); split_hi_lo(v.y, h1, l1); split_hi_lo(v.z, h2, l2); split_hi_lo(v.w, h3, l3);
  const int off = (k >> 2) * UM_A_SLAB + r * 16;
  *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(h0, h1, h2, h3);
  *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(l0, l1, l2, l3);
}
// single element variant
__device__ __forceinline__ void store_a_split1(uint8_t* a_hi, uint8_t* a_lo, int r, int k, float v) {
  uint32_t h, l;
  split_hi_lo(v, h, l);
  const int off = (k >> 2) * UM_A_SLAB + r * 16 + (k & 3) * 4;
  *reinterpret_cast<uint32_t*>(a_hi + off) = h;
  *reinterpret_cast<uint32_t*>(a_lo + off) = l;
}
#endif  // __CUDACC__

}  // namespace umma
}  // namespace lsr
