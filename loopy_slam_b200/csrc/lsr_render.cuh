// Shared pieces of the fused render kernels: tile geometry, weight re-layout ("packed" scratch),
// saved-activation layout, and the block-cooperative FP32 tile GEMM.
#pragma once
#include "lsr_common.cuh"

namespace lsr {

#ifndef LSR_TILE_M
#define LSR_TILE_M 64
#endif
constexpr int TILE_M = LSR_TILE_M;   // sample rows per tile (64: two CTAs = 16 warps per SM; 128: one CTA)
constexpr int NT = 256;              // threads per CTA
constexpr int CTAS_PER_SM = TILE_M == 64 ? 2 : 1;
constexpr int RSW = NT / 16;         // row stride of the WIDE thread map (16 threads across the columns)
constexpr int RSN = NT / 8;          // row stride of the NARROW thread map (8 threads across the columns)
constexpr int TMA = TILE_M / RSW;    // rows per thread of a WIDE-mapped ACTIVATION tile (TILE_M rows)
constexpr int TMNA = TILE_M / RSN;   // same, NARROW map
constexpr int TMW = 128 / RSW;       // rows per thread of a WIDE-mapped 128-row weight-gradient tile
constexpr int TMN128 = 128 / RSN;    // same, NARROW map
constexpr int TMW32 = 32 / RSW;      // 32-row outputs, WIDE map
constexpr int TMN32 = 32 / RSN;      // 32-row outputs, NARROW map
constexpr int KC = 16;         // contraction rows per streamed chunk
constexpr int NSTAGE = 3;      // cp.async ring depth
constexpr int HG = 32;         // geometry decoder hidden width  (decoder.py:566)
constexpr int HC = 128;        // colour decoder hidden width    (decoder.py:561,569)
constexpr int EG = 93;         // geometry Fourier features      (decoder.py:151)
constexpr int EGP = 96;        //   padded to a multiple of 4
constexpr int EC = 20;         // colour Fourier mapping size -> 40 features (decoder.py:393)
constexpr int ECC = 40;
constexpr int ER = 10;         // rel-pos Fourier mapping size -> 20 features (decoder.py:402)
constexpr int QD = 52;         // rel-pos MLP input = 20 + 32    (decoder.py:310)
constexpr int QDP = 56;
constexpr float TWO_PI_F = 6.2831855f;   // float32(2*math.pi), decoder.py:38

// leading dimensions of the shared-memory activation tiles (all == 4 mod 32 or 12 mod 32 so that
// the two/four row groups of a warp land in different banks; all multiples of 4 for float4)
constexpr int XLD = 172;       // [e'(40) | h(128)] colour, [e(96) | h(32)] geometry, Q(56), u(128)
constexpr int CLD = 36;        // interpolated feature c (32)
constexpr int DLD = 132;       // backward: dA / dH tile (128)
constexpr int ELD = 44;        // backward: e' (40) and dE'
constexpr int QLD = 60;        // backward: Q_k (56)
constexpr int GLD = 100;       // backward: geometry e (96)

// ------------------------------------------------------------------ packed weights (scratch)
// Transposed ([in][out], "contraction-major") copies for the forward GEMMs and zero-padded copies of
// the two geometry matrices whose input width is not a multiple of 4.  Offsets in floats.
struct Packed {
  static constexpr int gB = 0;                       // [3][96]
  static constexpr int gW0t = gB + 3 * EGP;          // [96][32]
  static constexpr int gW1t = gW0t + EGP * HG;       // [32][32]
  static constexpr int gW2t = gW1t + HG * HG;
  static constexpr int gW3t = gW2t + HG * HG;        // [128][32]  rows 0..92 emb, 93..95 zero, 96..127 h
  static constexpr int gW4t = gW3t + 128 * HG;
  static constexpr int gUt = gW4t + HG * HG;         // 5 x [32][32]
  static constexpr int gW0n = gUt + 5 * CDIM * HG;   // [32][96]
  static constexpr int gW3n = gW0n + HG * EGP;       // [32][128]
  static constexpr int cW0t = gW3n + HG * 128;       // [40][128]
  static constexpr int cW1t = cW0t + ECC * HC;       // [128][128]
  static constexpr int cW2t = cW1t + HC * HC;
  static constexpr int cW3t = cW2t + HC * HC;        // [168][128]
  static constexpr int cW4t = cW3t + (ECC + HC) * HC;
  static constexpr int cUt = cW4t + HC * HC;         // 5 x [32][128]
  static constexpr int V1t = cUt + 5 * CDIM * HC;    // [56][128]  rows 52..55 zero
  static constexpr int V2t = V1t + QDP * HC;         // [128][32]
  static constexpr int total = V2t + HC * CDIM;
};

struct PackJob {
  int src, n_out, n_in;      // source (out,in) row-major at blob + src
  int dst, dst_rows, dst_ld; // destination extent
  int gap_at, gap;           // input index k >= gap_at is shifted by +gap in the destination
  int transpose;             // 1: dst[k'][n]  0: dst[n][k']
};
constexpr int MAX_PACK_JOBS = 32;
struct PackJobs { PackJob j[MAX_PACK_JOBS]; int n; };

// ------------------------------------------------------------------ saved activations (global)
// Row p = ray*S + s.  Offsets in floats from the saved base.
struct SavedLayout {
  size_t idx, w, D, misc, cg, cc, gs, gh, occ, cs, ch, u, sp, rgbs, outraw, total;
  size_t P, Pp;
};
__host__ __device__ inline SavedLayout saved_layout(int64_t R, int S, int stage, int flags) {
  SavedLayout L;
  const size_t P = (size_t)R * S;
  const size_t Pp = align_up(P, 128) + 128;   // row pitch of every plane (independent of the tile height)
  L.P = P;
  L.Pp = Pp;
  size_t o = 0;
  L.idx = o;   o += Pp * KNN;
  L.w = o;     o += Pp * KNN;
  L.D = o;     o += Pp * KNN;
  L.misc = o;  o += Pp * 4;            // z, has, wsum, cnt
  L.cg = o;    o += Pp * CDIM;
  L.gs = o;    o += 5 * Pp * HG;
  L.gh = o;    o += 5 * Pp * HG;
  L.occ = o;   o += Pp;
  L.cc = o; L.cs = o; L.ch = o; L.u = o; L.sp = o; L.rgbs = o; L.outraw = o;
  if (stage == LSR_STAGE_COLOR) {
    L.cc = o;      o += Pp * CDIM;
    L.cs = o;      o += 5 * Pp * HC;
    L.ch = o;      o += 5 * Pp * HC;
    L.rgbs = o;    o += Pp * 4;
    L.outraw = o;  o += Pp * 4;
    if (flags & LSR_FLAG_REL_POS) {
      L.u = o;     o += Pp * HC;
      L.sp = o;    o += Pp * KNN * HC;
    }
  }
  L.total = o;
  return L;
}

#ifdef __CUDACC__
// ------------------------------------------------------------------ small math
__device__ __forceinline__ float softplus100(float x) {
  // torch.nn.Softplus(beta=100, threshold=20): x if 100x > 20 else log1p(exp(100x))/100
  const float y = 100.f * x;
#ifdef LSR_ACCURATE_MATH
  return y > 20.f ? x : log1pf(expf(y)) * 0.01f;
#else
  // MUFU path: |abs err| ~ 1e-8 on values that are added to O(0.1) activations
  return y > 20.f ? x : __logf(1.f + __expf(y)) * 0.01f;
#endif
}
// d softplus100 / dx = sigmoid(100x) = 1 - exp(-100*softplus100(x))
__device__ __forceinline__ float softplus100_grad_from_out(float sp) {
#ifdef LSR_ACCURATE_MATH
  return 1.f - expf(-100.f * sp);
#else
  return 1.f - __expf(-100.f * sp);
#endif
}
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }
// DRAM -> L2 prefetch of `nrows` rows of `row_floats` floats (one request per 128-byte line)
__device__ __forceinline__ void prefetch_rows_l2(const float* base, int nrows, int row_floats) {
  const int lines_per_row = row_floats / 32;
  for (int l = threadIdx.x; l < nrows * lines_per_row; l += NT)
    prefetch_l2(base + (size_t)(l / lines_per_row) * row_floats + (l % lines_per_row) * 32);
}

// ------------------------------------------------------------------ tile GEMM
// acc[i][j] += sum_{c < Kc} A(r_i, c) * B[c][col_j]
//   thread (tx = tid % TXN, ty = tid / TXN):  r_i = ty + (NT/TXN) * i,  col_j = (j/4)*TXN*4 + tx*4 + j%4
//   A in shared memory: A_ROWMAJOR ? A[r*lda + c] : A[c*lda + r]
//   B row-major [Kc][ldb] with the contraction index as the row:
//     B_SMEM  : read in place from shared memory
//     !B_SMEM : global memory, streamed through a 3-stage cp.async ring of KC-row chunks in sBuf
//   Columns >= ncols_valid (multiple of 4) contribute zeros.  All NT threads must call; ends with
//   __syncthreads() so callers may immediately overwrite A or reuse sBuf.
template <int TM, int TXN, int NCG, bool A_ROWMAJOR, bool B_SMEM>
__device__ __forceinline__ void tile_gemm(float (&acc)[TM][NCG * 4], const float* __restrict__ A, int lda,
                                          int Kc, const float* __restrict__ B, int ldb, int ncols_valid,
                                          float* sBuf) {
  constexpr int NCOLS = TXN * 4 * NCG;
  constexpr int RS = NT / TXN;
  const int tid = threadIdx.x;
  const int tx = tid % TXN, ty = tid / TXN;
  const float* Ap = A_ROWMAJOR ? (A + (size_t)ty * lda) : (A + ty);

  if constexpr (B_SMEM) {
    __syncthreads();   // A / B tiles written by the caller must be visible
    for (int k = 0; k < Kc; ++k) {
      float a[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = A_ROWMAJOR ? Ap[i * RS * lda + k] : Ap[k * lda + i * RS];
#pragma unroll
      for (int g = 0; g < NCG; ++g) {
        const int col = g * TXN * 4 + tx * 4;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < ncols_valid) b = *reinterpret_cast<const float4*>(B + (size_t)k * ldb + col);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          acc[i][g * 4 + 0] = fmaf(a[i], b.x, acc[i][g * 4 + 0]);
          acc[i][g * 4 + 1] = fmaf(a[i], b.y, acc[i][g * 4 + 1]);
          acc[i][g * 4 + 2] = fmaf(a[i], b.z, acc[i][g * 4 + 2]);
          acc[i][g * 4 + 3] = fmaf(a[i], b.w, acc[i][g * 4 + 3]);
        }
      }
    }
    __syncthreads();
    return;
  } else {
    constexpr int PIECES = KC * NCOLS / 4;            // float4 pieces per chunk
    constexpr int PPR = NCOLS / 4;                    // pieces per chunk row
    const int nchunks = (Kc + KC - 1) / KC;
    auto prefetch = [&](int chunk) {
      if (chunk < nchunks) {
        float* dst = sBuf + (chunk % NSTAGE) * (KC * NCOLS);
        const int k0 = chunk * KC;
        for (int p = tid; p < PIECES; p += NT) {
          const int row = p / PPR, c4 = p % PPR;
          float* d = dst + row * NCOLS + c4 * 4;
          if (k0 + row < Kc && c4 * 4 < ncols_valid) {
            cp_async16(d, B + (size_t)(k0 + row) * ldb + c4 * 4);
          } else {
            *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      cp_async_commit();
    };
    prefetch(0);
    prefetch(1);
    for (int c = 0; c < nchunks; ++c) {
      cp_async_wait<1>();
      __syncthreads();          // chunk c visible to all; everyone is done with chunk c-1's buffer
      prefetch(c + 2);
      const float* sb = sBuf + (c % NSTAGE) * (KC * NCOLS);
      const int k0 = c * KC;
      const int kmax = min(KC, Kc - k0);
      if (kmax == KC) {
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
          float a[TM];
#pragma unroll
          for (int i = 0; i < TM; ++i)
            a[i] = A_ROWMAJOR ? Ap[i * RS * lda + k0 + kk] : Ap[(k0 + kk) * lda + i * RS];
#pragma unroll
          for (int g = 0; g < NCG; ++g) {
            const float4 b = *reinterpret_cast<const float4*>(sb + kk * NCOLS + g * TXN * 4 + tx * 4);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
              acc[i][g * 4 + 0] = fmaf(a[i], b.x, acc[i][g * 4 + 0]);
              acc[i][g * 4 + 1] = fmaf(a[i], b.y, acc[i][g * 4 + 1]);
              acc[i][g * 4 + 2] = fmaf(a[i], b.z, acc[i][g * 4 + 2]);
              acc[i][g * 4 + 3] = fmaf(a[i], b.w, acc[i][g * 4 + 3]);
            }
          }
        }
      } else {
        for (int kk = 0; kk < kmax; ++kk) {
          float a[TM];
#pragma unroll
          for (int i = 0; i < TM; ++i)
            a[i] = A_ROWMAJOR ? Ap[i * RS * lda + k0 + kk] : Ap[(k0 + kk) * lda + i * RS];
#pragma unroll
          for (int g = 0; g < NCG; ++g) {
            const float4 b = *reinterpret_cast<const float4*>(sb + kk * NCOLS + g * TXN * 4 + tx * 4);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
              acc[i][g * 4 + 0] = fmaf(a[i], b.x, acc[i][g * 4 + 0]);
              acc[i][g * 4 + 1] = fmaf(a[i], b.y, acc[i][g * 4 + 1]);
              acc[i][g * 4 + 2] = fmaf(a[i], b.z, acc[i][g * 4 + 2]);
              acc[i][g * 4 + 3] = fmaf(a[i], b.w, acc[i][g * 4 + 3]);
            }
          }
        }
      }
    }
    cp_async_wait<0>();
    __syncthreads();
  }
}

template <int TM, int NC>
__device__ __forceinline__ void zero_acc(float (&acc)[TM][NC]) {
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[i][j] = 0.f;
}

// thread -> tile coordinates of the two mappings used everywhere
struct WideMap {    // 16 x (NT/16) threads, 8 rows x (4 + 4) cols per thread -> TILE_M x 128
  int tx, ty;
  __device__ WideMap() : tx(threadIdx.x & 15), ty(threadIdx.x >> 4) {}
  __device__ int row(int i) const { return ty + RSW * i; }
  __device__ int col(int g) const { return g * 64 + tx * 4; }   // first of 4 consecutive columns
};
struct NarrowMap {  // 8 x (NT/8) threads, 4 rows x 4 cols per thread -> TILE_M x 32
  int tx, ty;
  __device__ NarrowMap() : tx(threadIdx.x & 7), ty(threadIdx.x >> 3) {}
  __device__ int row(int i) const { return ty + RSN * i; }
  __device__ int col() const { return tx * 4; }
};
#endif  // __CUDACC__

}  // namespace lsr
