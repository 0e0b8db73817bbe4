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
#ifndef LSR_KC
#define LSR_KC 16
#endif
#ifndef LSR_NSTAGE
#define LSR_NSTAGE 3
#endif
constexpr int KC = LSR_KC;         // contraction rows per streamed chunk (multiple of 8)
constexpr int NSTAGE = LSR_NSTAGE; // cp.async ring depth (>= 2)
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
constexpr int GLD = 96;        // backward: geometry e (96), only staged when geometry weights train

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

__host__ __device__ inline size_t scratch_legacy_floats() { return (size_t)Packed::total; }

// ------------------------------------------------------------------ saved activations (global)
// Row p = ray*S + s.  Offsets in floats from the saved base.
//
// Two kinds of planes:
//  * row-major planes [Pp][width] (k-NN results, geometry activations, rel-pos activations);
//  * "T-planes" (colour trunk): tile-blocked and TRANSPOSED, tile = the forward's 128-row tile
//    (rows_per_tile = floor(128 / S) * S valid rows).  A T-plane of F features holds per tile
//    [4 atom columns q = row / 32][F features][32 rows] floats, the 16-byte chunks of each 128-byte feature
//    line XOR-swizzled with (f % 8): element (row, f) at  q * F * 32 + f * 32 + ((((row % 32) / 4) ^ (f % 8)) * 4) + row % 4.
//    This is byte-for-byte the UMMA SWIZZLE_128B K-major operand image with K = rows, so
//      - a row-owning thread (lane = row) reads / writes one feature of 32 rows as ONE coalesced 128-byte line,
//      - the backward bulk-copies an atom column (F * 128 bytes, contiguous) into shared memory and the tensor
//        core contracts over the ROWS directly (weight-gradient GEMMs), tools/umma_sw128_probe.cu test 1.
struct SavedLayout {
  size_t idx, w, D, misc, occ, rgbs, outraw, light_end, cg, gs, gh, cst, cc1t, ect, ut, spt, qt, total;
  size_t P, Pp;
  int ntiles, rays_per_tile;
};
constexpr int TP_ROWS = 128;          // rows of a T-plane tile
constexpr int TP_C1 = 40;             // features of the [c (32) | 1 | 0 x 7] plane
constexpr int TP_Q = 64;              // lines of the transposed rel-pos MLP input [sin 10 | cos 10 | feature 32 | 1 | 0 x 11]
__host__ __device__ inline size_t tplane_tile_floats(int F) { return (size_t)F * TP_ROWS; }
__host__ __device__ inline int tplane_off(int F, int row, int f) {
  const int q = row >> 5, j = row & 31;
  return q * (F * 32) + f * 32 + ((((j >> 2) ^ (f & 7)) << 2) | (j & 3));
}
__host__ __device__ inline SavedLayout saved_layout(int64_t R, int S, int stage, int flags) {
  SavedLayout L;
  const size_t P = (size_t)R * S;
  const size_t Pp = align_up(P, 128) + 128;   // row pitch of every row-major plane (independent of the tile height)
  L.P = P;
  L.Pp = Pp;
  L.rays_per_tile = TP_ROWS / S;
  L.ntiles = (int)((R + L.rays_per_tile - 1) / L.rays_per_tile);
  const size_t nt = (size_t)(L.ntiles > 0 ? L.ntiles : 1);
  size_t o = 0;
  L.idx = o;   o += Pp * KNN;
  L.w = o;     o += Pp * KNN;
  L.D = o;     o += Pp * KNN;
  L.misc = o;  o += Pp * 4;            // z, has, wsum, cnt
  L.occ = o;   o += Pp;
  L.rgbs = o; L.outraw = o;
  if (stage == LSR_STAGE_COLOR) {
    L.rgbs = o;    o += Pp * 4;
    L.outraw = o;  o += Pp * 4;
  }
  L.light_end = o;                     // LSR_FLAG_SAVE_LIGHT (forward-only decode): nothing behind this point is written
  L.cg = o;    o += Pp * CDIM;
  L.gs = o;    o += 5 * Pp * HG;
  L.gh = o;    o += 5 * Pp * HG;
  L.cst = o; L.cc1t = o; L.ect = o; L.ut = o; L.spt = o; L.qt = o;
  if (stage == LSR_STAGE_COLOR) {
    L.cst = o;     o += 5 * nt * tplane_tile_floats(HC);    // softplus outputs s_l          [layer][tile]
    // (the layer outputs h_l = s_l + U_l c + u_l are NOT saved: the backward needs them only inside Z^T h, which is
    //  Z^T s + (Z^T [c | 1]) [U | u]^T -- the second factor is accumulated anyway for the fc_c gradients)
    L.cc1t = o;    o += nt * tplane_tile_floats(TP_C1);     // [c | 1 | 0]
    L.ect = o;     o += nt * tplane_tile_floats(ECC);       // colour Fourier features e' = [sin | cos]
    if (flags & LSR_FLAG_REL_POS) {   // rel-pos neighbour MLP: u = sum_k w_k softplus(.), the 8 softplus outputs, the 8 inputs Q_k
      L.ut = o;    o += nt * tplane_tile_floats(HC);
      L.spt = o;   o += nt * KNN * tplane_tile_floats(HC);
      L.qt = o;    o += nt * KNN * tplane_tile_floats(TP_Q);
    }
  }
  L.total = o;
  return L;
}

// ------------------------------------------------------------------ scratch (global, caller-sized)
// [legacy packed weights | UMMA packed weights (forward) | k-NN results | backward: packed transposed weights,
//  per-call accumulators of the row-contracted extras, per-row hand-over planes]
constexpr int UMMA_PACKED_FLOATS_MAX = 2 * (93 * 32 + 3 * 32 * 32 + 128 * 32 + 5 * 32 * 32 +                  // geometry
                                            (40 + 128 + 128 + 168 + 128) * 128 + 5 * 32 * 128 + 56 * 128 +   // colour
                                            128 * 32 + 128 * 16) + 8192;                                     // V2, head, padding
constexpr int BWD_PACK_FLOATS_MAX = 4 * 128 * 128 + (32 + 80 + 32 + 32 + 48) * 128 + 128 * 32 + 64 * 128 + 128 + 4096;   // W^hT x 4, extras, V2^T, V1^T, P_out
constexpr int BWD_ACC_SLOTS = 80;                 // per colour layer: [e' (40) | c (32) | 1 | pad (7)] x 128 outputs
constexpr int BWD_ACC_FLOATS = 5 * BWD_ACC_SLOTS * 128 + 3 * TP_C1;   // + M_out [3][40]
constexpr int GEO_PACK_FLOATS = 14336 + 32;        // geometry backward: per-layer B blocks [W_l^h | P_l (| W^e)] + v = w_out U_4 (lsr_geo_bwd_umma.cu)
struct ScratchLayout {
  size_t legacy, umma, knn_idx, knn_rem, knn_w, knn_pos, knn_hw, fwd_end, bwd_pack, bwd_acc, bwd_dc, bwd_dp, bwd_dwh, bwd_dqt, bwd_gpack,
      bwd_gdc, bwd_gde, bwd_gdh, total;
};
__host__ __device__ inline ScratchLayout scratch_layout(int64_t n_rays, int S) {
  ScratchLayout L;
  const size_t Pp = align_up((size_t)n_rays * S, 128) + 128;
  size_t o = 0;
  L.legacy = o;   o = align_up(o + (size_t)Packed::total * sizeof(float), 256);
  L.umma = o;     o = align_up(o + (size_t)UMMA_PACKED_FLOATS_MAX * sizeof(float), 256);
  L.knn_idx = o;  o = align_up(o + Pp * KNN * 4, 256);
  L.knn_rem = o;  o = align_up(o + Pp * KNN * 4, 256);
  L.knn_w = o;    o = align_up(o + Pp * KNN * 4, 256);
  L.knn_pos = o;  o = align_up(o + Pp * 16, 256);
  L.knn_hw = o;   o = align_up(o + Pp * 8, 256);
  L.fwd_end = o + 256;                                      // LSR_FLAG_FWD_ONLY: nothing behind this point is needed
  L.bwd_pack = o; o = align_up(o + (size_t)BWD_PACK_FLOATS_MAX * sizeof(float), 256);
  L.bwd_acc = o;  o = align_up(o + (size_t)BWD_ACC_FLOATS * sizeof(float), 256);
  L.bwd_dc = o;   o = align_up(o + Pp * CDIM * 4, 256);     // dL/dc (colour feature) per sample row
  L.bwd_dp = o;   o = align_up(o + Pp * 16, 256);           // dL/dp contribution of the colour Fourier features
  L.bwd_dwh = o;  o = align_up(o + Pp * KNN * 4, 256);      // dL/d(normalised IDW weight) of the rel-pos path (tracker)
  {   // d[sin | cos] of the rel-pos Fourier features: [tile][8][20][128]
    const size_t rpt = (size_t)(128 / S), nt = ((size_t)n_rays + rpt - 1) / rpt;
    L.bwd_dqt = o;  o = align_up(o + (nt > 0 ? nt : 1) * KNN * 20 * 128 * 4, 256);
  }
  L.bwd_gpack = o; o = align_up(o + (size_t)GEO_PACK_FLOATS * sizeof(float), 256);
  L.bwd_gdc = o;   o = align_up(o + Pp * CDIM * 4, 256);    // geometry chain hand-over: dL/dc^g, dL/de (96), dL/dh_l (5 x 32)
  L.bwd_gde = o;   o = align_up(o + Pp * EGP * 4, 256);
  L.bwd_gdh = o;   o = align_up(o + 5 * Pp * HG * 4, 256);
  L.total = o + 256;
  return L;
}

#ifdef __CUDACC__
#ifdef LSR_PHASE_TIMING
// debug-only per-phase cycle accounting (tools/phase_timing.py); never compiled into the product build
extern __device__ unsigned long long lsr_phase_cycles[2][16];
#define LSR_PHASE_BEGIN() long long _pt_last = clock64();
#define LSR_PHASE(kernel, k)                                                         \
  do {                                                                               \
    __syncthreads();                                                                 \
    if (threadIdx.x == 0) {                                                          \
      const long long _n = clock64();                                                \
      atomicAdd(&lsr_phase_cycles[kernel][k], (unsigned long long)(_n - _pt_last));  \
      _pt_last = _n;                                                                 \
    }                                                                                \
  } while (0)
#else
#define LSR_PHASE_BEGIN()
#define LSR_PHASE(kernel, k)
#endif
// ------------------------------------------------------------------ small math
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float softplus100(float x) {
  // torch.nn.Softplus(beta=100, threshold=20): x if 100x > 20 else log1p(exp(100x))/100
#ifdef LSR_ACCURATE_MATH
  const float y = 100.f * x;
  return y > 20.f ? x : log1pf(expf(y)) * 0.01f;
#else
  // two MUFU ops: log(1 + e^(100x))/100 = lg2(1 + 2^(x * 100 log2 e)) * (ln 2 / 100); absolute error
  // ~1e-8 on values that are added to O(0.1) activations (the 1e-4 parity bar is relative to those)
  const float e = ex2_approx(x * 144.26950408889634f);
  const float sp = lg2_approx(1.f + e) * 0.0069314718055994531f;
  return x > 0.2f ? x : sp;
#endif
}
// d softplus100 / dx = sigmoid(100x) = 1 - exp(-100*softplus100(x))
__device__ __forceinline__ float softplus100_grad_from_out(float sp) {
#ifdef LSR_ACCURATE_MATH
  return 1.f - expf(-100.f * sp);
#else
  return 1.f - ex2_approx(sp * -144.26950408889634f);
#endif
}
// sin / cos of a Fourier-feature argument (|x| up to ~1e3 rad): two-term Cody-Waite reduction by 2*pi in fp32
// FMAs (exact for |k| < 2^12), then the MUFU units on the reduced argument in [-pi, pi] (max abs error
// 2^-21.4 there, CUDA math API) -- ~10 instructions instead of the ~70 of the full-range sinf / sincosf.
// LSR_ACCURATE_MATH falls back to the library functions.
__device__ __forceinline__ float reduce_2pi(float x) {
  const float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(-k, 6.2831854820251465f, x);          // float(2*pi)
  return fmaf(-k, -1.7484556000744083e-07f, r);         // 2*pi - float(2*pi)
}
__device__ __forceinline__ float sin_ff(float x) {
#ifdef LSR_ACCURATE_MATH
  return sinf(x);
#else
  return __sinf(reduce_2pi(x));
#endif
}
__device__ __forceinline__ float cos_ff(float x) {
#ifdef LSR_ACCURATE_MATH
  return cosf(x);
#else
  return __cosf(reduce_2pi(x));
#endif
}
__device__ __forceinline__ void sincos_ff(float x, float* s, float* c) {
#ifdef LSR_ACCURATE_MATH
  sincosf(x, s, c);
#else
  const float r = reduce_2pi(x);
  *s = __sinf(r);
  *c = __cosf(r);
#endif
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

// Compositing backward of ONE ray (Renderer.py:184-201, common.py:382-422): dL/d(occupancy logit) of its S samples -> dOcc[s],
// and (colour stage, dOut != nullptr) dL/d(rgb of sample s) -> dOut[4 s + {0,1,2}].  prow0 = global row of the ray's first sample.
__device__ __forceinline__ void composite_bwd_ray(const LsrParams& prm, bool color, const float* __restrict__ sv, const SavedLayout& SL,
                                                  size_t prow0, int ray, const float* __restrict__ gt_depth,
                                                  const float* __restrict__ g_depth, const float* __restrict__ g_var,
                                                  const float* __restrict__ g_rgb, float* dOcc, float* dOut) {
  const int S = prm.n_surface;
  const float g = gt_depth[ray];
  const float coef = prm.sigmoid_coef;
  const bool nz = g > 0.f;
  const float gD = (nz || (prm.flags & LSR_FLAG_SAMPLE_NEAR_PCL)) ? g_depth[ray] : 0.f;   // Renderer.py:197-198
  const float gV = g_var ? g_var[ray] : 0.f;
  float gC[3] = {0.f, 0.f, 0.f};
  if (color && g_rgb && (nz || !(prm.flags & LSR_FLAG_SKIP_ZERO_DEPTH))) {
    gC[0] = g_rgb[3 * ray + 0]; gC[1] = g_rgb[3 * ray + 1]; gC[2] = g_rgb[3 * ray + 2];
  }
  float al[8], Tv[8], wv[8], zv[8], rg[8][3];
  float T = 1.f, sw = 0.f, swz = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    if (s < S) {
      const float4 mi = reinterpret_cast<const float4*>(sv + SL.misc)[prow0 + s];   // (z, has, sum w, cnt)
      const float occ = mi.y > 0.5f ? sv[SL.occ + prow0 + s] : -100.f;
      const float alpha = sigmoidf_acc(coef * occ);
      al[s] = alpha; Tv[s] = T;
      const float w = alpha * T;
      wv[s] = w;
      T = T * ((1.f - alpha) + 1e-10f);
      zv[s] = mi.x;
      float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
      if (color) rs = reinterpret_cast<const float4*>(sv + SL.rgbs)[prow0 + s];
      rg[s][0] = rs.x; rg[s][1] = rs.y; rg[s][2] = rs.z;
      sw += w; swz += w * zv[s];
      c0 += w * rs.x; c1 += w * rs.y; c2 += w * rs.z;
    }
  }
  const float wsum = sw + 1e-10f;
  const float depth = swz / wsum;
  const float m0 = c0 / wsum, m1 = c1 / wsum, m2 = c2 / wsum;
  float dvar_ddepth = 0.f;
#pragma unroll
  for (int s = 0; s < 8; ++s)
    if (s < S) dvar_ddepth += -2.f * wv[s] * (zv[s] - depth);
  const float gDe = gD + gV * dvar_ddepth;
  float dwv[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    if (s < S) {
      const float dz = zv[s] - depth;
      dwv[s] = (gDe * dz + gC[0] * (rg[s][0] - m0) + gC[1] * (rg[s][1] - m1) + gC[2] * (rg[s][2] - m2)) / wsum + gV * dz * dz;
    }
  }
  float suffix = 0.f;
#pragma unroll
  for (int s = 7; s >= 0; --s) {
    if (s < S) {
      const float dalpha = Tv[s] * dwv[s] - suffix / ((1.f - al[s]) + 1e-10f);
      suffix += wv[s] * dwv[s];
      dOcc[s] = coef * al[s] * (1.f - al[s]) * dalpha;
      if (dOut) {
        const float f = wv[s] / wsum;
        dOut[4 * s + 0] = gC[0] * f; dOut[4 * s + 1] = gC[1] * f; dOut[4 * s + 2] = gC[2] * f;
      }
    }
  }
}

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
//
// Plain FP32 FFMA: the only callers left are the weight gradients of a TRAINABLE geometry decoder (render_bwd_kernel), a rare
// path; every other contraction of the library runs on tcgen05 (lsr_render_fwd.cu, lsr_render_bwd_umma.cu, lsr_geo_bwd_umma.cu).
// (Rounds 1-2 had an mma.sync 3xTF32 body here; it left with the last kernel that used it.)
constexpr int SB_LD_PAD = 8;                         // chunk / staging row pitch = cols + 8 (== 8 mod 32)
constexpr int SB_FLOATS = NSTAGE * KC * (128 + SB_LD_PAD);

// Feature rows may live in two places: the full (N,C) table, or -- for the rows the mapper is optimising --
// a compact (n_sel,C) leaf block addressed through row_remap[id] >= 0 (replaces the per-iteration
// table[indices] = leaf index_put of src/Mapper.py:581-582 and the gather in its backward).
__device__ __forceinline__ const float* feat_row(const float* __restrict__ table, const float* __restrict__ leaf,
                                                 const int32_t* __restrict__ remap, int idx) {
  if (remap != nullptr) {
    const int j = __ldg(remap + idx);
    if (j >= 0) return leaf + (size_t)j * CDIM;
  }
  return table + (size_t)idx * CDIM;
}
__device__ __forceinline__ const float* feat_row_cached(const float* __restrict__ table, const float* __restrict__ leaf,
                                                        int idx, int rem) {   // rem = row_remap[idx] looked up once
  return rem >= 0 ? leaf + (size_t)rem * CDIM : table + (size_t)idx * CDIM;
}
// gradient row of point idx, or nullptr when the row is not trainable (remap given and remap[idx] < 0)
__device__ __forceinline__ float* grad_row(float* __restrict__ d, const int32_t* __restrict__ remap, int idx) {
  if (remap != nullptr) {
    const int j = __ldg(remap + idx);
    return j >= 0 ? d + (size_t)j * CDIM : nullptr;
  }
  return d + (size_t)idx * CDIM;
}

template <int TM, int TXN, int NCG, bool A_ROWMAJOR, bool B_SMEM>
__device__ __forceinline__ void tile_gemm(float (&acc)[TM][NCG * 4], const float* __restrict__ A, int lda,
                                          int Kc, const float* __restrict__ B, int ldb, int ncols_valid,
                                          float* sBuf, int mvalid = 1 << 30) {
  (void)mvalid;
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
#pragma unroll
    for (int st = 0; st < NSTAGE - 1; ++st) prefetch(st);
    for (int c = 0; c < nchunks; ++c) {
      cp_async_wait<NSTAGE - 2>();
      __syncthreads();          // chunk c visible to all; everyone is done with chunk c-1's buffer
      prefetch(c + NSTAGE - 1);
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
