// Forward render path on sm_100a, two launches per ray batch:
//
//   sample_knn_kernel   z-sampling + exact grid k-NN + IDW weights, one warp per pair of sample rows at
//                       full occupancy (the walk is a chain of dependent global loads: it wants many warps,
//                       not the few a tensor-core tile kernel can keep resident);
//   render_fwd_kernel   one persistent CTA per SM, tiles of 128 sample rows (= floor(128/S) rays):
//                       IDW gather -> geometry MLP -> (rel-pos neighbour MLP) -> colour MLP -> colour head ->
//                       alpha compositing.  Every dense contraction runs on the 5th-generation tensor cores
//                       (tcgen05.mma kind::tf32, M = 128, accumulators in TMEM) as an error-compensated
//                       3xTF32 product; weights are streamed L2 -> shared memory by cp.async.bulk into an
//                       mbarrier ring, hidden activations never leave TMEM between layers (the epilogue
//                       warps read the accumulator with tcgen05.ld, apply bias / activation / skip, and write
//                       the next layer's A operand back with tcgen05.st).  See lsr_umma_prog.cuh for the roles.
//
// Reference semantics restated (math only):
//   Renderer.render_batch_ray   /root/reference/src/utils/Renderer.py:71-201
//   NICER / MLP_geometry / MLP_color   /root/reference/src/conv_onet/models/decoder.py:106-626
//   raw2outputs_nerf_color      /root/reference/src/common.py:382-422
#include <cstdlib>
#include <cstring>
#include "lsr_render.cuh"
#include "lsr_umma_prog.cuh"

namespace lsr {

using namespace umma;

#ifdef LSR_PHASE_TIMING
__device__ unsigned long long lsr_phase_cycles[2][16];
#endif

// linspace(start, end, S)[s] exactly as torch computes it in float32
__device__ __forceinline__ float linspace_f32(float start, float end, int S, int s) {
  if (S <= 1) return start;
  const float step = (end - start) / (float)(S - 1);
  return (s < S / 2) ? __fadd_rn(start, __fmul_rn(step, (float)s))
                     : __fsub_rn(end, __fmul_rn(step, (float)(S - 1 - s)));
}

// ------------------------------------------------------------------------------------------------
// scratch layout (bytes): [legacy packed weights (backward) | UMMA packed weights | k-NN results]
struct KnnScratch {   // per sample row p = ray * S + s, all planes of Pp rows
  int32_t* idx;       // [Pp][8] neighbour ids (-1: none)
  int32_t* rem;       // [Pp][8] leaf row of the neighbour (row_remap), -1 = read the table
  float* w;           // [Pp][8] normalised IDW weights
  float4* pos;        // [Pp]    (px, py, pz, z)
  float2* hw;         // [Pp]    (has_neighbors, sum of weights)
};
// (ScratchLayout / scratch_layout: lsr_render.cuh)

struct KnnArgs {
  LsrParams prm;
  const void* grid;
  const float *rays_o, *rays_d, *gt_depth;
  const double* r_query;
  const float* far_zero;
  const float* z_override;   // (R, S) z of zero-depth rays (sample_near_pcl), nullable
  int far_group;
  int R;
  const int32_t* remap;
  KnnScratch ks;
  float* saved;
  int stage;
};

// ------------------------------------------------------------------------------------------------ kernel 1
// one warp per sample row (LSR_KNN_NQ rows in lockstep; lane k < 8 ends up owning the k-th neighbour)
#ifndef LSR_KNN_NQ
#define LSR_KNN_NQ 1          // sample rows searched in lockstep by one warp (1 beats 2 and 4 at 32 warps/SM: 0.281 / 0.288 / 0.308 ms forward)
#endif
#ifndef LSR_KNN_BLOCKS
#define LSR_KNN_BLOCKS 4      // resident 256-thread blocks per SM the register budget is set for
#endif
__global__ void __launch_bounds__(256, LSR_KNN_BLOCKS) sample_knn_kernel(const __grid_constant__ KnnArgs a) {
  constexpr int NQ = LSR_KNN_NQ;
  const int S = a.prm.n_surface;
  const int P = a.R * S;
  const bool dynr = (a.prm.flags & LSR_FLAG_DYNAMIC_R) != 0;
  const bool save = a.saved != nullptr;
  const SavedLayout SL = saved_layout(a.R, S, a.stage, a.prm.flags);
  const GridHeader* gh_ = reinterpret_cast<const GridHeader*>(a.grid);
  const GridView gv = grid_view(a.grid, gh_->n_points, gh_->max_cells);
  __shared__ uint2 pend_s[8][NQ * KNN_PEND];
  const int lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pair * NQ < P; pair += warps_total) {
    const int m0 = pair * NQ;
    float px[NQ], py[NQ], pz[NQ], z[NQ], r2f[NQ], rr[NQ];
    double r2d[NQ];
    bool rowvalid[NQ];
    unsigned bD[NQ];
    int bI[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int m = m0 + q;
      rowvalid[q] = m < P;
      px[q] = py[q] = pz[q] = z[q] = r2f[q] = rr[q] = 0.f;
      r2d[q] = 0.0;
      if (rowvalid[q]) {
        const int ray = m / S, s = m - ray * S;
        const float g = a.gt_depth[ray];
        if (g > 0.f) {   // Renderer.py:140-150
          const float t = linspace_f32(0.f, 1.f, S, s);
          const float zn = __fmul_rn(a.prm.near_end_surface, g), zf = __fmul_rn(a.prm.far_end_surface, g);
          z[q] = __fadd_rn(__fmul_rn(zn, __fsub_rn(1.f, t)), __fmul_rn(zf, t));
        } else if (a.z_override) {   // Renderer.py:151-158: npc-guided z-range (sample_near_pcl)
          z[q] = a.z_override[m];
        } else {         // Renderer.py:162-163
          const float far = a.far_zero ? a.far_zero[ray / a.far_group] : a.prm.near_end;
          z[q] = linspace_f32(a.prm.near_end, far, S, s);
        }
        px[q] = __fadd_rn(a.rays_o[3 * ray + 0], __fmul_rn(a.rays_d[3 * ray + 0], z[q]));   // Renderer.py:167-168
        py[q] = __fadd_rn(a.rays_o[3 * ray + 1], __fmul_rn(a.rays_d[3 * ray + 1], z[q]));
        pz[q] = __fadd_rn(a.rays_o[3 * ray + 2], __fmul_rn(a.rays_d[3 * ray + 2], z[q]));
        const double r = dynr ? a.r_query[ray] : a.prm.radius_query;
        r2d[q] = r * r;
        r2f[q] = (float)r2d[q];
        rr[q] = (float)r * 1.00001f + 1e-7f;
      }
    }
    knn_warp_multi<NQ>(gv, px, py, pz, rr, rowvalid, dynr, r2f, r2d, bD, bI, pend_s[threadIdx.x >> 5]);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int m = m0 + q;
      if (!rowvalid[q]) continue;   // warp-uniform
      const bool vk = lane < KNN && bD[q] != KNN_INF;
      const float Dk = __uint_as_float(bD[q]);
      const bool strict = vk && (dynr ? ((double)Dk < r2d[q]) : (Dk < r2f[q]));   // neural_point.py:1701-1706
      const int ns = __popc(__ballot_sync(0xffffffffu, strict));
      const int cnt = __popc(__ballot_sync(0xffffffffu, vk));
      const float wraw = vk ? 1.0f / (Dk + 1e-10f) : 0.f;                        // decoder.py:210,217-220
      float wsum = 0.f;
#pragma unroll
      for (int k = 0; k < KNN; ++k) wsum += __shfl_sync(0xffffffffu, wraw, k);
      const float wn = wraw / fmaxf(wsum, 1e-12f);
      float wn_sum = 0.f;
#pragma unroll
      for (int k = 0; k < KNN; ++k) wn_sum += __shfl_sync(0xffffffffu, wn, k);
      const int has = (ns >= a.prm.min_nn_num) ? 1 : 0;                          // decoder.py:204
      if (lane < KNN) {
        a.ks.idx[(size_t)m * KNN + lane] = vk ? bI[q] : -1;
        a.ks.rem[(size_t)m * KNN + lane] = (vk && a.remap != nullptr) ? __ldg(a.remap + bI[q]) : -1;
        a.ks.w[(size_t)m * KNN + lane] = wn;
        if (save) {
          reinterpret_cast<int*>(a.saved + SL.idx)[(size_t)m * KNN + lane] = vk ? bI[q] : -1;
          a.saved[SL.w + (size_t)m * KNN + lane] = wn;
          a.saved[SL.D + (size_t)m * KNN + lane] = vk ? Dk : FLT_MAX;
        }
      }
      if (lane == 0) {
        a.ks.pos[m] = make_float4(px[q], py[q], pz[q], z[q]);
        a.ks.hw[m] = make_float2((float)has, wn_sum);
        if (save) reinterpret_cast<float4*>(a.saved + SL.misc)[m] = make_float4(z[q], (float)has, wn_sum, (float)cnt);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ kernel 2
constexpr int FNS = 2;                                   // weight ring stages
#ifndef LSR_FWD_CW
#define LSR_FWD_CW 16
#endif
constexpr int CW = LSR_FWD_CW;                           // compute / epilogue warps: 8 or 16 (4 lane quadrants x NCG column groups)
constexpr int NCG = CW / 4;                              // column groups of the 128-wide layers
constexpr int CPT = HC / NCG;                            // columns per thread in the 128-wide epilogues (64 / 32)
constexpr int STEPS = CPT / 16;                          // 16-column TMEM steps per thread
constexpr int TPR = CW / 4;                              // threads per row in the gather / operand-build phases
constexpr int F4T = 8 / TPR;                             // float4 of a 32-float feature row per thread
constexpr int NCT = CW * 32;                             // compute threads
constexpr int FT = NCT + 64;                             // + issuer warp + producer warp
static_assert(CW == 8 || CW == 16, "compute warps");
// dynamic shared memory carve-up (bytes from the 1024-aligned base)
constexpr int SM_RING = 0;
constexpr int SM_UNION = SM_RING + FNS * UM_STAGE_BYTES;
constexpr int SM_EG_HI = SM_UNION;                       // geometry Fourier features, 96 k = 24 slabs
constexpr int SM_EG_LO = SM_EG_HI + 24 * UM_A_SLAB;
constexpr int SM_EC_HI = SM_UNION;                       // colour Fourier features, 40 k = 10 slabs   } alive after the
constexpr int SM_EC_LO = SM_EC_HI + 10 * UM_A_SLAB;      //                                             } geometry MLP
constexpr int SM_Q_HI = SM_EC_LO + 10 * UM_A_SLAB;       // rel-pos MLP input Q_k, 56 k = 14 slabs      }
constexpr int SM_Q_LO = SM_Q_HI + 14 * UM_A_SLAB;
constexpr int SM_C_HI = SM_UNION + 48 * UM_A_SLAB;       // interpolated feature c, 32 k = 8 slabs
constexpr int SM_C_LO = SM_C_HI + 8 * UM_A_SLAB;
constexpr int SM_IDX = SM_C_LO + 8 * UM_A_SLAB;          // [128][8] i32
constexpr int SM_REM = SM_IDX + 128 * KNN * 4;
constexpr int SM_W = SM_REM + 128 * KNN * 4;
constexpr int SM_P = SM_W + 128 * KNN * 4;               // [128] float4 (px, py, pz, z)
constexpr int SM_OCC = SM_P + 128 * 16;
constexpr int SM_OCCP = SM_OCC + 128 * 4;                // [2][128] partial occupancy dot products
constexpr int SM_WSUM = SM_OCCP + 2 * 128 * 4;
constexpr int SM_HAS = SM_WSUM + 128 * 4;
constexpr int SM_RGB = SM_HAS + 128 * 4;                 // [128][4]
constexpr int SM_BIAS = SM_RGB + 128 * 16;               // bias table, see Bias
constexpr int SM_PIPE = SM_BIAS + 2200 * 4;
constexpr int FWD_SMEM_BYTES = SM_PIPE + 128;
static_assert(SM_Q_LO + 14 * UM_A_SLAB <= SM_C_HI && SM_EG_LO + 24 * UM_A_SLAB <= SM_C_HI, "union region");
static_assert(FWD_SMEM_BYTES <= 232448, "shared memory budget");
struct Bias {   // float offsets inside the shared bias table
  static constexpr int gb = 0, gu = 160, gow = 320, gob = 352;            // geometry: 5x32, 5x32, 32, 1
  static constexpr int cb = 356, cu = cb + 640, v1b = cu + 640, v2b = v1b + 128, cob = v2b + 32;   // colour
  static constexpr int gB = cob + 4;            // geometry Fourier matrix [3][96] (padded copy of embedder._B)
  static constexpr int cB = gB + 3 * EGP;       // colour Fourier matrix [3][20]
  static constexpr int rB = cB + 3 * EC;        // rel-pos Fourier matrix [3][10] (+2 pad)
  static constexpr int total = rB + 32;
};
static_assert(Bias::total <= 2200, "bias table");
// TMEM columns
constexpr uint32_t TM_ACC0 = 0, TM_ACC1 = 128, TM_AHI = 256, TM_ALO = 384;

struct FwdArgs {
  LsrParams prm;
  const float* cloud;
  const float* gt_depth;
  int R;
  const float *geo_feats, *col_feats;
  const float *geo_leaf, *col_leaf;
  LsrWeights w;
  const float* packed;     // legacy packed scratch (geometry Fourier matrix)
  const float* wpk;        // UMMA packed weights
  const float* affine;
  int stage;
  float *depth, *var, *rgb;
  uint8_t* valid;
  float* saved;
  KnnScratch ks;
  int rays_per_tile, ntiles;
  int n_ops;
  UOp ops[UM_MAX_OPS];
};

#ifdef LSR_PHASE_TIMING
#define FWD_PHASE(k)                                                                  \
  do {                                                                                \
    if (tid == 0) {                                                                   \
      const long long _n = clock64();                                                 \
      atomicAdd(&lsr_phase_cycles[0][k], (unsigned long long)(_n - _pt_last));        \
      _pt_last = _n;                                                                  \
    }                                                                                 \
  } while (0)
#else
#define FWD_PHASE(k)
#endif
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;\n" ::"n"(NCT) : "memory"); }

// Saved activations leave as T-planes (lsr_render.cuh: tplane_off): per 128-row tile [feature][row], so that the store of one
// feature by the 32 row-owner lanes of a warp is one full 128-byte line, and the backward can bulk-copy a plane as a
// ready-made K-major (K = rows) tensor-core operand.
__global__ void __launch_bounds__(FT, 1) render_fwd_kernel(const __grid_constant__ FwdArgs a) {
  // NB: no pointer re-alignment through integer casts here -- it makes the compiler lose the shared address
  // space and turns every LDS / STS below into a generic LD / ST (measured: 5x slower epilogues)
  extern __shared__ __align__(1024) uint8_t smem[];
  UPipe<FNS>* pipe = reinterpret_cast<UPipe<FNS>*>(smem + SM_PIPE);
  int* sIdx = reinterpret_cast<int*>(smem + SM_IDX);
  int* sRem = reinterpret_cast<int*>(smem + SM_REM);
  float* sW = reinterpret_cast<float*>(smem + SM_W);
  float4* sP = reinterpret_cast<float4*>(smem + SM_P);
  float* sOcc = reinterpret_cast<float*>(smem + SM_OCC);
  float* sOccP = reinterpret_cast<float*>(smem + SM_OCCP);
  float* sWsum = reinterpret_cast<float*>(smem + SM_WSUM);
  int* sHas = reinterpret_cast<int*>(smem + SM_HAS);
  float* sRgb = reinterpret_cast<float*>(smem + SM_RGB);
  float* sBias = reinterpret_cast<float*>(smem + SM_BIAS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.prm.n_surface;
  const float* __restrict__ blob = a.w.blob;
  const bool color = a.stage == LSR_STAGE_COLOR;
  const bool relpos = color && (a.prm.flags & LSR_FLAG_REL_POS) != 0;
  const bool save = a.saved != nullptr;
  const bool save_full = save && !(a.prm.flags & LSR_FLAG_SAVE_LIGHT);   // light: only k-NN results, occ, rgb (eval_points)
  const SavedLayout SL = saved_layout(a.R, S, a.stage, a.prm.flags);
  const size_t Pp = SL.Pp;

  // one-time setup: bias table, TMEM, barriers
  for (int i = tid; i < Bias::total; i += FT) {
    float v = 0.f;
    if (i < Bias::gu) v = blob[a.w.g_lin_b[i / 32] + i % 32];
    else if (i < Bias::gow) v = blob[a.w.g_fc_b[(i - Bias::gu) / 32] + i % 32];
    else if (i < Bias::gob) v = blob[a.w.g_out_w + (i - Bias::gow)];
    else if (i == Bias::gob) v = blob[a.w.g_out_b];
    else if (color && i >= Bias::cb && i < Bias::cu) v = blob[a.w.c_lin_b[(i - Bias::cb) / 128] + (i - Bias::cb) % 128];
    else if (color && i >= Bias::cu && i < Bias::v1b) v = blob[a.w.c_fc_b[(i - Bias::cu) / 128] + (i - Bias::cu) % 128];
    else if (relpos && i >= Bias::v1b && i < Bias::v2b) v = blob[a.w.c_nb1_b + (i - Bias::v1b)];
    else if (relpos && i >= Bias::v2b && i < Bias::cob) v = blob[a.w.c_nb2_b + (i - Bias::v2b)];
    else if (color && i >= Bias::cob && i < Bias::cob + 3) v = blob[a.w.c_out_b + (i - Bias::cob)];
    else if (i >= Bias::gB && i < Bias::cB) v = a.packed[Packed::gB + (i - Bias::gB)];
    else if (color && i >= Bias::cB && i < Bias::rB) v = blob[a.w.c_B + (i - Bias::cB)];
    else if (relpos && i >= Bias::rB && i < Bias::rB + 3 * ER) v = blob[a.w.c_Brel + (i - Bias::rB)];
    sBias[i] = v;
  }
  if (warp == CW) tmem_alloc(&pipe->tmem_base, 512);
  if (tid == 0) pipe_init<FNS>(pipe, CW);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = pipe->tmem_base;

  if (warp == CW + 1) {
    // ================================================================ producer: weight chunks -> ring
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x)
        producer_tile<FNS>(a.ops, a.n_ops, a.wpk, smem + SM_RING, pipe, it);
    }
  } else if (warp == CW) {
    // ================================================================ issuer: tcgen05.mma
    uint32_t it = 0, a_par = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x)
      issuer_tile<FNS>(a.ops, a.n_ops, smem_u32(smem), smem_u32(smem + SM_RING), pipe, tb, it, a_par);
  } else {
    // ================================================================ compute / epilogue warps
    EpiSync es;
    const int row = 32 * (warp & 3) + lane;       // sample row of the tile = TMEM lane
    const int cg = warp >> 2;                     // column group of this warp: columns [CPT * cg, +CPT) of a 128-wide layer
    const bool narrow = cg < 2;                   // the 32-wide layers (geometry, V2) are handled by column groups 0, 1 (16 each)
    const uint32_t lane_base = 32u * (warp & 3);
    const float4* sB4 = reinterpret_cast<const float4*>(sBias);

    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      const int r0 = tile * a.rays_per_tile;
      const int nr = min(a.rays_per_tile, a.R - r0);
      const int nrows = nr * S;
      const size_t p0 = (size_t)r0 * S;
      const bool rv = row < nrows;
      const size_t prow = p0 + row;
      LSR_PHASE_BEGIN();

      // ---------------------------------------------------------------- neighbour lists of the tile
      for (int i = tid; i < 128 * KNN; i += NCT) {
        const int m = i >> 3;
        const bool ok = m < nrows;
        sIdx[i] = ok ? a.ks.idx[p0 * KNN + i] : -1;
        sRem[i] = ok ? a.ks.rem[p0 * KNN + i] : -1;
        sW[i] = ok ? a.ks.w[p0 * KNN + i] : 0.f;
      }
      if (tid < 128) {
        const bool ok = tid < nrows;
        sP[tid] = ok ? a.ks.pos[p0 + tid] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float2 hw = ok ? a.ks.hw[p0 + tid] : make_float2(0.f, 0.f);
        sHas[tid] = hw.x > 0.5f ? 1 : 0;
        sWsum[tid] = hw.y;
      }
      bar_compute();
      FWD_PHASE(0);

      // ---------------------------------------------------------------- geometry feature (IDW gather) -> c
      {
        const int m = tid & 127, qg = tid >> 7;   // TPR threads per row, F4T float4 each
        float4 acc[F4T];
#pragma unroll
        for (int i = 0; i < F4T; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sHas[m]) {
#pragma unroll
          for (int k = 0; k < KNN; ++k) {
            const int idx = sIdx[m * KNN + k];
            if (idx >= 0) {
              const float w = sW[m * KNN + k];
              const float4* fr = reinterpret_cast<const float4*>(feat_row_cached(a.geo_feats, a.geo_leaf, idx, sRem[m * KNN + k])) + qg * F4T;
#pragma unroll
              for (int i = 0; i < F4T; ++i) {
                const float4 f = __ldg(fr + i);
                acc[i].x = fmaf(w, f.x, acc[i].x); acc[i].y = fmaf(w, f.y, acc[i].y);
                acc[i].z = fmaf(w, f.z, acc[i].z); acc[i].w = fmaf(w, f.w, acc[i].w);
              }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < F4T; ++i) {
          store_a_split(smem + SM_C_HI, smem + SM_C_LO, m, (qg * F4T + i) * 4, acc[i]);
          if (save_full && m < nrows) reinterpret_cast<float4*>(a.saved + SL.cg)[(p0 + m) * 8 + qg * F4T + i] = acc[i];
        }
      }
      // ---------------------------------------------------------------- geometry Fourier features -> e
      for (int it = tid; it < 128 * (EGP / 4); it += NCT) {
        const int m = it & 127, g = it >> 7;
        const float4 p = sP[m];
        const float t0 = TWO_PI_F * p.x, t1 = TWO_PI_F * p.y, t2 = TWO_PI_F * p.z;
        float v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int j = g * 4 + t;
          v[t] = 0.f;
          if (j < EG) {
            const float arg = fmaf(t2, sBias[Bias::gB + 2 * EGP + j], fmaf(t1, sBias[Bias::gB + EGP + j], t0 * sBias[Bias::gB + j]));
            v[t] = sin_ff(arg);
          }
        }
        store_a_split(smem + SM_EG_HI, smem + SM_EG_LO, m, g * 4, make_float4(v[0], v[1], v[2], v[3]));
      }
      es.signal_a(pipe);
      FWD_PHASE(1);

      // ---------------------------------------------------------------- geometry MLP (decoder.py:275-284)
      // h = relu(W x + b) + (U c + u); accumulators: W x in columns [0,32), U c in [32,64); this thread: 16 columns
      {
        float occ_part = 0.f;
#pragma unroll 1
        for (int li = 0; li < 5; ++li) {
          es.wait_d(pipe, 0);
          if (!narrow) {                       // warp-uniform: the other column groups only keep the handshake counts
            if (li < 4) es.signal_a_tmem(pipe);
            continue;
          }
          uint32_t x1[16], x2[16];
          tmem_ld16(tmem_addr(tb, lane_base, TM_ACC0 + 16 * cg), x1);
          tmem_ld16(tmem_addr(tb, lane_base, TM_ACC0 + 32 + 16 * cg), x2);
          tmem_wait_ld();
          float* gs = a.saved + SL.gs + ((size_t)li * Pp + prow) * HG + 16 * cg;
          float* gh = a.saved + SL.gh + ((size_t)li * Pp + prow) * HG + 16 * cg;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 bb = sB4[(Bias::gb + li * 32 + 16 * cg + j) >> 2];
            const float4 uu = sB4[(Bias::gu + li * 32 + 16 * cg + j) >> 2];
            const float bv[4] = {bb.x, bb.y, bb.z, bb.w}, uv[4] = {uu.x, uu.y, uu.z, uu.w};
            float s[4], h[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              s[t] = fmaxf(__uint_as_float(x1[j + t]) + bv[t], 0.f);
              h[t] = s[t] + (__uint_as_float(x2[j + t]) + uv[t]);
              split_hi_lo(h[t], x1[j + t], x2[j + t]);
            }
            if (save_full && rv) {
              *reinterpret_cast<float4*>(gs + j) = make_float4(s[0], s[1], s[2], s[3]);
              *reinterpret_cast<float4*>(gh + j) = make_float4(h[0], h[1], h[2], h[3]);
            }
            if (li == 4) {
              const float4 ow = sB4[(Bias::gow + 16 * cg + j) >> 2];
              occ_part = fmaf(h[0], ow.x, occ_part); occ_part = fmaf(h[1], ow.y, occ_part);
              occ_part = fmaf(h[2], ow.z, occ_part); occ_part = fmaf(h[3], ow.w, occ_part);
            }
          }
          if (li < 4) {
            tmem_st16(tmem_addr(tb, lane_base, TM_AHI + 16 * cg), x1);
            tmem_st16(tmem_addr(tb, lane_base, TM_ALO + 16 * cg), x2);
            es.signal_a_tmem(pipe);
          }
        }
        if (narrow) sOccP[cg * 128 + row] = occ_part;
        bar_compute();
        if (tid < 128) {   // occupancy logit (decoder.py:284)
          const float o = (sBias[Bias::gob] + sOccP[tid]) + sOccP[128 + tid];
          sOcc[tid] = o;
          if (save && tid < nrows) a.saved[SL.occ + p0 + tid] = o;
        }
      }

      FWD_PHASE(2);
      if (color) {
        // T-planes of this tile (lsr_render.cuh): read back by the tcgen05 backward as row-contraction operands
        float* tp_c1 = a.saved + SL.cc1t + (size_t)tile * tplane_tile_floats(TP_C1);
        float* tp_ec = a.saved + SL.ect + (size_t)tile * tplane_tile_floats(ECC);
        // -------------------------------------------------------------- colour feature
        if (relpos) {   // decoder.py:477-488: c = V2 . sum_k w_k softplus(V1 q_k + v1) + v2 * sum_k w_k
          // zero the K padding (columns 52..55) of Q once per tile
          if (tid < 128) store_a_split(smem + SM_Q_HI, smem + SM_Q_LO, tid, QD, make_float4(0.f, 0.f, 0.f, 0.f));
          auto build_q = [&](int k) {
            constexpr int HG2 = TPR / 2;                 // thread groups (of 128) per half of the work
            const int m = tid & 127, grp = tid >> 7;
            const int idx = sIdx[m * KNN + k];
            if (grp < HG2) {
              // relative-position Fourier features of row m: q[0..9] = sin, q[10..19] = cos; this thread: ER / HG2 frequencies
              constexpr int NJ = ER / HG2;
              float sn[NJ], cs[NJ];
#pragma unroll
              for (int j = 0; j < NJ; ++j) { sn[j] = 0.f; cs[j] = 0.f; }
              if (idx >= 0) {
                const float4 p = sP[m];
                const float t0 = TWO_PI_F * __fsub_rn(__ldg(a.cloud + 3 * (size_t)idx + 0), p.x);
                const float t1 = TWO_PI_F * __fsub_rn(__ldg(a.cloud + 3 * (size_t)idx + 1), p.y);
                const float t2 = TWO_PI_F * __fsub_rn(__ldg(a.cloud + 3 * (size_t)idx + 2), p.z);
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                  const int jj = grp * NJ + j;
                  const float arg = fmaf(t2, sBias[Bias::rB + 2 * ER + jj], fmaf(t1, sBias[Bias::rB + ER + jj], t0 * sBias[Bias::rB + jj]));
                  sincos_ff(arg, &sn[j], &cs[j]);
                }
              }
#pragma unroll
              for (int j = 0; j < NJ; ++j) {
                store_a_split1(smem + SM_Q_HI, smem + SM_Q_LO, m, grp * NJ + j, sn[j]);
                store_a_split1(smem + SM_Q_HI, smem + SM_Q_LO, m, ER + grp * NJ + j, cs[j]);
              }
              if (save_full) {   // Q_k^T for the backward's row-contracted dV1 GEMM (lane = row: one line per store)
                float* tq_ = a.saved + SL.qt + ((size_t)tile * KNN + k) * tplane_tile_floats(TP_Q);
                const bool ok = m < nrows;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                  tq_[tplane_off(TP_Q, m, grp * NJ + j)] = ok ? sn[j] : 0.f;
                  tq_[tplane_off(TP_Q, m, ER + grp * NJ + j)] = ok ? cs[j] : 0.f;
                }
                if (grp == 0) {   // ones line (bias gradient) + zero padding up to 64 lines
#pragma unroll
                  for (int f = QD; f < TP_Q; ++f) tq_[tplane_off(TP_Q, m, f)] = (ok && f == QD) ? 1.f : 0.f;
                }
              }
            } else {
              // neighbour feature row of row m: q[20..51]; this thread: 8 / HG2 float4
              constexpr int NF = 8 / HG2;
              const int f0 = (grp - HG2) * NF;
              float4 f[NF];
#pragma unroll
              for (int i = 0; i < NF; ++i) f[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (idx >= 0) {
                const float4* fr = reinterpret_cast<const float4*>(feat_row_cached(a.col_feats, a.col_leaf, idx, sRem[m * KNN + k])) + f0;
#pragma unroll
                for (int i = 0; i < NF; ++i) f[i] = __ldg(fr + i);
              }
#pragma unroll
              for (int i = 0; i < NF; ++i) store_a_split(smem + SM_Q_HI, smem + SM_Q_LO, m, 2 * ER + (f0 + i) * 4, f[i]);
              if (save_full) {
                float* tq_ = a.saved + SL.qt + ((size_t)tile * KNN + k) * tplane_tile_floats(TP_Q);
                const bool ok = m < nrows;
#pragma unroll
                for (int i = 0; i < NF; ++i) {
                  const float fv[4] = {f[i].x, f[i].y, f[i].z, f[i].w};
#pragma unroll
                  for (int t = 0; t < 4; ++t) tq_[tplane_off(TP_Q, m, 2 * ER + (f0 + i) * 4 + t)] = ok ? fv[t] : 0.f;
                }
              }
            }
          };
          float uf[CPT];
#pragma unroll
          for (int i = 0; i < CPT; ++i) uf[i] = 0.f;
          build_q(0);
          es.signal_a(pipe);
#pragma unroll 1
          for (int k = 0; k < KNN; ++k) {
            es.wait_d(pipe, k & 1);               // GEMM k done: Q is free, accumulator (k & 1) is ready
            if (k + 1 < KNN) { build_q(k + 1); es.signal_a(pipe); }   // GEMM k+1 runs under this epilogue
            const uint32_t accb = (k & 1) ? TM_ACC1 : TM_ACC0;
            const float wk = sW[row * KNN + k];
            float* tsp = a.saved + SL.spt + ((size_t)tile * KNN + k) * tplane_tile_floats(HC) + (row >> 5) * (HC * 32) + (row & 3);
            const int tchunk_r = (row & 31) >> 2;
            uint32_t x[2][16];
            tmem_ld16(tmem_addr(tb, lane_base, accb + CPT * cg), x[0]);
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < STEPS; ++c) {
              if (c + 1 < STEPS) tmem_ld16(tmem_addr(tb, lane_base, accb + CPT * cg + 16 * (c + 1)), x[(c + 1) & 1]);
              uint32_t (&xc)[16] = x[c & 1];
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 bb = sB4[(Bias::v1b + CPT * cg + 16 * c + j) >> 2];
                const float s0 = softplus100(__uint_as_float(xc[j + 0]) + bb.x), s1 = softplus100(__uint_as_float(xc[j + 1]) + bb.y);
                const float s2 = softplus100(__uint_as_float(xc[j + 2]) + bb.z), s3 = softplus100(__uint_as_float(xc[j + 3]) + bb.w);
                uf[16 * c + j + 0] = fmaf(wk, s0, uf[16 * c + j + 0]); uf[16 * c + j + 1] = fmaf(wk, s1, uf[16 * c + j + 1]);
                uf[16 * c + j + 2] = fmaf(wk, s2, uf[16 * c + j + 2]); uf[16 * c + j + 3] = fmaf(wk, s3, uf[16 * c + j + 3]);
                if (save_full) {
                  const float sv4[4] = {s0, s1, s2, s3};
#pragma unroll
                  for (int t = 0; t < 4; ++t)
                    tsp[(CPT * cg + 16 * c + j + t) * 32 + ((tchunk_r ^ ((j + t) & 7)) << 2)] = rv ? sv4[t] : 0.f;
                }
              }
              if (c + 1 < STEPS) tmem_wait_ld();
            }
          }
          // u -> TMEM A operand of the V2 GEMM
          {
            float* tu = a.saved + SL.ut + (size_t)tile * tplane_tile_floats(HC) + (row >> 5) * (HC * 32) + (row & 3);
            const int tchunk_r = (row & 31) >> 2;
#pragma unroll
            for (int c = 0; c < STEPS; ++c) {
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) split_hi_lo(uf[16 * c + j], hi[j], lo[j]);
              if (save_full) {
#pragma unroll
                for (int j = 0; j < 16; ++j) tu[(CPT * cg + 16 * c + j) * 32 + ((tchunk_r ^ (j & 7)) << 2)] = rv ? uf[16 * c + j] : 0.f;
              }
              tmem_st16(tmem_addr(tb, lane_base, TM_AHI + CPT * cg + 16 * c), hi);
              tmem_st16(tmem_addr(tb, lane_base, TM_ALO + CPT * cg + 16 * c), lo);
            }
          }
          es.signal_a_tmem(pipe);
          es.wait_d(pipe, 0);
          if (narrow) {
            uint32_t x[16];
            tmem_ld16(tmem_addr(tb, lane_base, TM_ACC0 + 16 * cg), x);
            tmem_wait_ld();
            const float ws = sWsum[row];
            const bool has = sHas[row] != 0;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 v2 = sB4[(Bias::v2b + 16 * cg + j) >> 2];
              float4 cc = make_float4(0.f, 0.f, 0.f, 0.f);
              if (has) cc = make_float4(fmaf(v2.x, ws, __uint_as_float(x[j])), fmaf(v2.y, ws, __uint_as_float(x[j + 1])),
                                        fmaf(v2.z, ws, __uint_as_float(x[j + 2])), fmaf(v2.w, ws, __uint_as_float(x[j + 3])));
              store_a_split(smem + SM_C_HI, smem + SM_C_LO, row, 16 * cg + j, cc);
              if (save_full) {   // [c | 1 | 0] T-plane: lane = row, one 128-byte line per feature
                const float cv[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) tp_c1[tplane_off(TP_C1, row, 16 * cg + j + t)] = rv ? cv[t] : 0.f;
              }
            }
          } else if (cg == 2 && save_full) {
#pragma unroll
            for (int f = CDIM; f < TP_C1; ++f) tp_c1[tplane_off(TP_C1, row, f)] = (rv && f == CDIM) ? 1.f : 0.f;
          }
        } else {        // decoder.py:476,487-488: plain IDW interpolation of the colour features
          const int m = tid & 127, qg = tid >> 7;
          float4 acc[F4T];
#pragma unroll
          for (int i = 0; i < F4T; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (sHas[m]) {
#pragma unroll
            for (int k = 0; k < KNN; ++k) {
              const int idx = sIdx[m * KNN + k];
              if (idx >= 0) {
                const float w = sW[m * KNN + k];
                const float4* fr = reinterpret_cast<const float4*>(feat_row_cached(a.col_feats, a.col_leaf, idx, sRem[m * KNN + k])) + qg * F4T;
#pragma unroll
                for (int i = 0; i < F4T; ++i) {
                  const float4 f = __ldg(fr + i);
                  acc[i].x = fmaf(w, f.x, acc[i].x); acc[i].y = fmaf(w, f.y, acc[i].y);
                  acc[i].z = fmaf(w, f.z, acc[i].z); acc[i].w = fmaf(w, f.w, acc[i].w);
                }
              }
            }
          }
#pragma unroll
          for (int i = 0; i < F4T; ++i) {
            store_a_split(smem + SM_C_HI, smem + SM_C_LO, m, (qg * F4T + i) * 4, acc[i]);
            if (save_full) {
              const float cv[4] = {acc[i].x, acc[i].y, acc[i].z, acc[i].w};
#pragma unroll
              for (int t = 0; t < 4; ++t) tp_c1[tplane_off(TP_C1, m, (qg * F4T + i) * 4 + t)] = m < nrows ? cv[t] : 0.f;
            }
          }
          if (save_full && qg == 0) {
#pragma unroll
            for (int f = CDIM; f < TP_C1; ++f) tp_c1[tplane_off(TP_C1, m, f)] = (m < nrows && f == CDIM) ? 1.f : 0.f;
          }
        }
        FWD_PHASE(3);
        // -------------------------------------------------------------- colour Fourier features e' = [sin | cos]
        for (int it = tid; it < 128 * (EC / 4); it += NCT) {
          const int m = it & 127, g = it >> 7;
          const float4 p = sP[m];
          const float t0 = TWO_PI_F * p.x, t1 = TWO_PI_F * p.y, t2 = TWO_PI_F * p.z;
          float sn[4], cs[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int j = g * 4 + t;
            const float arg = fmaf(t2, sBias[Bias::cB + 2 * EC + j], fmaf(t1, sBias[Bias::cB + EC + j], t0 * sBias[Bias::cB + j]));
            sincos_ff(arg, &sn[t], &cs[t]);
          }
          store_a_split(smem + SM_EC_HI, smem + SM_EC_LO, m, g * 4, make_float4(sn[0], sn[1], sn[2], sn[3]));
          store_a_split(smem + SM_EC_HI, smem + SM_EC_LO, m, EC + g * 4, make_float4(cs[0], cs[1], cs[2], cs[3]));
          if (save_full) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              tp_ec[tplane_off(ECC, m, g * 4 + t)] = m < nrows ? sn[t] : 0.f;
              tp_ec[tplane_off(ECC, m, EC + g * 4 + t)] = m < nrows ? cs[t] : 0.f;
            }
          }
        }
        es.signal_a(pipe);

        // -------------------------------------------------------------- colour trunk (decoder.py:515-533)
        // h = softplus(W x + b) + (U c + u); W x in ACC0, U c in ACC1; this thread: 64 columns in 4 steps of 16
#pragma unroll 1
        for (int li = 0; li < 5; ++li) {
          es.wait_d(pipe, 0);
          // s_l T-plane: this thread's row inside the tile's atom column, chunk swizzle folded per feature
          float* tps = a.saved + SL.cst + ((size_t)li * SL.ntiles + tile) * tplane_tile_floats(HC) + (row >> 5) * (HC * 32) + (row & 3);
          const int tchunk = (row & 31) >> 2;
          uint32_t v1[2][16], v2[2][16];
          tmem_ld16(tmem_addr(tb, lane_base, TM_ACC0 + CPT * cg), v1[0]);
          tmem_ld16(tmem_addr(tb, lane_base, TM_ACC1 + CPT * cg), v2[0]);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < STEPS; ++c) {
            const int col0 = CPT * cg + 16 * c;
            if (c + 1 < STEPS) {   // the next 16 columns fly while these are processed
              tmem_ld16(tmem_addr(tb, lane_base, TM_ACC0 + col0 + 16), v1[(c + 1) & 1]);
              tmem_ld16(tmem_addr(tb, lane_base, TM_ACC1 + col0 + 16), v2[(c + 1) & 1]);
            }
            uint32_t (&x1)[16] = v1[c & 1];
            uint32_t (&x2)[16] = v2[c & 1];
            float hk[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 bb = sB4[(Bias::cb + li * HC + col0 + j) >> 2], uu = sB4[(Bias::cu + li * HC + col0 + j) >> 2];
              const float bv[4] = {bb.x, bb.y, bb.z, bb.w}, uv[4] = {uu.x, uu.y, uu.z, uu.w};
              float s[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                s[t] = softplus100(__uint_as_float(x1[j + t]) + bv[t]);
                hk[j + t] = s[t] + (__uint_as_float(x2[j + t]) + uv[t]);
                split_hi_lo(hk[j + t], x1[j + t], x2[j + t]);
              }
              if (save_full) {   // lane = row: every store instruction of the warp fills one 128-byte line
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const int f = col0 + j + t;
                  const int off = f * 32 + ((tchunk ^ ((j + t) & 7)) << 2);
                  tps[off] = rv ? s[t] : 0.f;
                }
              }
            }
            tmem_st16(tmem_addr(tb, lane_base, TM_AHI + col0), x1);
            tmem_st16(tmem_addr(tb, lane_base, TM_ALO + col0), x2);
            if (c + 1 < STEPS) tmem_wait_ld();
          }
          es.signal_a_tmem(pipe);   // li == 4: feeds the colour head GEMM
        }
        FWD_PHASE(4);
        // -------------------------------------------------------------- colour head (decoder.py:533-546)
        es.wait_d(pipe, 0);
        if (cg == 0) {
          uint32_t x[16];
          tmem_ld16(tmem_addr(tb, lane_base, TM_ACC0), x);
          tmem_wait_ld();
          const float o0 = __uint_as_float(x[0]) + sBias[Bias::cob + 0], o1 = __uint_as_float(x[1]) + sBias[Bias::cob + 1],
                      o2 = __uint_as_float(x[2]) + sBias[Bias::cob + 2];
          float r0c = o0, r1c = o1, r2c = o2;
          if (a.prm.rgb_mode == LSR_RGB_SIGMOID) {
            r0c = sigmoidf_acc(o0); r1c = sigmoidf_acc(o1); r2c = sigmoidf_acc(o2);
          } else if (a.prm.rgb_mode == LSR_RGB_AFFINE_SIGMOID) {   // out @ A + t  (decoder.py:538-540)
            const float* Af = a.affine;
            r0c = sigmoidf_acc(fmaf(o2, Af[6], fmaf(o1, Af[3], o0 * Af[0])) + Af[9]);
            r1c = sigmoidf_acc(fmaf(o2, Af[7], fmaf(o1, Af[4], o0 * Af[1])) + Af[10]);
            r2c = sigmoidf_acc(fmaf(o2, Af[8], fmaf(o1, Af[5], o0 * Af[2])) + Af[11]);
          }
          sRgb[row * 4 + 0] = r0c; sRgb[row * 4 + 1] = r1c; sRgb[row * 4 + 2] = r2c;
          if (save && rv) {
            reinterpret_cast<float4*>(a.saved + SL.rgbs)[prow] = make_float4(r0c, r1c, r2c, 0.f);
            reinterpret_cast<float4*>(a.saved + SL.outraw)[prow] = make_float4(o0, o1, o2, 0.f);
          }
        }
      } else {
        if (tid < 128) { sRgb[tid * 4 + 0] = 0.f; sRgb[tid * 4 + 1] = 0.f; sRgb[tid * 4 + 2] = 0.f; }
      }
      tc_fence_before();
      bar_compute();
      FWD_PHASE(5);

      // ---------------------------------------------------------------- compositing (common.py:402-422)
      if (tid < nr) {
        const int ray = r0 + tid;
        const float g = a.gt_depth[ray];
        const float coef = a.prm.sigmoid_coef;
        float T = 1.f, sw = 0.f, swz = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
        float wv[8], zv[8];
        int nhas = 0;
        for (int s = 0; s < S; ++s) {
          const int m = tid * S + s;
          const int has = sHas[m];
          nhas += has;
          const float occ = has ? sOcc[m] : -100.f;            // Renderer.py:184-186
          const float alpha = sigmoidf_acc(coef * occ);
          const float w = alpha * T;
          T = T * ((1.f - alpha) + 1e-10f);
          const float z = sP[m].w;
          wv[s] = w; zv[s] = z;
          sw += w;
          swz += w * z;
          c0 += w * sRgb[m * 4 + 0]; c1 += w * sRgb[m * 4 + 1]; c2 += w * sRgb[m * 4 + 2];
        }
        const float wsum = sw + 1e-10f;
        float depth = swz / wsum;
        float var = 0.f;
        for (int s = 0; s < S; ++s) { const float t = zv[s] - depth; var += (wv[s] * t) * t; }
        float o0 = c0 / wsum, o1 = c1 / wsum, o2 = c2 / wsum;
        if (!(g > 0.f)) {                                      // Renderer.py:197-200
          if (!(a.prm.flags & LSR_FLAG_SAMPLE_NEAR_PCL)) depth = 0.f;
          if (a.prm.flags & LSR_FLAG_SKIP_ZERO_DEPTH) { o0 = 0.f; o1 = 0.f; o2 = 0.f; }
        }
        a.depth[ray] = depth;
        a.var[ray] = var;
        a.rgb[3 * ray + 0] = o0; a.rgb[3 * ray + 1] = o1; a.rgb[3 * ray + 2] = o2;
        a.valid[ray] = (nhas >= S / 2 + 1) ? 1 : 0;            // decoder.py:259-260
      }
      bar_compute();
      FWD_PHASE(6);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CW) tmem_dealloc(tb, 512);
}

// The GEMM program of one tile (must mirror the order of the epilogue code above) + the weight re-layout jobs.
static void build_program(const LsrWeights* w, int stage, int flags, UProgram* P) {
  UBuilder B(P);
  const bool color = stage == LSR_STAGE_COLOR;
  const bool relpos = color && (flags & LSR_FLAG_REL_POS);
  // ---- geometry (hidden 32): accumulators W x -> columns [0,32), U c -> [32,64)
  const int gW0 = B.weights(w->g_lin_w[0], EG, 0, EG, HG, HG);
  const int gW1 = B.weights(w->g_lin_w[1], HG, 0, HG, HG, HG);
  const int gW2 = B.weights(w->g_lin_w[2], HG, 0, HG, HG, HG);
  const int gW3e = B.weights(w->g_lin_w[3], EG + HG, 0, EG, HG, HG);
  const int gW3h = B.weights(w->g_lin_w[3], EG + HG, EG, HG, HG, HG);
  const int gW4 = B.weights(w->g_lin_w[4], HG, 0, HG, HG, HG);
  int gU[5];
  for (int i = 0; i < 5; ++i) gU[i] = B.weights(w->g_fc_w[i], CDIM, 0, CDIM, HG, HG);
  for (int li = 0; li < 5; ++li) {
    if (li == 0) B.gemm(gW0, false, SM_EG_HI, SM_EG_LO, TM_ACC0, true, true, 0);
    else if (li == 3) {
      B.gemm(gW3e, false, SM_EG_HI, SM_EG_LO, TM_ACC0, true, true, 0);
      B.gemm(gW3h, true, TM_AHI, TM_ALO, TM_ACC0, false, false, 0);
    } else B.gemm(li == 1 ? gW1 : li == 2 ? gW2 : gW4, true, TM_AHI, TM_ALO, TM_ACC0, true, true, 0);
    B.gemm(gU[li], false, SM_C_HI, SM_C_LO, TM_ACC0 + 32, true, false, 1);
  }
  if (!color) return;
  if (relpos) {
    const int V1 = B.weights(w->c_nb1_w, QD, 0, QD, HC, HC);
    const int V2 = B.weights(w->c_nb2_w, HC, 0, HC, CDIM, CDIM);
    for (int k = 0; k < KNN; ++k) B.gemm(V1, false, SM_Q_HI, SM_Q_LO, (k & 1) ? TM_ACC1 : TM_ACC0, true, true, 1 + (k & 1));
    B.gemm(V2, true, TM_AHI, TM_ALO, TM_ACC0, true, true, 1);
  }
  const int cW0 = B.weights(w->c_lin_w[0], ECC, 0, ECC, HC, HC);
  const int cW1 = B.weights(w->c_lin_w[1], HC, 0, HC, HC, HC);
  const int cW2 = B.weights(w->c_lin_w[2], HC, 0, HC, HC, HC);
  const int cW3e = B.weights(w->c_lin_w[3], ECC + HC, 0, ECC, HC, HC);
  const int cW3h = B.weights(w->c_lin_w[3], ECC + HC, ECC, HC, HC, HC);
  const int cW4 = B.weights(w->c_lin_w[4], HC, 0, HC, HC, HC);
  int cU[5];
  for (int i = 0; i < 5; ++i) cU[i] = B.weights(w->c_fc_w[i], CDIM, 0, CDIM, HC, HC);
  const int cOut = B.weights(w->c_out_w, HC, 0, HC, 16, 3);
  for (int li = 0; li < 5; ++li) {
    if (li == 0) B.gemm(cW0, false, SM_EC_HI, SM_EC_LO, TM_ACC0, true, true, 0);
    else if (li == 3) {
      B.gemm(cW3e, false, SM_EC_HI, SM_EC_LO, TM_ACC0, true, true, 0);
      B.gemm(cW3h, true, TM_AHI, TM_ALO, TM_ACC0, false, false, 0);
    } else B.gemm(li == 1 ? cW1 : li == 2 ? cW2 : cW4, true, TM_AHI, TM_ALO, TM_ACC0, true, true, 0);
    B.gemm(cU[li], false, SM_C_HI, SM_C_LO, TM_ACC1, true, false, 1);
  }
  B.gemm(cOut, true, TM_AHI, TM_ALO, TM_ACC0, true, true, 1);
}

// One launch re-lays-out everything: blockIdx.y < n_umma -> UMMA chunk layout (forward), the rest -> the three
// plain copies the backward still reads (padded geometry Fourier matrix and the two odd-width geometry matrices).
struct UPackJobs { UPackJob j[UM_MAX_JOBS]; PackJob legacy[3]; int n_umma; };
__global__ void pack_umma_jobs_kernel(const float* __restrict__ blob, float* __restrict__ packed, float* __restrict__ legacy,
                                      const __grid_constant__ UPackJobs jobs) {
  if ((int)blockIdx.y >= jobs.n_umma) {
    const PackJob jb = jobs.legacy[blockIdx.y - jobs.n_umma];
    const int total = jb.dst_rows * jb.dst_ld;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
      const int n = e / jb.dst_ld, kp = e % jb.dst_ld;   // never transposed
      int k = kp;
      bool ok = true;
      if (kp >= jb.gap_at) {
        if (kp < jb.gap_at + jb.gap) ok = false;
        k = kp - jb.gap;
      }
      float v = 0.f;
      if (ok && k < jb.n_in && n < jb.n_out) v = blob[jb.src + n * jb.n_in + k];
      legacy[jb.dst + e] = v;
    }
    return;
  }
  const UPackJob jb = jobs.j[blockIdx.y];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < jb.total / 2; e += gridDim.x * blockDim.x)
    pack_umma_element(jb, e, blob, packed);
}

// far bound of the z-range of zero-depth rays: min(5*mean, 1.2*max) per group of rays
__global__ void far_bound_kernel(const float* __restrict__ g, int64_t R, int64_t group, float* __restrict__ out) {
  const int64_t b = (int64_t)blockIdx.x * group;
  const int64_t e = b + group < R ? b + group : R;
  double sum = 0.0;
  float mx = -INFINITY;
  for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) { const float v = g[i]; sum += (double)v; mx = fmaxf(mx, v); }
  __shared__ double ssum[32];
  __shared__ float smax[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = sum; smax[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0; float m = -INFINITY;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { s += ssum[w]; m = fmaxf(m, smax[w]); }
    const float mean = (float)(s / (double)(e - b));
    const float far_bb = fminf(5.f * mean, m * 1.2f);
    out[blockIdx.x] = (m > 0.f) ? fminf(fmaxf(far_bb, 0.f), m * 1.2f) : far_bb;
  }
}

// host side ------------------------------------------------------------------------------------
static void add_job(PackJobs& J, int src, int n_out, int n_in, int dst, int dst_rows, int dst_ld, int transpose,
                    int gap_at = 1 << 30, int gap = 0) {
  if (J.n >= MAX_PACK_JOBS) return;   // guarded by the static job list in launch_pack (25 jobs)
  PackJob& j = J.j[J.n++];
  j.src = src; j.n_out = n_out; j.n_in = n_in; j.dst = dst; j.dst_rows = dst_rows; j.dst_ld = dst_ld;
  j.gap_at = gap_at; j.gap = gap; j.transpose = transpose;
}

int check_weights(const LsrWeights* w) {
  if (!w || !w->blob) return LSR_ERR_ARG;
  const int32_t* offs = &w->g_fc_w[0];
  const int n = (int)((&w->c_out_b - &w->g_fc_w[0]) + 1);
  for (int i = 0; i < n; ++i)
    if (offs[i] < 0 || (offs[i] & 3) || offs[i] >= w->n_elems) return LSR_ERR_ARG;
  return LSR_OK;
}

int check_params(const LsrParams* p) {
  if (!p) return LSR_ERR_ARG;
  if (p->nn_num != KNN || p->c_dim != CDIM) return LSR_ERR_UNSUPPORTED;
  if (p->n_surface < 1 || p->n_surface > 8) return LSR_ERR_UNSUPPORTED;
  if (p->rgb_mode < 0 || p->rgb_mode > 2) return LSR_ERR_ARG;
  return LSR_OK;
}

// Rays per tile: the largest tile that fits TILE_M rows, shrunk so that the tiles fill whole waves of
// (SMs x CTAs/SM) evenly -- short tiles skip their unused 16-row MMA tiles, so two even waves of 3/4
// tiles beat one full wave plus a 40 % one.  Forward and backward must agree (saved-row order).
int balanced_rays_per_tile(int64_t n_rays, int n_surface, int nsm) {
  const int rmax = TILE_M / n_surface;
  if (const char* e = getenv("LSR_DEBUG_RAYS_PER_TILE")) {   // tuning experiments only
    const int v = atoi(e);
    if (v >= 1 && v <= rmax) return v;
  }
  const int64_t slots = (int64_t)nsm * CTAS_PER_SM;
  const int64_t waves = (n_rays + rmax * slots - 1) / (rmax * slots);
  int64_t rpt = (n_rays + waves * slots - 1) / (waves * slots);
  if (rpt < 1) rpt = 1;
  if (rpt > rmax) rpt = rmax;
  return (int)rpt;
}

int sm_count() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  cached = n;
  return n;
}

}  // namespace lsr

using namespace lsr;

// Host-logic introspection for the CPU test-suite: the GEMM program a (stage, flags) pair compiles to.
// out[0..7] = {n_ops, n_jobs, packed_floats, ops that wait for a_ready, commits to d_ready[0], commits to d_ready[1],
//              largest chunk bytes (hi + lo), capacity of the packed buffer in floats}
extern "C" int lsr_debug_program_stats(const LsrWeights* w, int stage, int flags, int64_t* out) {
  if (!w || !out) return LSR_ERR_ARG;
  UProgram P;
  build_program(w, stage, flags, &P);
  int waits = 0, c0 = 0, c1 = 0;
  int64_t maxb = 0;
  for (int i = 0; i < P.n_ops; ++i) {
    waits += (P.ops[i].flags & UOP_WAIT_A) ? 1 : 0;
    c0 += P.ops[i].commit_d == 1;
    c1 += P.ops[i].commit_d == 2;
    if (2 * (int64_t)P.ops[i].half_bytes > maxb) maxb = 2 * (int64_t)P.ops[i].half_bytes;
  }
  out[0] = P.n_ops; out[1] = P.n_jobs; out[2] = P.packed_floats; out[3] = waits; out[4] = c0; out[5] = c1;
  out[6] = maxb; out[7] = UMMA_PACKED_FLOATS_MAX;
  return LSR_OK;
}

extern "C" int lsr_render_workspace_bytes(const LsrParams* prm, int64_t n_rays, int stage, size_t* saved_bytes,
                                          size_t* scratch_bytes) {
  int rc = check_params(prm);
  if (rc) return rc;
  if (n_rays < 0 || n_rays > (1ll << 27)) return LSR_ERR_ARG;
  if (saved_bytes) {
    const SavedLayout SLh = saved_layout(n_rays, prm->n_surface, stage, prm->flags);
    *saved_bytes = ((prm->flags & LSR_FLAG_SAVE_LIGHT) ? SLh.light_end : SLh.total) * sizeof(float);
  }
  if (scratch_bytes) {
    const ScratchLayout CLh = scratch_layout(n_rays, prm->n_surface);
    *scratch_bytes = (prm->flags & LSR_FLAG_FWD_ONLY) ? CLh.fwd_end : CLh.total;
  }
  return LSR_OK;
}

extern "C" int lsr_far_bound(const float* gt_depth, int64_t n_rays, int64_t group, float* far_out, lsr_stream_t stream) {
  if (n_rays < 0 || group < 1 || (n_rays > 0 && (!gt_depth || !far_out))) return LSR_ERR_ARG;
  if (n_rays == 0) return LSR_OK;
  const int64_t ng = (n_rays + group - 1) / group;
  far_bound_kernel<<<(unsigned)ng, 256, 0, stream>>>(gt_depth, n_rays, group, far_out);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  return LSR_OK;
}

extern "C" int lsr_render_fwd(const LsrParams* prm, const void* grid_ws, const float* cloud_pos, int64_t n_points,
                              const float* rays_o, const float* rays_d, const float* gt_depth,
                              const double* r_query, const float* far_zero, int64_t far_group,
                              const float* z_zero_depth, int64_t n_rays,
                              const float* geo_feats, const float* col_feats, const int32_t* row_remap,
                              const float* geo_leaf, const float* col_leaf, const LsrWeights* w,
                              const float* exposure_affine, int stage, float* depth, float* var, float* rgb,
                              uint8_t* valid, void* saved, void* scratch, lsr_stream_t stream) {
  int rc = check_params(prm);
  if (rc) return rc;
  rc = check_weights(w);
  if (rc) return rc;
  if (stage != LSR_STAGE_GEOMETRY && stage != LSR_STAGE_COLOR) return LSR_ERR_ARG;
  if (!grid_ws || !scratch || n_rays < 0 || n_rays > (1ll << 27) || n_points < 0) return LSR_ERR_ARG;
  if (n_rays == 0) return LSR_OK;
  if (!rays_o || !rays_d || !gt_depth || !depth || !var || !rgb || !valid) return LSR_ERR_ARG;
  if (n_points > 0 && (!cloud_pos || !geo_feats)) return LSR_ERR_ARG;
  if (n_points > 0 && stage == LSR_STAGE_COLOR && !col_feats) return LSR_ERR_ARG;
  if (row_remap && (!geo_leaf || (stage == LSR_STAGE_COLOR && !col_leaf))) return LSR_ERR_ARG;
  if ((prm->flags & LSR_FLAG_DYNAMIC_R) && !r_query) return LSR_ERR_ARG;
  if ((prm->flags & LSR_FLAG_SAMPLE_NEAR_PCL) && !z_zero_depth) return LSR_ERR_ARG;
  if (prm->rgb_mode == LSR_RGB_AFFINE_SIGMOID && !exposure_affine) return LSR_ERR_ARG;
  const int nsm = sm_count();
  if (nsm <= 0) return LSR_ERR_CUDA;

  const ScratchLayout CL = scratch_layout(n_rays, prm->n_surface);
  char* sbase = (char*)scratch;
  // GEMM program of a tile + weights -> chunked (hi, lo) UMMA layout
  UProgram P;
  build_program(w, stage, prm->flags, &P);
  if (P.packed_floats > UMMA_PACKED_FLOATS_MAX || P.n_ops > UM_MAX_OPS || P.n_jobs > UM_MAX_JOBS) return LSR_ERR_UNSUPPORTED;
  {
    UPackJobs J;
    memcpy(J.j, P.jobs, sizeof(UPackJob) * P.n_jobs);
    J.n_umma = P.n_jobs;
    PackJobs L;   // what the backward (and the bias table above) reads from the legacy scratch
    L.n = 0;
    add_job(L, w->g_B, 3, EG, Packed::gB, 3, EGP, 0);
    add_job(L, w->g_lin_w[0], HG, EG, Packed::gW0n, HG, EGP, 0);
    add_job(L, w->g_lin_w[3], HG, EG + HG, Packed::gW3n, HG, 128, 0, EG, EGP - EG);
    memcpy(J.legacy, L.j, sizeof(PackJob) * 3);
    pack_umma_jobs_kernel<<<dim3(8, P.n_jobs + 3), 256, 0, stream>>>(w->blob, (float*)(sbase + CL.umma),
                                                                      (float*)(sbase + CL.legacy), J);
    LSR_LAUNCHED(1);
    LSR_CUDA_CHECK(cudaGetLastError());
  }

  KnnScratch ks;
  ks.idx = (int32_t*)(sbase + CL.knn_idx);
  ks.rem = (int32_t*)(sbase + CL.knn_rem);
  ks.w = (float*)(sbase + CL.knn_w);
  ks.pos = (float4*)(sbase + CL.knn_pos);
  ks.hw = (float2*)(sbase + CL.knn_hw);
  {
    KnnArgs k;
    k.prm = *prm;
    k.grid = grid_ws;
    k.rays_o = rays_o; k.rays_d = rays_d; k.gt_depth = gt_depth; k.r_query = r_query; k.far_zero = far_zero;
    k.z_override = (prm->flags & LSR_FLAG_SAMPLE_NEAR_PCL) ? z_zero_depth : nullptr;
    k.far_group = (int)(far_group > 0 ? (far_group < (1ll << 30) ? far_group : (1ll << 30)) : 1);
    k.R = (int)n_rays;
    k.remap = row_remap;
    k.ks = ks;
    k.saved = (float*)saved;
    k.stage = stage;
    const int64_t pairs = (n_rays * prm->n_surface + LSR_KNN_NQ - 1) / LSR_KNN_NQ;
    const int64_t blocks = (pairs + 7) / 8;
    const int kgrid = (int)(blocks < (int64_t)nsm * 8 ? blocks : (int64_t)nsm * 8);
    sample_knn_kernel<<<kgrid, 256, 0, stream>>>(k);
    LSR_LAUNCHED(1);
    LSR_CUDA_CHECK(cudaGetLastError());
  }

  FwdArgs a;
  a.prm = *prm;
  a.cloud = cloud_pos;
  a.gt_depth = gt_depth;
  a.R = (int)n_rays;
  a.geo_feats = geo_feats; a.col_feats = col_feats;
  a.geo_leaf = geo_leaf; a.col_leaf = col_leaf;
  a.w = *w;
  a.packed = (const float*)(sbase + CL.legacy);
  a.wpk = (const float*)(sbase + CL.umma);
  a.affine = exposure_affine;
  a.stage = stage;
  a.depth = depth; a.var = var; a.rgb = rgb; a.valid = valid;
  a.saved = (float*)saved;
  a.ks = ks;
  a.rays_per_tile = UM_M / prm->n_surface;
  a.ntiles = (int)((n_rays + a.rays_per_tile - 1) / a.rays_per_tile);
  a.n_ops = P.n_ops;
  memcpy(a.ops, P.ops, sizeof(UOp) * P.n_ops);
  LSR_SMEM_ATTR_ONCE(render_fwd_kernel, FWD_SMEM_BYTES);
  const int grid = a.ntiles < nsm ? a.ntiles : nsm;
  render_fwd_kernel<<<grid, FT, FWD_SMEM_BYTES, stream>>>(a);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  return LSR_OK;
}
