// Geometry-decoder backward chain on tcgen05 / TMEM (sm_100a).
//
// Autograd of MLP_geometry.forward (/root/reference/src/conv_onet/models/decoder.py:265-288) through the five hidden
// layers of width 32, for one 128-row tile per CTA:
//     dH_4 = dOcc (x) w_out                                  (rank-1, registers)
//     dA_l = dH_l * [s_l > 0]                                 (ReLU mask from the saved layer outputs)
//     [dH_{l-1} | dC_l] = dA_l . [W_l^h | P_l]   P_l = W_l^h U_{l-1}      (l = 4..1; ONE MMA group, N = 64)
//     dE = dA_3 . W_3^e + dA_0 . W_0                          (rides on the l = 3 group as 96 more columns; l = 0: N = 96)
//     dL/dc = dOcc (x) (w_out U_4) + sum_l dC_l               (P_l folds the fc_c path into the dX GEMM)
// Every product is the error-compensated 3xTF32 sum A_lo.B_hi + A_hi.B_lo + A_hi.B_hi (A = dA as a TMEM operand: lane = row,
// hi in columns [0,32), lo in [32,64); B = the weight block in the canonical K-major shared-memory layout, raw fp32 image
// = hi because the tensor core truncates, lo image built once per CTA).  The chain is four dependent round trips of
// 12 MMAs each -- latency, not throughput -- so a tile takes ~10 k cycles where the mma.sync version of this chain inside
// render_bwd_kernel spent most of that kernel's 100 us in CTA barriers around its small GEMMs.
// Results go to scratch planes the remaining backward kernel (IDW scatter, Fourier backward, geometry weight gradients)
// reads: dL/dc (P x 32), dE (P x 96) and, when the geometry decoder trains, dH_l (5 x P x 32).
#include "lsr_render.cuh"
#include "lsr_umma.cuh"

namespace lsr {
using namespace umma;

// float offsets of the per-layer B blocks inside the pack: block of N rows, K = 32: element (n, k) at ((k / 4) * N + n) * 4 + k % 4
__host__ __device__ constexpr int gpk_off(int l) { return l == 0 ? 11264 : l == 1 ? 9216 : l == 2 ? 7168 : l == 3 ? 2048 : 0; }
__host__ __device__ constexpr int gpk_n(int l) { return l == 0 ? 96 : l == 3 ? 160 : 64; }
constexpr int GPK_V = 14336;                                // v = w_out U_4 (32)
static_assert(GPK_V + 32 == GEO_PACK_FLOATS, "geometry backward pack size");

// one block per operand row (448 rows over the five layers) + one for v; 32 threads = the contraction index k = o
__global__ void __launch_bounds__(32) geo_bwd_prep_kernel(const float* __restrict__ blob, float* __restrict__ pack,
                                                          const __grid_constant__ LsrWeights w) {
  const int k = threadIdx.x;
  int r = blockIdx.x;
  if (r == 448) {                      // v[c] = sum_i w_out[i] U_4[i][c]
    float s = 0.f;
    for (int i = 0; i < HG; ++i) s = fmaf(blob[w.g_out_w + i], blob[w.g_fc_w[4] + i * CDIM + k], s);
    pack[GPK_V + k] = s;
    return;
  }
  int l = 4;
  if (r < 64) l = 4;
  else if (r < 224) { l = 3; r -= 64; }
  else if (r < 288) { l = 2; r -= 224; }
  else if (r < 352) { l = 1; r -= 288; }
  else { l = 0; r -= 352; }
  const int n = r, N = gpk_n(l);
  const int ld = l == 0 ? EG : (l == 3 ? EG + HG : HG), hoff = l == 3 ? EG : 0;
  const float* W = blob + w.g_lin_w[l] + (size_t)k * ld;      // row o = k of W_l
  float v = 0.f;
  if (l == 0) {
    v = n < EG ? W[n] : 0.f;
  } else if (n < HG) {
    v = W[hoff + n];
  } else if (n < 2 * HG) {
    const int c = n - HG;
    const float* U = blob + w.g_fc_w[l - 1];
    for (int i = 0; i < HG; ++i) v = fmaf(W[hoff + i], U[i * CDIM + c], v);
  } else {
    const int e = n - 2 * HG;
    v = e < EG ? W[e] : 0.f;
  }
  pack[gpk_off(l) + ((k >> 2) * N + n) * 4 + (k & 3)] = v;
}

struct GeoArgs {
  LsrParams prm;
  LsrWeights w;
  const float* saved;
  const float* gpack;
  const float *gt_depth, *g_depth, *g_var, *g_rgb;
  int R, stage, gflags;
  float *gdc, *gde, *gdh;
  int ntiles, rays_per_tile;
};

// Two independent 128-thread groups per CTA, each walking its own tiles with its own TMEM columns and mbarrier: the chain of a
// tile is latency (five dependent MMA round trips), so a second tile in flight on the SM hides most of it, and the 112 KB of
// weight images are shared.
constexpr int GB_GROUPS = 2;
constexpr int GB_NT = 128 * GB_GROUPS;
constexpr int GSM_WHI = 0;
constexpr int GSM_WLO = GSM_WHI + GPK_V * 4;
constexpr int GSM_DOCC = GSM_WLO + GPK_V * 4;       // [group][128]
constexpr int GSM_TAB = GSM_DOCC + GB_GROUPS * 128 * 4;   // w_out (32) | v (32)
constexpr int GSM_BAR = GSM_TAB + 64 * 4;
constexpr int GEO_SMEM_BYTES = GSM_BAR + 64;
constexpr uint32_t GT_AHI = 0, GT_ALO = 32, GT_D = 64, GT_DE = 128, GT_GROUP = 224;   // TMEM columns per group (512 allocated)

__device__ __forceinline__ void bar_group(int g) { asm volatile("bar.sync %0, 128;\n" ::"r"(1 + g) : "memory"); }

__global__ void __launch_bounds__(GB_NT, 1) geo_bwd_umma_kernel(const __grid_constant__ GeoArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* sWhi = reinterpret_cast<float*>(smem + GSM_WHI);
  float* sWlo = reinterpret_cast<float*>(smem + GSM_WLO);
  float* sTab = reinterpret_cast<float*>(smem + GSM_TAB);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + GSM_BAR + 32);

  const int grp = threadIdx.x >> 7, tid = threadIdx.x & 127, warp = tid >> 5;     // tid / warp: inside the group
  float* sDOcc = reinterpret_cast<float*>(smem + GSM_DOCC) + grp * 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + GSM_BAR) + grp;
  const int S = a.prm.n_surface;
  const float* __restrict__ sv = a.saved;
  const bool color = a.stage == LSR_STAGE_COLOR;
  const bool g_gw = (a.gflags & LSR_GRAD_GEO_W) != 0 && a.gdh != nullptr;
  const bool need_e = a.gde != nullptr;
  const SavedLayout SL = saved_layout(a.R, S, a.stage, a.prm.flags);
  const size_t Pp = SL.Pp;

  // weights: raw image (= hi operand, the tensor core reads the top 19 bits) and lo residual image
  {
    constexpr int PER = GPK_V / 4 / GB_NT;          // 14 float4 per thread: every load in flight before the first use
    static_assert(PER * GB_NT * 4 == GPK_V, "pack size");
    float4 v[PER];
#pragma unroll
    for (int t = 0; t < PER; ++t) v[t] = __ldg(reinterpret_cast<const float4*>(a.gpack) + threadIdx.x + t * GB_NT);
#pragma unroll
    for (int t = 0; t < PER; ++t) {
      uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
      split_hi_lo(v[t].x, h0, l0); split_hi_lo(v[t].y, h1, l1); split_hi_lo(v[t].z, h2, l2); split_hi_lo(v[t].w, h3, l3);
      reinterpret_cast<float4*>(sWhi)[threadIdx.x + t * GB_NT] = v[t];
      reinterpret_cast<uint4*>(sWlo)[threadIdx.x + t * GB_NT] = make_uint4(l0, l1, l2, l3);
    }
  }
  if (threadIdx.x < 32) { sTab[threadIdx.x] = a.w.blob[a.w.g_out_w + threadIdx.x]; sTab[32 + threadIdx.x] = a.gpack[GPK_V + threadIdx.x]; }
  if (threadIdx.x < 32) tmem_alloc(tslot, 512);
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tslot + (uint32_t)grp * GT_GROUP;
  const uint32_t lane_base = 32u * (uint32_t)warp;
  const uint32_t whi = smem_u32(sWhi), wlo = smem_u32(sWlo);
  uint32_t parity = 0;

  // one layer's MMA group: D[dcol, dcol + N) (+)= A . B_l^T, 3 passes x 4 K-steps of 8
  auto issue = [&](int l, uint32_t dcol, uint32_t accumulate) {
    const uint32_t N = (uint32_t)gpk_n(l), lbo = N * 16u, idesc = idesc_tf32(128, (int)N);
    const uint32_t bh = whi + (uint32_t)gpk_off(l) * 4u, bl = wlo + (uint32_t)gpk_off(l) * 4u;
    uint32_t acc = accumulate;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
      const uint32_t acol = pass == 0 ? GT_ALO : GT_AHI;
      const uint32_t b = pass == 1 ? bl : bh;
#pragma unroll
      for (int k8 = 0; k8 < 4; ++k8) {
        mma_ts(tb + dcol, tb + acol + 8u * (uint32_t)k8, smem_desc(b + (uint32_t)k8 * 2u * lbo, lbo, 128u), idesc, acc);
        acc = 1u;
      }
    }
  };

  for (int tile = blockIdx.x + grp * (int)gridDim.x; tile < a.ntiles; tile += GB_GROUPS * (int)gridDim.x) {
    const int r0 = tile * a.rays_per_tile;
    const int nr = min(a.rays_per_tile, a.R - r0);
    const int nrows = nr * S;
    const size_t p0 = (size_t)r0 * S;
    const int row = tid;
    const bool rv = row < nrows;

    // ReLU masks of all five layers up front (one DRAM round trip instead of five): bit j of mbits[l] = s_l[row][j] > 0
    uint32_t mbits[5];
#pragma unroll
    for (int l = 0; l < 5; ++l) {
      const float4* srow = reinterpret_cast<const float4*>(sv + SL.gs + ((size_t)l * Pp + p0 + row) * HG);
      uint32_t m = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rv) s4 = __ldg(srow + j);
        m |= (s4.x > 0.f ? 1u : 0u) << (4 * j) | (s4.y > 0.f ? 2u : 0u) << (4 * j) | (s4.z > 0.f ? 4u : 0u) << (4 * j) |
             (s4.w > 0.f ? 8u : 0u) << (4 * j);
      }
      mbits[l] = m;
    }

    // ---- compositing backward (Renderer.py:184-201, common.py:382-422): d(occupancy logit) per sample row
    sDOcc[tid] = 0.f;
    bar_group(grp);
    if (tid < nr) composite_bwd_ray(a.prm, color, sv, SL, p0 + (size_t)tid * S, r0 + tid, a.gt_depth, a.g_depth, a.g_var, a.g_rgb,
                                    sDOcc + tid * S, nullptr);
    bar_group(grp);
    const float docc = sDOcc[row];
    float dH[32], dC[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { dH[j] = docc * sTab[j]; dC[j] = docc * sTab[32 + j]; }

#pragma unroll
    for (int l = 4; l >= 0; --l) {
      if (g_gw && rv) {
        float4* dst = reinterpret_cast<float4*>(a.gdh + ((size_t)l * Pp + p0 + row) * HG);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = make_float4(dH[4 * j], dH[4 * j + 1], dH[4 * j + 2], dH[4 * j + 3]);
      }
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) split_hi_lo((mbits[l] >> j) & 1u ? dH[j] : 0.f, hi[j], lo[j]);
      tmem_st32(tmem_addr(tb, lane_base, GT_AHI), hi);
      tmem_st32(tmem_addr(tb, lane_base, GT_ALO), lo);
      tmem_wait_st();
      tc_fence_before();
      bar_group(grp);
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        if (l == 0) issue(0, GT_DE, 1u);          // dE += dA_0 . W_0
        else issue(l, GT_D, 0u);                  // [dH_{l-1} | dC_l (| dE)] = dA_l . [W_l^h | P_l (| W_3^e)]
        mma_commit(bar);
      }
      mbar_wait(bar, parity);
      parity ^= 1u;
      tc_fence_after();
      if (l >= 1) {
        uint32_t x[32], y[32];
        tmem_ld32(tmem_addr(tb, lane_base, GT_D), x);
        tmem_ld32(tmem_addr(tb, lane_base, GT_D + 32), y);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) { dH[j] = __uint_as_float(x[j]); dC[j] += __uint_as_float(y[j]); }
      }
    }
    if (rv) {
      float4* dst = reinterpret_cast<float4*>(a.gdc + (p0 + row) * CDIM);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = make_float4(dC[4 * j], dC[4 * j + 1], dC[4 * j + 2], dC[4 * j + 3]);
    }
    if (need_e) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint32_t x[32];
        tmem_ld32(tmem_addr(tb, lane_base, GT_DE + 32 * c), x);
        tmem_wait_ld();
        if (rv) {
          float4* dst = reinterpret_cast<float4*>(a.gde + (p0 + row) * EGP + 32 * c);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(x[4 * j]), __uint_as_float(x[4 * j + 1]), __uint_as_float(x[4 * j + 2]),
                                 __uint_as_float(x[4 * j + 3]));
        }
      }
    }
    tc_fence_before();
    bar_group(grp);      // every TMEM read of this tile is done before the group's next tile overwrites the columns
    tc_fence_after();
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(*tslot, 512);
}

int sm_count();

// gdc / gde / gdh: scratch planes (ScratchLayout::bwd_gdc / bwd_gde / bwd_gdh); gde / gdh may be null when not needed
int launch_geo_bwd(const LsrParams* prm, const LsrWeights* w, const float* gt_depth, int64_t n_rays, int stage, const void* saved,
                   float* gpack, const float* g_depth, const float* g_var, const float* g_rgb, int grad_flags, float* gdc, float* gde,
                   float* gdh, cudaStream_t stream) {
  const int nsm = sm_count();
  if (nsm <= 0) return LSR_ERR_CUDA;
  geo_bwd_prep_kernel<<<449, 32, 0, stream>>>(w->blob, gpack, *w);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  GeoArgs a;
  a.prm = *prm; a.w = *w; a.saved = (const float*)saved; a.gpack = gpack;
  a.gt_depth = gt_depth; a.g_depth = g_depth; a.g_var = g_var; a.g_rgb = g_rgb;
  a.R = (int)n_rays; a.stage = stage; a.gflags = grad_flags;
  a.gdc = gdc; a.gde = gde; a.gdh = gdh;
  const SavedLayout SL = saved_layout(n_rays, prm->n_surface, stage, prm->flags);
  a.rays_per_tile = SL.rays_per_tile;
  a.ntiles = SL.ntiles;
  LSR_SMEM_ATTR_ONCE(geo_bwd_umma_kernel, GEO_SMEM_BYTES);
  const int grid = a.ntiles < nsm ? a.ntiles : nsm;
  geo_bwd_umma_kernel<<<grid, GB_NT, GEO_SMEM_BYTES, stream>>>(a);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  return LSR_OK;
}

}  // namespace lsr
