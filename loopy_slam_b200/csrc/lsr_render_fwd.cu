// Fused forward render kernel: z-sampling -> grid k-NN -> IDW gather -> geometry MLP ->
// (rel-pos neighbour MLP) -> colour MLP -> alpha compositing, one persistent CTA per SM, one
// tile of TILE_M sample rows (= floor(128/S) rays) at a time, activations resident in shared memory.
//
// Reference semantics restated (math only):
//   Renderer.render_batch_ray   /root/reference/src/utils/Renderer.py:71-201
//   NICER / MLP_geometry / MLP_color   /root/reference/src/conv_onet/models/decoder.py:106-626
//   raw2outputs_nerf_color      /root/reference/src/common.py:382-422
#include <cstdlib>
// the forward kernel has shared memory to spare: 32-row weight chunks, two stages -> half the barriers per GEMM
#define LSR_RING_FASTPATH 1
#define LSR_KC 32
#define LSR_NSTAGE 2
#include "lsr_render.cuh"

namespace lsr {

#ifdef LSR_PHASE_TIMING
__device__ unsigned long long lsr_phase_cycles[2][16];
#endif

struct FwdArgs {
  LsrParams prm;
  const void* grid;
  const float* cloud;
  const float *rays_o, *rays_d, *gt_depth;
  const double* r_query;
  const float* far_zero;
  int far_group;
  int R;
  const float *geo_feats, *col_feats;
  const int32_t* remap;
  const float *geo_leaf, *col_leaf;
  LsrWeights w;
  const float* packed;
  const float* affine;
  int stage;
  float *depth, *var, *rgb;
  uint8_t* valid;
  float* saved;
  int rays_per_tile, ntiles;
};

__global__ void pack_weights_kernel(const float* __restrict__ blob, float* __restrict__ packed, PackJobs jobs) {
  const PackJob jb = jobs.j[blockIdx.y];
  const int total = jb.dst_rows * jb.dst_ld;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    int kp, n;
    if (jb.transpose) { kp = e / jb.dst_ld; n = e % jb.dst_ld; }
    else              { n = e / jb.dst_ld; kp = e % jb.dst_ld; }
    int k = kp;
    bool ok = true;
    if (kp >= jb.gap_at) {
      if (kp < jb.gap_at + jb.gap) ok = false;
      k = kp - jb.gap;
    }
    float v = 0.f;
    if (ok && k < jb.n_in && n < jb.n_out) v = blob[jb.src + n * jb.n_in + k];
    packed[jb.dst + e] = v;
  }
}

// linspace(start, end, S)[s] exactly as torch computes it in float32
__device__ __forceinline__ float linspace_f32(float start, float end, int S, int s) {
  if (S <= 1) return start;
  const float step = (end - start) / (float)(S - 1);
  return (s < S / 2) ? __fadd_rn(start, __fmul_rn(step, (float)s))
                     : __fsub_rn(end, __fmul_rn(step, (float)(S - 1 - s)));
}

constexpr int FWD_SMEM_FLOATS = TILE_M * XLD + TILE_M * CLD + SB_FLOATS + TILE_M * KNN * 3 +
                                TILE_M * 4 + TILE_M * 3 + TILE_M * 4;

__global__ void __launch_bounds__(NT, CTAS_PER_SM) render_fwd_kernel(const __grid_constant__ FwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* sX = smem;
  float* sC = sX + TILE_M * XLD;
  float* sB = sC + TILE_M * CLD;
  int* sIdx = reinterpret_cast<int*>(sB + SB_FLOATS);
  float* sW = reinterpret_cast<float*>(sIdx + TILE_M * KNN);
  float* sP = sW + TILE_M * KNN;          // [m][4] = px,py,pz,z
  float* sOcc = sP + TILE_M * 4;
  float* sWsum = sOcc + TILE_M;
  int* sHas = reinterpret_cast<int*>(sWsum + TILE_M);
  float* sRgb = reinterpret_cast<float*>(sHas + TILE_M);   // [m][4]
  int* sRem = reinterpret_cast<int*>(sRgb + TILE_M * 4);   // [m][k] leaf row of neighbour k (row_remap), -1 = table

  const int tid = threadIdx.x;
  const int S = a.prm.n_surface;
  const float* __restrict__ blob = a.w.blob;
  const float* __restrict__ packed = a.packed;
  const bool color = a.stage == LSR_STAGE_COLOR;
  const bool relpos = (a.prm.flags & LSR_FLAG_REL_POS) != 0;
  const bool dynr = (a.prm.flags & LSR_FLAG_DYNAMIC_R) != 0;
  const bool save = a.saved != nullptr;
  const SavedLayout SL = saved_layout(a.R, S, a.stage, a.prm.flags);
  const size_t Pp = SL.Pp;   // row pitch of the per-layer saved planes
  const GridHeader* gh_ = reinterpret_cast<const GridHeader*>(a.grid);
  const GridView gv = grid_view(a.grid, gh_->n_points, gh_->max_cells);
  const WideMap wm;
  const NarrowMap nm;

  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int r0 = tile * a.rays_per_tile;
    const int nr = min(a.rays_per_tile, a.R - r0);
    const int nrows = nr * S;
    const size_t p0 = (size_t)r0 * S;

    LSR_PHASE_BEGIN();
    // ---------------------------------------------------------------- A: sample points + k-NN
    // one warp per PAIR of sample rows, searched in lockstep (lane k < 8 ends up owning the k-th neighbour)
    constexpr int NQ = 2;
    for (int m0 = (tid >> 5) * NQ; m0 < TILE_M; m0 += (NT / 32) * NQ) {
      const int lane = tid & 31;
      float px[NQ], py[NQ], pz[NQ], z[NQ], r2f[NQ], rr[NQ];
      double r2d[NQ];
      bool rowvalid[NQ];
      unsigned bD[NQ];
      int bI[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int m = m0 + q;
        rowvalid[q] = m < nrows;
        px[q] = py[q] = pz[q] = z[q] = r2f[q] = rr[q] = 0.f;
        r2d[q] = 0.0;
        if (rowvalid[q]) {
          const int rl = m / S, s = m - rl * S, ray = r0 + rl;
          const float g = a.gt_depth[ray];
          if (g > 0.f) {   // Renderer.py:140-150
            const float t = linspace_f32(0.f, 1.f, S, s);
            const float zn = __fmul_rn(a.prm.near_end_surface, g), zf = __fmul_rn(a.prm.far_end_surface, g);
            z[q] = __fadd_rn(__fmul_rn(zn, __fsub_rn(1.f, t)), __fmul_rn(zf, t));
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
      knn_warp_multi<NQ>(gv, px, py, pz, rr, rowvalid, dynr, r2f, r2d, bD, bI);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int m = m0 + q;
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
        const int has = (rowvalid[q] && ns >= a.prm.min_nn_num) ? 1 : 0;           // decoder.py:204
        if (lane < KNN) {
          sIdx[m * KNN + lane] = vk ? bI[q] : -1;
          sRem[m * KNN + lane] = (vk && a.remap != nullptr) ? __ldg(a.remap + bI[q]) : -1;
          sW[m * KNN + lane] = wn;
          if (save && rowvalid[q]) {
            reinterpret_cast<int*>(a.saved + SL.idx)[(p0 + m) * KNN + lane] = vk ? bI[q] : -1;
            a.saved[SL.w + (p0 + m) * KNN + lane] = wn;
            a.saved[SL.D + (p0 + m) * KNN + lane] = vk ? Dk : FLT_MAX;
          }
        }
        if (lane == 0) {
          *reinterpret_cast<float4*>(sP + m * 4) = make_float4(px[q], py[q], pz[q], z[q]);
          sHas[m] = has;
          sWsum[m] = wn_sum;
          if (save && rowvalid[q])
            reinterpret_cast<float4*>(a.saved + SL.misc)[p0 + m] = make_float4(z[q], (float)has, wn_sum, (float)cnt);
        }
      }
    }
    __syncthreads();

    LSR_PHASE(0, 0);   // knn
    // ---------------------------------------------------------------- B: geometry feature (IDW gather)
    for (int it = tid; it < TILE_M * 8; it += NT) {
      const int m = it >> 3, q = it & 7;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (sHas[m]) {
#pragma unroll
        for (int k = 0; k < KNN; ++k) {
          const int idx = sIdx[m * KNN + k];
          if (idx >= 0) {
            const float w = sW[m * KNN + k];
            const float4 f = __ldg(reinterpret_cast<const float4*>(feat_row_cached(a.geo_feats, a.geo_leaf, idx, sRem[m * KNN + k])) + q);
            acc.x = fmaf(w, f.x, acc.x); acc.y = fmaf(w, f.y, acc.y);
            acc.z = fmaf(w, f.z, acc.z); acc.w = fmaf(w, f.w, acc.w);
          }
        }
      }
      *reinterpret_cast<float4*>(sC + m * CLD + q * 4) = acc;
      if (save && m < nrows) reinterpret_cast<float4*>(a.saved + SL.cg)[(p0 + m) * 8 + q] = acc;
    }
    // ---------------------------------------------------------------- C: geometry Fourier features
    for (int it = tid; it < TILE_M * EGP; it += NT) {
      const int m = it / EGP, j = it - m * EGP;
      float v = 0.f;
      if (j < EG) {
        const float t0 = TWO_PI_F * sP[m * 4 + 0], t1 = TWO_PI_F * sP[m * 4 + 1], t2 = TWO_PI_F * sP[m * 4 + 2];
        const float arg = fmaf(t2, packed[Packed::gB + 2 * EGP + j],
                               fmaf(t1, packed[Packed::gB + EGP + j], t0 * packed[Packed::gB + j]));
        v = sinf(arg);
      }
      sX[m * XLD + j] = v;
    }
    LSR_PHASE(0, 1);   // gather + fourier
    // geometry MLP: h = relu(W x + b) + U c + u  (decoder.py:275-283); epilogues work directly on the MMA
    // accumulator fragments (rows g / g+8, column pairs 2t, 2t+1)
    {
      typedef FragTile<TILE_M, HG> FG;
      FG f;
#pragma unroll 1
      for (int li = 0; li < 5; ++li) {
        const float* A = (li == 0 || li == 3) ? sX : sX + EGP;
        const int Kc = (li == 0) ? EGP : (li == 3 ? 128 : HG);
        const int wt = li == 0 ? Packed::gW0t : li == 1 ? Packed::gW1t : li == 2 ? Packed::gW2t
                     : li == 3 ? Packed::gW3t : Packed::gW4t;
        f.zero();
        if (li == 0) mma_core<TILE_M, HG, true, false>(f.c, A, XLD, Kc, packed + wt, HG, HG, sB, nrows);
        else         mma_core<TILE_M, HG, true, false, true>(f.c, A, XLD, Kc, packed + wt, HG, HG, sB, nrows);
        gemm_prefetch<HG>(packed + Packed::gUt + li * CDIM * HG, HG, CDIM, HG, sB);   // fc_c weights fly during the epilogue
#pragma unroll
        for (int j = 0; j < FG::NJ; ++j) {
          const int col = FG::col(j);
          const float2 b = *reinterpret_cast<const float2*>(blob + a.w.g_lin_b[li] + col);
#pragma unroll
          for (int i = 0; i < FG::MI; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int r = FG::row(i, h);
              const float v0 = fmaxf(f.c[i][j][2 * h] + b.x, 0.f), v1 = fmaxf(f.c[i][j][2 * h + 1] + b.y, 0.f);
              f.c[i][j][2 * h] = v0; f.c[i][j][2 * h + 1] = v1;
              if (save && r < nrows)
                *reinterpret_cast<float2*>(a.saved + SL.gs + ((size_t)li * Pp + p0 + r) * HG + col) = make_float2(v0, v1);
            }
        }
        mma_core<TILE_M, HG, true, false, true>(f.c, sC, CLD, CDIM, packed + Packed::gUt + li * CDIM * HG, HG, HG, sB, nrows);
        if (li < 4) {   // next layer's weights
          const int kn = (li + 1 == 3) ? 128 : HG;
          const int wn_ = li + 1 == 1 ? Packed::gW1t : li + 1 == 2 ? Packed::gW2t : li + 1 == 3 ? Packed::gW3t : Packed::gW4t;
          gemm_prefetch<HG>(packed + wn_, HG, kn, HG, sB);
        }
#pragma unroll
        for (int j = 0; j < FG::NJ; ++j) {
          const int col = FG::col(j);
          const float2 u = *reinterpret_cast<const float2*>(blob + a.w.g_fc_b[li] + col);
#pragma unroll
          for (int i = 0; i < FG::MI; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int r = FG::row(i, h);
              const float2 hv = make_float2(f.c[i][j][2 * h] + u.x, f.c[i][j][2 * h + 1] + u.y);
              *reinterpret_cast<float2*>(sX + r * XLD + EGP + col) = hv;
              if (save && r < nrows)
                *reinterpret_cast<float2*>(a.saved + SL.gh + ((size_t)li * Pp + p0 + r) * HG + col) = hv;
            }
        }
      }
      __syncthreads();
      if (tid < TILE_M) {   // occupancy logit (decoder.py:284)
        float o = blob[a.w.g_out_b];
#pragma unroll 8
        for (int k = 0; k < HG; ++k) o = fmaf(sX[tid * XLD + EGP + k], blob[a.w.g_out_w + k], o);
        sOcc[tid] = o;
        if (save && tid < nrows) a.saved[SL.occ + p0 + tid] = o;
      }
      __syncthreads();
    }

    LSR_PHASE(0, 2);   // geometry MLP
    if (color) {
      // -------------------------------------------------------------- D: colour feature
      if (relpos) {   // decoder.py:477-488
        typedef FragTile<TILE_M, HC> FW;
        FW uf, f;
        uf.zero();
        for (int m = tid; m < TILE_M; m += NT)
          *reinterpret_cast<float4*>(sX + m * XLD + QD) = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int k = 0; k < KNN; ++k) {
          gemm_prefetch<HC>(packed + Packed::V1t, HC, QDP, HC, sB);   // V1 chunk flies while Q_k is built
          for (int it = tid; it < TILE_M * ER; it += NT) {
            const int m = it / ER, j = it - m * ER;
            const int idx = sIdx[m * KNN + k];
            float sn = 0.f, cs = 0.f;
            if (idx >= 0) {
              const float t0 = TWO_PI_F * __fsub_rn(__ldg(a.cloud + 3 * (size_t)idx + 0), sP[m * 4 + 0]);
              const float t1 = TWO_PI_F * __fsub_rn(__ldg(a.cloud + 3 * (size_t)idx + 1), sP[m * 4 + 1]);
              const float t2 = TWO_PI_F * __fsub_rn(__ldg(a.cloud + 3 * (size_t)idx + 2), sP[m * 4 + 2]);
              const float arg = fmaf(t2, blob[a.w.c_Brel + 2 * ER + j],
                                     fmaf(t1, blob[a.w.c_Brel + ER + j], t0 * blob[a.w.c_Brel + j]));
              sincosf(arg, &sn, &cs);
            }
            sX[m * XLD + j] = sn;
            sX[m * XLD + ER + j] = cs;
          }
          for (int it = tid; it < TILE_M * 8; it += NT) {
            const int m = it >> 3, q = it & 7;
            const int idx = sIdx[m * KNN + k];
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx >= 0) f = __ldg(reinterpret_cast<const float4*>(feat_row_cached(a.col_feats, a.col_leaf, idx, sRem[m * KNN + k])) + q);
            *reinterpret_cast<float4*>(sX + m * XLD + 2 * ER + q * 4) = f;
          }
          f.zero();
          mma_core<TILE_M, HC, true, false, true>(f.c, sX, XLD, QDP, packed + Packed::V1t, HC, HC, sB, nrows);
#pragma unroll
          for (int j = 0; j < FW::NJ; ++j) {
            const int col = FW::col(j);
            const float2 b = *reinterpret_cast<const float2*>(blob + a.w.c_nb1_b + col);
#pragma unroll
            for (int i = 0; i < FW::MI; ++i)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int r = FW::row(i, h);
                const float wk = sW[r * KNN + k];
                const float s0 = softplus100(f.c[i][j][2 * h] + b.x), s1 = softplus100(f.c[i][j][2 * h + 1] + b.y);
                uf.c[i][j][2 * h] = fmaf(wk, s0, uf.c[i][j][2 * h]);
                uf.c[i][j][2 * h + 1] = fmaf(wk, s1, uf.c[i][j][2 * h + 1]);
                if (save && r < nrows)
                  *reinterpret_cast<float2*>(a.saved + SL.sp + ((p0 + r) * KNN + k) * HC + col) = make_float2(s0, s1);
              }
          }
        }
        gemm_prefetch<CDIM>(packed + Packed::V2t, CDIM, HC, CDIM, sB);
#pragma unroll
        for (int j = 0; j < FW::NJ; ++j) {
          const int col = FW::col(j);
#pragma unroll
          for (int i = 0; i < FW::MI; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int r = FW::row(i, h);
              const float2 u = make_float2(uf.c[i][j][2 * h], uf.c[i][j][2 * h + 1]);
              *reinterpret_cast<float2*>(sX + r * XLD + col) = u;
              if (save && r < nrows) *reinterpret_cast<float2*>(a.saved + SL.u + (p0 + r) * HC + col) = u;
            }
        }
        typedef FragTile<TILE_M, CDIM> FC;
        FC c4;
        c4.zero();
        mma_core<TILE_M, CDIM, true, false, true>(c4.c, sX, XLD, HC, packed + Packed::V2t, CDIM, CDIM, sB, nrows);
#pragma unroll
        for (int j = 0; j < FC::NJ; ++j) {
          const int col = FC::col(j);
          const float2 v2 = *reinterpret_cast<const float2*>(blob + a.w.c_nb2_b + col);
#pragma unroll
          for (int i = 0; i < FC::MI; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int r = FC::row(i, h);
              const float ws = sWsum[r];
              float2 cc = make_float2(0.f, 0.f);
              if (sHas[r]) cc = make_float2(fmaf(v2.x, ws, c4.c[i][j][2 * h]), fmaf(v2.y, ws, c4.c[i][j][2 * h + 1]));
              *reinterpret_cast<float2*>(sC + r * CLD + col) = cc;
              if (save && r < nrows) *reinterpret_cast<float2*>(a.saved + SL.cc + (p0 + r) * CDIM + col) = cc;
            }
        }
      } else {        // decoder.py:476,487-488
        for (int it = tid; it < TILE_M * 8; it += NT) {
          const int m = it >> 3, q = it & 7;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          if (sHas[m]) {
#pragma unroll
            for (int k = 0; k < KNN; ++k) {
              const int idx = sIdx[m * KNN + k];
              if (idx >= 0) {
                const float w = sW[m * KNN + k];
                const float4 f = __ldg(reinterpret_cast<const float4*>(feat_row_cached(a.col_feats, a.col_leaf, idx, sRem[m * KNN + k])) + q);
                acc.x = fmaf(w, f.x, acc.x); acc.y = fmaf(w, f.y, acc.y);
                acc.z = fmaf(w, f.z, acc.z); acc.w = fmaf(w, f.w, acc.w);
              }
            }
          }
          *reinterpret_cast<float4*>(sC + m * CLD + q * 4) = acc;
          if (save && m < nrows) reinterpret_cast<float4*>(a.saved + SL.cc)[(p0 + m) * 8 + q] = acc;
        }
      }
      __syncthreads();
      LSR_PHASE(0, 3);   // rel-pos neighbour MLP / colour gather
      // -------------------------------------------------------------- E: colour trunk (decoder.py:515-533)
      for (int it = tid; it < TILE_M * EC; it += NT) {
        const int m = it / EC, j = it - m * EC;
        const float t0 = TWO_PI_F * sP[m * 4 + 0], t1 = TWO_PI_F * sP[m * 4 + 1], t2 = TWO_PI_F * sP[m * 4 + 2];
        const float arg = fmaf(t2, blob[a.w.c_B + 2 * EC + j], fmaf(t1, blob[a.w.c_B + EC + j], t0 * blob[a.w.c_B + j]));
        float sn, cs;
        sincosf(arg, &sn, &cs);
        sX[m * XLD + j] = sn;
        sX[m * XLD + EC + j] = cs;
      }
      {
        typedef FragTile<TILE_M, HC> FW;
        FW f;
#pragma unroll 1
        for (int li = 0; li < 5; ++li) {
          const float* A = (li == 0 || li == 3) ? sX : sX + ECC;
          const int Kc = (li == 0) ? ECC : (li == 3 ? ECC + HC : HC);
          const int wt = li == 0 ? Packed::cW0t : li == 1 ? Packed::cW1t : li == 2 ? Packed::cW2t
                       : li == 3 ? Packed::cW3t : Packed::cW4t;
          f.zero();
          if (li == 0) mma_core<TILE_M, HC, true, false>(f.c, A, XLD, Kc, packed + wt, HC, HC, sB, nrows);
          else         mma_core<TILE_M, HC, true, false, true>(f.c, A, XLD, Kc, packed + wt, HC, HC, sB, nrows);
          gemm_prefetch<HC>(packed + Packed::cUt + li * CDIM * HC, HC, CDIM, HC, sB);   // fc_c weights fly during the epilogue
#pragma unroll
          for (int j = 0; j < FW::NJ; ++j) {
            const int col = FW::col(j);
            const float2 b = *reinterpret_cast<const float2*>(blob + a.w.c_lin_b[li] + col);
#pragma unroll
            for (int i = 0; i < FW::MI; ++i)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int r = FW::row(i, h);
                const float s0 = softplus100(f.c[i][j][2 * h] + b.x), s1 = softplus100(f.c[i][j][2 * h + 1] + b.y);
                f.c[i][j][2 * h] = s0; f.c[i][j][2 * h + 1] = s1;
                if (save && r < nrows)
                  *reinterpret_cast<float2*>(a.saved + SL.cs + ((size_t)li * Pp + p0 + r) * HC + col) = make_float2(s0, s1);
              }
          }
          mma_core<TILE_M, HC, true, false, true>(f.c, sC, CLD, CDIM, packed + Packed::cUt + li * CDIM * HC, HC, HC, sB, nrows);
          if (li < 4) {   // next layer's weights
            const int kn = (li + 1 == 3) ? ECC + HC : HC;
            const int wn_ = li + 1 == 1 ? Packed::cW1t : li + 1 == 2 ? Packed::cW2t : li + 1 == 3 ? Packed::cW3t : Packed::cW4t;
            gemm_prefetch<HC>(packed + wn_, HC, kn, HC, sB);
          }
#pragma unroll
          for (int j = 0; j < FW::NJ; ++j) {
            const int col = FW::col(j);
            const float2 u = *reinterpret_cast<const float2*>(blob + a.w.c_fc_b[li] + col);
#pragma unroll
            for (int i = 0; i < FW::MI; ++i)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int r = FW::row(i, h);
                const float2 hv = make_float2(f.c[i][j][2 * h] + u.x, f.c[i][j][2 * h + 1] + u.y);
                *reinterpret_cast<float2*>(sX + r * XLD + ECC + col) = hv;
                if (save && r < nrows)
                  *reinterpret_cast<float2*>(a.saved + SL.ch + ((size_t)li * Pp + p0 + r) * HC + col) = hv;
              }
          }
        }
      }
      __syncthreads();
      LSR_PHASE(0, 4);   // colour trunk
      // colour head (decoder.py:533-546)
      for (int it = tid; it < TILE_M * 3; it += NT) {
        const int m = it % TILE_M, ch = it / TILE_M;
        float o = blob[a.w.c_out_b + ch];
        const float* wrow = blob + a.w.c_out_w + ch * HC;
        const float* hrow = sX + m * XLD + ECC;
#pragma unroll 8
        for (int k = 0; k < HC; ++k) o = fmaf(hrow[k], wrow[k], o);
        sRgb[m * 4 + ch] = o;
      }
      __syncthreads();
      if (tid < TILE_M) {
        const float o0 = sRgb[tid * 4 + 0], o1 = sRgb[tid * 4 + 1], o2 = sRgb[tid * 4 + 2];
        float r0c = o0, r1c = o1, r2c = o2;
        if (a.prm.rgb_mode == LSR_RGB_SIGMOID) {
          r0c = sigmoidf_acc(o0); r1c = sigmoidf_acc(o1); r2c = sigmoidf_acc(o2);
        } else if (a.prm.rgb_mode == LSR_RGB_AFFINE_SIGMOID) {   // out @ A + t  (decoder.py:538-540)
          const float* Af = a.affine;
          r0c = sigmoidf_acc(fmaf(o2, Af[6], fmaf(o1, Af[3], o0 * Af[0])) + Af[9]);
          r1c = sigmoidf_acc(fmaf(o2, Af[7], fmaf(o1, Af[4], o0 * Af[1])) + Af[10]);
          r2c = sigmoidf_acc(fmaf(o2, Af[8], fmaf(o1, Af[5], o0 * Af[2])) + Af[11]);
        }
        sRgb[tid * 4 + 0] = r0c; sRgb[tid * 4 + 1] = r1c; sRgb[tid * 4 + 2] = r2c;
        if (save && tid < nrows) {
          reinterpret_cast<float4*>(a.saved + SL.rgbs)[p0 + tid] = make_float4(r0c, r1c, r2c, 0.f);
          reinterpret_cast<float4*>(a.saved + SL.outraw)[p0 + tid] = make_float4(o0, o1, o2, 0.f);
        }
      }
      __syncthreads();
    } else {
      if (tid < TILE_M) { sRgb[tid * 4 + 0] = 0.f; sRgb[tid * 4 + 1] = 0.f; sRgb[tid * 4 + 2] = 0.f; }
      __syncthreads();
    }

    LSR_PHASE(0, 5);   // colour head
    // ---------------------------------------------------------------- F: compositing (common.py:402-422)
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
        const float z = sP[m * 4 + 3];
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
        depth = 0.f;
        if (a.prm.flags & LSR_FLAG_SKIP_ZERO_DEPTH) { o0 = 0.f; o1 = 0.f; o2 = 0.f; }
      }
      a.depth[ray] = depth;
      a.var[ray] = var;
      a.rgb[3 * ray + 0] = o0; a.rgb[3 * ray + 1] = o1; a.rgb[3 * ray + 2] = o2;
      a.valid[ray] = (nhas >= S / 2 + 1) ? 1 : 0;            // decoder.py:259-260
    }
    __syncthreads();
    LSR_PHASE(0, 6);   // compositing
  }
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

int launch_pack(const LsrWeights* w, float* packed, cudaStream_t st) {
  PackJobs J;
  J.n = 0;
  add_job(J, w->g_B, 3, EG, Packed::gB, 3, EGP, 0);
  add_job(J, w->g_lin_w[0], HG, EG, Packed::gW0t, EGP, HG, 1);
  add_job(J, w->g_lin_w[1], HG, HG, Packed::gW1t, HG, HG, 1);
  add_job(J, w->g_lin_w[2], HG, HG, Packed::gW2t, HG, HG, 1);
  add_job(J, w->g_lin_w[3], HG, EG + HG, Packed::gW3t, 128, HG, 1, EG, EGP - EG);
  add_job(J, w->g_lin_w[4], HG, HG, Packed::gW4t, HG, HG, 1);
  for (int i = 0; i < 5; ++i) add_job(J, w->g_fc_w[i], HG, CDIM, Packed::gUt + i * CDIM * HG, CDIM, HG, 1);
  add_job(J, w->g_lin_w[0], HG, EG, Packed::gW0n, HG, EGP, 0);
  add_job(J, w->g_lin_w[3], HG, EG + HG, Packed::gW3n, HG, 128, 0, EG, EGP - EG);
  add_job(J, w->c_lin_w[0], HC, ECC, Packed::cW0t, ECC, HC, 1);
  add_job(J, w->c_lin_w[1], HC, HC, Packed::cW1t, HC, HC, 1);
  add_job(J, w->c_lin_w[2], HC, HC, Packed::cW2t, HC, HC, 1);
  add_job(J, w->c_lin_w[3], HC, ECC + HC, Packed::cW3t, ECC + HC, HC, 1);
  add_job(J, w->c_lin_w[4], HC, HC, Packed::cW4t, HC, HC, 1);
  for (int i = 0; i < 5; ++i) add_job(J, w->c_fc_w[i], HC, CDIM, Packed::cUt + i * CDIM * HC, CDIM, HC, 1);
  add_job(J, w->c_nb1_w, HC, QD, Packed::V1t, QDP, HC, 1);
  add_job(J, w->c_nb2_w, CDIM, HC, Packed::V2t, HC, CDIM, 1);
  pack_weights_kernel<<<dim3(8, J.n), 256, 0, st>>>(w->blob, packed, J);
  return cudaGetLastError() == cudaSuccess ? LSR_OK : LSR_ERR_CUDA;
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

extern "C" int lsr_render_workspace_bytes(const LsrParams* prm, int64_t n_rays, int stage, size_t* saved_bytes,
                                          size_t* scratch_bytes) {
  int rc = check_params(prm);
  if (rc) return rc;
  if (n_rays < 0 || n_rays > (1ll << 27)) return LSR_ERR_ARG;
  if (saved_bytes) *saved_bytes = saved_layout(n_rays, prm->n_surface, stage, prm->flags).total * sizeof(float);
  if (scratch_bytes) *scratch_bytes = align_up((size_t)Packed::total * sizeof(float), 256) + 256;
  return LSR_OK;
}

extern "C" int lsr_far_bound(const float* gt_depth, int64_t n_rays, int64_t group, float* far_out, lsr_stream_t stream) {
  if (n_rays < 0 || group < 1 || (n_rays > 0 && (!gt_depth || !far_out))) return LSR_ERR_ARG;
  if (n_rays == 0) return LSR_OK;
  const int64_t ng = (n_rays + group - 1) / group;
  far_bound_kernel<<<(unsigned)ng, 256, 0, stream>>>(gt_depth, n_rays, group, far_out);
  LSR_CUDA_CHECK(cudaGetLastError());
  return LSR_OK;
}

extern "C" int lsr_render_fwd(const LsrParams* prm, const void* grid_ws, const float* cloud_pos, int64_t n_points,
                              const float* rays_o, const float* rays_d, const float* gt_depth,
                              const double* r_query, const float* far_zero, int64_t far_group, int64_t n_rays,
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
  if (prm->rgb_mode == LSR_RGB_AFFINE_SIGMOID && !exposure_affine) return LSR_ERR_ARG;
  const int nsm = sm_count();
  if (nsm <= 0) return LSR_ERR_CUDA;

  rc = launch_pack(w, (float*)scratch, stream);
  if (rc) return rc;

  FwdArgs a;
  a.prm = *prm;
  a.grid = grid_ws; a.cloud = cloud_pos;
  a.rays_o = rays_o; a.rays_d = rays_d; a.gt_depth = gt_depth; a.r_query = r_query; a.far_zero = far_zero;
  a.far_group = (int)(far_group > 0 ? (far_group < (1ll << 30) ? far_group : (1ll << 30)) : 1);
  a.R = (int)n_rays;
  a.geo_feats = geo_feats; a.col_feats = col_feats;
  a.remap = row_remap; a.geo_leaf = geo_leaf; a.col_leaf = col_leaf;
  a.w = *w;
  a.packed = (const float*)scratch;
  a.affine = exposure_affine;
  a.stage = stage;
  a.depth = depth; a.var = var; a.rgb = rgb; a.valid = valid;
  a.saved = (float*)saved;
  a.rays_per_tile = balanced_rays_per_tile(n_rays, prm->n_surface, nsm);
  a.ntiles = (int)((n_rays + a.rays_per_tile - 1) / a.rays_per_tile);
  const size_t smem = FWD_SMEM_FLOATS * sizeof(float);
  LSR_CUDA_CHECK(cudaFuncSetAttribute(render_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int slots = nsm * CTAS_PER_SM;
  const int grid = a.ntiles < slots ? a.ntiles : slots;
  render_fwd_kernel<<<grid, NT, smem, stream>>>(a);
  LSR_CUDA_CHECK(cudaGetLastError());
  return LSR_OK;
}
