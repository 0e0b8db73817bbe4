// Fused backward render kernel: compositing backward -> colour MLP backward (+ rel-pos neighbour
// MLP) -> geometry MLP backward -> IDW / Fourier backward -> feature scatter-add, decoder weight
// gradients (per-tile partial GEMMs + red.global.add.v4), ray gradients.  Mirrors the autograd
// graph of Renderer.render_batch_ray (/root/reference/src/utils/Renderer.py:71-201) that the
// reference differentiates with loss.backward() (src/Mapper.py:722, src/Tracker.py:193).
// Math: SURVEY.md Appendix A ("Backward of step 10") + the chain rule through Appendix A steps 2-7.
#include <mutex>

#include "lsr_render.cuh"

namespace lsr {

#ifdef LSR_PHASE_TIMING
extern __device__ unsigned long long lsr_phase_cycles[2][16];
#endif

struct BwdArgs {
  LsrParams prm;
  const float* cloud;
  const float *rays_o, *rays_d, *gt_depth;
  int R;
  const float *geo_feats, *col_feats;
  const int32_t* remap;
  const float *geo_leaf, *col_leaf;
  LsrWeights w;
  const float* packed;
  const float* affine;
  int stage, is_tracker;
  const float* saved;
  const float *g_depth, *g_var, *g_rgb;
  int gflags;
  float *d_geo, *d_col, *d_w, *d_affine, *d_ro, *d_rd;
  const float *ext_dc, *ext_dp, *ext_dwh;   // colour backward hand-over (scratch planes)
  const float *ext_gdc, *ext_gde, *ext_gdh; // geometry chain hand-over (lsr_geo_bwd_umma.cu): dL/dc^g (P x 32), dL/de (P x 96), dL/dh_l (5 x P x 32)
  int rays_per_tile, ntiles;
};

constexpr int E2_FLOATS = 2 * TILE_M * ELD;   // {sE, sDE} | {sQ, sDQ}
static_assert(TILE_M * (QLD + 24) <= E2_FLOATS && TILE_M * (CLD + GLD) <= TILE_M * DLD, "smem aliasing");
constexpr int DQLD = 24;
constexpr int RED_FLOATS = 320;           // 3*96 dB_geo + 30 dB_rel
constexpr int BWD_SMEM_FLOATS = TILE_M * DLD + SB_FLOATS + 2 * TILE_M * CLD + E2_FLOATS + 3 * TILE_M * KNN +
                                2 * TILE_M * 4 + TILE_M * KNN + 3 * TILE_M + TILE_M * 4 + RED_FLOATS;

__device__ __forceinline__ float half_warp_sum(float v) {   // sum over the 16 lanes sharing ty
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// dF[idx_k[m]] += w_k * dC[m]  and (tracker) dw_hat_k += dC[m] . F[idx_k[m]]
__device__ __forceinline__ void scatter_idw(const float* sDC, const int* sIdx, const float* sW, const int* sHas,
                                            float* sDWh, const float* __restrict__ feats,
                                            const float* __restrict__ leaf, const int32_t* __restrict__ remap,
                                            float* __restrict__ d_feats, bool want_feat, bool want_w) {
  for (int it = threadIdx.x; it < TILE_M * 64; it += NT) {
    const int q = it & 7, k = (it >> 3) & 7, m = it >> 6;
    const int idx = sIdx[m * KNN + k];
    const bool ok = idx >= 0 && sHas[m];
    const float4 dc = *reinterpret_cast<const float4*>(sDC + m * CLD + q * 4);
    if (want_feat && ok) {
      const float wk = sW[m * KNN + k];
      float* dst = grad_row(d_feats, remap, idx);
      if (dst) red_add_v4(dst + q * 4, wk * dc.x, wk * dc.y, wk * dc.z, wk * dc.w);
    }
    if (want_w) {
      float part = 0.f;
      if (ok) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(feat_row(feats, leaf, remap, idx)) + q);
        part = dc.x * f.x + dc.y * f.y + dc.z * f.z + dc.w * f.w;
      }
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      if (q == 0 && ok) sDWh[m * KNN + k] += part;
    }
  }
}

__global__ void __launch_bounds__(NT, CTAS_PER_SM) render_bwd_kernel(const __grid_constant__ BwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* sD = smem;                                   // [128][DLD] (colour) / [128][CLD] view (geometry)
  float* sB = sD + TILE_M * DLD;
  float* sC = sB + SB_FLOATS;
  float* sDC = sC + TILE_M * CLD;
  float* sE2 = sDC + TILE_M * CLD;
  int* sIdx = reinterpret_cast<int*>(sE2 + E2_FLOATS);
  float* sW = reinterpret_cast<float*>(sIdx + TILE_M * KNN);
  float* sD8 = sW + TILE_M * KNN;
  float* sP = sD8 + TILE_M * KNN;
  float* sDP = sP + TILE_M * 4;
  float* sDWh = sDP + TILE_M * 4;
  int* sHas = reinterpret_cast<int*>(sDWh + TILE_M * KNN);
  float* sWsum = reinterpret_cast<float*>(sHas + TILE_M);
  float* sDOcc = sWsum + TILE_M;
  float* sDOut = sDOcc + TILE_M;                      // [m][4]
  float* sRed = sDOut + TILE_M * 4;
  float* sE = sE2;                                    // [128][ELD]
  float* sDE = sE2 + TILE_M * ELD;
  float* sQ = sE2;                                    // [128][QLD]
  float* sDQ = sE2 + TILE_M * QLD;                    // [128][DQLD]
  float* sEg = sD + TILE_M * CLD;                     // [TILE_M][GLD], behind the [TILE_M][CLD] view of sD (geometry phase)

  const int tid = threadIdx.x;
  const int S = a.prm.n_surface;
  const float* __restrict__ blob = a.w.blob;
  const float* __restrict__ packed = a.packed;
  const float* __restrict__ sv = a.saved;
  const bool color = a.stage == LSR_STAGE_COLOR;
  const bool relpos = (a.prm.flags & LSR_FLAG_REL_POS) != 0;
  const bool g_gf = (a.gflags & LSR_GRAD_GEO_FEATS) && a.d_geo;
  const bool g_cf = (a.gflags & LSR_GRAD_COL_FEATS) && a.d_col && color;
  const bool g_gw = (a.gflags & LSR_GRAD_GEO_W) && a.d_w;
  const bool g_gb = ((a.gflags & (LSR_GRAD_GEO_B | LSR_GRAD_GEO_W)) != 0) && a.d_w;
  const bool g_cw = (a.gflags & LSR_GRAD_COL_W) && a.d_w && color;
  const bool g_ry = (a.gflags & LSR_GRAD_RAYS) && a.d_ro && a.d_rd;
  const bool g_af = (a.gflags & LSR_GRAD_AFFINE) && a.d_affine && a.prm.rgb_mode == LSR_RGB_AFFINE_SIGMOID;
  const bool trk = a.is_tracker && g_ry;              // neighbour weights depend on p
  const SavedLayout SL = saved_layout(a.R, S, a.stage, a.prm.flags);
  const size_t Pp = SL.Pp;
  const WideMap wm;
  const NarrowMap nm;
  float* __restrict__ dW = a.d_w;

  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int r0 = tile * a.rays_per_tile;
    const int nr = min(a.rays_per_tile, a.R - r0);
    const int nrows = nr * S;
    const size_t p0 = (size_t)r0 * S;

    LSR_PHASE_BEGIN();
    // ------------------------------------------------------------ 0: per-row state
    if (tid < TILE_M) {
      const int m = tid;
      const bool rv = m < nrows;
      float4 mi = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rv) mi = reinterpret_cast<const float4*>(sv + SL.misc)[p0 + m];
#pragma unroll
      for (int k = 0; k < KNN; ++k) {
        sIdx[m * KNN + k] = rv ? reinterpret_cast<const int*>(sv + SL.idx)[(p0 + m) * KNN + k] : -1;
        sW[m * KNN + k] = rv ? sv[SL.w + (p0 + m) * KNN + k] : 0.f;
        sD8[m * KNN + k] = rv ? sv[SL.D + (p0 + m) * KNN + k] : FLT_MAX;
        sDWh[m * KNN + k] = 0.f;
      }
      float px = 0.f, py = 0.f, pz = 0.f;
      if (rv) {
        const int ray = r0 + m / S;
        px = __fadd_rn(a.rays_o[3 * ray + 0], __fmul_rn(a.rays_d[3 * ray + 0], mi.x));
        py = __fadd_rn(a.rays_o[3 * ray + 1], __fmul_rn(a.rays_d[3 * ray + 1], mi.x));
        pz = __fadd_rn(a.rays_o[3 * ray + 2], __fmul_rn(a.rays_d[3 * ray + 2], mi.x));
      }
      sP[m * 4 + 0] = px; sP[m * 4 + 1] = py; sP[m * 4 + 2] = pz; sP[m * 4 + 3] = mi.x;
      sHas[m] = (rv && mi.y > 0.5f) ? 1 : 0;
      sWsum[m] = mi.z;
      sDP[m * 4 + 0] = 0.f; sDP[m * 4 + 1] = 0.f; sDP[m * 4 + 2] = 0.f; sDP[m * 4 + 3] = 0.f;
      sDOcc[m] = 0.f;
      sDOut[m * 4 + 0] = 0.f; sDOut[m * 4 + 1] = 0.f; sDOut[m * 4 + 2] = 0.f; sDOut[m * 4 + 3] = 0.f;
    }
    for (int i = tid; i < RED_FLOATS; i += NT) sRed[i] = 0.f;
    __syncthreads();

    // ------------------------------------------------------------ 1: compositing backward
    // (only the output-layer gradients of a TRAINABLE geometry decoder still need dL/d(occupancy logit) here: the colour head
    //  and both MLP chains take it inside their tensor-core kernels)
    if (g_gw && tid < nr)
      composite_bwd_ray(a.prm, color, sv, SL, p0 + (size_t)tid * S, r0 + tid, a.gt_depth, a.g_depth, a.g_var, a.g_rgb, sDOcc + tid * S,
                        sDOut + tid * S * 4);
    __syncthreads();

    LSR_PHASE(1, 0);   // load state + compositing backward
    if (color) {
      // The colour head + trunk backward ran on the tensor cores (lsr_render_bwd_umma.cu); it left dL/dc and the
      // colour-Fourier part of dL/dp per sample row.
      for (int it = tid; it < TILE_M * 8; it += NT) {
        const int m = it >> 3, q = it & 7;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!relpos && m < nrows && sHas[m]) v = reinterpret_cast<const float4*>(a.ext_dc)[(p0 + m) * 8 + q];   // rel-pos: scattered by the trunk kernel
        *reinterpret_cast<float4*>(sDC + m * CLD + q * 4) = v;
      }
      if (g_ry && tid < nrows) {
        const float4 v = reinterpret_cast<const float4*>(a.ext_dp)[p0 + tid];
        sDP[tid * 4 + 0] += v.x; sDP[tid * 4 + 1] += v.y; sDP[tid * 4 + 2] += v.z;
      }
      __syncthreads();

      LSR_PHASE(1, 3);   // fourier bwd + dC
      if (relpos) {
        // the rel-pos neighbour MLP backward (feature scatter, dV1 / dV2 / dB_rel, pose part) ran in the tcgen05 kernel;
        // what is left for the shared IDW-weight backward below is its d(w_hat) per neighbour
        if (trk) {
          for (int i = tid; i < nrows * KNN; i += NT) sDWh[i] += a.ext_dwh[p0 * KNN + i];
        }
      } else {
        scatter_idw(sDC, sIdx, sW, sHas, sDWh, a.col_feats, a.col_leaf, a.remap, a.d_col, g_cf, trk);
      }
      __syncthreads();
    }

    LSR_PHASE(1, 4);   // rel-pos backward / colour scatter
    // -------------------------------------------------------------- geometry MLP backward
    {
      if (g_gw) {
        for (int it = tid; it < TILE_M * 8; it += NT) {
          const int m = it >> 3, q = it & 7;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m < nrows) v = reinterpret_cast<const float4*>(sv + SL.cg)[(p0 + m) * 8 + q];
          *reinterpret_cast<float4*>(sC + m * CLD + q * 4) = v;
        }
        for (int it = tid; it < TILE_M * EGP; it += NT) {
          const int m = it / EGP, j = it - m * EGP;
          float v = 0.f;
          if (j < EG) {
            const float t0 = TWO_PI_F * sP[m * 4 + 0], t1 = TWO_PI_F * sP[m * 4 + 1], t2 = TWO_PI_F * sP[m * 4 + 2];
            v = sin_ff(fmaf(t2, packed[Packed::gB + 2 * EGP + j], fmaf(t1, packed[Packed::gB + EGP + j], t0 * packed[Packed::gB + j])));
          }
          sEg[m * GLD + j] = v;
        }
      }
      __syncthreads();
      if (g_gw) {
        if (tid < HG) {
          float s = 0.f;
          for (int m = 0; m < nrows; ++m) s = fmaf(sDOcc[m], sv[SL.gh + ((size_t)4 * Pp + p0 + m) * HG + tid], s);
          atomicAdd(dW + a.w.g_out_w + tid, s);
        } else if (tid == HG) {
          float s = 0.f;
          for (int m = 0; m < nrows; ++m) s += sDOcc[m];
          atomicAdd(dW + a.w.g_out_b, s);
        }
      }
      // The dX chain (dH_l, dL/dc^g, dL/de) ran on the tensor cores (geo_bwd_umma_kernel); pick up its results.
      float dCacc[TMNA][4], dEacc[TMA][8];
#pragma unroll
      for (int i = 0; i < TMNA; ++i) {
        const int r = nm.row(i);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nrows) v = *reinterpret_cast<const float4*>(a.ext_gdc + (p0 + r) * CDIM + nm.col());
        dCacc[i][0] = v.x; dCacc[i][1] = v.y; dCacc[i][2] = v.z; dCacc[i][3] = v.w;
      }
      zero_acc(dEacc);
      if (g_gb || g_ry) {
#pragma unroll
        for (int i = 0; i < TMA; ++i) {
          const int r = wm.row(i);
          if (r < nrows) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              if (wm.col(g) < EGP) {
                const float4 v = *reinterpret_cast<const float4*>(a.ext_gde + (p0 + r) * EGP + wm.col(g));
                dEacc[i][g * 4 + 0] = v.x; dEacc[i][g * 4 + 1] = v.y; dEacc[i][g * 4 + 2] = v.z; dEacc[i][g * 4 + 3] = v.w;
              }
            }
          }
        }
      }
      if (g_gw) {
        // geometry decoder weight gradients (only when mapping.fix_geo_decoder is off): per layer dU = dH^T c, dW = dA^T h_prev
#pragma unroll 1
        for (int li = 4; li >= 0; --li) {
          float dHg[TMNA][4];
#pragma unroll
          for (int i = 0; i < TMNA; ++i) {
            const int r = nm.row(i);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nrows) v = *reinterpret_cast<const float4*>(a.ext_gdh + ((size_t)li * Pp + p0 + r) * HG + nm.col());
            dHg[i][0] = v.x; dHg[i][1] = v.y; dHg[i][2] = v.z; dHg[i][3] = v.w;
            *reinterpret_cast<float4*>(sD + r * CLD + nm.col()) = v;
          }
          __syncthreads();
          if (tid < HG) {
            float s = 0.f;
            for (int m = 0; m < nrows; ++m) s += sD[m * CLD + tid];
            atomicAdd(dW + a.w.g_fc_b[li] + tid, s);
          }
          {
            float au[TMN32][4];
            zero_acc(au);
            tile_gemm<TMN32, 8, 1, false, true>(au, sD, CLD, nrows, sC, CLD, CDIM, sB);
#pragma unroll
            for (int i = 0; i < TMN32; ++i)
              red_add_v4(dW + a.w.g_fc_w[li] + nm.row(i) * CDIM + nm.col(), au[i][0], au[i][1], au[i][2], au[i][3]);
          }
#pragma unroll
          for (int i = 0; i < TMNA; ++i) {
            const int r = nm.row(i);
            float4 dA = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nrows) {
              const float4 s4 = *reinterpret_cast<const float4*>(sv + SL.gs + ((size_t)li * Pp + p0 + r) * HG + nm.col());
              dA = make_float4(s4.x > 0.f ? dHg[i][0] : 0.f, s4.y > 0.f ? dHg[i][1] : 0.f, s4.z > 0.f ? dHg[i][2] : 0.f,
                               s4.w > 0.f ? dHg[i][3] : 0.f);
            }
            *reinterpret_cast<float4*>(sD + r * CLD + nm.col()) = dA;
          }
          __syncthreads();
          if (tid < HG) {
            float s = 0.f;
            for (int m = 0; m < nrows; ++m) s += sD[m * CLD + tid];
            atomicAdd(dW + a.w.g_lin_b[li] + tid, s);
          }
          if (li == 1 || li == 2 || li == 4) {
            float aw[TMN32][4];
            zero_acc(aw);
            tile_gemm<TMN32, 8, 1, false, false>(aw, sD, CLD, nrows, sv + SL.gh + ((size_t)(li - 1) * Pp + p0) * HG, HG, HG, sB);
#pragma unroll
            for (int i = 0; i < TMN32; ++i)
              red_add_v4(dW + a.w.g_lin_w[li] + nm.row(i) * HG + nm.col(), aw[i][0], aw[i][1], aw[i][2], aw[i][3]);
          } else {
            if (li == 3) {
              float aw[TMN32][4];
              zero_acc(aw);
              tile_gemm<TMN32, 8, 1, false, false>(aw, sD, CLD, nrows, sv + SL.gh + ((size_t)2 * Pp + p0) * HG, HG, HG, sB);
#pragma unroll
              for (int i = 0; i < TMN32; ++i) {
                float* d = dW + a.w.g_lin_w[3] + nm.row(i) * (EG + HG) + EG + nm.col();   // 125-float rows: unaligned
                atomicAdd(d + 0, aw[i][0]); atomicAdd(d + 1, aw[i][1]); atomicAdd(d + 2, aw[i][2]); atomicAdd(d + 3, aw[i][3]);
              }
            }
            float ae2[TMW32][8];
            zero_acc(ae2);
            tile_gemm<TMW32, 16, 2, false, true>(ae2, sD, CLD, nrows, sEg, GLD, EGP, sB);
            const int ldw = li == 3 ? EG + HG : EG;
#pragma unroll
            for (int g = 0; g < 2; ++g)
#pragma unroll
              for (int i = 0; i < TMW32; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int col = wm.col(g) + j;
                  if (col < EG) atomicAdd(dW + a.w.g_lin_w[li] + wm.row(i) * ldw + col, ae2[i][g * 4 + j]);
                }
          }
          __syncthreads();
        }
      }
      // geometry Fourier backward: e_j = sin(arg_j)
      if (g_gb || g_ry) {
        float dpr[TMA][3];
#pragma unroll
        for (int i = 0; i < TMA; ++i) { dpr[i][0] = 0.f; dpr[i][1] = 0.f; dpr[i][2] = 0.f; }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = wm.col(g) + j;
            if (col < EG) {
              const float b0 = packed[Packed::gB + col], b1 = packed[Packed::gB + EGP + col],
                          b2 = packed[Packed::gB + 2 * EGP + col];
              float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
              for (int i = 0; i < TMA; ++i) {
                const int r = wm.row(i);
                if (r < nrows) {
                  const float t0 = TWO_PI_F * sP[r * 4 + 0], t1 = TWO_PI_F * sP[r * 4 + 1], t2 = TWO_PI_F * sP[r * 4 + 2];
                  const float dar = dEacc[i][g * 4 + j] * cos_ff(fmaf(t2, b2, fmaf(t1, b1, t0 * b0)));
                  s0 = fmaf(t0, dar, s0); s1 = fmaf(t1, dar, s1); s2 = fmaf(t2, dar, s2);
                  dpr[i][0] = fmaf(b0, dar, dpr[i][0]); dpr[i][1] = fmaf(b1, dar, dpr[i][1]);
                  dpr[i][2] = fmaf(b2, dar, dpr[i][2]);
                }
              }
              if (g_gb) {
                atomicAdd(&sRed[col], s0); atomicAdd(&sRed[EGP + col], s1); atomicAdd(&sRed[2 * EGP + col], s2);
              }
            }
          }
        }
        if (g_ry) {
#pragma unroll
          for (int i = 0; i < TMA; ++i) {
            const float v0 = half_warp_sum(dpr[i][0]), v1 = half_warp_sum(dpr[i][1]), v2 = half_warp_sum(dpr[i][2]);
            if (wm.tx == 0) {
              const int r = wm.row(i);
              sDP[r * 4 + 0] += TWO_PI_F * v0; sDP[r * 4 + 1] += TWO_PI_F * v1; sDP[r * 4 + 2] += TWO_PI_F * v2;
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < TMNA; ++i) {
        const int r = nm.row(i);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sHas[r]) v = make_float4(dCacc[i][0], dCacc[i][1], dCacc[i][2], dCacc[i][3]);
        *reinterpret_cast<float4*>(sDC + r * CLD + nm.col()) = v;
      }
      __syncthreads();
      scatter_idw(sDC, sIdx, sW, sHas, sDWh, a.geo_feats, a.geo_leaf, a.remap, a.d_geo, g_gf, trk);
      __syncthreads();
    }

    LSR_PHASE(1, 5);   // geometry backward + scatter
    // -------------------------------------------------------------- IDW weight backward (tracker)
    if (trk && tid < TILE_M && sHas[tid]) {
      const int m = tid;
      float wr[KNN], Wt = 0.f, dot = 0.f;
#pragma unroll
      for (int k = 0; k < KNN; ++k) {
        wr[k] = sIdx[m * KNN + k] >= 0 ? 1.0f / (sD8[m * KNN + k] + 1e-10f) : 0.f;
        Wt += wr[k];
        dot = fmaf(sDWh[m * KNN + k], sW[m * KNN + k], dot);
      }
      const float denom = fmaxf(Wt, 1e-12f);
      if (!(Wt > 1e-12f)) dot = 0.f;
      float q0 = 0.f, q1 = 0.f, q2 = 0.f;
#pragma unroll
      for (int k = 0; k < KNN; ++k) {
        const int idx = sIdx[m * KNN + k];
        if (idx >= 0) {
          const float dw = (sDWh[m * KNN + k] - dot) / denom;
          const float dD = -wr[k] * wr[k] * dw;       // w = 1/(D+eps)
          const float e0 = __ldg(a.cloud + 3 * (size_t)idx + 0) - sP[m * 4 + 0];
          const float e1 = __ldg(a.cloud + 3 * (size_t)idx + 1) - sP[m * 4 + 1];
          const float e2 = __ldg(a.cloud + 3 * (size_t)idx + 2) - sP[m * 4 + 2];
          q0 = fmaf(-2.f * e0, dD, q0); q1 = fmaf(-2.f * e1, dD, q1); q2 = fmaf(-2.f * e2, dD, q2);
        }
      }
      sDP[m * 4 + 0] += q0; sDP[m * 4 + 1] += q1; sDP[m * 4 + 2] += q2;
    }
    __syncthreads();
    // -------------------------------------------------------------- ray gradients + parameter reductions
    if (g_ry && tid < nr) {
      float o0 = 0.f, o1 = 0.f, o2 = 0.f, e0 = 0.f, e1 = 0.f, e2 = 0.f;
      for (int s = 0; s < S; ++s) {
        const int m = tid * S + s;
        const float z = sP[m * 4 + 3];
        o0 += sDP[m * 4 + 0]; o1 += sDP[m * 4 + 1]; o2 += sDP[m * 4 + 2];
        e0 = fmaf(z, sDP[m * 4 + 0], e0); e1 = fmaf(z, sDP[m * 4 + 1], e1); e2 = fmaf(z, sDP[m * 4 + 2], e2);
      }
      const int ray = r0 + tid;
      a.d_ro[3 * ray + 0] = o0; a.d_ro[3 * ray + 1] = o1; a.d_ro[3 * ray + 2] = o2;
      a.d_rd[3 * ray + 0] = e0; a.d_rd[3 * ray + 1] = e1; a.d_rd[3 * ray + 2] = e2;
    }
    if (g_gb) {
      for (int i = tid; i < 3 * EGP; i += NT) {
        const int c = i / EGP, j = i - c * EGP;
        if (j < EG) atomicAdd(dW + a.w.g_B + c * EG + j, sRed[i]);
      }
    }
    __syncthreads();
  }
}

int check_weights(const LsrWeights* w);
int check_params(const LsrParams* p);
int sm_count();
int balanced_rays_per_tile(int64_t n_rays, int n_surface, int nsm);
int launch_geo_bwd(const LsrParams* prm, const LsrWeights* w, const float* gt_depth, int64_t n_rays, int stage, const void* saved,
                   float* gpack, const float* g_depth, const float* g_var, const float* g_rgb, int grad_flags, float* gdc, float* gde,
                   float* gdh, cudaStream_t stream);
int launch_trunk_bwd(const LsrParams* prm, const LsrWeights* w, const float* gt_depth, int64_t n_rays, const float* affine,
                     const void* saved, void* scratch, const float* g_depth, const float* g_var, const float* g_rgb, int grad_flags,
                     float* d_weights, float* d_affine, const float* cloud_pos, const int32_t* row_remap, float* d_col_feats,
                     int is_tracker, cudaStream_t stream, int phase, cudaEvent_t ev_after_trunk);

// Side stream of the colour-stage backward.  trunk_bwd_umma_kernel keeps one CTA per SM and 45 of the 148 SMs get a second
// tile, so the other SMs idle for the second half of it; the geometry chain (independent of the colour trunk) is enqueued on a
// second stream right behind it and fills those SMs, and the finalize kernel runs beside the scatter kernel.
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t e_fork = nullptr, e_geo = nullptr, e_trunk = nullptr, e_fin = nullptr;
  int state = 0;   // 0: not created, 1: ok, -1: creation failed (fall back to one stream)
  std::mutex mu;   // the fork / join events are reused: enqueueing a backward is atomic per device
};
static SideStream* side_stream() {
  static SideStream side[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& S = side[dev];
  std::lock_guard<std::mutex> lock(S.mu);
  if (S.state == 0) {
    bool ok = cudaStreamCreateWithFlags(&S.s, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&S.e_fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&S.e_geo, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&S.e_trunk, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&S.e_fin, cudaEventDisableTiming) == cudaSuccess;
    S.state = ok ? 1 : -1;
  }
  return S.state == 1 ? &S : nullptr;
}

}  // namespace lsr

using namespace lsr;

#ifdef LSR_PHASE_TIMING
extern "C" int lsr_debug_phase_cycles(unsigned long long* out32) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out32, lsr_phase_cycles, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : 3;
}
#endif

extern "C" int lsr_render_bwd(const LsrParams* prm, const void* grid_ws, const float* cloud_pos, int64_t n_points,
                              const float* rays_o, const float* rays_d, const float* gt_depth,
                              const double* r_query, int64_t n_rays, const float* geo_feats,
                              const float* col_feats, const int32_t* row_remap, const float* geo_leaf,
                              const float* col_leaf, const LsrWeights* w, const float* exposure_affine, int stage,
                              int is_tracker, const void* saved, void* scratch, const float* g_depth,
                              const float* g_var, const float* g_rgb, int grad_flags, float* d_geo_feats,
                              float* d_col_feats, float* d_weights, float* d_exposure_affine, float* d_rays_o,
                              float* d_rays_d, lsr_stream_t stream) {
  (void)grid_ws; (void)r_query;
  int rc = check_params(prm);
  if (rc) return rc;
  rc = check_weights(w);
  if (rc) return rc;
  if (stage != LSR_STAGE_GEOMETRY && stage != LSR_STAGE_COLOR) return LSR_ERR_ARG;
  if (n_rays < 0 || n_rays > (1ll << 27) || n_points < 0) return LSR_ERR_ARG;
  if (n_rays == 0) return LSR_OK;
  if (!saved || !scratch || !rays_o || !rays_d || !gt_depth || !g_depth) return LSR_ERR_ARG;
  if (prm->flags & (LSR_FLAG_SAVE_LIGHT | LSR_FLAG_FWD_ONLY)) return LSR_ERR_ARG;   // no activations / no backward scratch to differentiate with
  if (n_points > 0 && (!cloud_pos || !geo_feats)) return LSR_ERR_ARG;
  if (n_points > 0 && stage == LSR_STAGE_COLOR && !col_feats) return LSR_ERR_ARG;
  if (row_remap && (!geo_leaf || (stage == LSR_STAGE_COLOR && !col_leaf))) return LSR_ERR_ARG;
  if (prm->rgb_mode == LSR_RGB_AFFINE_SIGMOID && !exposure_affine) return LSR_ERR_ARG;
  if ((grad_flags & LSR_GRAD_GEO_FEATS) && !d_geo_feats) return LSR_ERR_ARG;
  if ((grad_flags & LSR_GRAD_COL_FEATS) && !d_col_feats) return LSR_ERR_ARG;
  if ((grad_flags & (LSR_GRAD_GEO_W | LSR_GRAD_GEO_B | LSR_GRAD_COL_W)) && !d_weights) return LSR_ERR_ARG;
  if ((grad_flags & LSR_GRAD_RAYS) && (!d_rays_o || !d_rays_d)) return LSR_ERR_ARG;
  if ((grad_flags & LSR_GRAD_AFFINE) && !d_exposure_affine) return LSR_ERR_ARG;
  const int nsm = sm_count();
  if (nsm <= 0) return LSR_ERR_CUDA;

  const ScratchLayout CL = scratch_layout(n_rays, prm->n_surface);
  SideStream* side = stage == LSR_STAGE_COLOR ? side_stream() : nullptr;
  std::unique_lock<std::mutex> side_lock;
  if (side) side_lock = std::unique_lock<std::mutex>(side->mu);
  if (side) {   // fork: the side stream starts behind everything already enqueued on the caller's stream
    LSR_CUDA_CHECK(cudaEventRecord(side->e_fork, stream));
    LSR_CUDA_CHECK(cudaStreamWaitEvent(side->s, side->e_fork, 0));
  }
  // With the rel-pos decoder and no pose gradient the scatter kernel needs nothing from the colour trunk (its feature scatter and
  // Fourier pose terms happen in the trunk kernel): the whole geometry side -- chain + scatter kernel -- runs on the side stream.
  const bool geo_side_only = side && (prm->flags & LSR_FLAG_REL_POS) && !(grad_flags & LSR_GRAD_RAYS) && !is_tracker;
  if (stage == LSR_STAGE_COLOR) {   // trunk kernel, rel-pos trig kernel (+ finalize when there is no side stream)
    rc = launch_trunk_bwd(prm, w, gt_depth, n_rays, exposure_affine, saved, scratch, g_depth, g_var, g_rgb, grad_flags, d_weights,
                          d_exposure_affine, cloud_pos, row_remap, d_col_feats, is_tracker, stream, side ? 1 : 0,
                          side ? side->e_trunk : nullptr);
    if (rc) return rc;
  }
  {
    const bool need_e = (grad_flags & (LSR_GRAD_GEO_B | LSR_GRAD_GEO_W | LSR_GRAD_RAYS)) != 0;
    rc = launch_geo_bwd(prm, w, gt_depth, n_rays, stage, saved, (float*)((char*)scratch + CL.bwd_gpack), g_depth, g_var, g_rgb,
                        grad_flags, (float*)((char*)scratch + CL.bwd_gdc), need_e ? (float*)((char*)scratch + CL.bwd_gde) : nullptr,
                        (grad_flags & LSR_GRAD_GEO_W) ? (float*)((char*)scratch + CL.bwd_gdh) : nullptr, side ? side->s : stream);
    if (rc) return rc;
    if (side && !geo_side_only) {
      LSR_CUDA_CHECK(cudaEventRecord(side->e_geo, side->s));
      LSR_CUDA_CHECK(cudaStreamWaitEvent(stream, side->e_geo, 0));          // the scatter kernel below needs the geometry planes
    }
  }
  cudaStream_t scatter_stream = geo_side_only ? side->s : stream;
  BwdArgs a;
  a.ext_gdc = (const float*)((const char*)scratch + CL.bwd_gdc);
  a.ext_gde = (const float*)((const char*)scratch + CL.bwd_gde);
  a.ext_gdh = (const float*)((const char*)scratch + CL.bwd_gdh);
  a.ext_dc = (const float*)((const char*)scratch + CL.bwd_dc);
  a.ext_dp = (const float*)((const char*)scratch + CL.bwd_dp);
  a.ext_dwh = (const float*)((const char*)scratch + CL.bwd_dwh);
  a.prm = *prm;
  a.cloud = cloud_pos;
  a.rays_o = rays_o; a.rays_d = rays_d; a.gt_depth = gt_depth;
  a.R = (int)n_rays;
  a.geo_feats = geo_feats; a.col_feats = col_feats;
  a.remap = row_remap; a.geo_leaf = geo_leaf; a.col_leaf = col_leaf;
  a.w = *w;
  a.packed = (const float*)scratch;
  a.affine = exposure_affine;
  a.stage = stage; a.is_tracker = is_tracker;
  a.saved = (const float*)saved;
  a.g_depth = g_depth; a.g_var = g_var; a.g_rgb = g_rgb;
  a.gflags = grad_flags;
  a.d_geo = d_geo_feats; a.d_col = d_col_feats; a.d_w = d_weights; a.d_affine = d_exposure_affine;
  a.d_ro = d_rays_o; a.d_rd = d_rays_d;
  a.rays_per_tile = balanced_rays_per_tile(n_rays, prm->n_surface, nsm);
  a.ntiles = (int)((n_rays + a.rays_per_tile - 1) / a.rays_per_tile);
  const size_t smem = BWD_SMEM_FLOATS * sizeof(float);
  LSR_SMEM_ATTR_ONCE(render_bwd_kernel, smem);
  const int slots = nsm * CTAS_PER_SM;
  const int grid = a.ntiles < slots ? a.ntiles : slots;
  render_bwd_kernel<<<grid, NT, smem, scatter_stream>>>(a);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  if (side) {   // finalize: behind the trunk kernel, beside the rel-pos trig kernel / the scatter kernel; then join
    LSR_CUDA_CHECK(cudaStreamWaitEvent(side->s, side->e_trunk, 0));
    rc = launch_trunk_bwd(prm, w, gt_depth, n_rays, exposure_affine, saved, scratch, g_depth, g_var, g_rgb, grad_flags, d_weights,
                          d_exposure_affine, cloud_pos, row_remap, d_col_feats, is_tracker, side->s, 2, nullptr);
    if (rc) return rc;
    LSR_CUDA_CHECK(cudaEventRecord(side->e_fin, side->s));
    LSR_CUDA_CHECK(cudaStreamWaitEvent(stream, side->e_fin, 0));   // everything is ordered on the caller's stream again
  }
  return LSR_OK;
}
