// Backward of the colour trunk (MLP_color.forward, /root/reference/src/conv_onet/models/decoder.py:515-546, differentiated by
// loss.backward() at src/Mapper.py:722 / src/Tracker.py:193) on the 5th-generation tensor cores.
//
// Per 128-row tile (the forward's tile), layer l = 4 .. 0, with  Z_l = dL/dh_l * softplus'(pre_l)  (128 rows x 128):
//
//   dX  GEMM  [G_{l-1} | dE' | dC] = Z_l . [W_l^h | W_l^e | P_l]        contraction over the OUTPUT FEATURES of layer l
//             A = Z_l (hi, lo) in TMEM (lane = row), B = transposed weight chunks streamed L2 -> shared memory.
//             P_l = W_l^h . U_{l-1} folds the fc_c skip of the layer below into the same GEMM
//             (dC = sum_l G_l U_l = G_4 U_4 + sum_l Z_{l+1} (W_{l+1}^h U_l)).
//   dW  GEMMs dW_l^h ^T = h_{l-1}^T . Z_l   and   [E_l | M_l] = Z_l^T . [e' | c | 1]   contraction over the ROWS
//             both operands are 128-byte-swizzled K-major images with K = rows: Z_l^T is written by the row-owning
//             threads (4-byte stores, one conflict-free 128-byte line per warp instruction), h_{l-1}^T / e'^T / [c|1]^T
//             are the forward's T-planes (lsr_render.cuh) bulk-copied as they lie.  M_l carries the bias gradients
//             (ones column) and, through  dU_{l-1} = W_l^h^T (Z_l^T c),  the fc_c weight gradients: a tiny finalize
//             kernel applies W^T once per call instead of a second row-contracted GEMM per layer and tile.
//
// Every product is an error-compensated 3xTF32 sum (hi.lo + lo.hi + hi.hi, fp32 accumulate in TMEM).  Operands
// arrive in shared memory as plain fp32 (half the L2 traffic of pre-split copies) and are split into (hi, lo) IN
// PLACE by two converter warps before the issuing warp may read them.
//
// Roles of the 640-thread CTA: 16 compute / epilogue warps (TMEM lane quadrant x 32-column group), 1 issuer,
// 1 producer (cp.async.bulk), 2 converters.  TMEM: [0,256) Z hi | lo, [256,384) G / dW^h^T, [384,464) extras.
// Measured building blocks: tools/umma_sw128_probe.cu (swizzled K = rows operands: exact),
// tools/umma_bwd_probe.cu (TS chain + row-contracted GEMMs vs fp64: 2e-6), tools/umma_rate_probe.cu (issue cost).
#include <cstring>
#include "lsr_render.cuh"
#include "lsr_umma_prog.cuh"

namespace lsr {

using namespace umma;

// ------------------------------------------------------------------------------------------------ packed weights
// B-operand chunk format (canonical no-swizzle K-major, fp32): chunk of kc contraction values, element (n, k) at
// ((k / 4) * n_pad + n) * 4 + k % 4 floats; all chunks of a job but the last hold `kc` values.
struct BJob {
  int32_t type;        // 0: B[n][k] = blob[w + k * ld + col0 + n]      1: B[n][k] = sum_i blob[w + k * ld + col0 + i] * blob[u + i * 32 + n]
  int32_t w, ld, col0, u;
  int32_t n0, n_valid, n_pad;   // rows [n0, n0 + n_valid) of an n_pad-row operand
  int32_t k_total, kc;
  int32_t dst;         // float offset in the backward pack buffer
};
constexpr int MAX_BJOBS = 16;
struct BJobs { BJob j[MAX_BJOBS]; int n; int pout_dst; int w_out, u4; };

__global__ void bwd_prep_kernel(const float* __restrict__ blob, float* __restrict__ pack, const __grid_constant__ BJobs jobs) {
  if ((int)blockIdx.y == jobs.n) {   // P_out[o][c] = sum_i W_out[o][i] U_4[i][c]   (3 x 32)
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 3 * CDIM; e += gridDim.x * blockDim.x) {
      const int o = e / CDIM, c = e % CDIM;
      float s = 0.f;
      for (int i = 0; i < HC; ++i) s = fmaf(blob[jobs.w_out + o * HC + i], blob[jobs.u4 + i * CDIM + c], s);
      pack[jobs.pout_dst + e] = s;
    }
    return;
  }
  const BJob jb = jobs.j[blockIdx.y];
  auto put = [&](int n, int k, float v) {
    const int c = k / jb.kc, kl = k - c * jb.kc;
    pack[jb.dst + (size_t)c * jb.n_pad * jb.kc + ((kl >> 2) * jb.n_pad + jb.n0 + n) * 4 + (kl & 3)] = v;
  };
  if (jb.type == 0) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < jb.n_valid * jb.k_total; e += gridDim.x * blockDim.x) {
      const int n = e % jb.n_valid, k = e / jb.n_valid;
      put(n, k, blob[jb.w + (size_t)k * jb.ld + jb.col0 + n]);
    }
  } else {
    // P[n = c][k = o] = sum_i W[o][col0 + i] U[i][c]: block handles output features k = blockIdx.x, +gridDim.x, ...;
    // 256 threads = 32 c x 8 slices of i, the W row staged in shared memory
    __shared__ float wrow[HC];
    __shared__ float part[8][CDIM];
    const int c = threadIdx.x & 31, sl = threadIdx.x >> 5;
    for (int k = blockIdx.x; k < jb.k_total; k += gridDim.x) {
      if (threadIdx.x < HC) wrow[threadIdx.x] = blob[jb.w + (size_t)k * jb.ld + jb.col0 + threadIdx.x];
      __syncthreads();
      float v = 0.f;
#pragma unroll
      for (int i = sl * 16; i < sl * 16 + 16; ++i) v = fmaf(wrow[i], blob[jb.u + i * CDIM + c], v);
      part[sl][c] = v;
      __syncthreads();
      if (sl == 0) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += part[j][c];
        put(c, k, t);
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------ ring program
enum { K_DXW = 0, K_DXE = 1, K_DWH = 2, K_DWE = 3, K_DV1 = 4, K_DV2 = 5 };
enum { RF_WAIT_A = 1, RF_WAIT_B = 2, RF_COMMIT_D0 = 4, RF_COMMIT_D1 = 8, RF_FIRST = 16 };
struct ROp {             // 32 bytes; one ring chunk = one bulk-copied (two-part) fp32 block + the MMAs that consume it
  uint32_t src0, src1;   // float offsets: weights -> backward pack buffer; T-planes -> saved buffer (+ tile * stride)
  uint32_t stride0, stride1;   // per-tile stride in floats (0 for weights)
  uint16_t bytes0, bytes1;     // part 1 lands right behind part 0
  uint8_t kind, flags, nk8, q; // q: atom column (row-contracted kinds) / first TMEM A column / 8 (dX kinds)
  uint16_t n;            // N of the MMAs
  uint16_t pad;
  uint32_t pad2;
};
static_assert(sizeof(ROp) == 32, "ROp");
constexpr int MAX_ROPS = 128;

constexpr int BNS = 2;                                   // ring stages
constexpr int BCW = 16;                                  // compute warps
constexpr int BNCT = BCW * 32;
constexpr int BCONV = 2;                                 // converter warps
constexpr int BT = BNCT + 64 + 32 * BCONV;               // + issuer + producer + converters
constexpr int W_ISSUER = BCW, W_PRODUCER = BCW + 1, W_CONV0 = BCW + 2;
// shared memory (bytes from the 1024-aligned base)
constexpr int SB_RING = 0;
constexpr int SB_ZT = SB_RING + BNS * UM_STAGE_BYTES;    // Z^T: hi 4 atom columns x 16 KB, then lo
constexpr int ZT_ATOM = HC * 128;                        // 16 KB: [128 features][32 rows x 4 B]
constexpr int ZT_LO = 4 * ZT_ATOM;
constexpr int SB_DOUT = SB_ZT + 2 * ZT_LO;               // [128][4] d(colour logits)
constexpr int SB_HAS = SB_DOUT + 128 * 16;
constexpr int SB_DP = SB_HAS + 128 * 4;                  // [128][4]
constexpr int SB_DWH = SB_DP + 128 * 16;                 // [128][8] d(normalised IDW weight) of the rel-pos path (tracker)
constexpr int SB_BREL = SB_DWH + 128 * KNN * 4;          // [32] per-tile partial of d B_rel
constexpr int SB_TAB = SB_BREL + 32 * 4;           // W_out [3][128] | P_out [3][32] | c_B [3][20] | B_rel | v2
constexpr int TAB_WOUT = 0, TAB_POUT = 3 * HC, TAB_CB = TAB_POUT + 3 * CDIM, TAB_BREL = TAB_CB + 3 * EC + 4, TAB_V2B = TAB_BREL + 32,
              TAB_TOTAL = TAB_V2B + CDIM;   // ... | rel-pos Fourier matrix [3][10] | v2 bias [32]
constexpr int SB_PIPE = SB_TAB + TAB_TOTAL * 4;
constexpr int BWD_UMMA_SMEM = SB_PIPE + 256;
static_assert(2 * BNCT == 128 * KNN, "sDWh zeroing");
static_assert(BWD_UMMA_SMEM <= 232448, "shared memory budget");
static_assert(SB_ZT % 1024 == 0 && ZT_ATOM % 1024 == 0, "swizzle atoms need 1024-byte alignment");
static_assert(HC * (HC + 4) * 4 + 80 * HC * 4 <= 2 * ZT_LO, "weight-gradient staging image fits the Z^T region");
constexpr uint32_t TMB_ZHI = 0, TMB_ZLO = 128, TMB_G = 256, TMB_V1 = 320, TMB_EX = 384;
constexpr uint32_t SW128_HIWORD = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B

struct BPipe {
  uint64_t full[BNS], conv[BNS], empty[BNS];
  uint64_t a_ready, b_ready, d_ready[2];
  uint32_t tmem_base;
  uint32_t pad;
};
static_assert(sizeof(BPipe) <= 256, "pipe");

struct TrunkArgs {
  LsrParams prm;
  LsrWeights w;
  const float* gt_depth;
  const float *g_depth, *g_var, *g_rgb;
  const float* affine;
  float* d_affine;
  int R;
  const float* saved;
  const float* pack;       // backward pack buffer
  float* acc;              // [5][80][128] + M_out [3][40]
  float* d_w;
  float *out_dc, *out_dp;  // [Pp][32], [Pp][4]
  float* out_dwh;          // [Pp][8] d(normalised IDW weight) from the rel-pos path (tracker)
  float* out_dqt;          // [tile][8][20][128] d[sin | cos] of the rel-pos Fourier features (-> relpos_trig_bwd_kernel)
  const float* cloud;      // rel-pos: neighbour positions
  const float4* knn_pos;   // forward scratch: sample positions (px, py, pz, z)
  const int32_t* remap;
  float* d_col;
  int is_tracker;
  int gflags;
  int ntiles, rays_per_tile;
  int n_ops;
  ROp ops[MAX_ROPS];
};

#ifdef LSR_TRACE
// bring-up only (tools/trace_bwd.py builds a separate .so with -DLSR_TRACE): clock64 stamps of CTA 0, first tile
__device__ long long lsr_trace[3][512];
#define TRC(role, k) do { if (blockIdx.x == 0 && tile == (int)blockIdx.x && (threadIdx.x & 31) == 0) lsr_trace[role][k] = clock64(); } while (0)
#else
#define TRC(role, k)
#endif
__device__ __forceinline__ void bar_compute_b() { asm volatile("bar.sync 1, %0;\n" ::"n"(BNCT) : "memory"); }
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;\n" ::"l"(addr), "f"(v) : "memory");
}
// single tcgen05.mma with descriptors assembled from 32-bit halves inside the asm (keeps them in the uniform datapath)
__device__ __forceinline__ void mma_ts2(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hiword, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}\n" ::"r"(d),
      "r"(a_tmem), "r"(b_lo), "r"(b_hiword), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ss2(uint32_t d, uint32_t a_lo, uint32_t a_hiword, uint32_t b_lo, uint32_t b_hiword, uint32_t idesc,
                                        uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 ad, bd;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 ad, {%1, %2};\n\t"
      "mov.b64 bd, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %5, p;\n\t}\n" ::"r"(d),
      "r"(a_lo), "r"(a_hiword), "r"(b_lo), "r"(b_hiword), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};\n" ::"r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(addr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr)
               : "memory");
}

__global__ void __launch_bounds__(BT, 1) trunk_bwd_umma_kernel(const __grid_constant__ TrunkArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  BPipe* pipe = reinterpret_cast<BPipe*>(smem + SB_PIPE);
  float* sDOut = reinterpret_cast<float*>(smem + SB_DOUT);
  int* sHas = reinterpret_cast<int*>(smem + SB_HAS);
  float* sDP = reinterpret_cast<float*>(smem + SB_DP);
  float* sTab = reinterpret_cast<float*>(smem + SB_TAB);
  float* sDWh = reinterpret_cast<float*>(smem + SB_DWH);
  float* sBrel = reinterpret_cast<float*>(smem + SB_BREL);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.prm.n_surface;
  const float* __restrict__ blob = a.w.blob;
  const float* __restrict__ sv = a.saved;
  const SavedLayout SL = saved_layout(a.R, S, LSR_STAGE_COLOR, a.prm.flags);
  const bool g_cw = (a.gflags & LSR_GRAD_COL_W) && a.d_w;
  const bool g_ry = (a.gflags & LSR_GRAD_RAYS) != 0;
  const bool relpos = (a.prm.flags & LSR_FLAG_REL_POS) != 0;
  const bool g_cf = (a.gflags & LSR_GRAD_COL_FEATS) && a.d_col;
  const bool trk = a.is_tracker && g_ry;
  const bool g_af = (a.gflags & LSR_GRAD_AFFINE) && a.d_affine && a.prm.rgb_mode == LSR_RGB_AFFINE_SIGMOID;

  for (int i = tid; i < TAB_TOTAL; i += BT) {
    float v = 0.f;
    if (i < TAB_POUT) v = blob[a.w.c_out_w + i];
    else if (i < TAB_CB) v = a.pack[BWD_PACK_FLOATS_MAX - 128 + (i - TAB_POUT)];
    else if (i < TAB_CB + 3 * EC) v = blob[a.w.c_B + (i - TAB_CB)];
    else if (i >= TAB_BREL && i < TAB_BREL + 3 * ER && (a.prm.flags & LSR_FLAG_REL_POS)) v = blob[a.w.c_Brel + (i - TAB_BREL)];
    else if (i >= TAB_V2B && (a.prm.flags & LSR_FLAG_REL_POS)) v = blob[a.w.c_nb2_b + (i - TAB_V2B)];
    sTab[i] = v;
  }
  if (warp == W_ISSUER) tmem_alloc(&pipe->tmem_base, 512);
  if (tid == 0) {
    for (int s = 0; s < BNS; ++s) { mbar_init(&pipe->full[s], 1); mbar_init(&pipe->conv[s], BCONV); mbar_init(&pipe->empty[s], 1); }
    mbar_init(&pipe->a_ready, BCW);
    mbar_init(&pipe->b_ready, BCW);
    mbar_init(&pipe->d_ready[0], 1);
    mbar_init(&pipe->d_ready[1], 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = pipe->tmem_base;

  if (warp == W_PRODUCER) {
    // ================================================================ producer: fp32 chunks -> ring
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        for (int i = 0; i < a.n_ops; ++i, ++it) {
          const uint32_t stage = it % BNS, use = it / BNS;
          if (use > 0) mbar_wait(&pipe->empty[stage], (use - 1) & 1);
          const ROp& op = a.ops[i];
          const bool tp = op.kind >= K_DWH;
          const float* s0 = (tp ? sv : a.pack) + op.src0 + (size_t)tile * op.stride0;
          uint8_t* dst = smem + SB_RING + stage * UM_STAGE_BYTES;
          mbar_arrive_expect_tx(&pipe->full[stage], (uint32_t)op.bytes0 + op.bytes1);
          bulk_g2s(dst, s0, op.bytes0, &pipe->full[stage]);
          if (op.bytes1) bulk_g2s(dst + op.bytes0, sv + op.src1 + (size_t)tile * op.stride1, op.bytes1, &pipe->full[stage]);
        }
      }
    }
  } else if (warp >= W_CONV0) {
    // ================================================================ converters: fp32 -> (hi, lo) in place
    const int cw = warp - W_CONV0;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      for (int i = 0; i < a.n_ops; ++i, ++it) {
        const uint32_t stage = it % BNS, use = it / BNS;
        mbar_wait(&pipe->full[stage], use & 1);
        const int n16 = ((int)a.ops[i].bytes0 + a.ops[i].bytes1) >> 4;
        uint8_t* base = smem + SB_RING + stage * UM_STAGE_BYTES;
#pragma unroll 4
        for (int e = cw * 32 + lane; e < n16; e += 32 * BCONV) {
          const float4 v = *reinterpret_cast<const float4*>(base + 16 * e);
          uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
          split_hi_lo(v.x, h0, l0); split_hi_lo(v.y, h1, l1); split_hi_lo(v.z, h2, l2); split_hi_lo(v.w, h3, l3);
          // the tensor core reads only the top 19 bits of an fp32 word (tools/umma_sw128_probe.cu test 4): the raw image IS
          // the hi operand, only the residuals have to be produced
          (void)h0; (void)h1; (void)h2; (void)h3;
          *reinterpret_cast<uint4*>(base + UM_STAGE_BYTES / 2 + 16 * e) = make_uint4(l0, l1, l2, l3);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pipe->conv[stage]);
        // DRAM -> L2 prefetch (LSU path: the TMA queue stays free for the ring) of what the NEXT tile of this CTA reads
        // from the saved planes -- s_l T-planes, [c|1], e' -- one request per 128-byte line, spread over the ops
        const int tt = tile + (int)gridDim.x;
        if (tt < a.ntiles) {
          constexpr int LINES = 5 * HC * 4 + (TP_C1 + ECC) * 4;
          const int per = (LINES + a.n_ops - 1) / a.n_ops;
          for (int e = i * per + cw * 32 + lane; e < min((i + 1) * per, LINES); e += 32 * BCONV) {
            const float* base;
            int ln = e;
            if (ln < 5 * HC * 4) {
              const int pl = ln / (HC * 4);
              ln -= pl * HC * 4;
              base = sv + SL.cst + ((size_t)pl * SL.ntiles + tt) * tplane_tile_floats(HC);
            } else if (ln < 5 * HC * 4 + TP_C1 * 4) {
              ln -= 5 * HC * 4;
              base = sv + SL.cc1t + (size_t)tt * tplane_tile_floats(TP_C1);
            } else {
              ln -= 5 * HC * 4 + TP_C1 * 4;
              base = sv + SL.ect + (size_t)tt * tplane_tile_floats(ECC);
            }
            prefetch_l2(base + (size_t)ln * 32);
          }
        }
      }
    }
  } else if (warp == W_ISSUER) {
    // ================================================================ issuer
    uint32_t it = 0, a_par = 0, b_par = 0;
    const uint32_t ring_addr = smem_u32(smem + SB_RING), zt_addr = smem_u32(smem + SB_ZT);
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      for (int i = 0; i < a.n_ops; ++i, ++it) {
        const uint32_t stage = it % BNS, use = it / BNS;
        const uint32_t flags = a.ops[i].flags;
        if (flags & RF_WAIT_A) { mbar_wait(&pipe->a_ready, a_par); a_par ^= 1; }
        if (flags & RF_WAIT_B) { mbar_wait(&pipe->b_ready, b_par); b_par ^= 1; }
        TRC(0, 3 * i);
        // Two-phase issue: the products that only need the RAW chunk (= its hi part) go out as soon as the bulk copy has
        // landed; the one product per K step that needs the residual image waits for the converters.
        const uint32_t kind = a.ops[i].kind, n = a.ops[i].n, q = a.ops[i].q;
        const int nk8 = a.ops[i].nk8;
        const uint32_t idesc = idesc_tf32(128, (int)n);
        const uint32_t st_hi = ring_addr + stage * UM_STAGE_BYTES, st_lo = st_hi + UM_STAGE_BYTES / 2;
        const uint32_t acc0 = (flags & RF_FIRST) ? 0u : 1u;
        mbar_wait(&pipe->full[stage], use & 1);
        tc_fence_after();
        TRC(0, 3 * i + 1);
        if (kind <= K_DXE) {
          // D[rows][n] (+)= Z[rows][K chunk] . B[n][K chunk]^T ; A from TMEM, B = ring chunk (no swizzle, LBO = n * 16)
          const uint32_t lbo_word = ((n * 16u) >> 4) << 16, step = (2u * n * 16u) >> 4;
          const uint32_t d = tb + (kind == K_DXW ? TMB_G : TMB_EX);
          if (elect_one()) {
            uint32_t bh = lbo_word | ((st_hi >> 4) & 0x3fffu);
            uint32_t ah = tb + TMB_ZHI + q * 8u, al = tb + TMB_ZLO + q * 8u;
            uint32_t acc = acc0;
#pragma unroll 4
            for (int k8 = 0; k8 < nk8; ++k8) {
              mma_ts2(d, al, bh, UM_DESC_HIWORD, idesc, acc);
              mma_ts2(d, ah, bh, UM_DESC_HIWORD, idesc, 1u);
              acc = 1u; ah += 8; al += 8; bh += step;
            }
          }
          __syncwarp();
          mbar_wait(&pipe->conv[stage], use & 1);
          tc_fence_after();
          if (elect_one()) {
            uint32_t bl = lbo_word | ((st_lo >> 4) & 0x3fffu);
            uint32_t ah = tb + TMB_ZHI + q * 8u;
#pragma unroll 4
            for (int k8 = 0; k8 < nk8; ++k8) {
              mma_ts2(d, ah, bl, UM_DESC_HIWORD, idesc, 1u);
              ah += 8; bl += step;
            }
          }
        } else {
          // row-contracted: D[out][feature] (+)= Z^T (atom column q, resident) . X^T (ring chunk); both operands
          // 128-byte-swizzled K-major with K = rows; K = 8 rows = 32 bytes inside the 128-byte line.
          // K_DV2 has the roles swapped: A = ring chunk (u^T, 128 lines), B = the resident image (dC^T, 32 lines).
          const uint32_t d = tb + (kind == K_DWH ? TMB_G : kind == K_DV1 ? TMB_V1 : TMB_EX);
          const uint32_t zh0 = ((zt_addr + q * ZT_ATOM) >> 4) & 0x3fffu, zl0 = ((zt_addr + ZT_LO + q * ZT_ATOM) >> 4) & 0x3fffu;
          const bool swap = kind == K_DV2;
          if (elect_one()) {
            uint32_t rh = (st_hi >> 4) & 0x3fffu, zh = zh0, zl = zl0;
            uint32_t acc = acc0;
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) {
              if (swap) {
                mma_ss2(d, rh, SW128_HIWORD, zl, SW128_HIWORD, idesc, acc);
                mma_ss2(d, rh, SW128_HIWORD, zh, SW128_HIWORD, idesc, 1u);
              } else {
                mma_ss2(d, zl, SW128_HIWORD, rh, SW128_HIWORD, idesc, acc);
                mma_ss2(d, zh, SW128_HIWORD, rh, SW128_HIWORD, idesc, 1u);
              }
              acc = 1u; rh += 2; zh += 2; zl += 2;
            }
          }
          __syncwarp();
          mbar_wait(&pipe->conv[stage], use & 1);
          tc_fence_after();
          if (elect_one()) {
            uint32_t rl = (st_lo >> 4) & 0x3fffu, zh = zh0;
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) {
              if (swap) mma_ss2(d, rl, SW128_HIWORD, zh, SW128_HIWORD, idesc, 1u);
              else mma_ss2(d, zh, SW128_HIWORD, rl, SW128_HIWORD, idesc, 1u);
              rl += 2; zh += 2;
            }
          }
        }
        if (elect_one()) {
          mma_commit(&pipe->empty[stage]);
          if (flags & RF_COMMIT_D0) mma_commit(&pipe->d_ready[0]);
          if (flags & RF_COMMIT_D1) mma_commit(&pipe->d_ready[1]);
        }
        __syncwarp();
        TRC(0, 3 * i + 2);
      }
    }
  } else {
    // ================================================================ compute / epilogue warps
    uint32_t d_par[2] = {0u, 0u};
    auto wait_d = [&](int which) {
      mbar_wait(&pipe->d_ready[which], d_par[which]);
      d_par[which] ^= 1;
      tc_fence_after();
    };
    auto signal = [&](uint64_t* bar) {   // this thread's TMEM accesses and shared-memory operand writes are done
      tmem_wait_st();
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    const int row = 32 * (warp & 3) + lane;     // sample row of the tile = TMEM lane
    const int cg = warp >> 2;                   // column group: columns [32 cg, 32 cg + 32) of a 128-wide tile
    const uint32_t lane_base = 32u * (warp & 3);
    const int tq = row >> 5, tj = row & 31, tchunk = tj >> 2;
    uint8_t* zt_row = smem + SB_ZT + tq * ZT_ATOM + (tj & 3) * 4;   // + f * 128 + ((tchunk ^ (f & 7)) << 4)

    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      const int r0 = tile * a.rays_per_tile;
      const int nr = min(a.rays_per_tile, a.R - r0);
      const int nrows = nr * S;
      const size_t p0 = (size_t)r0 * S;
      const bool rv = row < nrows;

      // ---------------------------------------------------------------- per-row state + compositing backward
      if (warp == 0) TRC(2, 0);
      sDWh[tid] = 0.f;
      sDWh[tid + BNCT] = 0.f;
      if (tid < 32) sBrel[tid] = 0.f;
      float4 h_rgbs = make_float4(0.f, 0.f, 0.f, 0.f), h_raw = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tid < 128) {   // all per-row loads of the head in flight together (one DRAM round trip)
        float4 mi = make_float4(0.f, 0.f, 0.f, 0.f);
        float occ = 0.f;
        if (tid < nrows) {
          mi = reinterpret_cast<const float4*>(sv + SL.misc)[p0 + tid];
          occ = sv[SL.occ + p0 + tid];
          h_rgbs = reinterpret_cast<const float4*>(sv + SL.rgbs)[p0 + tid];
          if (a.prm.rgb_mode == LSR_RGB_AFFINE_SIGMOID) h_raw = reinterpret_cast<const float4*>(sv + SL.outraw)[p0 + tid];
        }
        const int has = (tid < nrows && mi.y > 0.5f) ? 1 : 0;
        sHas[tid] = has;
        sDOut[tid * 4 + 0] = 0.f; sDOut[tid * 4 + 1] = 0.f; sDOut[tid * 4 + 2] = 0.f;
        sDOut[tid * 4 + 3] = has ? occ : -100.f;            // Renderer.py:184-186 (slot 3: occupancy logit for the compositing pass)
        sDP[tid * 4 + 0] = 0.f; sDP[tid * 4 + 1] = 0.f; sDP[tid * 4 + 2] = 0.f; sDP[tid * 4 + 3] = 0.f;
      }
      bar_compute_b();
      if (warp == 0) TRC(2, 4);
      if (tid < nr) {   // common.py:410-421 backward: only d(rgb_s) is needed here
        const int ray = r0 + tid;
        const float g = a.gt_depth[ray];
        const float coef = a.prm.sigmoid_coef;
        const bool nz = g > 0.f;
        float gC[3] = {0.f, 0.f, 0.f};
        if (a.g_rgb && (nz || !(a.prm.flags & LSR_FLAG_SKIP_ZERO_DEPTH))) {
          gC[0] = a.g_rgb[3 * ray + 0]; gC[1] = a.g_rgb[3 * ray + 1]; gC[2] = a.g_rgb[3 * ray + 2];
        }
        float wv[8];
        float T = 1.f, sw = 0.f;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (s < S) {
            const int m = tid * S + s;
            const float alpha = sigmoidf_acc(coef * sDOut[m * 4 + 3]);
            wv[s] = alpha * T;
            T = T * ((1.f - alpha) + 1e-10f);
            sw += wv[s];
          }
        }
        const float wsum = sw + 1e-10f;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (s < S) {
            const int m = tid * S + s;
            const float f = wv[s] / wsum;
            sDOut[m * 4 + 0] = gC[0] * f; sDOut[m * 4 + 1] = gC[1] * f; sDOut[m * 4 + 2] = gC[2] * f;
          }
        }
      }
      bar_compute_b();
      if (warp == 0) TRC(2, 5);
      // ---------------------------------------------------------------- colour head activation backward (decoder.py:533-546)
      if (tid < 128) {
        const int m = tid;
        float d0 = sDOut[m * 4 + 0], d1 = sDOut[m * 4 + 1], d2 = sDOut[m * 4 + 2];
        float da[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) da[k] = 0.f;
        if (m < nrows && a.prm.rgb_mode != LSR_RGB_RAW) {
          const float4 rs = h_rgbs;
          d0 *= rs.x * (1.f - rs.x); d1 *= rs.y * (1.f - rs.y); d2 *= rs.z * (1.f - rs.z);
          if (a.prm.rgb_mode == LSR_RGB_AFFINE_SIGMOID) {
            const float4 o = h_raw;
            const float* Af = a.affine;
            const float y0 = d0, y1 = d1, y2 = d2;
            da[0] = o.x * y0; da[1] = o.x * y1; da[2] = o.x * y2;
            da[3] = o.y * y0; da[4] = o.y * y1; da[5] = o.y * y2;
            da[6] = o.z * y0; da[7] = o.z * y1; da[8] = o.z * y2;
            da[9] = y0; da[10] = y1; da[11] = y2;
            d0 = Af[0] * y0 + Af[1] * y1 + Af[2] * y2;
            d1 = Af[3] * y0 + Af[4] * y1 + Af[5] * y2;
            d2 = Af[6] * y0 + Af[7] * y1 + Af[8] * y2;
          }
        }
        sDOut[m * 4 + 0] = d0; sDOut[m * 4 + 1] = d1; sDOut[m * 4 + 2] = d2;
        if (g_af) {
#pragma unroll
          for (int k = 0; k < 12; ++k) {
            float v = da[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) atomicAdd(a.d_affine + k, v);
          }
        }
      }
      bar_compute_b();
      if (warp == 0) TRC(2, 6);

      // ---------------------------------------------------------------- output_linear gradients + M_out = dOut^T [c | 1]
      // thread-per-feature-line FMAs: lane l reads 4 rows of the (swizzled) 128-byte lines of feature f
      if (g_cw) {
        const float* tph4 = sv + SL.cst + ((size_t)4 * SL.ntiles + tile) * tplane_tile_floats(HC);   // s_4; the (U_4 c + u_4) part of h_4 is added by the finalize kernel
        const float* tpc1 = sv + SL.cc1t + (size_t)tile * tplane_tile_floats(TP_C1);
        constexpr int NL = (HC + CDIM + 1 + BCW - 1) / BCW;   // feature lines per warp (11)
        float4 xs[NL];
#pragma unroll
        for (int it = 0; it < NL; ++it) {     // every load of this warp in flight before the first use
          const int f = warp + BCW * it;
          xs[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (f < HC + CDIM + 1) {
            const bool isc = f >= HC;
            const int ff = isc ? f - HC : f, F = isc ? TP_C1 : HC;
            xs[it] = __ldcs(reinterpret_cast<const float4*>((isc ? tpc1 : tph4) + (lane >> 3) * (F * 32) + ff * 32 + (lane & 7) * 4));
          }
        }
#pragma unroll
        for (int it = 0; it < NL; ++it) {
          const int f = warp + BCW * it;
          if (f >= HC + CDIM + 1) break;
          const bool isc = f >= HC;
          const int ff = isc ? f - HC : f;
          const int rbase = (lane >> 3) * 32 + (((lane & 7) ^ (ff & 7)) << 2);
          const float xv[4] = {xs[it].x, xs[it].y, xs[it].z, xs[it].w};
          float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float4 dv = *reinterpret_cast<const float4*>(sDOut + (rbase + t) * 4);
            s0 = fmaf(dv.x, xv[t], s0); s1 = fmaf(dv.y, xv[t], s1); s2 = fmaf(dv.z, xv[t], s2);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          }
          if (lane == 0) {
            if (!isc) {
              atomicAdd(a.d_w + a.w.c_out_w + f, s0); atomicAdd(a.d_w + a.w.c_out_w + HC + f, s1); atomicAdd(a.d_w + a.w.c_out_w + 2 * HC + f, s2);
            } else {
              float* mo = a.acc + 5 * BWD_ACC_SLOTS * 128;
              atomicAdd(mo + ff, s0); atomicAdd(mo + TP_C1 + ff, s1); atomicAdd(mo + 2 * TP_C1 + ff, s2);
            }
          }
        }
      }

      if (warp == 0) TRC(2, 1);
      // ---------------------------------------------------------------- G_4 = dOut . W_out, dC = dOut . (W_out U_4)
      float g[32];
      float dCacc[8], dEacc[16];
      {
        const float d0 = sDOut[row * 4 + 0], d1 = sDOut[row * 4 + 1], d2 = sDOut[row * 4 + 2];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = 32 * cg + j;
          g[j] = d0 * sTab[TAB_WOUT + c] + d1 * sTab[TAB_WOUT + HC + c] + d2 * sTab[TAB_WOUT + 2 * HC + c];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = 8 * cg + j;
          dCacc[j] = d0 * sTab[TAB_POUT + c] + d1 * sTab[TAB_POUT + CDIM + c] + d2 * sTab[TAB_POUT + 2 * CDIM + c];
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) dEacc[j] = 0.f;
      }

      // Weight-gradient flush of layer lw, in two steps so that only a register/shared-memory copy sits on the critical
      // path: (1) flush_load: TMEM accumulators -> staging image in the (idle) Z^T region, then the accumulator columns
      // go back to the issuer; (2) flush_store: staging -> global with coalesced red.global (one 128-byte line per
      // instruction) WHILE the dX MMAs of the next layer run.  (Measured: the atomics cost 5-8 k cycles per layer and
      // tile whichever way they are issued -- LSU scalar, LSU v4 or cp.reduce.async.bulk, the latter also delaying the
      // ring's bulk copies in the shared TMA queue.)
      constexpr int STP = HC + 4;                                        // staging row pitch (floats): conflict-free 16-byte row stores
      float* st_h = reinterpret_cast<float*>(smem + SB_ZT);              // [128 o][STP]  (i contiguous)
      float* st_e = st_h + HC * STP;                                     // [80 slots][128 o]
      auto flush_load = [&](int lw) {
        const bool has_e = lw == 0 || lw == 3;
        const int ncol = has_e ? 80 : 48;
        if (lw >= 1) {   // dW^h: lane = output o of layer lw, columns = features i of h_{lw-1}
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t x[16];
            tmem_ld16(tmem_addr(tb, lane_base, TMB_G + 32 * cg + 16 * h), x);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<uint4*>(st_h + row * STP + 32 * cg + 16 * h + j) = make_uint4(x[j], x[j + 1], x[j + 2], x[j + 3]);
          }
        }
        // extras: lane = output o, columns = [e' (40) |] c (32) | 1 | pad;  16-column groups cg, cg + 4
        for (int c0 = 16 * cg; c0 < ncol; c0 += 64) {
          uint32_t x[16];
          tmem_ld16(tmem_addr(tb, lane_base, TMB_EX + c0), x);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) st_e[(c0 + j + (has_e ? 0 : 40)) * HC + row] = __uint_as_float(x[j]);
        }
      };
      auto flush_store = [&](int lw) {
        const bool has_e = lw == 0 || lw == 3;
        bar_compute_b();                                   // staging image complete
        if (lw >= 1) {   // lane = feature i (= row index of this thread), this thread's 32 outputs o
          const int ldw = lw == 3 ? ECC + HC : HC;
          float* dst = a.d_w + a.w.c_lin_w[lw] + (lw == 3 ? ECC : 0) + row;
#pragma unroll 8
          for (int j = 0; j < 32; ++j) {
            const int o = 32 * cg + j;
            red_add_f32(dst + (size_t)o * ldw, st_h[o * STP + row]);
          }
        }
        {
          float* dse = a.acc + (size_t)lw * BWD_ACC_SLOTS * 128 + row;
          const int s0 = has_e ? 0 : 40, s1 = 40 + CDIM + 1;
          for (int slot = s0 + cg; slot < s1; slot += 4) red_add_f32(dse + (size_t)slot * HC, st_e[slot * HC + row]);
        }
        bar_compute_b();                                   // staging image read: Z^T may be rewritten
      };
#pragma unroll 1
      for (int l = 4; l >= 0; --l) {
        // ---- A: Z_l = G_l * softplus'(s_l) -> TMEM (hi, lo); kept in g[] for the transposed copy
        if (warp == 0) TRC(1, (4 - l) * 10 + 0);
        if (l >= 1) {   // DRAM -> L2 one layer ahead: s_{l-1} (epilogue loads) and h_{l-2} (row-contraction pieces of layer l - 1)
          prefetch_l2(sv + SL.cst + ((size_t)(l - 1) * SL.ntiles + tile) * tplane_tile_floats(HC) + (size_t)tid * 32);
        }
        {
          const float* tps = sv + SL.cst + ((size_t)l * SL.ntiles + tile) * tplane_tile_floats(HC) + tq * (HC * 32) + (tj & 3);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float sv16[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int f = 32 * cg + 16 * h + j;
              sv16[j] = __ldcs(tps + f * 32 + ((tchunk ^ (j & 7)) << 2));
            }
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float z = rv ? g[16 * h + j] * softplus100_grad_from_out(sv16[j]) : 0.f;
              g[16 * h + j] = z;
              split_hi_lo(z, hi[j], lo[j]);
            }
            tmem_st16(tmem_addr(tb, lane_base, TMB_ZHI + 32 * cg + 16 * h), hi);
            tmem_st16(tmem_addr(tb, lane_base, TMB_ZLO + 32 * cg + 16 * h), lo);
          }
        }
        // ---- B: weight-gradient accumulators of layer l + 1 (their MMAs also read Z^T: it is free afterwards)
        if (warp == 0) TRC(1, (4 - l) * 10 + 1);
        if (g_cw && l < 4) {
          wait_d(1);
          if (warp == 0) TRC(1, (4 - l) * 10 + 2);
          flush_load(l + 1);
        }
        if (warp == 0) TRC(1, (4 - l) * 10 + 3);
        // ---- C: Z_l is in TMEM, the accumulator columns are free -> dX MMAs of layer l
        signal(&pipe->a_ready);
        if (warp == 0) TRC(1, (4 - l) * 10 + 4);
        if (g_cw && l < 4) flush_store(l + 1);
        if (warp == 0) TRC(1, (4 - l) * 10 + 5);
        // ---- D: Z_l^T -> shared memory (row-contraction operand), lane = row: one 128-byte line per store instruction
        if (g_cw) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int f = 32 * cg + j;
            uint32_t hi, lo;
            split_hi_lo(g[j], hi, lo);
            uint8_t* p = zt_row + f * 128 + ((tchunk ^ (j & 7)) << 4);
            *reinterpret_cast<uint32_t*>(p) = __float_as_uint(g[j]);   // raw word == hi operand (the tensor core truncates)
            *reinterpret_cast<uint32_t*>(p + ZT_LO) = lo;
          }
        }
        // ---- E: results of the dX GEMMs
        if (warp == 0) TRC(1, (4 - l) * 10 + 6);
        wait_d(0);
        if (warp == 0) TRC(1, (4 - l) * 10 + 7);
        if (l >= 1) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t x[16];
            tmem_ld16(tmem_addr(tb, lane_base, TMB_G + 32 * cg + 16 * h), x);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) g[16 * h + j] = __uint_as_float(x[j]);
          }
        }
        {
          const bool has_e = l == 0 || l == 3;
          uint32_t x[8];
          if (has_e) {
            tmem_ld8(tmem_addr(tb, lane_base, TMB_EX + 8 * cg), x);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j) dEacc[j] += __uint_as_float(x[j]);
            if (cg == 0) {
              tmem_ld8(tmem_addr(tb, lane_base, TMB_EX + 32), x);
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 8; ++j) dEacc[8 + j] += __uint_as_float(x[j]);
            }
          }
          if (l >= 1) {
            tmem_ld8(tmem_addr(tb, lane_base, TMB_EX + (has_e ? ECC : 0) + 8 * cg), x);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j) dCacc[j] += __uint_as_float(x[j]);
          }
        }
        // ---- F: Z^T written, accumulator columns read -> weight-gradient MMAs of layer l
        if (warp == 0) TRC(1, (4 - l) * 10 + 8);
        if (g_cw) signal(&pipe->b_ready);
        else { tc_fence_before(); }
        if (warp == 0) TRC(1, (4 - l) * 10 + 9);
      }
      if (g_cw) {   // layer 0: extras only ([E_0 | M_0])
        wait_d(1);
        flush_load(0);
        flush_store(0);
      }
      if (warp == 0) TRC(2, 2);
      // ---------------------------------------------------------------- hand-over: dL/dc and the Fourier part of dL/dp
      if (rv) {
        float4* o = reinterpret_cast<float4*>(a.out_dc + (p0 + row) * CDIM + 8 * cg);
        o[0] = make_float4(dCacc[0], dCacc[1], dCacc[2], dCacc[3]);
        o[1] = make_float4(dCacc[4], dCacc[5], dCacc[6], dCacc[7]);
      }
      // ---------------------------------------------------------------- rel-pos neighbour MLP backward (decoder.py:307-323,477-488)
      //   c = V2 u + v2 sum_k w_k,  u = sum_k w_k softplus(V1 q_k + v1),  q_k = [sin(phi_k) | cos(phi_k) | F^c[I_k]],  phi_k = 2 pi (x_{I_k} - p) B_r
      if (relpos) {
        const size_t prow = p0 + row;
        const bool has = sHas[row] != 0;
        prefetch_l2(sv + SL.spt + ((size_t)tile * KNN) * tplane_tile_floats(HC) + (size_t)tid * 32);
        if (tid < TP_Q * 4) prefetch_l2(sv + SL.qt + ((size_t)tile * KNN) * tplane_tile_floats(TP_Q) + (size_t)tid * 32);
        if (g_cw) prefetch_l2(sv + SL.ut + (size_t)tile * tplane_tile_floats(HC) + (size_t)tid * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) if (!has || !rv) dCacc[j] = 0.f;    // rows without neighbours do not reach the features (decoder.py:489-490)
        {   // dC -> TMEM A operand (K = 32) of dU = dC V2, and (weights train) dC^T -> row-contraction operand of dV2 = dC^T u
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) split_hi_lo(dCacc[j], hi[j], lo[j]);
          tmem_st8(tmem_addr(tb, lane_base, TMB_ZHI + 8 * cg), hi);
          tmem_st8(tmem_addr(tb, lane_base, TMB_ZLO + 8 * cg), lo);
          if (g_cw) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = 8 * cg + j;
              uint8_t* pz = zt_row + c * 128 + ((tchunk ^ (j & 7)) << 4);
              *reinterpret_cast<uint32_t*>(pz) = __float_as_uint(dCacc[j]);
              *reinterpret_cast<uint32_t*>(pz + ZT_LO) = lo[j];
            }
            // d v2 = sum_r dC[r] * sum_k w_k[r]
            const float ws = rv ? reinterpret_cast<const float4*>(sv + SL.misc)[prow].z : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float v = dCacc[j] * ws;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
              if (lane == 0) atomicAdd(a.d_w + a.w.c_nb2_b + 8 * cg + j, v);
            }
          }
        }
        signal(&pipe->a_ready);
        if (g_cw) signal(&pipe->b_ready);
        float rt = 0.f;   // tracker: d w_hat_k += dC . v2 for every listed neighbour
        if (trk) {
#pragma unroll
          for (int j = 0; j < 8; ++j) rt = fmaf(dCacc[j], sTab[TAB_V2B + 8 * cg + j], rt);
        }
        // dU = dC V2: this thread's 32 hidden columns, kept for all 8 neighbours
        wait_d(0);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t x[16];
          tmem_ld16(tmem_addr(tb, lane_base, TMB_G + 32 * cg + 16 * h), x);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) g[16 * h + j] = __uint_as_float(x[j]);
        }
        if (g_cw) {   // dV2^T[hid][c] (lane = hid, columns = c): coalesced over the lanes for every c
          wait_d(1);
          uint32_t x[8];
          tmem_ld8(tmem_addr(tb, lane_base, TMB_EX + 8 * cg), x);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 8; ++j) red_add_f32(a.d_w + a.w.c_nb2_w + (size_t)(8 * cg + j) * HC + row, __uint_as_float(x[j]));
        }
        // Z_k = w_k dU * softplus'(.) for one neighbour: values in zk[], d w_hat_k partial = dU . softplus_k
        float zk[32];
        int idx = -1;
        auto compute_z = [&](int k) {
          if (k + 1 < KNN) {   // DRAM -> L2: what round k + 1 reads (512 + 256 lines: one or two per thread)
            prefetch_l2(sv + SL.spt + ((size_t)tile * KNN + k + 1) * tplane_tile_floats(HC) + (size_t)tid * 32);
            if (tid < TP_Q * 4) prefetch_l2(sv + SL.qt + ((size_t)tile * KNN + k + 1) * tplane_tile_floats(TP_Q) + (size_t)tid * 32);
          }
          const float wk = rv ? sv[SL.w + prow * KNN + k] : 0.f;
          const int idk = rv ? reinterpret_cast<const int*>(sv + SL.idx)[prow * KNN + k] : -1;
          const float* tsp = sv + SL.spt + ((size_t)tile * KNN + k) * tplane_tile_floats(HC) + tq * (HC * 32) + (tj & 3);
          float part = 0.f;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float sp16[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) sp16[j] = __ldcs(tsp + (32 * cg + 16 * h + j) * 32 + ((tchunk ^ (j & 7)) << 2));
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float du = g[16 * h + j];
              part = fmaf(du, sp16[j], part);
              zk[16 * h + j] = wk * du * softplus100_grad_from_out(sp16[j]);
            }
          }
          if (trk && idk >= 0) atomicAdd(&sDWh[row * KNN + k], part + rt);
          return idk;
        };
        int idx_next = compute_z(0);
#pragma unroll 1
        for (int k = 0; k < KNN; ++k) {
          if (warp == 0) TRC(2, 16 + 6 * k);
          idx = idx_next;
          // ---- Z_k -> TMEM (A of dQ_k = Z_k V1); the accumulator columns of dQ_{k-1} were read at the end of the last round
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) split_hi_lo(zk[16 * h + j], hi[j], lo[j]);
            tmem_st16(tmem_addr(tb, lane_base, TMB_ZHI + 32 * cg + 16 * h), hi);
            tmem_st16(tmem_addr(tb, lane_base, TMB_ZLO + 32 * cg + 16 * h), lo);
          }
          signal(&pipe->a_ready);
          if (warp == 0) TRC(2, 16 + 6 * k + 1);
          if (g_cw) {                             // Z_k^T for dV1 += Z_k^T Q_k, once the previous round's MMAs have read the old image
            if (k > 0) wait_d(1);
            if (warp == 0) TRC(2, 16 + 6 * k + 2);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              uint32_t hi, lo;
              split_hi_lo(zk[j], hi, lo);
              uint8_t* pz = zt_row + (32 * cg + j) * 128 + ((tchunk ^ (j & 7)) << 4);
              *reinterpret_cast<uint32_t*>(pz) = __float_as_uint(zk[j]);
              *reinterpret_cast<uint32_t*>(pz + ZT_LO) = lo;
            }
            signal(&pipe->b_ready);
          }
          if (warp == 0) TRC(2, 16 + 6 * k + 3);
          // ---- next round's Z values, computed while this round's MMAs run
          if (k + 1 < KNN) idx_next = compute_z(k + 1);
          // ---- dQ_k = Z_k V1: columns [0,20) -> Fourier / pose / B_rel, [20,52) -> the neighbour's colour feature row
          wait_d(0);
          if (warp == 0) TRC(2, 16 + 6 * k + 4);
          if (cg == 0) {
            // columns [0,20): gradients of [sin(phi_k) | cos(phi_k)].  Their chain rule (sincos recompute, d B_rel, pose part) is
            // elementwise work with gathers and reductions: it runs in relpos_trig_bwd_kernel at full occupancy; here only
            // a coalesced store (lane = row)
            uint32_t x[16], y[8];
            tmem_ld16(tmem_addr(tb, lane_base, TMB_G), x);
            tmem_ld8(tmem_addr(tb, lane_base, TMB_G + 16), y);
            tmem_wait_ld();
            if (g_cw || g_ry) {
              float* dq = a.out_dqt + (((size_t)tile * KNN + k) * (2 * ER)) * 128 + row;
#pragma unroll
              for (int j = 0; j < 16; ++j) dq[j * 128] = __uint_as_float(x[j]);
#pragma unroll
              for (int j = 0; j < 4; ++j) dq[(16 + j) * 128] = __uint_as_float(y[j]);
            }
          } else if (cg <= 2) {
            // feature columns 20 + 16 (cg - 1) .. + 15, fetched as the two aligned 16-column blocks around them
            uint32_t x[32];
            tmem_ld16(tmem_addr(tb, lane_base, TMB_G + 16 * cg), *reinterpret_cast<uint32_t(*)[16]>(&x[0]));
            tmem_ld16(tmem_addr(tb, lane_base, TMB_G + 16 * cg + 16), *reinterpret_cast<uint32_t(*)[16]>(&x[16]));
            tmem_wait_ld();
            if (g_cf && idx >= 0) {
              float* dst = grad_row(a.d_col, a.remap, idx);
              if (dst) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                  red_add_v4(dst + 16 * (cg - 1) + j, __uint_as_float(x[4 + j]), __uint_as_float(x[5 + j]), __uint_as_float(x[6 + j]),
                             __uint_as_float(x[7 + j]));
              }
            }
          }
          tc_fence_before();
          if (warp == 0) TRC(2, 16 + 6 * k + 5);
        }
        if (g_cw) {   // dV1^T[hid][q] accumulated over the 8 neighbours: lane = hid, columns = q (52) | bias | pad
          wait_d(1);
          float* stg = reinterpret_cast<float*>(smem + SB_ZT);            // [128 hid][68]: coalesced rows of d V1 afterwards
          uint32_t x[16];
          tmem_ld16(tmem_addr(tb, lane_base, TMB_V1 + 16 * cg), x);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) stg[row * 68 + 16 * cg + j] = __uint_as_float(x[j]);
          bar_compute_b();
          for (int e = tid; e < HC * QD; e += BNCT) red_add_f32(a.d_w + a.w.c_nb1_w + e, stg[(e / QD) * 68 + e % QD]);
          if (tid < HC) red_add_f32(a.d_w + a.w.c_nb1_b + tid, stg[tid * 68 + QD]);
          bar_compute_b();
        }
      }
      if (g_ry) {   // e' = [sin(arg) | cos(arg)], arg_j = 2 pi p . B_j  (decoder.py:34-43)
        const float* tpe = sv + SL.ect + (size_t)tile * tplane_tile_floats(ECC) + tq * (ECC * 32) + (tj & 3);
        float q0 = 0.f, q1 = 0.f, q2 = 0.f;
        const int ncols = cg == 0 ? 16 : 8;
        for (int jj = 0; jj < ncols; ++jj) {
          const int e = jj < 8 ? 8 * cg + jj : 32 + (jj - 8);
          const int j = e < EC ? e : e - EC, partner = e < EC ? e + EC : e - EC;
          const float pv = tpe[partner * 32 + ((tchunk ^ (partner & 7)) << 2)];
          const float dar = e < EC ? dEacc[jj] * pv : -dEacc[jj] * pv;
          q0 = fmaf(sTab[TAB_CB + j], dar, q0);
          q1 = fmaf(sTab[TAB_CB + EC + j], dar, q1);
          q2 = fmaf(sTab[TAB_CB + 2 * EC + j], dar, q2);
        }
        if (rv) {
          atomicAdd(&sDP[row * 4 + 0], TWO_PI_F * q0);
          atomicAdd(&sDP[row * 4 + 1], TWO_PI_F * q1);
          atomicAdd(&sDP[row * 4 + 2], TWO_PI_F * q2);
        }
      }
      tc_fence_before();
      bar_compute_b();
      if (warp == 0) TRC(2, 3);
      if (g_ry && tid < nrows)
        reinterpret_cast<float4*>(a.out_dp)[p0 + tid] = make_float4(sDP[tid * 4 + 0], sDP[tid * 4 + 1], sDP[tid * 4 + 2], 0.f);
      if (trk && relpos) {
        for (int i = tid; i < nrows * KNN; i += BNCT) a.out_dwh[p0 * KNN + i] = sDWh[i];
      }
      bar_compute_b();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_ISSUER) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------------------ rel-pos Fourier backward
// q_k[0:20] = [sin(phi_k) | cos(phi_k)], phi_k = 2 pi (x_{I_k} - p) B_r (decoder.py:34-43,477-480).  Given d q_k[0:20] from the
// tcgen05 kernel: d phi_j = dsin_j cos_j - dcos_j sin_j;  d B_r[c][j] += t_c d phi_j (t = 2 pi (x - p));  d p -= 2 pi B_r d phi.
// One block per (tile, neighbour), one thread per sample row (sincos recomputed exactly as the forward does).
struct TrigArgs {
  const float* saved; const float* dqt; const float4* knn_pos; const float* cloud; const float* blob;
  int brel; int R, S, rays_per_tile; int flags; int want_w, want_p;
  float* d_w; float* out_dp;
};
__global__ void __launch_bounds__(128) relpos_trig_bwd_kernel(const __grid_constant__ TrigArgs a) {
  const SavedLayout SL = saved_layout(a.R, a.S, LSR_STAGE_COLOR, a.flags);
  __shared__ float sB[3 * ER];
  __shared__ float sRed[4][3 * ER];
  const int tile = blockIdx.x, k = blockIdx.y, row = threadIdx.x, lane = row & 31, warp = row >> 5;   // block = (tile, neighbour)
  if (row < 3 * ER) sB[row] = a.blob[a.brel + row];
  __syncthreads();
  const int r0 = tile * a.rays_per_tile;
  const int nrows = min(a.rays_per_tile, a.R - r0) * a.S;
  const size_t p0 = (size_t)r0 * a.S;
  float dB[3 * ER];
#pragma unroll
  for (int i = 0; i < 3 * ER; ++i) dB[i] = 0.f;
  if (row < nrows) {
    const int idx = reinterpret_cast<const int*>(a.saved + SL.idx)[(p0 + row) * KNN + k];
    if (idx >= 0) {
      const float4 pp = a.knn_pos[p0 + row];
      const float t0 = TWO_PI_F * __fsub_rn(__ldg(a.cloud + 3 * (size_t)idx + 0), pp.x);
      const float t1 = TWO_PI_F * __fsub_rn(__ldg(a.cloud + 3 * (size_t)idx + 1), pp.y);
      const float t2 = TWO_PI_F * __fsub_rn(__ldg(a.cloud + 3 * (size_t)idx + 2), pp.z);
      const float* dq = a.dqt + (((size_t)tile * KNN + k) * (2 * ER)) * 128 + row;
      float q0 = 0.f, q1 = 0.f, q2 = 0.f;
#pragma unroll
      for (int j = 0; j < ER; ++j) {
        const float arg = fmaf(t2, sB[2 * ER + j], fmaf(t1, sB[ER + j], t0 * sB[j]));
        float sn, cs;
        sincos_ff(arg, &sn, &cs);
        const float dphi = dq[j * 128] * cs - dq[(ER + j) * 128] * sn;
        dB[j] = t0 * dphi; dB[ER + j] = t1 * dphi; dB[2 * ER + j] = t2 * dphi;
        q0 = fmaf(sB[j], dphi, q0); q1 = fmaf(sB[ER + j], dphi, q1); q2 = fmaf(sB[2 * ER + j], dphi, q2);
      }
      if (a.want_p) {   // rel = x - p; the 8 neighbour blocks of a row add into the same slot
        float* o = a.out_dp + (p0 + row) * 4;
        atomicAdd(o + 0, -TWO_PI_F * q0); atomicAdd(o + 1, -TWO_PI_F * q1); atomicAdd(o + 2, -TWO_PI_F * q2);
      }
    }
  }
  if (a.want_w) {
#pragma unroll
    for (int i = 0; i < 3 * ER; ++i) {
      float v = dB[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) sRed[warp][i] = v;
    }
    __syncthreads();
    if (row < 3 * ER) atomicAdd(a.d_w + a.brel + row, (sRed[0][row] + sRed[1][row]) + (sRed[2][row] + sRed[3][row]));
  }
}

// ------------------------------------------------------------------------------------------------ finalize
// acc[l][slot][o]: slot < 40: E_l = Z_l^T e' (layers 0, 3), slot 40 + c: M_l = Z_l^T c, slot 72: Z_l^T 1 = d b_l.
//   d c_lin_b[l]          = M_l[ones]
//   d c_lin_w[l][o][e]   += E_l[e][o]                          (l = 0, 3)
//   d c_fc_w[l-1][i][c]   = sum_o W_l[o][hoff + i] M_l[c][o]   (l = 1..4);  d c_fc_b[l-1][i] = sum_o W_l[o][hoff + i] M_l[ones][o]
//   d c_fc_w[4][i][c]     = sum_o W_out[o][i] M_out[o][c];     d c_fc_b[4][i] = sum_o W_out[o][i] M_out[o][ones];  d c_out_b = M_out[:, ones]
__global__ void __launch_bounds__(HC) trunk_bwd_finalize_kernel(const float* __restrict__ blob, const float* __restrict__ acc,
                                                                float* __restrict__ dW, const __grid_constant__ LsrWeights w) {
  // grid (33 + 2 + 128, 5): blockIdx.x = c < 33 -> column c of fc_c[l] (32: bias) for all 128 i (threadIdx.x);
  //                   33 -> lin bias of layer l; 34 -> e' columns of c_lin_w[l] (l = 0, 3) and the head bias;
  //                   35 + o -> the (U c + u) part of the h-columns of row o of c_lin_w[l] (l >= 1) / c_out_w (l == 0)
  const int l = blockIdx.y, c = blockIdx.x, i = threadIdx.x;
  const float* mo = acc + 5 * BWD_ACC_SLOTS * 128;
  __shared__ float sm[HC];
  if (c <= CDIM) {
    float s = 0.f;
    if (l == 4) {
      for (int o = 0; o < 3; ++o) s = fmaf(blob[w.c_out_w + o * HC + i], mo[o * TP_C1 + c], s);
    } else {
      const int lw = l + 1, ldw = lw == 3 ? ECC + HC : HC, off = lw == 3 ? ECC : 0;
      sm[i] = acc[((size_t)lw * BWD_ACC_SLOTS + 40 + c) * 128 + i];
      __syncthreads();
      const float* wr = blob + w.c_lin_w[lw] + off + i;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 8
      for (int o = 0; o < HC; o += 4) {
        s0 = fmaf(wr[(size_t)(o + 0) * ldw], sm[o + 0], s0);
        s1 = fmaf(wr[(size_t)(o + 1) * ldw], sm[o + 1], s1);
        s2 = fmaf(wr[(size_t)(o + 2) * ldw], sm[o + 2], s2);
        s3 = fmaf(wr[(size_t)(o + 3) * ldw], sm[o + 3], s3);
      }
      s = (s0 + s1) + (s2 + s3);
    }
    if (c < CDIM) dW[w.c_fc_w[l] + i * CDIM + c] += s;
    else dW[w.c_fc_b[l] + i] += s;
  } else if (c == CDIM + 1) {
    dW[w.c_lin_b[l] + i] += acc[((size_t)l * BWD_ACC_SLOTS + 40 + CDIM) * 128 + i];
  } else if (c >= CDIM + 3) {
    // The tensor-core pass contracted Z_l with the SAVED softplus outputs s_{l-1} only; the rest of h_{l-1} = s_{l-1} + U_{l-1} c +
    // u_{l-1} enters through M_l:  d c_lin_w[l][o][hoff + i] += sum_c M_l[c][o] U_{l-1}[i][c] + M_l[ones][o] u_{l-1}[i]   (l = 1..4),
    // and for the head (blockIdx.y == 0):  d c_out_w[o][i] += sum_c M_out[o][c] U_4[i][c] + M_out[o][ones] u_4[i].
    const int o = c - (CDIM + 3);
    if (l == 0 && o >= 3) return;
    const int lu = l == 0 ? 4 : l - 1;                       // the fc_c layer whose output is part of the contracted h
    sm[i] = 0.f;
    if (i <= CDIM) sm[i] = l == 0 ? mo[o * TP_C1 + i] : acc[((size_t)l * BWD_ACC_SLOTS + 40 + i) * 128 + o];
    __syncthreads();
    const float* ur = blob + w.c_fc_w[lu] + (size_t)i * CDIM;
    float s = sm[CDIM] * blob[w.c_fc_b[lu] + i];
#pragma unroll 8
    for (int cc = 0; cc < CDIM; ++cc) s = fmaf(sm[cc], ur[cc], s);
    if (l == 0) dW[w.c_out_w + o * HC + i] += s;
    else dW[w.c_lin_w[l] + (size_t)o * (l == 3 ? ECC + HC : HC) + (l == 3 ? ECC : 0) + i] += s;
  } else {
    if (l == 0 || l == 3) {
      const int ldw = l == 3 ? ECC + HC : ECC;
      for (int e = 0; e < ECC; ++e) dW[w.c_lin_w[l] + (size_t)i * ldw + e] += acc[((size_t)l * BWD_ACC_SLOTS + e) * 128 + i];
    }
    if (l == 4 && i < 3) dW[w.c_out_b + i] += mo[i * TP_C1 + CDIM];
  }
}

// ------------------------------------------------------------------------------------------------ host
static int add_bjob(BJobs& J, int type, int wsrc, int ld, int col0, int u, int n0, int n_valid, int n_pad, int kc, int dst,
                    int k_total = HC) {
  if (J.n >= MAX_BJOBS) return -1;
  BJob& j = J.j[J.n++];
  j.type = type; j.w = wsrc; j.ld = ld; j.col0 = col0; j.u = u; j.n0 = n0; j.n_valid = n_valid; j.n_pad = n_pad;
  j.k_total = k_total; j.kc = kc; j.dst = dst;
  return 0;
}

struct TrunkProgram { ROp ops[MAX_ROPS]; int n_ops; BJobs jobs; int pack_floats; };

// The ring program of one tile + the weight re-layout jobs.  Mirrors the epilogue code of trunk_bwd_umma_kernel.
static void build_trunk_program(const LsrWeights* w, const SavedLayout& SL, bool g_cw, bool relpos, TrunkProgram* P) {
  P->n_ops = 0;
  P->jobs.n = 0;
  int pk = 0;
  auto op = [&](int kind, uint32_t src0, uint32_t stride0, int bytes0, uint32_t src1, uint32_t stride1, int bytes1, int flags, int nk8,
                int q, int n) {
    ROp& o = P->ops[P->n_ops++];
    o.src0 = src0; o.src1 = src1; o.stride0 = stride0; o.stride1 = stride1;
    o.bytes0 = (uint16_t)bytes0; o.bytes1 = (uint16_t)bytes1;
    o.kind = (uint8_t)kind; o.flags = (uint8_t)flags; o.nk8 = (uint8_t)nk8; o.q = (uint8_t)q; o.n = (uint16_t)n;
    o.pad = 0; o.pad2 = 0;
  };
  const uint32_t tile_h = (uint32_t)tplane_tile_floats(HC), tile_c1 = (uint32_t)tplane_tile_floats(TP_C1),
                 tile_e = (uint32_t)tplane_tile_floats(ECC);
  for (int l = 4; l >= 0; --l) {
    const int ldw = l == 0 ? ECC : (l == 3 ? ECC + HC : HC), hoff = l == 3 ? ECC : 0;
    const bool has_e = l == 0 || l == 3;
    bool first_dx = true;
    // ---- dX: G_{l-1} = Z_l W_l^h  (four chunks of K = 32 output features)
    if (l >= 1) {
      add_bjob(P->jobs, 0, w->c_lin_w[l], ldw, hoff, 0, 0, HC, HC, 32, pk);
      for (int c = 0; c < 4; ++c) {
        op(K_DXW, (uint32_t)(pk + c * HC * 32), 0, HC * 32 * 4, 0, 0, 0, (first_dx ? RF_WAIT_A : 0) | (c == 0 ? RF_FIRST : 0), 4, c * 4, HC);
        first_dx = false;
      }
      pk += HC * HC;
    }
    // ---- dX extras: [dE' |] dC
    {
      const int n_valid = (has_e ? ECC : 0) + (l >= 1 ? CDIM : 0);
      const int n_pad = (n_valid + 15) / 16 * 16;
      const int kc = (4096 / n_pad) / 8 * 8 < HC ? (4096 / n_pad) / 8 * 8 : HC;
      if (has_e) add_bjob(P->jobs, 0, w->c_lin_w[l], ldw, 0, 0, 0, ECC, n_pad, kc, pk);
      if (l >= 1) add_bjob(P->jobs, 1, w->c_lin_w[l], ldw, hoff, w->c_fc_w[l - 1], has_e ? ECC : 0, CDIM, n_pad, kc, pk);
      int off = pk;
      for (int k0 = 0; k0 < HC; k0 += kc) {
        const int kk = HC - k0 < kc ? HC - k0 : kc;
        const bool last = k0 + kc >= HC;
        op(K_DXE, (uint32_t)off, 0, n_pad * kk * 4, 0, 0, 0, (first_dx ? RF_WAIT_A : 0) | (k0 == 0 ? RF_FIRST : 0) | (last ? RF_COMMIT_D0 : 0),
           kk / 8, k0 / 8, n_pad);
        first_dx = false;
        off += n_pad * kk;
      }
      pk += n_pad * HC;
    }
    if (!g_cw) continue;
    // ---- dW: Z_l^T . h_{l-1}^T per atom column, then Z_l^T . [e' | c | 1]^T
    bool first_dw = true;
    if (l >= 1) {
      for (int q = 0; q < 4; ++q) {
        op(K_DWH, (uint32_t)(SL.cst + (size_t)(l - 1) * SL.ntiles * tile_h + q * (HC * 32)), tile_h, HC * 128, 0, 0, 0,
           (first_dw ? RF_WAIT_B : 0) | (q == 0 ? RF_FIRST : 0), 4, q, HC);
        first_dw = false;
      }
    }
    for (int q = 0; q < 4; ++q) {
      const int flags = (first_dw ? RF_WAIT_B : 0) | (q == 0 ? RF_FIRST : 0) | (q == 3 ? RF_COMMIT_D1 : 0);
      if (has_e)
        op(K_DWE, (uint32_t)(SL.ect + q * (ECC * 32)), tile_e, ECC * 128, (uint32_t)(SL.cc1t + q * (TP_C1 * 32)), tile_c1, TP_C1 * 128,
           flags, 4, q, 80);
      else
        op(K_DWE, (uint32_t)(SL.cc1t + q * (TP_C1 * 32)), tile_c1, TP_C1 * 128, 0, 0, 0, flags, 4, q, 48);
      first_dw = false;
    }
  }
  if (relpos) {
    // ---- rel-pos neighbour MLP: dU = dC V2 (K = 32), dV2^T = u^T . dC^T, then per neighbour dQ_k = Z_k V1 and dV1 += Z_k^T Q_k
    add_bjob(P->jobs, 0, w->c_nb2_w, HC, 0, 0, 0, HC, HC, 32, pk, CDIM);          // B[n = hid][k = c] = V2[c][hid]
    op(K_DXW, (uint32_t)pk, 0, HC * 32 * 4, 0, 0, 0, RF_WAIT_A | RF_FIRST | RF_COMMIT_D0, 4, 0, HC);
    pk += HC * 32;
    if (g_cw) {
      for (int q = 0; q < 4; ++q)
        op(K_DV2, (uint32_t)(SL.ut + q * (HC * 32)), tile_h, HC * 128, 0, 0, 0,
           (q == 0 ? RF_WAIT_B | RF_FIRST : 0) | (q == 3 ? RF_COMMIT_D1 : 0), 4, q, CDIM);
    }
    const int v1t = pk;
    add_bjob(P->jobs, 0, w->c_nb1_w, QD, 0, 0, 0, QD, 64, 64, pk, HC);             // B[n = q][k = hid] = V1[hid][q]
    pk += 64 * HC;
    const uint32_t tile_q = (uint32_t)tplane_tile_floats(TP_Q);
    for (int k = 0; k < KNN; ++k) {
      op(K_DXW, (uint32_t)v1t, 0, 64 * 64 * 4, 0, 0, 0, RF_WAIT_A | RF_FIRST, 8, 0, 64);
      op(K_DXW, (uint32_t)(v1t + 64 * 64), 0, 64 * 64 * 4, 0, 0, 0, RF_COMMIT_D0, 8, 8, 64);
      if (g_cw) {
        for (int q = 0; q < 4; ++q)
          op(K_DV1, (uint32_t)(SL.qt + (size_t)k * tile_q + q * (TP_Q * 32)), KNN * tile_q, TP_Q * 128, 0, 0, 0,
             (q == 0 ? RF_WAIT_B : 0) | ((k == 0 && q == 0) ? RF_FIRST : 0) | (q == 3 ? RF_COMMIT_D1 : 0), 4, q, 64);
      }
    }
  }
  P->pack_floats = pk;
  P->jobs.pout_dst = BWD_PACK_FLOATS_MAX - 128;
  P->jobs.w_out = w->c_out_w;
  P->jobs.u4 = w->c_fc_w[4];
}

int sm_count();

// Colour-trunk backward of one ray batch: weight re-layout, tcgen05 kernel, finalize.  Writes dL/dc (out_dc) and the
// Fourier part of dL/dp (out_dp) per sample row for the remaining (rel-pos / geometry) backward.
int launch_trunk_bwd(const LsrParams* prm, const LsrWeights* w, const float* gt_depth, int64_t n_rays, const float* affine,
                     const void* saved, void* scratch, const float* g_depth, const float* g_var, const float* g_rgb, int grad_flags,
                     float* d_weights, float* d_affine, const float* cloud_pos, const int32_t* row_remap, float* d_col_feats,
                     int is_tracker, cudaStream_t stream, int phase, cudaEvent_t ev_after_trunk) {
  // phase 0: weight images, trunk, trig, finalize on `stream`; 1: without the finalize kernel; 2: only the finalize kernel (the
  // caller orders it behind the trunk kernel -- ev_after_trunk -- and may put it on another stream)
  const SavedLayout SL = saved_layout(n_rays, prm->n_surface, LSR_STAGE_COLOR, prm->flags);
  const ScratchLayout CL = scratch_layout(n_rays, prm->n_surface);
  if (SL.total >= (1ull << 32)) return LSR_ERR_UNSUPPORTED;   // ROp offsets are 32-bit float indices
  char* sbase = (char*)scratch;
  const bool g_cw = (grad_flags & LSR_GRAD_COL_W) && d_weights;
  if (phase == 2) {
    if (g_cw) {
      trunk_bwd_finalize_kernel<<<dim3(CDIM + 3 + HC, 5), HC, 0, stream>>>(w->blob, (const float*)(sbase + CL.bwd_acc), d_weights, *w);
      LSR_LAUNCHED(1);
      LSR_CUDA_CHECK(cudaGetLastError());
    }
    return LSR_OK;
  }
  static thread_local TrunkProgram P;   // host scratch, too large for the stack of a small thread
  const bool relpos = (prm->flags & LSR_FLAG_REL_POS) != 0;
  build_trunk_program(w, SL, g_cw, relpos, &P);
  if (P.pack_floats > BWD_PACK_FLOATS_MAX - 128 || P.n_ops > MAX_ROPS) return LSR_ERR_UNSUPPORTED;
  float* pack = (float*)(sbase + CL.bwd_pack);
  float* acc = (float*)(sbase + CL.bwd_acc);
  LSR_CUDA_CHECK(cudaMemsetAsync(pack, 0, (size_t)BWD_PACK_FLOATS_MAX * sizeof(float), stream));
  if (g_cw) LSR_CUDA_CHECK(cudaMemsetAsync(acc, 0, (size_t)BWD_ACC_FLOATS * sizeof(float), stream));
  bwd_prep_kernel<<<dim3(32, P.jobs.n + 1), 256, 0, stream>>>(w->blob, pack, P.jobs);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());

  static thread_local TrunkArgs a;
  a.prm = *prm;
  a.w = *w;
  a.gt_depth = gt_depth;
  a.g_depth = g_depth; a.g_var = g_var; a.g_rgb = g_rgb;
  a.affine = affine;
  a.d_affine = d_affine;
  a.R = (int)n_rays;
  a.saved = (const float*)saved;
  a.pack = pack;
  a.acc = acc;
  a.d_w = d_weights;
  a.out_dc = (float*)(sbase + CL.bwd_dc);
  a.out_dp = (float*)(sbase + CL.bwd_dp);
  a.out_dwh = (float*)(sbase + CL.bwd_dwh);
  a.out_dqt = (float*)(sbase + CL.bwd_dqt);
  a.cloud = cloud_pos;
  a.knn_pos = (const float4*)(sbase + CL.knn_pos);
  a.remap = row_remap;
  a.d_col = d_col_feats;
  a.is_tracker = is_tracker;
  a.gflags = grad_flags;
  a.ntiles = SL.ntiles;
  a.rays_per_tile = SL.rays_per_tile;
  a.n_ops = P.n_ops;
  memcpy(a.ops, P.ops, sizeof(ROp) * P.n_ops);
  const int nsm = sm_count();
  if (nsm <= 0) return LSR_ERR_CUDA;
  LSR_SMEM_ATTR_ONCE(trunk_bwd_umma_kernel, BWD_UMMA_SMEM);
  const int grid = a.ntiles < nsm ? a.ntiles : nsm;
  trunk_bwd_umma_kernel<<<grid, BT, BWD_UMMA_SMEM, stream>>>(a);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  if (ev_after_trunk) LSR_CUDA_CHECK(cudaEventRecord(ev_after_trunk, stream));
  if (relpos && (g_cw || (grad_flags & LSR_GRAD_RAYS))) {
    TrigArgs t;
    t.saved = (const float*)saved; t.dqt = a.out_dqt; t.knn_pos = a.knn_pos; t.cloud = cloud_pos; t.blob = w->blob;
    t.brel = w->c_Brel; t.R = (int)n_rays; t.S = prm->n_surface; t.rays_per_tile = SL.rays_per_tile; t.flags = prm->flags;
    t.want_w = g_cw ? 1 : 0; t.want_p = (grad_flags & LSR_GRAD_RAYS) ? 1 : 0;
    t.d_w = d_weights; t.out_dp = a.out_dp;
    relpos_trig_bwd_kernel<<<dim3(a.ntiles, KNN), 128, 0, stream>>>(t);
    LSR_LAUNCHED(1);
    LSR_CUDA_CHECK(cudaGetLastError());
  }
  if (g_cw && phase == 0) {
    trunk_bwd_finalize_kernel<<<dim3(CDIM + 3 + HC, 5), HC, 0, stream>>>(w->blob, acc, d_weights, *w);
    LSR_LAUNCHED(1);
    LSR_CUDA_CHECK(cudaGetLastError());
  }
  return LSR_OK;
}

}  // namespace lsr

#ifdef LSR_TRACE
extern "C" int lsr_debug_trace(long long* out) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, lsr::lsr_trace, sizeof(long long) * 3 * 512) == cudaSuccess ? 0 : 3;
}
extern "C" int lsr_debug_trace_ops(const LsrWeights* w, int64_t n_rays, int S, int flags, int g_cw, unsigned char* kinds) {
  static lsr::TrunkProgram P;
  lsr::build_trunk_program(w, lsr::saved_layout(n_rays, S, LSR_STAGE_COLOR, flags), g_cw != 0, (flags & LSR_FLAG_REL_POS) != 0, &P);
  for (int i = 0; i < P.n_ops; ++i) { kinds[2 * i] = P.ops[i].kind; kinds[2 * i + 1] = P.ops[i].flags; }
  return P.n_ops;
}
#endif
