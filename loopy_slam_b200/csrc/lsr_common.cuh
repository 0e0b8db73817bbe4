// Shared device/host definitions of the lsr library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include "../../include/lsr.h"

// every kernel launch of the library bumps this counter (lsr_launch_count; bench.py reports it as gpu_launches)
namespace lsr { extern long long g_launch_count; }
// cudaFuncAttributeMaxDynamicSharedMemorySize of a kernel, set once per device (it was ~2 us of host time per call)
#define LSR_SMEM_ATTR_ONCE(kernel, bytes)                                                                        \
  do {                                                                                                           \
    static unsigned long long done_mask_ = 0ull;                                                                 \
    int dev_ = 0;                                                                                                \
    LSR_CUDA_CHECK(cudaGetDevice(&dev_));                                                                        \
    if (!((done_mask_ >> (dev_ & 63)) & 1ull)) {                                                                 \
      LSR_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));   \
      done_mask_ |= 1ull << (dev_ & 63);                                                                         \
    }                                                                                                            \
  } while (0)
#define LSR_LAUNCHED(n) (__atomic_fetch_add(&lsr::g_launch_count, (long long)(n), __ATOMIC_RELAXED))

#define LSR_CUDA_CHECK(expr)                          \
  do {                                                \
    cudaError_t _e = (expr);                          \
    if (_e != cudaSuccess) return LSR_ERR_CUDA;       \
  } while (0)

namespace lsr {

constexpr int KNN = 8;        // pointcloud.nn_num
constexpr int CDIM = 32;      // model.c_dim
constexpr uint32_t GRID_MAGIC = 0x4c535247u;  // "LSRG"

// ---------------------------------------------------------------------------------- grid
// Device-resident header at the start of the grid workspace.  The host never reads it (no
// sync); every kernel that walks the grid loads it from global memory.
struct GridHeader {
  uint32_t magic;
  int32_t n_points;
  float origin[3];
  float cell;        // effective cell edge (>= requested)
  float inv_cell;
  int32_t dims[3];
  int32_t ncells;
  int32_t bmin[3];   // order-preserving int encodings of the bbox (atomicMin/Max targets)
  int32_t bmax[3];
  int32_t max_cells;
  int32_t pad[8];
};
static_assert(sizeof(GridHeader) <= 128, "grid header");

struct GridLayout {   // byte offsets inside the workspace, derived from (N, max_cells) only
  size_t header, cell_start, cursor, sorted, cell_of_point, block_sums, total;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline GridLayout grid_layout(int64_t n, int64_t max_cells) {
  GridLayout L;
  size_t o = 0;
  L.header = o;          o = align_up(o + sizeof(GridHeader), 256);
  L.cell_start = o;      o = align_up(o + sizeof(int32_t) * (size_t)(max_cells + 1), 256);
  L.cursor = o;          o = align_up(o + sizeof(int32_t) * (size_t)(max_cells + 1), 256);
  L.sorted = o;          o = align_up(o + sizeof(float4) * (size_t)(n > 0 ? n : 1), 256);
  L.cell_of_point = o;   o = align_up(o + sizeof(int32_t) * (size_t)(n > 0 ? n : 1), 256);
  L.block_sums = o;      o = align_up(o + sizeof(int32_t) * (size_t)((max_cells + 1) / 1024 + 2), 256);
  L.total = o;
  return L;
}

struct GridView {     // what the walking kernels need
  const GridHeader* hdr;
  const int32_t* cell_start;
  const float4* sorted;   // (x, y, z, int-bits id), counting-sorted by cell (x fastest)
};

__host__ __device__ inline GridView grid_view(const void* ws, int64_t n, int64_t max_cells) {
  GridLayout L = grid_layout(n, max_cells);
  const char* b = (const char*)ws;
  GridView g;
  g.hdr = (const GridHeader*)(b + L.header);
  g.cell_start = (const int32_t*)(b + L.cell_start);
  g.sorted = (const float4*)(b + L.sorted);
  return g;
}

#ifdef __CUDACC__
__device__ __forceinline__ int cell_coord_raw(float x, float origin, float inv_cell) {
  // monotone non-decreasing in x: each step (sub, mul, floor) is monotone under IEEE rounding
  return (int)floorf(__fmul_rn(__fsub_rn(x, origin), inv_cell));
}

// squared distance exactly as the oracle: ((dx*dx + dy*dy) + dz*dz), every op rounded (no FMA)
__device__ __forceinline__ float sqdist_rn(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Exact <=8 nearest points with D <= r^2 around a query, ordered by (D, id) lexicographically, found
// warp-cooperatively: lanes fetch the (cy,cz) rows of the cell neighbourhood in parallel (each row is one
// contiguous range of the counting-sorted point array), the candidates of all rows are enumerated 32 at a
// time (coalesced float4 loads inside a row), and the running top-8 lives in lanes 0..7 (lane k = k-th
// nearest).  Selection = 8 rounds of REDUX min over (D bits, id).
// dyn: compare in double against r2d (the reference evaluates D < r^2 in float64 when the radius is a
// float64 tensor); else in float against r2f.
constexpr unsigned KNN_INF = 0x7f800000u;   // +inf bits; squared distances are >= 0 so bit order == value order
constexpr int KNN_NOID = 0x7fffffff;
constexpr int KNN_PEND = 64;               // pending-candidate slots per query of knn_warp_multi (warp-private shared memory)

// Conservative lower bound of |p - x| over every x binned into cell c along one axis (0 when p may be
// inside).  Points land in cell c when floor(fl(fl(x - o) * inv)) == c, i.e. x in [o + c*cell, o + (c+1)*cell)
// up to a few roundings; `slack` covers those.
// The first / last cell of an axis also hold whatever the binning clamped into them, so they are
// unbounded on the outer side.
__device__ __forceinline__ float slab_dist_lb(float p, float o, float cell, int c, int dim) {
  const float lo = o + (float)c * cell, hi = lo + cell;
  const float slack = 1e-6f * (fabsf(p) + fabsf(o) + fabsf(hi - o)) + 1e-7f;
  const float below = c > 0 ? (lo - slack) - p : 0.f;         // > 0: p is below the slab
  const float above = c < dim - 1 ? p - (hi + slack) : 0.f;   // > 0: p is above the slab
  return fmaxf(0.f, fmaxf(below, above));
}

// NQ independent queries per warp, advanced in lockstep: the cell-range loads and the candidate loads of
// all NQ queries are issued back to back before any of them is consumed, so one global round trip is
// shared by NQ searches (the walk is a chain of dependent loads: header -> ranges -> candidates).
// Per-query results are identical to knn_warp.
template <int NQ>
__device__ __forceinline__ void knn_warp_multi(const GridView& g, const float (&px)[NQ], const float (&py)[NQ],
                                               const float (&pz)[NQ], const float (&rr)[NQ], const bool (&active)[NQ],
                                               bool dyn, const float (&r2f)[NQ], const double (&r2d)[NQ],
                                               unsigned (&bestD)[NQ], int (&bestI)[NQ], uint2* __restrict__ pend) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  // In-radius candidates that beat the current 8th best are only APPENDED to a warp-private list in shared memory
  // (pend: KNN_PEND entries per query); the 8-round selection -- by far the most expensive part of the walk, and it was run
  // for nearly every 32-candidate chunk -- runs when 32 of them are waiting or at the end: typically once per query
  // (a query ball holds ~15-25 points).  The result does not depend on the processing order: (D, id) is a total order.
  int npend[NQ];
  unsigned b7D[NQ];
  int b7I[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) { bestD[q] = KNN_INF; bestI[q] = KNN_NOID; npend[q] = 0; b7D[q] = KNN_INF; b7I[q] = KNN_NOID; }
  auto flush = [&](int q) {
    __syncwarp();
    for (int base = 0; base < npend[q]; base += 32) {
      unsigned cD = KNN_INF;
      int cI = KNN_NOID;
      if (base + lane < npend[q]) { const uint2 e = pend[q * KNN_PEND + base + lane]; cD = e.x; cI = (int)e.y; }
      unsigned oD = lane < KNN ? bestD[q] : KNN_INF;
      int oI = lane < KNN ? bestI[q] : KNN_NOID;
      unsigned nD = KNN_INF;
      int nI = KNN_NOID;
#pragma unroll
      for (int k = 0; k < KNN; ++k) {
        const bool c_lt = (cD < oD) || (cD == oD && cI < oI);
        const unsigned lD = c_lt ? cD : oD;
        const int lI = c_lt ? cI : oI;
        const unsigned mD = __reduce_min_sync(FULL, lD);
        const int mI = __reduce_min_sync(FULL, lD == mD ? lI : KNN_NOID);
        if (lane == k) { nD = mD; nI = mI; }
        if (cD == mD && cI == mI) { cD = KNN_INF; cI = KNN_NOID; }
        else if (oD == mD && oI == mI) { oD = KNN_INF; oI = KNN_NOID; }
      }
      bestD[q] = nD;
      bestI[q] = nI;
    }
    npend[q] = 0;
    b7D[q] = __shfl_sync(FULL, bestD[q], KNN - 1);
    b7I[q] = __shfl_sync(FULL, bestI[q], KNN - 1);
    __syncwarp();
  };
  const GridHeader* h = g.hdr;
  if (h->n_points <= 0) return;
  const float inv = h->inv_cell, ce = h->cell;
  const float ox = h->origin[0], oy = h->origin[1], oz = h->origin[2];
  const int dx = h->dims[0], dy = h->dims[1], dz = h->dims[2];
  int x0[NQ], x1[NQ], y0[NQ], z0[NQ], ny[NQ], nranges[NQ];
  int maxr = 0;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    x0[q] = max(cell_coord_raw(px[q] - rr[q], ox, inv), 0);
    x1[q] = min(cell_coord_raw(px[q] + rr[q], ox, inv), dx - 1);
    y0[q] = max(cell_coord_raw(py[q] - rr[q], oy, inv), 0);
    const int y1 = min(cell_coord_raw(py[q] + rr[q], oy, inv), dy - 1);
    z0[q] = max(cell_coord_raw(pz[q] - rr[q], oz, inv), 0);
    const int z1 = min(cell_coord_raw(pz[q] + rr[q], oz, inv), dz - 1);
    ny[q] = y1 - y0[q] + 1;
    const bool ok = active[q] && x0[q] <= x1[q] && y0[q] <= y1 && z0[q] <= z1;
    nranges[q] = ok ? ny[q] * (z1 - z0[q] + 1) : 0;
    maxr = max(maxr, nranges[q]);
  }
  for (int rc0 = 0; rc0 < maxr; rc0 += 32) {
    const int r = rc0 + lane;
    int beg[NQ], cnt[NQ], inc[NQ], T[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      beg[q] = 0;
      cnt[q] = 0;
      if (r < nranges[q]) {
        const int cz = z0[q] + r / ny[q], cy = y0[q] + r % ny[q];
        // rows of cells the query ball cannot reach are skipped, the others are clipped to the chord of
        // the ball at the row's minimum (y,z) offset: never drops a point with D <= r^2 (bounds are
        // conservative), only trims the candidate list
        const float dyl = slab_dist_lb(py[q], oy, ce, cy, dy), dzl = slab_dist_lb(pz[q], oz, ce, cz, dz);
        const float rem = rr[q] * rr[q] - dyl * dyl - dzl * dzl;
        if (rem >= 0.f) {
          const float hx = sqrtf(rem) * 1.000001f + 1e-7f;
          const int xa = max(x0[q], cell_coord_raw(px[q] - hx, ox, inv));
          const int xb = min(x1[q], cell_coord_raw(px[q] + hx, ox, inv));
          if (xa <= xb) {
            const int base = (cz * dy + cy) * dx;
            beg[q] = __ldg(g.cell_start + base + xa);
            cnt[q] = __ldg(g.cell_start + base + xb + 1);
          }
        }
      }
    }
    int maxT = 0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      cnt[q] -= beg[q];
      inc[q] = cnt[q];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, inc[q], o);
        if (lane >= o) inc[q] += v;
      }
      T[q] = __shfl_sync(FULL, inc[q], 31);
      maxT = max(maxT, T[q]);
    }
    for (int j0 = 0; j0 < maxT; j0 += 32) {
      const int j = j0 + lane;
      float4 q4[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        int rsel = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          const int v = __shfl_sync(FULL, inc[q], rsel + step - 1);
          if (v <= j) rsel += step;
        }
        rsel = min(rsel, 31);
        const int inc_sel = __shfl_sync(FULL, inc[q], rsel), cnt_sel = __shfl_sync(FULL, cnt[q], rsel);
        const int beg_sel = __shfl_sync(FULL, beg[q], rsel);
        q4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < T[q]) q4[q] = __ldg(g.sorted + beg_sel + (j - (inc_sel - cnt_sel)));
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        unsigned cD = KNN_INF;
        int cI = KNN_NOID;
        if (j < T[q]) {
          const float D = sqdist_rn(q4[q].x, q4[q].y, q4[q].z, px[q], py[q], pz[q]);
          const bool outside = dyn ? ((double)D > r2d[q]) : (D > r2f[q]);
          if (!outside) { cD = __float_as_uint(D); cI = __float_as_int(q4[q].w); }
        }
        const bool better = (cD < b7D[q]) || (cD == b7D[q] && cI < b7I[q]);   // cD == INF (outside / no candidate) never is
        const unsigned bal = __ballot_sync(FULL, better);
        if (bal) {
          if (better) pend[q * KNN_PEND + npend[q] + __popc(bal & ((1u << lane) - 1u))] = make_uint2(cD, (unsigned)cI);
          npend[q] += __popc(bal);
          if (npend[q] > 32) flush(q);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < NQ; ++q)
    if (npend[q] > 0) flush(q);
}
#endif  // __CUDACC__


}  // namespace lsr
