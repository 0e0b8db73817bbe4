// Shared device/host definitions of the lsr library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include "../../include/lsr.h"

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

struct Knn8 {
  float D[KNN];
  int I[KNN];
  int cnt;
};

__device__ __forceinline__ bool knn_less(float d, int i, float D, int I) {
  return d < D || (d == D && i < I);
}

// Exact <=8 nearest points with D <= r^2 around p, by (D, id) lexicographic order.
// dyn: compare in double against r2d (reference evaluates D < r^2 in float64 when the radius is a
// float64 tensor); else in float against r2f.
__device__ __forceinline__ void knn_walk(const GridView& g, float px, float py, float pz, float rr,
                                         bool dyn, float r2f, double r2d, Knn8& out) {
#pragma unroll
  for (int k = 0; k < KNN; ++k) { out.D[k] = INFINITY; out.I[k] = 0x7fffffff; }
  out.cnt = 0;
  const GridHeader* h = g.hdr;
  const int n = h->n_points;
  if (n <= 0) return;
  const float inv = h->inv_cell;
  const float ox = h->origin[0], oy = h->origin[1], oz = h->origin[2];
  const int dx = h->dims[0], dy = h->dims[1], dz = h->dims[2];
  int x0 = max(cell_coord_raw(px - rr, ox, inv), 0), x1 = min(cell_coord_raw(px + rr, ox, inv), dx - 1);
  int y0 = max(cell_coord_raw(py - rr, oy, inv), 0), y1 = min(cell_coord_raw(py + rr, oy, inv), dy - 1);
  int z0 = max(cell_coord_raw(pz - rr, oz, inv), 0), z1 = min(cell_coord_raw(pz + rr, oz, inv), dz - 1);
  if (x0 > x1 || y0 > y1 || z0 > z1) return;
  int cnt = 0;
  for (int cz = z0; cz <= z1; ++cz) {
    for (int cy = y0; cy <= y1; ++cy) {
      const int base = (cz * dy + cy) * dx;
      const int beg = __ldg(g.cell_start + base + x0);
      const int end = __ldg(g.cell_start + base + x1 + 1);
      for (int j = beg; j < end; ++j) {
        const float4 q = __ldg(g.sorted + j);
        const float D = sqdist_rn(q.x, q.y, q.z, px, py, pz);
        const bool outside = dyn ? ((double)D > r2d) : (D > r2f);
        const int id = __float_as_int(q.w);
        if (!outside && knn_less(D, id, out.D[KNN - 1], out.I[KNN - 1])) {
          out.D[KNN - 1] = D;
          out.I[KNN - 1] = id;
#pragma unroll
          for (int t = KNN - 1; t > 0; --t) {
            if (knn_less(out.D[t], out.I[t], out.D[t - 1], out.I[t - 1])) {
              float td = out.D[t]; out.D[t] = out.D[t - 1]; out.D[t - 1] = td;
              int ti = out.I[t];   out.I[t] = out.I[t - 1]; out.I[t - 1] = ti;
            }
          }
          cnt = min(cnt + 1, KNN);
        }
      }
    }
  }
  out.cnt = cnt;
}
// Warp-cooperative version of knn_walk: the 32 lanes of one warp search for ONE query point.
// Lanes fetch the (cy,cz) cell ranges in parallel, the candidates of all ranges are enumerated 32 at
// a time (coalesced float4 loads inside a range), and the running top-8 by (D, id) lives in lanes
// 0..7 (lane k = k-th nearest).  Selection = 8 rounds of REDUX min over (D bits, id).  Same exact
// arithmetic, predicate and tie-break as knn_walk / the oracle.
constexpr unsigned KNN_INF = 0x7f800000u;   // +inf bits; squared distances are >= 0 so bit order == value order
constexpr int KNN_NOID = 0x7fffffff;

__device__ __forceinline__ void knn_warp(const GridView& g, float px, float py, float pz, float rr, bool dyn,
                                         float r2f, double r2d, unsigned& bestD, int& bestI) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  bestD = KNN_INF;
  bestI = KNN_NOID;
  const GridHeader* h = g.hdr;
  if (h->n_points <= 0) return;
  const float inv = h->inv_cell;
  const float ox = h->origin[0], oy = h->origin[1], oz = h->origin[2];
  const int dx = h->dims[0], dy = h->dims[1], dz = h->dims[2];
  const int x0 = max(cell_coord_raw(px - rr, ox, inv), 0), x1 = min(cell_coord_raw(px + rr, ox, inv), dx - 1);
  const int y0 = max(cell_coord_raw(py - rr, oy, inv), 0), y1 = min(cell_coord_raw(py + rr, oy, inv), dy - 1);
  const int z0 = max(cell_coord_raw(pz - rr, oz, inv), 0), z1 = min(cell_coord_raw(pz + rr, oz, inv), dz - 1);
  if (x0 > x1 || y0 > y1 || z0 > z1) return;
  const int ny = y1 - y0 + 1, nranges = ny * (z1 - z0 + 1);
  for (int rc0 = 0; rc0 < nranges; rc0 += 32) {
    const int r = rc0 + lane;
    int beg = 0, cnt = 0;
    if (r < nranges) {
      const int cz = z0 + r / ny, cy = y0 + r % ny;
      const int base = (cz * dy + cy) * dx;
      beg = __ldg(g.cell_start + base + x0);
      cnt = __ldg(g.cell_start + base + x1 + 1) - beg;
    }
    int inc = cnt;   // inclusive prefix sum of the range sizes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL, inc, o);
      if (lane >= o) inc += v;
    }
    const int T = __shfl_sync(FULL, inc, 31);
    for (int j0 = 0; j0 < T; j0 += 32) {
      const int j = j0 + lane;
      // range of candidate j = number of ranges whose inclusive prefix is <= j: binary search over the
      // (non-decreasing) prefix held one-per-lane; lanes beyond the last range hold T > j
      int rsel = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int v = __shfl_sync(FULL, inc, rsel + step - 1);
        if (v <= j) rsel += step;
      }
      rsel = min(rsel, 31);
      const int inc_sel = __shfl_sync(FULL, inc, rsel), cnt_sel = __shfl_sync(FULL, cnt, rsel);
      const int beg_sel = __shfl_sync(FULL, beg, rsel);
      unsigned cD = KNN_INF;
      int cI = KNN_NOID;
      if (j < T) {
        const float4 q4 = __ldg(g.sorted + beg_sel + (j - (inc_sel - cnt_sel)));
        const float D = sqdist_rn(q4.x, q4.y, q4.z, px, py, pz);
        const bool outside = dyn ? ((double)D > r2d) : (D > r2f);
        if (!outside) { cD = __float_as_uint(D); cI = __float_as_int(q4.w); }
      }
      const unsigned b7D = __shfl_sync(FULL, bestD, KNN - 1);
      const int b7I = __shfl_sync(FULL, bestI, KNN - 1);
      if (!__any_sync(FULL, (cD < b7D) || (cD == b7D && cI < b7I))) continue;
      unsigned oD = lane < KNN ? bestD : KNN_INF;
      int oI = lane < KNN ? bestI : KNN_NOID;
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
      bestD = nD;
      bestI = nI;
    }
  }
}
#endif  // __CUDACC__

}  // namespace lsr
