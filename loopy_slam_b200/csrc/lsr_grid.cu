// Uniform hash-grid neighbour index: build (counting sort by cell) + exact radius-limited 8-NN.
// Replaces faiss GpuIndexIVFFlat train/add/search (/root/reference/src/neural_point.py:67-72,
// 1382-1392, 1623-1627, 1659-1708).  The grid is exact; FAISS IVF (nprobe 4 of nlist 400) is not.
#include "lsr_common.cuh"

namespace lsr {

__device__ __forceinline__ int32_t f2ord(float f) {   // order-preserving float -> int
  int32_t i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int32_t i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

__global__ void grid_init_kernel(GridHeader* h, int32_t n, int32_t max_cells) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    h->magic = GRID_MAGIC;
    h->n_points = n;
    h->max_cells = max_cells;
    for (int c = 0; c < 3; ++c) { h->bmin[c] = 0x7fffffff; h->bmax[c] = (int32_t)0x80000000; }
  }
}

__global__ void grid_bbox_kernel(const float* __restrict__ pos, int n, GridHeader* h) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = pos[3 * (size_t)i + c];
      mn[c] = fminf(mn[c], v);
      mx[c] = fmaxf(mx[c], v);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (mn[c] <= mx[c]) {
        atomicMin(&h->bmin[c], f2ord(mn[c]));
        atomicMax(&h->bmax[c], f2ord(mx[c]));
      }
    }
  }
}

__global__ void grid_setup_kernel(GridHeader* h, float cell) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float lo[3], hi[3];
  for (int c = 0; c < 3; ++c) {
    lo[c] = h->n_points > 0 ? ord2f(h->bmin[c]) : 0.f;
    hi[c] = h->n_points > 0 ? ord2f(h->bmax[c]) : 0.f;
    if (!(lo[c] <= hi[c])) { lo[c] = 0.f; hi[c] = 0.f; }   // NaN guard
  }
  float ce = fmaxf(cell, 1e-6f);
  int d[3];
  for (int it = 0; it < 64; ++it) {
    float inv = 1.0f / ce;
    double prod = 1.0;
    for (int c = 0; c < 3; ++c) {
      float span = __fmul_rn(__fsub_rn(hi[c], lo[c]), inv);
      d[c] = (span < 2.0e9f) ? (int)floorf(span) + 1 : 0x7fffffff;
      prod *= (double)d[c];
    }
    if (prod <= (double)h->max_cells) break;
    ce *= 2.0f;
  }
  h->cell = ce;
  h->inv_cell = 1.0f / ce;
  for (int c = 0; c < 3; ++c) { h->origin[c] = lo[c]; h->dims[c] = d[c]; }
  h->ncells = d[0] * d[1] * d[2];
}

__device__ __forceinline__ int point_cell(const GridHeader* h, float x, float y, float z) {
  int cx = min(max(cell_coord_raw(x, h->origin[0], h->inv_cell), 0), h->dims[0] - 1);
  int cy = min(max(cell_coord_raw(y, h->origin[1], h->inv_cell), 0), h->dims[1] - 1);
  int cz = min(max(cell_coord_raw(z, h->origin[2], h->inv_cell), 0), h->dims[2] - 1);
  return (cz * h->dims[1] + cy) * h->dims[0] + cx;
}

__global__ void grid_count_kernel(const float* __restrict__ pos, int n, const GridHeader* h,
                                  int32_t* counts, int32_t* cell_of_point) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = point_cell(h, pos[3 * (size_t)i], pos[3 * (size_t)i + 1], pos[3 * (size_t)i + 2]);
  cell_of_point[i] = c;
  atomicAdd(&counts[c], 1);
}

// ---- exclusive scan over (max_cells + 1) int32, three passes, 1024 elements per block
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[SCAN_BLOCK / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int w = lane < SCAN_BLOCK / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    if (lane < SCAN_BLOCK / 32) warp_sums[lane] = w;   // inclusive
  }
  __syncthreads();
  int warp_off = wid > 0 ? warp_sums[wid - 1] : 0;
  if (total) *total = warp_sums[SCAN_BLOCK / 32 - 1];
  __syncthreads();
  return warp_off + x - v;
}

__global__ void scan_pass1(int32_t* data, int n, int32_t* block_sums) {
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = (base + k < n) ? data[base + k] : 0; s += v[k]; }
  int tot;
  int off = block_exclusive_scan(s, &tot);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) data[base + k] = off;
    off += v[k];
  }
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void scan_pass2(int32_t* block_sums, int nb) {   // single block
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += SCAN_BLOCK) {
    int i = base + threadIdx.x;
    int v = i < nb ? block_sums[i] : 0;
    int tot;
    int off = block_exclusive_scan(v, &tot);
    int c = carry;
    if (i < nb) block_sums[i] = off + c;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + tot;
    __syncthreads();
  }
}

__global__ void scan_pass3(int32_t* data, int n, const int32_t* block_sums, int32_t* copy) {
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  const int add = block_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) {
      int v = data[base + k] + add;
      data[base + k] = v;
      copy[base + k] = v;
    }
  }
}

__global__ void grid_scatter_kernel(const float* __restrict__ pos, int n, const int32_t* cell_of_point,
                                    int32_t* cursor, float4* sorted) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int slot = atomicAdd(&cursor[cell_of_point[i]], 1);
  sorted[slot] = make_float4(pos[3 * (size_t)i], pos[3 * (size_t)i + 1], pos[3 * (size_t)i + 2],
                             __int_as_float(i));
}

// one warp per pair of queries (knn_warp_multi<2>); lane k < 8 holds the k-th neighbour
constexpr int KQ_NT = 256, KQ_NQ = 2;
__global__ void __launch_bounds__(KQ_NT) knn_query_kernel(const void* ws, const float* __restrict__ q,
                                                          const double* __restrict__ r_dyn, double r_fixed, int64_t P,
                                                          float* __restrict__ D, int64_t* __restrict__ I,
                                                          int32_t* __restrict__ nnum) {
  const GridHeader* h = (const GridHeader*)ws;
  const GridView g = grid_view(ws, h->n_points, h->max_cells);
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * KQ_NT + threadIdx.x) >> 5;
  const int64_t i0 = warp * KQ_NQ;
  if (i0 >= P) return;
  const bool dyn = r_dyn != nullptr;
  float px[KQ_NQ], py[KQ_NQ], pz[KQ_NQ], rr[KQ_NQ], r2f[KQ_NQ];
  double r2d[KQ_NQ];
  bool act[KQ_NQ];
  unsigned bD[KQ_NQ];
  int bI[KQ_NQ];
#pragma unroll
  for (int t = 0; t < KQ_NQ; ++t) {
    const int64_t i = i0 + t;
    act[t] = i < P;
    px[t] = py[t] = pz[t] = rr[t] = r2f[t] = 0.f;
    r2d[t] = 0.0;
    if (act[t]) {
      const double r = dyn ? r_dyn[i] : r_fixed;
      r2d[t] = r * r;
      r2f[t] = (float)r2d[t];
      rr[t] = (float)r * 1.00001f + 1e-7f;
      px[t] = q[3 * i]; py[t] = q[3 * i + 1]; pz[t] = q[3 * i + 2];
    }
  }
  __shared__ uint2 pend_s[KQ_NT / 32][KQ_NQ * KNN_PEND];
  knn_warp_multi<KQ_NQ>(g, px, py, pz, rr, act, dyn, r2f, r2d, bD, bI, pend_s[threadIdx.x >> 5]);
#pragma unroll
  for (int t = 0; t < KQ_NQ; ++t) {
    const int64_t i = i0 + t;
    if (!act[t]) continue;
    const bool ok = lane < KNN && bD[t] != KNN_INF;
    const float Dk = __uint_as_float(bD[t]);
    const bool strict = ok && (dyn ? ((double)Dk < r2d[t]) : (Dk < r2f[t]));
    const int ns = __popc(__ballot_sync(0xffffffffu, strict));
    if (lane < KNN) {
      D[i * KNN + lane] = ok ? Dk : FLT_MAX;
      I[i * KNN + lane] = ok ? (int64_t)bI[t] : (int64_t)-1;
    }
    if (lane == 0) nnum[i] = ns;
  }
}

}  // namespace lsr

using namespace lsr;

extern "C" int lsr_grid_workspace_bytes(int64_t n_points, int64_t max_cells, size_t* out_bytes) {
  if (!out_bytes || n_points < 0 || max_cells < 1 || max_cells > (1ll << 30) || n_points > (1ll << 30))
    return LSR_ERR_ARG;
  *out_bytes = grid_layout(n_points, max_cells).total;
  return LSR_OK;
}

extern "C" int lsr_grid_build(const float* cloud_pos, int64_t n, float cell, int64_t max_cells, void* ws,
                              size_t ws_bytes, lsr_stream_t stream) {
  if (!ws || n < 0 || (n > 0 && !cloud_pos) || !(cell > 0.f) || max_cells < 1 || max_cells > (1ll << 30) ||
      n > (1ll << 30))
    return LSR_ERR_ARG;
  const GridLayout L = grid_layout(n, max_cells);
  if (ws_bytes < L.total) return LSR_ERR_WORKSPACE;
  char* b = (char*)ws;
  GridHeader* h = (GridHeader*)(b + L.header);
  int32_t* cell_start = (int32_t*)(b + L.cell_start);
  int32_t* cursor = (int32_t*)(b + L.cursor);
  float4* sorted = (float4*)(b + L.sorted);
  int32_t* cop = (int32_t*)(b + L.cell_of_point);
  int32_t* bsum = (int32_t*)(b + L.block_sums);
  const int nscan = (int)(max_cells + 1);
  const int nb = (nscan + SCAN_TILE - 1) / SCAN_TILE;

  grid_init_kernel<<<1, 32, 0, stream>>>(h, (int32_t)n, (int32_t)max_cells);
  LSR_LAUNCHED(1);
  if (n > 0) {
    int blocks = (int)((n + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    grid_bbox_kernel<<<blocks, 256, 0, stream>>>(cloud_pos, (int)n, h);
    LSR_LAUNCHED(1);
  }
  grid_setup_kernel<<<1, 32, 0, stream>>>(h, cell);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaMemsetAsync(cell_start, 0, sizeof(int32_t) * (size_t)nscan, stream));
  if (n > 0) {
    grid_count_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(cloud_pos, (int)n, h, cell_start, cop);
    LSR_LAUNCHED(1);
  }
  scan_pass1<<<nb, SCAN_BLOCK, 0, stream>>>(cell_start, nscan, bsum);
  LSR_LAUNCHED(1);
  scan_pass2<<<1, SCAN_BLOCK, 0, stream>>>(bsum, nb);
  LSR_LAUNCHED(1);
  scan_pass3<<<nb, SCAN_BLOCK, 0, stream>>>(cell_start, nscan, bsum, cursor);
  LSR_LAUNCHED(1);
  if (n > 0) {
    grid_scatter_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(cloud_pos, (int)n, cop, cursor, sorted);
    LSR_LAUNCHED(1);
  }
  LSR_CUDA_CHECK(cudaGetLastError());
  return LSR_OK;
}

extern "C" int lsr_knn_query(const void* grid_ws, const float* q, const double* r_dyn, double r_fixed,
                             int64_t P, float* D, int64_t* I, int32_t* nnum, lsr_stream_t stream) {
  if (!grid_ws || P < 0 || (P > 0 && (!q || !D || !I || !nnum))) return LSR_ERR_ARG;
  if (P == 0) return LSR_OK;
  const int64_t per_block = (KQ_NT / 32) * KQ_NQ;
  knn_query_kernel<<<(unsigned)((P + per_block - 1) / per_block), KQ_NT, 0, stream>>>(grid_ws, q, r_dyn, r_fixed, P, D, I, nnum);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  return LSR_OK;
}
