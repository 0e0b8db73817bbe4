/* Plain-C restatement of the radius-limited exact k-NN used as the CPU neighbour search of the CPU
 * baseline and as an independent cross-check of the CUDA grid walk.  TEST INFRASTRUCTURE (see
 * oracle/__init__.py).  Contract = find_neighbors_faiss (/root/reference/src/neural_point.py:1659-1708)
 * restricted to the entries that matter downstream: the <= K nearest cloud points with D <= r^2,
 * ascending by (D, id); D = (dx*dx + dy*dy) + dz*dz in float32 (compile with -ffp-contract=off);
 * missing entries I = -1, D = FLT_MAX; nnum = #{D < r^2}.  r2 is passed per query as double when
 * `r_dyn` is given (float64 radii, compared in double), else r_fixed^2 rounded to float.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC knn_grid.c -o liboracle_knn.so
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define K 8

typedef struct {
  float origin[3], cell, inv;
  int dims[3];
  int64_t n;
  int32_t* start; /* ncells + 1 */
  int32_t* order; /* point ids sorted by cell */
  const float* pos;
} Grid;

static int coord(float x, float o, float inv) { return (int)floorf((x - o) * inv); }

Grid* oracle_grid_build(const float* pos, int64_t n, float cell) {
  Grid* g = (Grid*)calloc(1, sizeof(Grid));
  float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  for (int c = 0; c < 3; ++c) { lo[c] = FLT_MAX; hi[c] = -FLT_MAX; }
  for (int64_t i = 0; i < n; ++i)
    for (int c = 0; c < 3; ++c) {
      float v = pos[3 * i + c];
      if (v < lo[c]) lo[c] = v;
      if (v > hi[c]) hi[c] = v;
    }
  if (n == 0) for (int c = 0; c < 3; ++c) { lo[c] = 0; hi[c] = 0; }
  for (;;) {
    double prod = 1;
    for (int c = 0; c < 3; ++c) {
      g->dims[c] = (int)floorf((hi[c] - lo[c]) / cell) + 1;
      prod *= g->dims[c];
    }
    if (prod <= (double)(1 << 26)) break;
    cell *= 2;
  }
  g->cell = cell; g->inv = 1.0f / cell; g->n = n; g->pos = pos;
  for (int c = 0; c < 3; ++c) g->origin[c] = lo[c];
  int64_t nc = (int64_t)g->dims[0] * g->dims[1] * g->dims[2];
  g->start = (int32_t*)calloc(nc + 1, sizeof(int32_t));
  g->order = (int32_t*)malloc(sizeof(int32_t) * (n > 0 ? n : 1));
  int32_t* cof = (int32_t*)malloc(sizeof(int32_t) * (n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) {
    int cc[3];
    for (int c = 0; c < 3; ++c) {
      cc[c] = coord(pos[3 * i + c], g->origin[c], g->inv);
      if (cc[c] < 0) cc[c] = 0;
      if (cc[c] >= g->dims[c]) cc[c] = g->dims[c] - 1;
    }
    cof[i] = (cc[2] * g->dims[1] + cc[1]) * g->dims[0] + cc[0];
    g->start[cof[i] + 1]++;
  }
  for (int64_t c = 0; c < nc; ++c) g->start[c + 1] += g->start[c];
  int32_t* cur = (int32_t*)malloc(sizeof(int32_t) * (nc + 1));
  memcpy(cur, g->start, sizeof(int32_t) * (nc + 1));
  for (int64_t i = 0; i < n; ++i) g->order[cur[cof[i]]++] = (int32_t)i;
  free(cur); free(cof);
  return g;
}

void oracle_grid_free(Grid* g) {
  if (!g) return;
  free(g->start); free(g->order); free(g);
}

static int less(float d, int64_t i, float D, int64_t I) { return d < D || (d == D && i < I); }

void oracle_knn_query(const Grid* g, const float* q, const double* r_dyn, double r_fixed, int64_t P, float* D,
                      int64_t* I, int32_t* nnum) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t p = 0; p < P; ++p) {
    float bd[K]; int64_t bi[K]; int cnt = 0;
    for (int k = 0; k < K; ++k) { bd[k] = INFINITY; bi[k] = INT64_MAX; }
    const double r = r_dyn ? r_dyn[p] : r_fixed;
    const double r2d = r * r; const float r2f = (float)r2d;
    const float rr = (float)r * 1.00001f + 1e-7f;
    const float px = q[3 * p], py = q[3 * p + 1], pz = q[3 * p + 2];
    int lo[3], hi[3]; const float pp[3] = {px, py, pz};
    int empty = g->n == 0;
    for (int c = 0; c < 3; ++c) {
      lo[c] = coord(pp[c] - rr, g->origin[c], g->inv); if (lo[c] < 0) lo[c] = 0;
      hi[c] = coord(pp[c] + rr, g->origin[c], g->inv); if (hi[c] > g->dims[c] - 1) hi[c] = g->dims[c] - 1;
      if (lo[c] > hi[c]) empty = 1;
    }
    if (!empty)
      for (int cz = lo[2]; cz <= hi[2]; ++cz)
        for (int cy = lo[1]; cy <= hi[1]; ++cy) {
          int base = (cz * g->dims[1] + cy) * g->dims[0];
          for (int j = g->start[base + lo[0]]; j < g->start[base + hi[0] + 1]; ++j) {
            int64_t id = g->order[j];
            float dx = g->pos[3 * id] - px, dy = g->pos[3 * id + 1] - py, dz = g->pos[3 * id + 2] - pz;
            float d = (dx * dx + dy * dy) + dz * dz;
            int outside = r_dyn ? ((double)d > r2d) : (d > r2f);
            if (!outside && less(d, id, bd[K - 1], bi[K - 1])) {
              bd[K - 1] = d; bi[K - 1] = id;
              for (int t = K - 1; t > 0 && less(bd[t], bi[t], bd[t - 1], bi[t - 1]); --t) {
                float td = bd[t]; bd[t] = bd[t - 1]; bd[t - 1] = td;
                int64_t ti = bi[t]; bi[t] = bi[t - 1]; bi[t - 1] = ti;
              }
              if (cnt < K) cnt++;
            }
          }
        }
    int ns = 0;
    for (int k = 0; k < K; ++k) {
      int ok = k < cnt;
      D[p * K + k] = ok ? bd[k] : FLT_MAX;
      I[p * K + k] = ok ? bi[k] : -1;
      if (ok) ns += r_dyn ? ((double)bd[k] < r2d) : (bd[k] < r2f);
    }
    nnum[p] = ns;
  }
}
