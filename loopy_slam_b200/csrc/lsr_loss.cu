// Fused mapper / tracker losses of the render hot path: one pass over the (R,)
// render outputs produces the scalar loss terms AND the upstream gradients dL/d(depth), dL/d(rgb) that
// lsr_render_bwd consumes -- replacing the ~30 elementwise / reduction launches the reference's inline
// PyTorch expressions (src/Mapper.py:689-720, src/Tracker.py:171-191) cost per iteration.
//
// Sums are accumulated per thread in fp32, per block in fp64, and across blocks with fp64 atomics, so
// the result is order-independent to ~1e-16 relative; the last block to finish converts to fp32.
#include "lsr_common.cuh"

namespace lsr {

constexpr int LOSS_NT = 256;

struct LossScratch {   // 32 bytes at the head of the caller's scratch; zeroed by the entry point
  double sum[3];       // mapper: geo, colour, -   tracker: geo, colour, sum(tmp)
  unsigned int done;   // blocks finished (last-block finalisation)
  unsigned int pad;
};

__device__ __forceinline__ float sgnf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

template <int NV>
__device__ __forceinline__ void block_accumulate(const float (&v)[NV], double* dst) {
  __shared__ double sh[NV][LOSS_NT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    double x = (double)v[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sh[q][warp] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < LOSS_NT / 32; ++w) s += sh[threadIdx.x][w];
    atomicAdd(dst + threadIdx.x, s);
  }
}

// returns true in exactly one block: the last one to arrive, after every block's sums are visible
__device__ __forceinline__ bool last_block(LossScratch* sc) {
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&sc->done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) __threadfence();
  return last;
}

// ---- mapper: geo = sum |gt - depth| over (gt > 0 & valid & !nan(depth)); colour = sum |gt_rgb - rgb| over the
// same rays (src/Mapper.py:689-693,713-717); loss = geo + w_color * colour in stage 'color'.
__global__ void __launch_bounds__(LOSS_NT) mapper_loss_kernel(
    const float* __restrict__ depth, const float* __restrict__ rgb, const uint8_t* __restrict__ valid,
    const float* __restrict__ gt_depth, const float* __restrict__ gt_rgb, int64_t R, int use_color, float w_color,
    LossScratch* sc, float* __restrict__ loss3, float* __restrict__ d_depth, float* __restrict__ d_rgb) {
  float acc[2] = {0.f, 0.f};
  for (int64_t i = (int64_t)blockIdx.x * LOSS_NT + threadIdx.x; i < R; i += (int64_t)gridDim.x * LOSS_NT) {
    const float d = depth[i], g = gt_depth[i];
    const bool m = (g > 0.f) && (valid == nullptr || valid[i] != 0) && !isnan(d);
    const float e = g - d;
    acc[0] += m ? fabsf(e) : 0.f;
    d_depth[i] = m ? -sgnf(e) : 0.f;
    if (use_color) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float ec = gt_rgb[3 * i + c] - rgb[3 * i + c];
        acc[1] += m ? fabsf(ec) : 0.f;
        d_rgb[3 * i + c] = m ? -w_color * sgnf(ec) : 0.f;
      }
    } else if (d_rgb != nullptr) {
      d_rgb[3 * i] = 0.f; d_rgb[3 * i + 1] = 0.f; d_rgb[3 * i + 2] = 0.f;
    }
  }
  block_accumulate<2>(acc, sc->sum);
  if (last_block(sc) && threadIdx.x == 0) {
    const volatile double* s = sc->sum;
    const float geo = (float)s[0], col = (float)s[1];
    loss3[0] = use_color ? geo + w_color * col : geo;
    loss3[1] = geo;
    loss3[2] = col;
  }
}

// ---- tracker (src/Tracker.py:171-191).  Pass 1: tmp_i (the outlier statistic) and its sum.
__global__ void __launch_bounds__(LOSS_NT) tracker_resid_kernel(
    const float* __restrict__ depth, const float* __restrict__ var, const float* __restrict__ gt_depth, int64_t R,
    int handle_dynamic, LossScratch* sc, float* __restrict__ tmp) {
  float acc[1] = {0.f};
  for (int64_t i = (int64_t)blockIdx.x * LOSS_NT + threadIdx.x; i < R; i += (int64_t)gridDim.x * LOSS_NT) {
    float t = fabsf(gt_depth[i] - depth[i]);
    if (handle_dynamic) t = t / sqrtf(var[i] + 1e-10f);
    tmp[i] = t;
    acc[0] += t;    // NaN propagates like torch.mean
  }
  block_accumulate<1>(acc, sc->sum + 2);
}

// Pass 2: mask = (tmp < thr) & (gt > 0) & !nan(depth) & !nan(var); thr = 10*mean(tmp) (handle_dynamic) or the
// caller's 10*median(tmp) (*thr_in).  geo = sum clamp(|gt-depth|/sqrt(var+1e-10), 0, 1e3); colour = sum |gt_rgb-rgb|;
// loss = geo + w_color*colour when use_color.  var is detached (no gradient).
__global__ void __launch_bounds__(LOSS_NT) tracker_loss_kernel(
    const float* __restrict__ depth, const float* __restrict__ var, const float* __restrict__ rgb,
    const float* __restrict__ gt_depth, const float* __restrict__ gt_rgb, const float* __restrict__ tmp, int64_t R,
    const float* __restrict__ thr_in, int use_color, float w_color, LossScratch* sc, float* __restrict__ loss3,
    float* __restrict__ d_depth, float* __restrict__ d_rgb, uint8_t* __restrict__ mask_out) {
  const float thr = thr_in ? *thr_in : 10.f * ((float)(*(const volatile double*)(sc->sum + 2)) / (float)R);
  float acc[2] = {0.f, 0.f};
  for (int64_t i = (int64_t)blockIdx.x * LOSS_NT + threadIdx.x; i < R; i += (int64_t)gridDim.x * LOSS_NT) {
    const float d = depth[i], g = gt_depth[i], v = var[i];
    const bool m = (tmp[i] < thr) && (g > 0.f) && !isnan(d) && !isnan(v);
    const float e = g - d;
    const float inv = 1.f / sqrtf(v + 1e-10f);
    const float xr = fabsf(e) / sqrtf(v + 1e-10f);
    acc[0] += m ? fminf(fmaxf(xr, 0.f), 1e3f) : 0.f;
    // clamp passes the gradient on [min, max] (closed), zero outside
    d_depth[i] = (m && xr >= 0.f && xr <= 1e3f) ? -sgnf(e) * inv : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float ec = gt_rgb[3 * i + c] - rgb[3 * i + c];
      acc[1] += m ? fabsf(ec) : 0.f;
      d_rgb[3 * i + c] = (m && use_color) ? -w_color * sgnf(ec) : 0.f;
    }
    if (mask_out) mask_out[i] = m ? 1 : 0;
  }
  block_accumulate<2>(acc, sc->sum);
  if (last_block(sc) && threadIdx.x == 0) {
    const volatile double* s = sc->sum;
    const float geo = (float)s[0], col = (float)s[1];
    loss3[0] = use_color ? geo + w_color * col : geo;
    loss3[1] = geo;
    loss3[2] = col;
  }
}

static int loss_grid(int64_t R) {
  int64_t b = (R + LOSS_NT - 1) / LOSS_NT;
  return (int)(b < 1 ? 1 : (b > 148 * 4 ? 148 * 4 : b));
}

}  // namespace lsr

using namespace lsr;

extern "C" {

int lsr_loss_scratch_bytes(size_t* bytes) {
  if (!bytes) return LSR_ERR_ARG;
  *bytes = sizeof(LossScratch);
  return LSR_OK;
}

int lsr_mapper_loss(const float* depth, const float* rgb, const uint8_t* valid, const float* gt_depth,
                    const float* gt_rgb, int64_t n_rays, int stage, float w_color, void* scratch, float* loss3,
                    float* d_depth, float* d_rgb, lsr_stream_t stream) {
  if (n_rays < 0 || !scratch || !loss3) return LSR_ERR_ARG;
  if (stage != LSR_STAGE_GEOMETRY && stage != LSR_STAGE_COLOR) return LSR_ERR_ARG;
  const int use_color = stage == LSR_STAGE_COLOR;
  if (n_rays > 0 && (!depth || !gt_depth || !d_depth)) return LSR_ERR_ARG;
  if (n_rays > 0 && use_color && (!rgb || !gt_rgb || !d_rgb)) return LSR_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  LSR_CUDA_CHECK(cudaMemsetAsync(scratch, 0, sizeof(LossScratch), st));
  mapper_loss_kernel<<<loss_grid(n_rays), LOSS_NT, 0, st>>>(depth, rgb, valid, gt_depth, gt_rgb, n_rays, use_color,
                                                            w_color, (LossScratch*)scratch, loss3, d_depth, d_rgb);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  return LSR_OK;
}

int lsr_tracker_resid(const float* depth, const float* var, const float* gt_depth, int64_t n_rays,
                      int handle_dynamic, void* scratch, float* tmp, lsr_stream_t stream) {
  if (n_rays < 0 || !scratch) return LSR_ERR_ARG;
  if (n_rays > 0 && (!depth || !var || !gt_depth || !tmp)) return LSR_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  LSR_CUDA_CHECK(cudaMemsetAsync(scratch, 0, sizeof(LossScratch), st));
  tracker_resid_kernel<<<loss_grid(n_rays), LOSS_NT, 0, st>>>(depth, var, gt_depth, n_rays, handle_dynamic,
                                                              (LossScratch*)scratch, tmp);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  return LSR_OK;
}

int lsr_tracker_loss(const float* depth, const float* var, const float* rgb, const float* gt_depth,
                     const float* gt_rgb, const float* tmp, int64_t n_rays, const float* thr, int use_color,
                     float w_color, void* scratch, float* loss3, float* d_depth, float* d_rgb, uint8_t* mask_out,
                     lsr_stream_t stream) {
  if (n_rays < 0 || !scratch || !loss3) return LSR_ERR_ARG;
  if (n_rays > 0 && (!depth || !var || !rgb || !gt_depth || !gt_rgb || !tmp || !d_depth || !d_rgb)) return LSR_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  tracker_loss_kernel<<<loss_grid(n_rays), LOSS_NT, 0, st>>>(depth, var, rgb, gt_depth, gt_rgb, tmp, n_rays, thr,
                                                             use_color, w_color, (LossScratch*)scratch, loss3, d_depth,
                                                             d_rgb, mask_out);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  return LSR_OK;
}

}  // extern "C"
