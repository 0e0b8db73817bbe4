// Pixel -> ray sampling and quaternion pose kernels (rows a1/a2 of SURVEY.md section 8a).
//   get_samples / get_sample_uv / select_uv / get_rays_from_uv   /root/reference/src/common.py:104-138,160-172,237-259
//   quad2rotation / get_camera_from_tensor                       /root/reference/src/common.py:301-343
// The reference rebuilds a full H x W meshgrid per call and launches ~10 tiny kernels; here one
// launch turns n pixel indices into rays + gathered depth/colour.
#include <cstring>
#include "lsr_common.cuh"

namespace lsr {

__global__ void sample_rays_kernel(const float* __restrict__ depth_img, const float* __restrict__ color_img, int H,
                                   int W, float fx, float fy, float cx, float cy, const float* __restrict__ c2w,
                                   int ld, const int64_t* __restrict__ pix, int64_t n, int H0, int H1, int W0,
                                   int W1, float* __restrict__ rays_o, float* __restrict__ rays_d,
                                   float* __restrict__ depth, float* __restrict__ color, int64_t* __restrict__ i_out,
                                   int64_t* __restrict__ j_out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int ww = W1 - W0;
  int64_t p = pix[t];
  const int64_t tot = (int64_t)ww * (H1 - H0);
  p = p < 0 ? 0 : (p >= tot ? tot - 1 : p);   // select_uv clamps (common.py:131)
  const int h = (int)(p / ww), w = (int)(p - (int64_t)h * ww);
  const int col = W0 + w, row = H0 + h;
  const float fi = (float)col, fj = (float)row;
  // dirs = [(i-cx)/fx, -(j-cy)/fy, -1]   (common.py:113-114)
  const float d0 = __fdiv_rn(__fsub_rn(fi, cx), fx);
  const float d1 = -__fdiv_rn(__fsub_rn(fj, cy), fy);
  const float d2 = -1.0f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {   // rays_d = sum(dirs * R[a,:])  (common.py:117)
    const float r0 = c2w[a * ld + 0], r1 = c2w[a * ld + 1], r2 = c2w[a * ld + 2];
    rays_d[3 * t + a] = __fadd_rn(__fadd_rn(__fmul_rn(d0, r0), __fmul_rn(d1, r1)), __fmul_rn(d2, r2));
    rays_o[3 * t + a] = c2w[a * ld + 3];
  }
  const size_t lin = (size_t)row * W + col;
  if (depth) depth[t] = depth_img ? depth_img[lin] : 0.f;
  if (color && color_img) {
    color[3 * t + 0] = color_img[3 * lin + 0];
    color[3 * t + 1] = color_img[3 * lin + 1];
    color[3 * t + 2] = color_img[3 * lin + 2];
  }
  if (i_out) i_out[t] = col;
  if (j_out) j_out[t] = row;
}

// get_samples(..., depth_filter=True[, depth_limit]) in ONE launch: the n picks are sampled as above and the ones with
// depth > 0 (and < depth_limit) are COMPACTED in their original order (common.py:249-255 keeps the order: boolean-mask
// indexing).  One block walks the picks in chunks of its size with an order-preserving ballot / prefix scan; the number
// of kept samples goes to *count (the host reads it once to size the returned views instead of syncing on six
// boolean-mask gathers).
__global__ void __launch_bounds__(1024) sample_rays_filtered_kernel(
    const float* __restrict__ depth_img, const float* __restrict__ color_img, int H, int W, float fx, float fy, float cx, float cy,
    const float* __restrict__ c2w, int ld, const int64_t* __restrict__ pix, int64_t n, int H0, int H1, int W0, int W1,
    float depth_limit, float* __restrict__ rays_o, float* __restrict__ rays_d, float* __restrict__ depth,
    float* __restrict__ color, int64_t* __restrict__ i_out, int64_t* __restrict__ j_out, int32_t* __restrict__ count,
    unsigned long long* __restrict__ host_word, uint32_t ticket) {
  __shared__ int warp_cnt[32];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  const int ww = W1 - W0;
  const int64_t tot = (int64_t)ww * (H1 - H0);
  float R[12];
#pragma unroll
  for (int a = 0; a < 3; ++a) { R[4 * a] = c2w[a * ld]; R[4 * a + 1] = c2w[a * ld + 1]; R[4 * a + 2] = c2w[a * ld + 2]; R[4 * a + 3] = c2w[a * ld + 3]; }
  for (int64_t t0 = 0; t0 < n; t0 += blockDim.x) {
    const int64_t t = t0 + threadIdx.x;
    bool keep = false;
    int col = 0, row = 0;
    float dv = 0.f;
    size_t lin = 0;
    if (t < n) {
      int64_t p = pix[t];
      p = p < 0 ? 0 : (p >= tot ? tot - 1 : p);
      const int h = (int)(p / ww), w = (int)(p - (int64_t)h * ww);
      col = W0 + w; row = H0 + h;
      lin = (size_t)row * W + col;
      dv = depth_img[lin];
      keep = dv > 0.f && (!(depth_limit > 0.f) || dv < depth_limit);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int k = 0; k < nw; ++k) { const int c = warp_cnt[k]; if (k < warp) before += c; total += c; }
    const int base = base_s;
    if (keep) {
      const int64_t o = base + before + __popc(bal & ((1u << lane) - 1u));
      const float d0 = __fdiv_rn(__fsub_rn((float)col, cx), fx);
      const float d1 = -__fdiv_rn(__fsub_rn((float)row, cy), fy);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        rays_d[3 * o + a] = __fadd_rn(__fadd_rn(__fmul_rn(d0, R[4 * a]), __fmul_rn(d1, R[4 * a + 1])), __fmul_rn(-1.0f, R[4 * a + 2]));
        rays_o[3 * o + a] = R[4 * a + 3];
      }
      depth[o] = dv;
      if (color && color_img) {
        color[3 * o + 0] = color_img[3 * lin + 0]; color[3 * o + 1] = color_img[3 * lin + 1]; color[3 * o + 2] = color_img[3 * lin + 2];
      }
      i_out[o] = col;
      j_out[o] = row;
    }
    __syncthreads();
    if (threadIdx.x == 0) base_s = base + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (count) *count = base_s;
    // host-mapped result word of lsr_sample_rays_filtered_sync: (ticket << 32) | count in ONE 8-byte store, so the spinning
    // host thread can never see a count without its ticket
    if (host_word) *reinterpret_cast<volatile unsigned long long*>(host_word) = ((unsigned long long)ticket << 32) | (uint32_t)base_s;
  }
}

__global__ void sample_rays_bwd_kernel(const float* __restrict__ g_o, const float* __restrict__ g_d,
                                       const int64_t* __restrict__ ip, const int64_t* __restrict__ jp, int64_t n,
                                       float fx, float fy, float cx, float cy, float* __restrict__ d_c2w) {
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const float d0 = __fdiv_rn(__fsub_rn((float)ip[t], cx), fx);
    const float d1 = -__fdiv_rn(__fsub_rn((float)jp[t], cy), fy);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float gd = g_d ? g_d[3 * t + a] : 0.f;
      acc[a * 4 + 0] = fmaf(gd, d0, acc[a * 4 + 0]);
      acc[a * 4 + 1] = fmaf(gd, d1, acc[a * 4 + 1]);
      acc[a * 4 + 2] -= gd;
      acc[a * 4 + 3] += g_o ? g_o[3 * t + a] : 0.f;
    }
  }
#pragma unroll
  for (int k = 0; k < 12; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 12; ++k) atomicAdd(d_c2w + k, acc[k]);
  }
}

__global__ void pose_fwd_kernel(const float* __restrict__ cam, float* __restrict__ c2w) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float qr = cam[0], qi = cam[1], qj = cam[2], qk = cam[3];
  const float s = 2.0f / (qr * qr + qi * qi + qj * qj + qk * qk);   // common.py:313
  c2w[0] = 1.f - s * (qj * qj + qk * qk);
  c2w[1] = s * (qi * qj - qk * qr);
  c2w[2] = s * (qi * qk + qj * qr);
  c2w[3] = cam[4];
  c2w[4] = s * (qi * qj + qk * qr);
  c2w[5] = 1.f - s * (qi * qi + qk * qk);
  c2w[6] = s * (qj * qk - qi * qr);
  c2w[7] = cam[5];
  c2w[8] = s * (qi * qk - qj * qr);
  c2w[9] = s * (qj * qk + qi * qr);
  c2w[10] = 1.f - s * (qi * qi + qj * qj);
  c2w[11] = cam[6];
}

__global__ void pose_bwd_kernel(const float* __restrict__ cam, const float* __restrict__ G, float* __restrict__ d_cam) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float qr = cam[0], qi = cam[1], qj = cam[2], qk = cam[3];
  const float s = 2.0f / (qr * qr + qi * qi + qj * qj + qk * qk);
  const float G00 = G[0], G01 = G[1], G02 = G[2], G10 = G[4], G11 = G[5], G12 = G[6], G20 = G[8], G21 = G[9],
              G22 = G[10];
  // R = I + s*M(q);  dL/dq_x = -s^2 q_x <G,M> + s <G, dM/dq_x>
  const float GM = G00 * (-(qj * qj + qk * qk)) + G01 * (qi * qj - qk * qr) + G02 * (qi * qk + qj * qr) +
                   G10 * (qi * qj + qk * qr) + G11 * (-(qi * qi + qk * qk)) + G12 * (qj * qk - qi * qr) +
                   G20 * (qi * qk - qj * qr) + G21 * (qj * qk + qi * qr) + G22 * (-(qi * qi + qj * qj));
  const float dr = -G01 * qk + G02 * qj + G10 * qk - G12 * qi - G20 * qj + G21 * qi;
  const float di = G01 * qj + G02 * qk + G10 * qj - 2.f * G11 * qi - G12 * qr + G20 * qk + G21 * qr - 2.f * G22 * qi;
  const float dj = -2.f * G00 * qj + G01 * qi + G02 * qr + G10 * qi + G12 * qk - G20 * qr + G21 * qk - 2.f * G22 * qj;
  const float dk = -2.f * G00 * qk - G01 * qr + G02 * qi + G10 * qr - 2.f * G11 * qk + G12 * qj + G20 * qi + G21 * qj;
  d_cam[0] = -s * s * qr * GM + s * dr;
  d_cam[1] = -s * s * qi * GM + s * di;
  d_cam[2] = -s * s * qj * GM + s * dj;
  d_cam[3] = -s * s * qk * GM + s * dk;
  d_cam[4] = G[3];
  d_cam[5] = G[7];
  d_cam[6] = G[11];
}

// Per-frame dynamic radius maps: grey -> Sobel magnitude -> clip -> piecewise
// linear map to r_add / r_query, float64 like the reference's numpy path
// (/root/reference/src/Tracker.py:243-258, src/Mapper.py:854-869: skimage rgb2gray + sobel_h/sobel_v
// (reflect borders, smoothing [1,2,1]/4) + scipy interp1d on [0, 0.01, thr]).
template <typename T>
__global__ void dynamic_radius_kernel(const T* __restrict__ color, int H, int W, double thr, double r_add_max,
                                      double r_add_min, double ratio, double* __restrict__ r_add,
                                      double* __restrict__ r_query) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  auto refl = [](int i, int n) { return i < 0 ? -i - 1 : (i >= n ? 2 * n - 1 - i : i); };   // scipy 'reflect'
  double g[3][3];
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const size_t p = ((size_t)refl(y + dy, H) * W + refl(x + dx, W)) * 3;
      g[dy + 1][dx + 1] = 0.2125 * (double)color[p] + 0.7154 * (double)color[p + 1] + 0.0721 * (double)color[p + 2];
    }
  const double gy = ((g[2][0] - g[0][0]) + 2.0 * (g[2][1] - g[0][1]) + (g[2][2] - g[0][2])) * 0.25;
  const double gx = ((g[0][2] - g[0][0]) + 2.0 * (g[1][2] - g[1][0]) + (g[2][2] - g[2][0])) * 0.25;
  double m = sqrt(gx * gx + gy * gy);
  m = fmin(fmax(m, 0.0), thr);
  auto map = [&](double hi, double lo) {   // interp1d([0, 0.01, thr], [hi, hi, lo])
    if (m <= 0.01) return hi;
    return (lo - hi) / (thr - 0.01) * (m - 0.01) + hi;
  };
  const size_t o = (size_t)y * W + x;
  r_add[o] = map(r_add_max, r_add_min);
  r_query[o] = map(ratio * r_add_max, ratio * r_add_min);
}

// ------------------------------------------------------------------------------------------------ frustum selection
// Mapper.get_mask_from_c2w (/root/reference/src/Mapper.py:165-217): project every neural point into the current frame
// (x axis flipped, :188-189), look the sensor depth up bilinearly at (u, v) (cv2.remap INTER_LINEAR, constant 0 border),
// replace zero lookups by the maximum lookup (:210-211), keep points inside the edge-cropped image whose camera depth
// -z lies in [0, depth + 0.5].  Pass 1: projection + lookup + global max; pass 2: the mask.  float64 projection like the
// reference (numpy float64 points, float32 c2w inverse promoted).
struct FrustumArgs {
  const float* cloud; int64_t n;
  double w2c[12];
  const float* depth; int H, W;
  double fx, fy, cx, cy; int edge;
  float* u; float* v; float* negz; float* ds; unsigned int* maxbits; uint8_t* mask;
};
__device__ __forceinline__ float depth_at(const float* __restrict__ d, int H, int W, int x, int y) {
  return (x >= 0 && x < W && y >= 0 && y < H) ? d[(size_t)y * W + x] : 0.f;
}
__global__ void frustum_project_kernel(const __grid_constant__ FrustumArgs a) {
  float lmax = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    const double px = a.cloud[3 * i], py = a.cloud[3 * i + 1], pz = a.cloud[3 * i + 2];
    double cx_ = a.w2c[0] * px + a.w2c[1] * py + a.w2c[2] * pz + a.w2c[3];
    const double cy_ = a.w2c[4] * px + a.w2c[5] * py + a.w2c[6] * pz + a.w2c[7];
    const double cz_ = a.w2c[8] * px + a.w2c[9] * py + a.w2c[10] * pz + a.w2c[11];
    cx_ = -cx_;                                                       // :188-189
    const double z = cz_ + 1e-5;                                      // uv = K @ cam; z = uv[2] + 1e-5  (:190-191)
    const float u = (float)((a.fx * cx_ + a.cx * cz_) / z), v = (float)((a.fy * cy_ + a.cy * cz_) / z);   // :192-193 (float32 maps)
    // bilinear lookup, pixel centres at integer coordinates, zeros outside the image
    const float fu = floorf(u), fv = floorf(v);
    float dsv = 0.f;
    if (isfinite(u) && isfinite(v) && fu >= -1.f && fv >= -1.f && fu <= (float)a.W && fv <= (float)a.H) {
      const int x0 = (int)fu, y0 = (int)fv;
      const float ax = u - fu, ay = v - fv;
      const float d00 = depth_at(a.depth, a.H, a.W, x0, y0), d10 = depth_at(a.depth, a.H, a.W, x0 + 1, y0);
      const float d01 = depth_at(a.depth, a.H, a.W, x0, y0 + 1), d11 = depth_at(a.depth, a.H, a.W, x0 + 1, y0 + 1);
      dsv = (d00 * (1.f - ax) + d10 * ax) * (1.f - ay) + (d01 * (1.f - ax) + d11 * ax) * ay;
    }
    a.u[i] = u; a.v[i] = v; a.negz[i] = (float)(-z); a.ds[i] = dsv;
    lmax = fmaxf(lmax, dsv);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if ((threadIdx.x & 31) == 0) atomicMax(a.maxbits, __float_as_uint(lmax));     // depths >= 0: bit order == value order
}
__global__ void frustum_mask_kernel(const __grid_constant__ FrustumArgs a) {
  const float dmax = __uint_as_float(*a.maxbits);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    const float u = a.u[i], v = a.v[i], nz = a.negz[i];
    float d = a.ds[i];
    if (d == 0.f) d = dmax;                                           // :210-211
    const bool in = u < (float)(a.W - a.edge) && u > (float)a.edge && v < (float)(a.H - a.edge) && v > (float)a.edge;
    a.mask[i] = (in && 0.f <= nz && nz <= d + 0.5f) ? 1 : 0;          // :213
  }
}

}  // namespace lsr

using namespace lsr;

extern "C" int lsr_frustum_scratch_bytes(int64_t n_points, size_t* bytes) {
  if (n_points < 0 || !bytes) return LSR_ERR_ARG;
  *bytes = (size_t)(n_points > 0 ? n_points : 1) * 16 + 256;
  return LSR_OK;
}

extern "C" int lsr_frustum_mask(const float* cloud_pos, int64_t n_points, const double* w2c12_host, const float* depth_img,
                                int32_t H, int32_t W, double fx, double fy, double cx, double cy, int32_t edge, void* scratch,
                                uint8_t* mask_out, lsr_stream_t stream) {
  if (n_points < 0 || !w2c12_host || !depth_img || H <= 0 || W <= 0 || !scratch || (n_points > 0 && (!cloud_pos || !mask_out)))
    return LSR_ERR_ARG;
  if (n_points == 0) return LSR_OK;
  FrustumArgs a;
  a.cloud = cloud_pos; a.n = n_points;
  for (int i = 0; i < 12; ++i) a.w2c[i] = w2c12_host[i];
  a.depth = depth_img; a.H = H; a.W = W; a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy; a.edge = edge;
  char* sb = (char*)scratch;
  a.maxbits = (unsigned int*)sb;
  a.u = (float*)(sb + 256); a.v = a.u + n_points; a.negz = a.v + n_points; a.ds = a.negz + n_points;
  a.mask = mask_out;
  LSR_CUDA_CHECK(cudaMemsetAsync(a.maxbits, 0, sizeof(unsigned int), stream));
  int blocks = (int)((n_points + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  frustum_project_kernel<<<blocks, 256, 0, stream>>>(a);
  LSR_LAUNCHED(1);
  frustum_mask_kernel<<<blocks, 256, 0, stream>>>(a);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  return LSR_OK;
}

extern "C" int lsr_dynamic_radius(const float* color_f32, const double* color_f64, int32_t H, int32_t W, double thr,
                                  double r_add_max, double r_add_min, double ratio, double* r_add, double* r_query,
                                  lsr_stream_t stream) {
  if (H <= 0 || W <= 0 || (!color_f32 && !color_f64) || !r_add || !r_query || !(thr > 0.01)) return LSR_ERR_ARG;
  const dim3 blk(32, 8), grd((W + 31) / 32, (H + 7) / 8);
  if (color_f64)
    dynamic_radius_kernel<double><<<grd, blk, 0, stream>>>(color_f64, H, W, thr, r_add_max, r_add_min, ratio, r_add, r_query);
  else
    dynamic_radius_kernel<float><<<grd, blk, 0, stream>>>(color_f32, H, W, thr, r_add_max, r_add_min, ratio, r_add, r_query);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  return LSR_OK;
}

extern "C" int lsr_sample_rays(const float* depth_img, const float* color_img, int32_t H, int32_t W, float fx,
                               float fy, float cx, float cy, const float* c2w, int32_t c2w_ld, const int64_t* pix,
                               int64_t n, int32_t H0, int32_t H1, int32_t W0, int32_t W1, float* rays_o,
                               float* rays_d, float* depth, float* color, int64_t* i_out, int64_t* j_out,
                               lsr_stream_t stream) {
  if (n < 0 || H <= 0 || W <= 0 || H0 < 0 || W0 < 0 || H1 > H || W1 > W || H0 >= H1 || W0 >= W1 || c2w_ld < 4)
    return LSR_ERR_ARG;
  if (n == 0) return LSR_OK;
  if (!c2w || !pix || !rays_o || !rays_d) return LSR_ERR_ARG;
  sample_rays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(depth_img, color_img, H, W, fx, fy, cx, cy, c2w,
                                                                     c2w_ld, pix, n, H0, H1, W0, W1, rays_o, rays_d,
                                                                     depth, color, i_out, j_out);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  return LSR_OK;
}

extern "C" int lsr_sample_rays_filtered(const float* depth_img, const float* color_img, int32_t H, int32_t W, float fx,
                                        float fy, float cx, float cy, const float* c2w, int32_t c2w_ld, const int64_t* pix,
                                        int64_t n, int32_t H0, int32_t H1, int32_t W0, int32_t W1, float depth_limit,
                                        float* rays_o, float* rays_d, float* depth, float* color, int64_t* i_out,
                                        int64_t* j_out, int32_t* count, lsr_stream_t stream) {
  if (n < 0 || H <= 0 || W <= 0 || H0 < 0 || W0 < 0 || H1 > H || W1 > W || H0 >= H1 || W0 >= W1 || c2w_ld < 4 || !count)
    return LSR_ERR_ARG;
  if (n == 0) { LSR_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int32_t), stream)); return LSR_OK; }
  if (!depth_img || !c2w || !pix || !rays_o || !rays_d || !depth || !i_out || !j_out) return LSR_ERR_ARG;
  sample_rays_filtered_kernel<<<1, 1024, 0, stream>>>(depth_img, color_img, H, W, fx, fy, cx, cy, c2w, c2w_ld, pix, n, H0, H1,
                                                      W0, W1, depth_limit, rays_o, rays_d, depth, color, i_out, j_out, count,
                                                      nullptr, 0u);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  return LSR_OK;
}

// The reference's get_samples returns tensors whose LENGTH is the number of kept pixels (src/common.py:249-255), so every call
// has to bring one integer back to the host.  Instead of a D2H copy + stream synchronize (~20 us of driver work per call, twelve
// calls per mapping iteration), the kernel stores the count together with a per-call ticket into a host-mapped pinned word and
// this function spins on it.  Everything else the call wrote stays ordered by the stream as usual.
namespace {
struct HostWords {
  unsigned long long* w = nullptr;   // 64 words, cudaHostAllocPortable | cudaHostAllocMapped (UVA: same pointer on every device)
  unsigned int next = 0;
};
HostWords g_words;
}  // namespace

extern "C" int lsr_sample_rays_filtered_sync(const float* depth_img, const float* color_img, int32_t H, int32_t W, float fx,
                                             float fy, float cx, float cy, const float* c2w, int32_t c2w_ld,
                                             const int64_t* pix, int64_t n, int32_t H0, int32_t H1, int32_t W0, int32_t W1,
                                             float depth_limit, float* rays_o, float* rays_d, float* depth, float* color,
                                             int64_t* i_out, int64_t* j_out, int32_t* count_host, lsr_stream_t stream) {
  if (n < 0 || H <= 0 || W <= 0 || H0 < 0 || W0 < 0 || H1 > H || W1 > W || H0 >= H1 || W0 >= W1 || c2w_ld < 4 || !count_host)
    return LSR_ERR_ARG;
  *count_host = 0;
  if (n == 0) return LSR_OK;
  if (!depth_img || !c2w || !pix || !rays_o || !rays_d || !depth || !i_out || !j_out) return LSR_ERR_ARG;
  if (!g_words.w) {
    void* p = nullptr;
    LSR_CUDA_CHECK(cudaHostAlloc(&p, 64 * sizeof(unsigned long long), cudaHostAllocPortable | cudaHostAllocMapped));
    memset(p, 0, 64 * sizeof(unsigned long long));
    unsigned long long* expect = nullptr;
    if (!__atomic_compare_exchange_n(&g_words.w, &expect, (unsigned long long*)p, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE))
      cudaFreeHost(p);                                   // another thread won the race
  }
  const unsigned int t = __atomic_add_fetch(&g_words.next, 1u, __ATOMIC_RELAXED);
  const uint32_t ticket = t | 0x80000000u;               // never 0: a fresh (zeroed) word cannot match
  unsigned long long* word = g_words.w + (t & 63u);
  sample_rays_filtered_kernel<<<1, 1024, 0, stream>>>(depth_img, color_img, H, W, fx, fy, cx, cy, c2w, c2w_ld, pix, n, H0, H1,
                                                      W0, W1, depth_limit, rays_o, rays_d, depth, color, i_out, j_out, nullptr,
                                                      word, ticket);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  for (unsigned long long spins = 1;; ++spins) {
    const unsigned long long v = __atomic_load_n(word, __ATOMIC_ACQUIRE);
    if ((uint32_t)(v >> 32) == ticket) { *count_host = (int32_t)(uint32_t)v; return LSR_OK; }
    if ((spins & 0xfffu) == 0) {                          // a faulted kernel never writes its ticket: ask the stream
      const cudaError_t q = cudaStreamQuery(stream);
      if (q == cudaSuccess) {
        const unsigned long long v2 = __atomic_load_n(word, __ATOMIC_ACQUIRE);
        if ((uint32_t)(v2 >> 32) == ticket) { *count_host = (int32_t)(uint32_t)v2; return LSR_OK; }
        return LSR_ERR_CUDA;
      }
      if (q != cudaErrorNotReady) return LSR_ERR_CUDA;
    }
  }
}

extern "C" int lsr_sample_rays_bwd(const float* d_rays_o, const float* d_rays_d, const int64_t* i_pix,
                                   const int64_t* j_pix, int64_t n, float fx, float fy, float cx, float cy,
                                   float* d_c2w, lsr_stream_t stream) {
  if (n < 0 || !d_c2w) return LSR_ERR_ARG;
  LSR_CUDA_CHECK(cudaMemsetAsync(d_c2w, 0, 12 * sizeof(float), stream));
  if (n == 0) return LSR_OK;
  if (!i_pix || !j_pix) return LSR_ERR_ARG;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 296) blocks = 296;
  sample_rays_bwd_kernel<<<blocks, 256, 0, stream>>>(d_rays_o, d_rays_d, i_pix, j_pix, n, fx, fy, cx, cy, d_c2w);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  return LSR_OK;
}

extern "C" int lsr_pose_fwd(const float* cam7, float* c2w12, lsr_stream_t stream) {
  if (!cam7 || !c2w12) return LSR_ERR_ARG;
  pose_fwd_kernel<<<1, 32, 0, stream>>>(cam7, c2w12);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  return LSR_OK;
}

extern "C" int lsr_pose_bwd(const float* cam7, const float* d_c2w12, float* d_cam7, lsr_stream_t stream) {
  if (!cam7 || !d_c2w12 || !d_cam7) return LSR_ERR_ARG;
  pose_bwd_kernel<<<1, 32, 0, stream>>>(cam7, d_c2w12, d_cam7);
  LSR_LAUNCHED(1);
  LSR_CUDA_CHECK(cudaGetLastError());
  return LSR_OK;
}

namespace lsr { long long g_launch_count = 0; }
extern "C" int lsr_version(void) { return 200; }
// kernels launched by this library since the last reset (host-side counter; reset != 0 zeroes it after reading)
extern "C" long long lsr_launch_count(int reset) {
  return reset ? __atomic_exchange_n(&lsr::g_launch_count, 0ll, __ATOMIC_RELAXED) : __atomic_load_n(&lsr::g_launch_count, __ATOMIC_RELAXED);
}

extern "C" const char* lsr_strerror(int code) {
  switch (code) {
    case LSR_OK: return "ok";
    case LSR_ERR_ARG: return "invalid argument";
    case LSR_ERR_WORKSPACE: return "workspace too small";
    case LSR_ERR_CUDA: return "CUDA runtime error";
    case LSR_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
  }
}

extern "C" int lsr_device_sm_count(int* out) {
  if (!out) return LSR_ERR_ARG;
  int dev = 0, n = 0;
  LSR_CUDA_CHECK(cudaGetDevice(&dev));
  LSR_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  *out = n;
  return LSR_OK;
}
