/*
 * lsr.h -- C ABI of the B200-native Loopy-SLAM neural-point renderer ("lsr").
 *
 * The reference (eriksandstroem/Loopy-SLAM @ 95eb3bd) has no FFI for this path: its render hot
 * path is pure PyTorch + faiss-gpu.  This header is the boundary a maintainer would bind with
 * ctypes to replace, per entry point:
 *
 *   lsr_grid_build / lsr_knn_query   NeuralPointCloud index lifecycle + find_neighbors_faiss
 *                                    src/neural_point.py:67-72,1382-1392,1623-1627 / :1659-1708
 *   lsr_sample_rays(_bwd)            get_samples -> get_sample_uv -> select_uv -> get_rays_from_uv
 *                                    src/common.py:237-259,160-172,123-138,104-120
 *   lsr_pose_fwd / lsr_pose_bwd      get_camera_from_tensor / quad2rotation  src/common.py:301-343
 *   lsr_dynamic_radius               per-frame radius maps  src/Tracker.py:243-258, src/Mapper.py:854-869
 *   lsr_render_fwd                   Renderer.render_batch_ray + eval_points + NICER.forward +
 *                                    raw2outputs_nerf_color
 *                                    src/utils/Renderer.py:24-201, src/conv_onet/models/decoder.py:573-626,
 *                                    src/common.py:382-422
 *   lsr_render_bwd                   the autograd backward of the above (loss.backward(),
 *                                    src/Mapper.py:722, src/Tracker.py:193)
 *   lsr_mapper_loss                  mapper loss + its gradient wrt (depth, rgb)   src/Mapper.py:689-693,713-720
 *   lsr_tracker_resid / _loss        tracker outlier statistic, loss + gradient     src/Tracker.py:171-191
 *
 * Conventions: all pointers are DEVICE pointers unless stated; fp32 row-major contiguous;
 * every call is asynchronous on `stream`; return value 0 = LSR_OK, otherwise an error code for
 * lsr_strerror().  The library allocates nothing: workspaces are caller-supplied and sized by
 * the *_bytes queries.  Thread-safe per workspace; not across workspaces sharing one stream.
 * No CPU fallback exists: without a CUDA device every compute entry returns LSR_ERR_CUDA.
 */
#ifndef LSR_H_
#define LSR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* lsr_stream_t;   /* == cudaStream_t */

enum {
  LSR_OK = 0,
  LSR_ERR_ARG = 1,        /* bad argument (null pointer, unsupported dimension, misaligned offset) */
  LSR_ERR_WORKSPACE = 2,  /* workspace too small */
  LSR_ERR_CUDA = 3,       /* CUDA runtime error (launch failure, no device) */
  LSR_ERR_UNSUPPORTED = 4 /* configuration outside the supported envelope */
};

/* stage (decoder.py:592-626) */
enum { LSR_STAGE_GEOMETRY = 0, LSR_STAGE_COLOR = 1 };

/* LsrParams.flags */
enum {
  LSR_FLAG_REL_POS = 1,          /* model.encode_rel_pos_in_col   (decoder.py:477-485)          */
  LSR_FLAG_DYNAMIC_R = 2,        /* use_dynamic_radius: per-ray float64 radii (Appendix D)      */
  LSR_FLAG_SKIP_ZERO_DEPTH = 4,  /* rendering.skip_zero_depth_pixel (Renderer.py:199-200)       */
  LSR_FLAG_SAMPLE_NEAR_PCL = 8,  /* rendering.sample_near_pcl: zero-depth rays take their z from  */
                                 /*   z_zero_depth and keep their rendered depth (Renderer.py:150-158,197-198) */
  LSR_FLAG_FWD_ONLY = 32,        /* no backward will follow: lsr_render_workspace_bytes sizes `scratch` without the backward's
                                    hand-over planes (render_img: 0.6 GB instead of 4 GB at 680 x 1200) */
  LSR_FLAG_SAVE_LIGHT = 16       /* forward-only decode (eval_points): `saved` holds only the k-NN results, occupancy
                                    logits and per-sample colours (148 B / sample); lsr_render_bwd must not be called */
};

/* LsrParams.rgb_mode: what happens to the colour head output (decoder.py:534-546) */
enum {
  LSR_RGB_SIGMOID = 0,           /* rgb = sigmoid(out)                                           */
  LSR_RGB_RAW = 1,               /* encode_exposure, exposure_feat None: pre-sigmoid (:541-542)  */
  LSR_RGB_AFFINE_SIGMOID = 2     /* encode_exposure, tracker: sigmoid(out @ A + t) (:535-540)    */
};

/* grad_flags of lsr_render_bwd: which gradient sinks are wanted (needs_input_grad pruning) */
enum {
  LSR_GRAD_GEO_FEATS = 1,   /* d_geo_feats  (N, C)                                               */
  LSR_GRAD_COL_FEATS = 2,   /* d_col_feats  (N, C)                                               */
  LSR_GRAD_GEO_W = 4,       /* geometry decoder weights (off when mapping.fix_geo_decoder)       */
  LSR_GRAD_GEO_B = 8,       /* geo_decoder.embedder._B (trains even with fix_geo_decoder,        */
                            /*   src/Mapper.py:536-540)                                          */
  LSR_GRAD_COL_W = 16,      /* colour decoder weights incl. embedder_rel_pos._B                  */
  LSR_GRAD_RAYS = 32,       /* d_rays_o, d_rays_d (tracker / BA: pose gradient)                  */
  LSR_GRAD_AFFINE = 64      /* d_exposure_affine (12)                                            */
};

typedef struct LsrParams {
  int32_t n_surface;        /* S  rendering.N_surface (1..8; 5 in every shipped config)          */
  int32_t nn_num;           /* K  pointcloud.nn_num   (must be 8)                                */
  int32_t min_nn_num;       /*    pointcloud.min_nn_num                                          */
  int32_t c_dim;            /* C  model.c_dim         (must be 32)                               */
  float near_end_surface;   /*    rendering.near_end_surface                                     */
  float far_end_surface;    /*    rendering.far_end_surface                                      */
  float near_end;           /*    rendering.near_end  (zero-depth rays)                          */
  float sigmoid_coef;       /*    renderer.sigmoid_coefficient                                   */
  double radius_query;      /*    pointcloud.radius_query (fixed radius, squared in double)      */
  int32_t flags;            /*    LSR_FLAG_*                                                     */
  int32_t rgb_mode;         /*    LSR_RGB_*                                                      */
} LsrParams;

/*
 * Decoder weights: ONE flat fp32 device blob plus element offsets of every tensor inside it
 * (nn.Linear weights are (out,in) row-major exactly as in NICER.state_dict()).  Every offset
 * must be a multiple of 4 elements (16 B) -- the host mirror lays the blob out that way and makes
 * the nn.Parameters views into it.  Gradients are returned in a blob of identical layout.
 * Appendix B of SURVEY.md lists the reference tensors these correspond to.
 */
typedef struct LsrWeights {
  const float* blob;
  int64_t n_elems;
  /* geo_decoder (decoder.py:106-288), hidden 32, embedding 93 */
  int32_t g_fc_w[5], g_fc_b[5];      /* fc_c.i            (32,32) / (32)                          */
  int32_t g_B;                       /* embedder._B       (3,93)                                  */
  int32_t g_lin_w[5], g_lin_b[5];    /* pts_linears.i     (32,93)(32,32)(32,32)(32,125)(32,32)    */
  int32_t g_out_w, g_out_b;          /* output_linear     (1,32) / (1)                            */
  /* color_decoder (decoder.py:345-546), hidden 128, embedding 2*20 */
  int32_t c_fc_w[5], c_fc_b[5];      /* fc_c.i            (128,32) / (128)                        */
  int32_t c_B;                       /* embedder._B       (3,20)  non-persistent, not trained     */
  int32_t c_Brel;                    /* embedder_rel_pos._B (3,10)                                */
  int32_t c_nb1_w, c_nb1_b;          /* mlp_col_neighbor.linear1 (128,52) / (128)                 */
  int32_t c_nb2_w, c_nb2_b;          /* mlp_col_neighbor.linear2 (32,128) / (32)                  */
  int32_t c_lin_w[5], c_lin_b[5];    /* pts_linears.i     (128,40)(128,128)x2(128,168)(128,128)   */
  int32_t c_out_w, c_out_b;          /* output_linear     (3,128) / (3)                           */
} LsrWeights;

int lsr_version(void);
const char* lsr_strerror(int code);
/* number of SMs of the current device (grid sizing is done inside; exposed for benches) */
int lsr_device_sm_count(int* out);
/* Number of CUDA kernels this library has launched since the last reset (host-side counter; bench.py: gpu_launches). */
long long lsr_launch_count(int reset);

/* ---------------------------------------------------------------- neighbour index (hash grid)
 * Uniform grid over the cloud's bounding box, cell edge >= `cell` (doubled until the grid has
 * <= max_cells cells), points counting-sorted by cell (x fastest).  Rebuilt per cloud change
 * (insertion, PGO move); cheap enough for <= ~1e6 points. */
int lsr_grid_workspace_bytes(int64_t n_points, int64_t max_cells, size_t* out_bytes);
int lsr_grid_build(const float* cloud_pos /* N x 3 */, int64_t n_points, float cell, int64_t max_cells,
                   void* grid_ws, size_t grid_ws_bytes, lsr_stream_t stream);

/* == find_neighbors_faiss: EXACT <=K nearest cloud points with D <= r^2, ascending by (D, id);
 * D = (dx*dx + dy*dy) + dz*dz in fp32 without FMA; missing entries: I = -1, D = FLT_MAX;
 * nnum = #{D < r^2} (strict).  r_dyn (P float64, nullable) overrides r_fixed per query. */
int lsr_knn_query(const void* grid_ws, const float* q /* P x 3 */, const double* r_dyn, double r_fixed,
                  int64_t n_query, float* D /* P x 8 */, int64_t* I /* P x 8 */, int32_t* nnum /* P */,
                  lsr_stream_t stream);

/* ---------------------------------------------------------------- pixel -> ray sampling
 * pix: n int64 indices into the (H1-H0) x (W1-W0) window, row-major (what torch.randint gives
 * select_uv).  c2w: 3x4 row-major with row stride c2w_ld floats (4 for a (3,4)/(4,4) tensor).
 * Outputs: rays_o/rays_d (n,3), depth (n), color (n,3), i = column (n) and j = row (n) as int64. */
int lsr_sample_rays(const float* depth_img, const float* color_img, int32_t H, int32_t W, float fx,
                    float fy, float cx, float cy, const float* c2w, int32_t c2w_ld, const int64_t* pix,
                    int64_t n, int32_t H0, int32_t H1, int32_t W0, int32_t W1, float* rays_o,
                    float* rays_d, float* depth, float* color, int64_t* i_out, int64_t* j_out,
                    lsr_stream_t stream);
/* get_samples(..., depth_filter=True, depth_limit) (src/common.py:249-255) in one launch: as lsr_sample_rays, but only the
 * picks with depth > 0 (and < depth_limit when depth_limit > 0) are written, compacted in their original order; *count
 * (device int32) receives how many.  Output buffers must hold n entries. */
int lsr_sample_rays_filtered(const float* depth_img, const float* color_img, int32_t H, int32_t W, float fx,
                             float fy, float cx, float cy, const float* c2w, int32_t c2w_ld, const int64_t* pix,
                             int64_t n, int32_t H0, int32_t H1, int32_t W0, int32_t W1, float depth_limit,
                             float* rays_o, float* rays_d, float* depth, float* color, int64_t* i_out,
                             int64_t* j_out, int32_t* count, lsr_stream_t stream);
/* As lsr_sample_rays_filtered, but RETURNS the count to the host (*count_host, a plain host int): the kernel posts it to a
 * host-mapped pinned word and the call spins on that word instead of copying + synchronising the stream.  This is the size
 * the reference's get_samples needs on the host to shape its return values (src/common.py:249-259). */
int lsr_sample_rays_filtered_sync(const float* depth_img, const float* color_img, int32_t H, int32_t W, float fx,
                                  float fy, float cx, float cy, const float* c2w, int32_t c2w_ld, const int64_t* pix,
                                  int64_t n, int32_t H0, int32_t H1, int32_t W0, int32_t W1, float depth_limit,
                                  float* rays_o, float* rays_d, float* depth, float* color, int64_t* i_out,
                                  int64_t* j_out, int32_t* count_host, lsr_stream_t stream);
/* d_c2w (3x4 row-major, 12 floats, overwritten) from d_rays_o/d_rays_d and the pixel coords */
int lsr_sample_rays_bwd(const float* d_rays_o, const float* d_rays_d, const int64_t* i_pix,
                        const int64_t* j_pix, int64_t n, float fx, float fy, float cx, float cy,
                        float* d_c2w, lsr_stream_t stream);
/* camera tensor [qw qx qy qz tx ty tz] -> c2w 3x4 (unnormalised quaternion, two_s = 2/|q|^2) */
int lsr_pose_fwd(const float* cam7, float* c2w12, lsr_stream_t stream);
int lsr_pose_bwd(const float* cam7, const float* d_c2w12, float* d_cam7, lsr_stream_t stream);

/* Frustum feature selection (Mapper.get_mask_from_c2w, src/Mapper.py:165-217): mask_out[i] = 1 when point i projects
 * inside the edge-cropped image of the frame with world-to-camera matrix w2c (3x4 row-major, 12 HOST doubles) and its
 * camera depth lies in [0, bilinear sensor depth + 0.5] (zero lookups count as the maximum lookup).  scratch: device,
 * lsr_frustum_scratch_bytes(n_points). */
int lsr_frustum_scratch_bytes(int64_t n_points, size_t* bytes);
int lsr_frustum_mask(const float* cloud_pos, int64_t n_points, const double* w2c12_host, const float* depth_img,
                     int32_t H, int32_t W, double fx, double fy, double cx, double cy, int32_t edge, void* scratch,
                     uint8_t* mask_out, lsr_stream_t stream);

/* Per-frame dynamic radius maps (use_dynamic_radius, src/Tracker.py:243-258, src/Mapper.py:854-869):
 * grey -> Sobel magnitude -> clip to [0, thr] -> linear map on [0, 0.01, thr]; r_add / r_query are (H,W)
 * float64 like the reference's numpy path.  Exactly one of color_f32 / color_f64 ((H,W,3)) is non-NULL. */
int lsr_dynamic_radius(const float* color_f32, const double* color_f64, int32_t H, int32_t W, double thr,
                       double r_add_max, double r_add_min, double ratio, double* r_add, double* r_query,
                       lsr_stream_t stream);

/* ---------------------------------------------------------------- fused render
 * Workspace sizes: `saved` holds the activations the backward needs (0 rows -> forward only),
 * `scratch` the re-laid-out weights (tcgen05 chunk layout for the forward, three plain copies for the backward)
 * and the per-sample k-NN results the forward's two kernels hand over (120 B per sample row).  The SAME scratch
 * must be passed to lsr_render_bwd of that forward. */
int lsr_render_workspace_bytes(const LsrParams* prm, int64_t n_rays, int stage, size_t* saved_bytes,
                               size_t* scratch_bytes);

/* far[g] = min(5*mean(d), 1.2*max(d)) over rays [g*group, (g+1)*group)  (Renderer.py:104-121) */
int lsr_far_bound(const float* gt_depth, int64_t n_rays, int64_t group, float* far_out, lsr_stream_t stream);

/* The forward: weight re-layout, then z-sampling + grid k-NN (sample_knn_kernel), then IDW gather, geometry MLP,
 * (rel-pos neighbour MLP,) colour MLP, alpha compositing on tcgen05 / TMEM (render_fwd_kernel); three launches.  r_query: per-ray float64 radii when LSR_FLAG_DYNAMIC_R.
 * z_zero_depth (nullable; required with LSR_FLAG_SAMPLE_NEAR_PCL): (R, S) sample depths; row r is used when
 *   gt_depth[r] <= 0 (NeuralPointCloud.sample_near_pcl, src/neural_point.py:1734-1786), the other rows are ignored.
 * far_zero: far bound of the z-range used for rays with gt_depth <= 0, one value per group of
 *   far_group consecutive rays (Renderer.py:102-121 batch statistic; see lsr_far_bound); nullable.
 * exposure_affine: 12 floats [A row-major 3x3 | t] for LSR_RGB_AFFINE_SIGMOID.
 * saved == NULL -> inference (render_img); valid: 1 byte per ray.
 * row_remap (nullable, int32 per point): rows with row_remap[id] = j >= 0 read their features from the
 *   compact leaf blocks geo_leaf / col_leaf (n_sel,C) row j instead of the tables -- the optimised
 *   sub-block of src/Mapper.py:502-505, WITHOUT the per-iteration table[indices] = leaf index_put of
 *   src/Mapper.py:581-582 (and without the (N,C) gradient tables + gather of its backward). */
int lsr_render_fwd(const LsrParams* prm, const void* grid_ws, const float* cloud_pos, int64_t n_points,
                   const float* rays_o, const float* rays_d, const float* gt_depth,
                   const double* r_query, const float* far_zero, int64_t far_group, const float* z_zero_depth,
                   int64_t n_rays, const float* geo_feats, const float* col_feats, const int32_t* row_remap,
                   const float* geo_leaf, const float* col_leaf, const LsrWeights* w,
                   const float* exposure_affine, int stage, float* depth, float* var, float* rgb,
                   uint8_t* valid, void* saved, void* scratch, lsr_stream_t stream);

/* Host-logic introspection (no GPU needed; used by the CPU test-suite): statistics of the per-tile tensor-core
 * GEMM program lsr_render_fwd builds for (stage, flags): out[0..7] = {weight chunks (= bulk copies = mbarrier ring
 * steps) per tile, weight matrices, packed floats, chunks that wait for an operand hand-over, completions signalled on
 * accumulator barrier 0, on barrier 1, largest chunk in bytes, capacity of the packed-weight scratch in floats}. */
int lsr_debug_program_stats(const LsrWeights* w, int stage, int flags, int64_t* out);

/* Backward of lsr_render_fwd for upstream gradients g_depth (R), g_var (R, nullable), g_rgb (R,3).
 * is_tracker: neighbour weights depend on the sample position (decoder.py:191-198).
 * Gradient buffers must be ZEROED by the caller; they are accumulated into with atomics:
 * d_geo_feats/d_col_feats (N,C), d_weights (layout of w->blob), d_exposure_affine (12),
 * d_rays_o/d_rays_d (R,3, plain stores).  Unwanted sinks may be NULL (and unset in grad_flags).
 * With row_remap, d_geo_feats / d_col_feats are the (n_sel,C) gradients of geo_leaf / col_leaf; rows that
 * are not remapped are not trainable and receive nothing. 
 * `scratch` must be the buffer the matching lsr_render_fwd call used (it holds the k-NN results).  In the colour stage part
 * of the work runs on a library-owned side stream; everything is ordered on `stream` again when the call returns.
 */
int lsr_render_bwd(const LsrParams* prm, const void* grid_ws, const float* cloud_pos, int64_t n_points,
                   const float* rays_o, const float* rays_d, const float* gt_depth,
                   const double* r_query, int64_t n_rays, const float* geo_feats,
                   const float* col_feats, const int32_t* row_remap, const float* geo_leaf,
                   const float* col_leaf, const LsrWeights* w, const float* exposure_affine,
                   int stage, int is_tracker, const void* saved, void* scratch, const float* g_depth,
                   const float* g_var, const float* g_rgb, int grad_flags, float* d_geo_feats,
                   float* d_col_feats, float* d_weights, float* d_exposure_affine, float* d_rays_o,
                   float* d_rays_d, lsr_stream_t stream);

/* ---------------------------------------------------------------- fused losses
 * One pass over the (R,) render outputs -> loss terms AND the upstream gradients lsr_render_bwd consumes.
 * loss3 = [loss, geo_loss, color_loss] (fp32, device).  d_depth (R) / d_rgb (R,3) are dloss/d(depth|rgb)
 * for an upstream gradient of 1 (the caller scales them by grad_output).  scratch: lsr_loss_scratch_bytes
 * bytes, owned by the call sequence (mapper: one call; tracker: resid then loss). */
int lsr_loss_scratch_bytes(size_t* bytes);

/* src/Mapper.py:689-693,713-720: mask = gt_depth > 0 & valid & !isnan(depth) (valid nullable = all true);
 * geo = sum |gt_depth - depth|, color = sum |gt_rgb - rgb| over the mask; loss = geo (+ w_color * color when
 * stage == LSR_STAGE_COLOR; rgb / gt_rgb / d_rgb may be NULL in stage geometry). */
int lsr_mapper_loss(const float* depth, const float* rgb, const uint8_t* valid, const float* gt_depth,
                    const float* gt_rgb, int64_t n_rays, int stage, float w_color, void* scratch, float* loss3,
                    float* d_depth, float* d_rgb, lsr_stream_t stream);

/* src/Tracker.py:175-180: tmp = |gt_depth - depth| (/ sqrt(var + 1e-10) when handle_dynamic), written to
 * tmp (R); its sum is kept in scratch for the 10*mean(tmp) threshold of lsr_tracker_loss. */
int lsr_tracker_resid(const float* depth, const float* var, const float* gt_depth, int64_t n_rays,
                      int handle_dynamic, void* scratch, float* tmp, lsr_stream_t stream);

/* src/Tracker.py:171-191: mask = tmp < thr & gt_depth > 0 & !isnan(depth) & !isnan(var), thr = *thr (device,
 * e.g. 10*median(tmp)) or 10*mean(tmp) from lsr_tracker_resid when thr == NULL; geo = sum clamp(|gt_depth -
 * depth| / sqrt(var + 1e-10), 0, 1e3), color = sum |gt_rgb - rgb|; loss = geo + (use_color ? w_color*color : 0).
 * var is detached (no gradient).  mask_out: nullable, 1 byte per ray. */
int lsr_tracker_loss(const float* depth, const float* var, const float* rgb, const float* gt_depth,
                     const float* gt_rgb, const float* tmp, int64_t n_rays, const float* thr, int use_color,
                     float w_color, void* scratch, float* loss3, float* d_depth, float* d_rgb, uint8_t* mask_out,
                     lsr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LSR_H_ */
