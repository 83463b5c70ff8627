/*
 * up3d.h -- C ABI of libunipre3d_b200.so: the B200 (sm_100a) render-loss hot path of UniPre3D.
 *
 * Drop-in boundary.  Plain pointers + sizes + a cudaStream_t; no torch types.
 * Every entry point returns 0 on success and a non-zero status on error, with a human-readable
 * message available from up3d_last_error() (thread-local).  The reference's own native ABI never
 * reports errors (it exit(-1)s the process: pointnet2_batch/src/sampling_gpu.cu:46-50,
 * ball_query.cpp:14-26); the external rasterizer throws std::runtime_error -- the Python host
 * layer turns a non-zero status into RuntimeError to keep that behaviour.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the name ends in _h.  The caller owns all memory,
 *     including workspaces whose size comes from the *_bytes() queries.  No hidden
 *     synchronisation, no allocation: every call only enqueues work on `stream`.
 *   - Thread-safe for distinct streams and distinct buffers.
 *   - float = IEEE fp32, indices int32, row-vector 4x4 matrices flat[16] exactly as the reference
 *     passes them (dataset/shapenet.py:303-316).
 */
#ifndef UP3D_H
#define UP3D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *up3d_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define UP3D_API __attribute__((visibility("default")))
#else
#define UP3D_API
#endif

UP3D_API const char *up3d_last_error(void);
UP3D_API int up3d_version(void);

/* ------------------------------------------------------------------------------------------
 * Point ops.  Replace the pybind functions of
 * /root/reference/openpoints/cpp/pointnet2_batch/src/pointnet2_api.cpp:10-24 .
 * ---------------------------------------------------------------------------------------- */

/* replaces furthest_point_sampling_wrapper (sampling.cpp:39-48; kernel sampling_gpu.cu:100-216).
 * xyz (B,N,3) -> idx (B,M) int32.  Index sequence (first index 0, tie-breaking of the reference's
 * block reduction for its block size opt_n_threads(N)) is reproduced bit-for-bit.
 * `temp` (B,N) floats is scratch (the reference's 1e10-filled buffer, subsample.py:93); it is
 * initialised internally and may be NULL when N <= up3d_fps_max_resident_points(). */
UP3D_API int up3d_fps(int B, int N, int M, const float *xyz, float *temp, int32_t *idx, up3d_stream_t stream);
UP3D_API int up3d_fps_max_resident_points(void);

/* replaces ball_query_wrapper_fast (ball_query.cpp:29-39; kernel ball_query_gpu.cu:15-51).
 * new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample) int32: first nsample indices (ascending)
 * with d^2 < radius^2, padded with the first hit, all-zero when there is no hit. */
UP3D_API int up3d_ball_query(int B, int N, int M, float radius, int nsample, const float *new_xyz, const float *xyz,
                    int32_t *idx, up3d_stream_t stream);

/* replaces group_points_wrapper_fast / group_points_grad_wrapper_fast (group_points.cpp:13-35;
 * kernels group_points_gpu.cu:14-72).  points (B,C,N), idx (B,M,K) -> out (B,C,M,K);
 * grad_out (B,C,M,K) -> grad_points (B,C,N) (zero-filled by the call, then scatter-added). */
UP3D_API int up3d_group_points(int B, int C, int N, int M, int K, const float *points, const int32_t *idx, float *out,
                      up3d_stream_t stream);
UP3D_API int up3d_group_points_grad(int B, int C, int N, int M, int K, const float *grad_out, const int32_t *idx,
                           float *grad_points, up3d_stream_t stream);

/* replaces gather_points_wrapper_fast / gather_points_grad_wrapper_fast (sampling.cpp:9-36;
 * kernels sampling_gpu.cu:15-70).  points (B,C,N), idx (B,M) -> out (B,C,M). */
UP3D_API int up3d_gather_points(int B, int C, int N, int M, const float *points, const int32_t *idx, float *out,
                       up3d_stream_t stream);
UP3D_API int up3d_gather_points_grad(int B, int C, int N, int M, const float *grad_out, const int32_t *idx,
                            float *grad_points, up3d_stream_t stream);

/* k nearest neighbours of each query among xyz (pointMLP's knn_point, openpoints/models/backbone/pointmlp.py:102-113,
 * and -- with K = 3 and dist != NULL -- the 3-NN search of PointNetFeaturePropagation, pointmlp.py:397-403, which
 * the reference does with a FULL sort of the (N x S) distance matrix).  xyz (B,N,D), query (B,S,D), D = 3 or 4
 * (pointMLP with in_channels = 4 measures distances over xyz+height) -> idx (B,S,K) int32 in ascending distance,
 * dist (B,S,K) squared distances |q|^2+|p|^2-2q.p or NULL.  K <= 64. */
UP3D_API int up3d_knn(int B, int N, int S, int K, int D, const float *xyz, const float *query, int32_t *idx,
                      float *dist, up3d_stream_t stream);

/* Fused SubsampleGroup tail (openpoints/models/layers/group_embed.py:39-57 + group.py:235-255):
 * centres = xyz[fps_idx]; idx = ball_query(radius, K, xyz, centres);
 * neighborhood[b, :, g, k] = xyz[b, idx[b,g,k], :] - centres[b, g, :]   -> (B,3,G,K)
 * One launch instead of gather + ball_query + transpose + group + subtract.
 * center (B,G,3) and idx (B,G,K) are also written (idx may be NULL). */
UP3D_API int up3d_subsample_group(int B, int N, int G, int K, float radius, const float *xyz, const int32_t *fps_idx,
                         float *center, float *neighborhood, int32_t *idx, up3d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Differentiable 3DGS rasterizer, batched over views.
 * Replaces _C.rasterize_gaussians / _C.rasterize_gaussians_backward of the external
 * diff_gaussian_rasterization package the reference binds at gaussian_renderer/__init__.py:8
 * (called per (object, view) from train_network.py:418-442).  One call renders ALL views of
 * ALL Gaussian sets (objects / scenes) of a step.
 *
 * Gaussian sets are concatenated: set s owns Gaussians [set_offsets[s], set_offsets[s+1]).
 * Views are grouped by set: set s owns views [set_view_start[s], set_view_start[s+1]).
 * "Records" are (view, Gaussian-of-its-set) pairs: view v owns records
 * [view_rec_start[v], view_rec_start[v+1]); n_records = view_rec_start[n_views].
 * ---------------------------------------------------------------------------------------- */
typedef struct up3d_raster_desc {
    int32_t n_sets;
    int32_t n_views;
    int32_t n_gaussians;   /* total over sets */
    int32_t n_records;     /* total (view, gaussian) pairs */
    int32_t max_set_size;  /* max Gaussians in one set */
    int32_t width, height;
    int32_t sh_degree;     /* active degree D (cfg.model.max_sh_degree) */
    int32_t sh_coeffs;     /* M = shs.size(1); 0 when colors_precomp is used */
    int32_t antialiasing;  /* reference passes True (gaussian_renderer/__init__.py:58) */
    float tanfovx, tanfovy;
    float scale_modifier;
    /* device int32 index arrays described above */
    const int32_t *set_offsets;     /* n_sets + 1 */
    const int32_t *set_view_start;  /* n_sets + 1 */
    const int32_t *view_set;        /* n_views */
    const int32_t *view_rec_start;  /* n_views + 1 */
} up3d_raster_desc;

/* bytes of the opaque forward state the caller must keep alive until the backward call
 * (the role of upstream's geomBuffer / binningBuffer / imgBuffer). */
UP3D_API size_t up3d_raster_state_bytes(const up3d_raster_desc *d);
/* bytes of transient scratch for forward / backward (may be reused across calls). */
UP3D_API size_t up3d_raster_scratch_bytes(const up3d_raster_desc *d);

/* Forward.  Inputs: means3D (n_gaussians,3), shs (n_gaussians,M,3) or NULL, colors_precomp
 * (n_gaussians,3) or NULL (exactly one), opacities (n_gaussians), scales (n_gaussians,3),
 * rotations (n_gaussians,4), viewmats/projmats (n_views,16), campos (n_views,3), bg (3).
 * Outputs: out_color (n_views,3,H,W), radii (n_records) int32, invdepth (n_views,1,H,W) or NULL. */
UP3D_API int up3d_raster_forward(const up3d_raster_desc *d, const float *means3D, const float *shs,
                        const float *colors_precomp, const float *opacities, const float *scales,
                        const float *rotations, const float *viewmats, const float *projmats, const float *campos,
                        const float *bg, float *out_color, int32_t *radii, float *invdepth, void *state,
                        void *scratch, up3d_stream_t stream);

/* Backward.  dL_dcolor (n_views,3,H,W).  Gradient outputs are overwritten (not accumulated):
 * dL_dmeans3D (n_gaussians,3), dL_dshs (n_gaussians,M,3) or NULL, dL_dcolors (n_gaussians,3) or NULL,
 * dL_dopacities (n_gaussians), dL_dscales (n_gaussians,3), dL_drotations (n_gaussians,4);
 * dL_dmeans2D (n_records,3) or NULL -- the per-render screen-space gradient upstream returns
 * for `means2D` (x*0.5W, y*0.5H, 0).  Sums over the views of a set are done in a fixed order
 * (deterministic given the blend partials). */
UP3D_API int up3d_raster_backward(const up3d_raster_desc *d, const float *means3D, const float *shs,
                         const float *colors_precomp, const float *opacities, const float *scales,
                         const float *rotations, const float *viewmats, const float *projmats, const float *campos,
                         const float *bg, const float *dL_dcolor, const void *state, void *scratch,
                         float *dL_dmeans3D, float *dL_dmeans2D, float *dL_dshs, float *dL_dcolors,
                         float *dL_dopacities, float *dL_dscales, float *dL_drotations, up3d_stream_t stream);

/* Debug accessors for the parity tests (materialise what the reference's binning buffer holds).
 * All outputs optional (NULL to skip):
 *   n_visible (n_views) int32            -- depth-sorted, un-culled Gaussians per view
 *   sorted_ids (n_records) int32         -- per view: local Gaussian ids in (depth, id) order
 *   depths (n_records) f32, xy (n_records,2), conic_opacity (n_records,4), rgb (n_records,3),
 *   rects (n_records,4) int32            -- per record, UNSORTED (indexed by local id), 0 if culled
 *   final_T (n_views,H,W) f32, n_contrib (n_views,H,W) int32 */
UP3D_API int up3d_raster_debug_state(const up3d_raster_desc *d, const void *state, int32_t *n_visible, int32_t *sorted_ids,
                            float *depths, float *xy, float *conic_opacity, float *rgb, int32_t *rects,
                            float *final_T, int32_t *n_contrib, up3d_stream_t stream);

/* Coarse-bin decision of the forward (views of large sets with small footprints are blended from per-bin candidate
 * lists instead of the whole view): bin_mode (n_views) int32 1 = binned, bin_total (n_views) int32 = sum over the
 * view's visible records of the 64x64-pixel bins they touch (0 when the configuration never bins). */
UP3D_API int up3d_raster_debug_bins(const up3d_raster_desc *d, const void *state, int32_t *bin_mode, int32_t *bin_total,
                                    up3d_stream_t stream);

/* Per-tile lists exactly as the reference's global (tile|depth) sort would produce them.
 * tile_counts (n_views, tiles) int32 is always written; when tile_lists != NULL the ids of tile t of
 * view v are written at tile_lists[view_list_start_h-style offsets]: the caller passes
 * tile_offsets (n_views*tiles + 1) int32 = exclusive scan of a previous tile_counts query. */
UP3D_API int up3d_raster_debug_tile_lists(const up3d_raster_desc *d, const void *state, int32_t *tile_counts,
                                 const int32_t *tile_offsets, int32_t *tile_lists, up3d_stream_t stream);

/* Per-kernel device timing for the benchmark's roofline leg: when enabled, forward/backward record CUDA events
 * on the launching stream around each kernel (do not enable while capturing a CUDA graph).  read() blocks on the
 * last event and returns milliseconds for [project, depth_sort, blend_forward, grad-clear, blend_backward,
 * geometry_backward] of the most recent forward / backward on this host thread (-1 where not recorded). */
/* (diagnostic, process-wide: while enabled only ONE stream may issue raster calls; untimed callers on other streams
 *  stay correct but are not measured.  All other entry points are thread-safe for distinct streams.) */
UP3D_API int up3d_raster_timing_enable(int enable);
UP3D_API int up3d_raster_timing_read(float *ms6);

/* ------------------------------------------------------------------------------------------
 * Fused loss (train_network.py:260-302 + utils/loss_utils.py:23-45): focal-L2 between
 * rendered (V,3,H,W) and gt (V,3,H,W); writes the scalar mean loss (loss_out[0]) and
 * dL/drendered (same shape as rendered, may be NULL) in one pass. bg (3) is the background colour the
 * reference compares gt against (isclose, atol 1e-6 + rtol 1e-5*|bg|).
 * loss_out must be an 8-byte-aligned buffer of 4 floats: [0] = result, [2..3] = fp64 accumulator scratch.
 * ---------------------------------------------------------------------------------------- */
UP3D_API int up3d_focal_l2_loss(int64_t n_images, int H, int W, const float *rendered, const float *gt, const float *bg,
                       float non_bg_rate, float bg_rate, float *loss_out, float *dL_drendered,
                       up3d_stream_t stream);
/* Same, with gt read in place: gt_is_u8 != 0 -> gt holds the 8-bit images as decoded from the dataset (value/255 is
 * applied per read, the division the reference's loader does on the host); image n = (object n / views_per_object,
 * view n % views_per_object) starts at gt + object*gt_object_stride + view*3*H*W elements -- the trainer's
 * gt_images[:, input_images:] slice (train_network.py:427,444) without a copy. */
UP3D_API int up3d_focal_l2_loss_strided(int64_t n_images, int H, int W, const float *rendered, const void *gt, int gt_is_u8,
                                        int64_t views_per_object, int64_t gt_object_stride, const float *bg,
                                        float non_bg_rate, float bg_rate, float *loss_out, float *dL_drendered,
                                        up3d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Transformer-encoder glue (openpoints/models/backbone/transformer.py:89-121 Block, 123-207
 * TransformerEncoder, 10-33 Mlp): everything between two GEMMs of a Block in ONE pass.
 * Rows are tokens: T = B*L rows of width C (C = 128*{1,2,3,4,6,8}); row r belongs to sample r / L.
 * act_bf16 selects the dtype of the GEMM-side activations (0: float, 1: __nv_bfloat16); the
 * residual stream, LayerNorm statistics and every parameter gradient are fp32.
 * ---------------------------------------------------------------------------------------- */

/* xs = x + scale[r/L]*delta + pos  (delta, scale, pos optional; scale NULL == 1: DropPath's per-sample
 * keep/(1-p) factor, transformer.py:119-120; pos: `block(x + pos)`, transformer.py:188);
 * y = LayerNorm(xs; gamma, beta, eps) (nn.LayerNorm semantics, biased variance); mean/rstd (T) saved for
 * the backward.  xs_out may be NULL; y may be NULL (then only the residual sum is written). */
UP3D_API int up3d_ln_fwd(int act_bf16, int T, int L, int C, const float *x, const void *delta, const float *scale,
                         const float *pos, const float *gamma, const float *beta, float eps, float *xs_out, void *y,
                         float *mean, float *rstd, up3d_stream_t stream);

/* dx = g_res + dLayerNorm(dy)  (g_res optional: the gradient arriving on the residual path);
 * dgamma/dbeta (C) += column sums (atomic accumulate: zero them once per step);
 * dpos (T,C) += dx when not NULL; dscaled (T,C, activation dtype) = scale[r/L]*dx when not NULL --
 * the dY of the Linear that produced the residual branch in front of this LayerNorm -- and
 * dbias (C) += column sums of dscaled (that Linear's bias gradient). */
UP3D_API int up3d_ln_bwd(int act_bf16, int T, int L, int C, const void *dy, const float *xs, const float *mean,
                         const float *rstd, const float *gamma, const float *g_res, const float *scale, float *dx,
                         float *dpos, void *dscaled, float *dgamma, float *dbeta, float *dbias, up3d_stream_t stream);

/* y = GELU(x), erf form (nn.GELU default, transformer.py:17), n elements (multiple of 4). */
UP3D_API int up3d_gelu_fwd(int act_bf16, int64_t n, const void *x, void *y, up3d_stream_t stream);
/* dx = dy * GELU'(pre) over (T,C); dbias (C, may be NULL) += column sums of dx (fc1's bias gradient). */
UP3D_API int up3d_gelu_bwd(int act_bf16, int T, int C, const void *dy, const void *pre, void *dx, float *dbias,
                           up3d_stream_t stream);
/* out (T,C, activation dtype) = scale[r/L] * g (fp32); dbias (C, may be NULL) += column sums of out. */
UP3D_API int up3d_scale_cast_colsum(int act_bf16, int T, int L, int C, const float *g, const float *scale, void *out,
                                    float *dbias, up3d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Multi-head self-attention of the transformer Blocks (openpoints/models/backbone/transformer.py:36-77:
 * softmax(q k^T * scale) v, no mask, attn_drop = 0), bf16 operands / fp32 accumulation, head_dim D = 64,
 * L <= up3d_attn_max_len() tokens per sample.  qkv (B*L, 3*H*D) is the fused qkv GEMM output (q | k | v, heads
 * contiguous inside each third, exactly the reshape(B,N,3,H,D) of transformer.py:58-62); o (B*L, H*D);
 * lse (B,H,L) fp32 = log-sum-exp of the scaled scores (saved for the backward).
 * Backward: dqkv (B*L, 3*H*D) from qkv, o, lse and do (B*L, H*D) in one launch (every output row owned by one CTA:
 * deterministic, no atomics).
 * ---------------------------------------------------------------------------------------- */
UP3D_API int up3d_attn_max_len(void);
UP3D_API int up3d_attn_fwd(int B, int L, int H, int D, float scale, const void *qkv, void *o, float *lse,
                           up3d_stream_t stream);
UP3D_API int up3d_attn_bwd(int B, int L, int H, int D, float scale, const void *qkv, const void *o, const float *lse,
                           const void *dout, void *dqkv, up3d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer step of train_network.py:333-352: `_check_and_clip_gradients` (368-390: skip the step when any
 * gradient is NaN/Inf, else clip_grad_norm_(max_norm=1.0)) + torch.optim.AdamW(eps=1e-15) (156-159) in two
 * launches over a multi-tensor work list: sum of squares -> (coef, found_inf) -> one read-modify-write pass
 * over param / exp_avg / exp_avg_sq (+ optional bf16 shadow of the new parameter values).
 *   work list: chunk c covers tensor chunk_tensor[c], elements [chunk_start[c], +up3d_adamw_chunk_elems());
 *   pointer tables (device arrays of device pointers, n_tensors each); bf16_shadows[i] may be NULL, and the
 *   table itself may be NULL; group[i] indexes lrs (device floats, so a captured CUDA graph sees LR changes).
 *   state: 8-byte aligned device buffer of 32 bytes, zero-initialised once:
 *          fp64 sum-of-squares accumulator | float step | float total_norm | float clip_coef | float found_inf |
 *          uint32 ticket | pad.
 * A non-finite total norm leaves parameters, moments and the step counter untouched.
 * grad_scale (> 0) multiplies every gradient on the fly (1/world when the gradient buffers hold the SUM over the
 * data-parallel ranks; 1 otherwise): the norm, the clipping and the update all see grad_scale * g.
 * up3d_adamw_step = up3d_grad_sumsq (1 launch) + up3d_adamw_apply (1 launch; its last CTA advances the step
 * counter and clears the accumulator), exported separately so each pass can be timed on its own.
 * ---------------------------------------------------------------------------------------- */
UP3D_API int up3d_adamw_chunk_elems(void);
UP3D_API int up3d_grad_sumsq(int n_tensors, int n_chunks, const int32_t *chunk_tensor, const int32_t *chunk_start,
                             const int64_t *numel, const float *const *grads, float grad_scale, void *state,
                             up3d_stream_t stream);
UP3D_API int up3d_adamw_apply(int n_tensors, int n_chunks, const int32_t *chunk_tensor, const int32_t *chunk_start,
                              const int64_t *numel, float *const *params, const float *const *grads, float *const *exp_avg,
                              float *const *exp_avg_sq, void *const *bf16_shadows, const int32_t *group, const float *lrs,
                              float beta1, float beta2, float eps, float weight_decay, float max_norm, float grad_scale,
                              void *state, up3d_stream_t stream);
UP3D_API int up3d_adamw_step(int n_tensors, int n_chunks, const int32_t *chunk_tensor, const int32_t *chunk_start,
                             const int64_t *numel, float *const *params, const float *const *grads, float *const *exp_avg,
                             float *const *exp_avg_sq, void *const *bf16_shadows, const int32_t *group, const float *lrs,
                             float beta1, float beta2, float eps, float weight_decay, float max_norm, float grad_scale,
                             void *state, up3d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Mini-PointNet of the tokenizer (openpoints/models/backbone/transformer.py:210-243 `Encoder`): the memory-bound
 * passes around its four dense GEMMs.  Rows are (group, k): r = g*K + k, g < Gt = B*G, R = Gt*K.
 * act_bf16 selects the dtype of the (R,C) / (Gt,C) activations (0 float, 1 __nv_bfloat16); statistics, partial
 * sums and parameter gradients are fp32.  BatchNorm statistics are written as per-CTA partial sums and merged in
 * fp64 by up3d_bn_reduce_finalize / up3d_bn_reduce_sums (deterministic, no atomics).
 * stats (4,C) = batch mean, rstd, a = gamma*rstd, d = beta - mean*a.
 * ---------------------------------------------------------------------------------------- */

/* first layer Conv1d(3,128) (transformer.py:215, 230): nb is the tokenizer's (B,3,G,K) neighbourhood tensor
 * (GK = G*K), W1 (128,3), b1 (128).  stats: one partial per tile of up3d_pn_stats_tile_rows() rows,
 * partials (ceil(R/tile), 3, 128) = [shift, sum (z-shift), sum (z-shift)^2] of z = W1 x + b1 (shift = z of the
 * tile's first row: no cancellation when |mean| >> std). */
UP3D_API int up3d_pn_stats_tile_rows(void);
UP3D_API int up3d_pn_conv1_stats(int R, int GK, const float *nb, const float *W1, const float *b1, float *partials,
                                 up3d_stream_t stream);
/* y1 (R,128) = ReLU(BatchNorm(W1 x + b1)) in one pass from the 3-channel input (transformer.py:215-217). */
UP3D_API int up3d_pn_conv1_bn_relu(int act_bf16, int R, int GK, const float *nb, const float *W1, const float *b1,
                                   const float *stats, void *y1, up3d_stream_t stream);
/* backward of the same three ops from dy1 (R,128).  pass 0: partials (n_partials,2,128) of sum dy, sum dy*xhat (dy masked
 * by the ReLU); pass 1 (sums (2,128) = reduced pass-0 partials): partials (n_partials,2,256) with row 0 =
 * [gW1[:,0] | gW1[:,1]], row 1 = [gW1[:,2] | gb1] (reduce with up3d_bn_reduce_sums; gW1/gb1 arguments are unused),
 * dz = a (dy - mean(dy) - xhat mean(dy xhat)), means = sums / count (count = the number of
 * rows the batch statistics were taken over: R, or R * world under SyncBatchNorm). */
UP3D_API int up3d_pn_conv1_bwd(int act_bf16, int pass, int R, int GK, const float *nb, const float *W1, const float *b1,
                               const float *stats, const void *dy1, const float *sums, double count, float *partials,
                               int n_partials, float *gW1, float *gb1, up3d_stream_t stream);
/* partial sums (n_partials, n_rows, C) -> sums (n_rows, C), n_rows = 1 or 2, fp64 accumulation. */
UP3D_API int up3d_bn_reduce_sums(int n_partials, int n_rows, int C, const float *partials, float *sums, up3d_stream_t stream);
/* forward statistics: partials (n_partials,3,C) = [shift, sum (z-shift), sum (z-shift)^2], partial p covering
 * min(rows_per_partial, total_rows - p*rows_per_partial) rows, merged pairwise in fp64 into the batch mean / M2 ->
 * triple_out (3,C) = [mean, 0, M2] (same format: a second-level merge across ranks for SyncBatchNorm; may be NULL)
 * and/or stats (4,C) plus nn.BatchNorm1d's running-statistics update (momentum, unbiased variance; running_* and
 * num_batches_tracked may be NULL). */
UP3D_API int up3d_bn_reduce_finalize(int n_partials, int C, const float *partials, int rows_per_partial, int64_t total_rows,
                                     const float *gamma, const float *beta, float eps, float momentum, float *running_mean,
                                     float *running_var, int64_t *num_batches_tracked, float *triple_out, float *stats,
                                     up3d_stream_t stream);
/* The group-tile passes below stream whole group tiles through a shared-memory ring (cp.async.bulk + mbarrier) when the
 * tile layout allows it; up3d_set_group_tile_staging(0) selects the direct-load variant instead (same arithmetic, bit-identical
 * results; UP3D_GT_STAGED=0 in the environment does the same).  Returns the previous setting.  No reference counterpart. */
UP3D_API int up3d_set_group_tile_staging(int on);

/* BatchNorm over z = zl + gpart[g] + bias (zl (R,C): the local half of Conv1d(512,512) on [global || local]
 * (transformer.py:236-238); gpart (Gt,C) fp32: the global half, one row per group; both optional):
 *   stats      : partials (ceil(Gt/gpc), 3, C) = [shift, sum (z-shift), sum (z-shift)^2], gpc*K rows per partial
 *   apply_relu : y (R,C) = ReLU(a z + d)
 *   bwd_reduce : partials of sum dy, sum dy*xhat (dy masked by the ReLU)
 *   bwd_apply  : dz (R,C) = a (dy - mean(dy) - xhat mean(dy xhat)) and dgroup (Gt,C) = sum_k dz (the gradient of gpart) */
UP3D_API int up3d_gbn_stats(int act_bf16, int Gt, int K, int C, int gpc, const void *zl, const float *gpart, const float *bias,
                            float *partials, up3d_stream_t stream);
UP3D_API int up3d_gbn_apply_relu(int act_bf16, int Gt, int K, int C, int gpc, const void *zl, const float *gpart,
                                 const float *bias, const float *stats, void *y, up3d_stream_t stream);
UP3D_API int up3d_gbn_bwd_reduce(int act_bf16, int Gt, int K, int C, int gpc, const void *dy, const void *zl, const float *gpart,
                                 const float *bias, const float *stats, float *partials, up3d_stream_t stream);
UP3D_API int up3d_gbn_bwd_apply(int act_bf16, int Gt, int K, int C, int gpc, const void *dy, const void *zl, const float *gpart,
                                const float *bias, const float *stats, const float *sums, double count, void *dz,
                                void *dgroup, up3d_stream_t stream);
/* per-group max-pool over the K rows (torch.max(dim), transformer.py:235,242): out (Gt,C), arg (Gt,C) int32 = first
 * arg-max row; its backward scatter dx (R,C) = dpooled at the arg-max row, 0 elsewhere; and
 * dx = dlocal + scatter(dpooled) with colsum_partials (ceil(Gt/gpc), C; may be NULL) = per-CTA column sums of dx. */
UP3D_API int up3d_group_max(int act_bf16, int Gt, int K, int C, const void *x, void *out, int32_t *arg, up3d_stream_t stream);
UP3D_API int up3d_group_max_scatter(int act_bf16, int Gt, int K, int C, const void *dpooled, const int32_t *arg, void *dx,
                                    up3d_stream_t stream);
UP3D_API int up3d_group_combine(int act_bf16, int Gt, int K, int C, int gpc, const void *dlocal, const void *dpooled,
                                const int32_t *arg, void *dx, float *colsum_partials, up3d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Splat head: GaussianSplatPredictor._process_network_output (model/gaussian_predictor.py:279-328; activations
 * 249-254) + the SH concatenation of render_predicted (gaussian_renderer/__init__.py:66-69) in one launch each way.
 * raw (B,P,11+3M) = the `final` MLP's output rows [xyz 3 | opacity 1 | scaling 3 | rotation 4 | sh (M,3)],
 * center (B,P,3).  Outputs: xyz (B,P,3) = tanh(raw)*offset_scale + center, opacity (B,P) = sigmoid,
 * scaling (B,P,3) = exp(clamp(raw,-1,20)) (isotropic: channel 0 broadcast), rotation (B,P,4) = raw / max(||raw||_P, 1e-6)
 * -- normalised over the P POINTS per component, the reference's F.normalize(dim=-1) on a (B,4,P) tensor --,
 * shs (B,P,M,3), rot_norm (B,4) (saved for the backward).  Backward: d_raw (B,P,11+3M) from the five output
 * gradients (any may be NULL = zero).
 * ---------------------------------------------------------------------------------------- */
UP3D_API int up3d_splat_head_fwd(int B, int P, int M, const float *raw, const float *center, float offset_scale,
                                 int isotropic, float *xyz, float *opacity, float *scaling, float *rotation, float *shs,
                                 float *rot_norm, up3d_stream_t stream);
UP3D_API int up3d_splat_head_bwd(int B, int P, int M, const float *raw, const float *rot_norm, float offset_scale,
                                 int isotropic, const float *d_xyz, const float *d_opacity, const float *d_scaling,
                                 const float *d_rotation, const float *d_shs, float *d_raw, up3d_stream_t stream);

/* FeatureFusion's geometry (fusion/feat_fusion.py:23-56, 88-131) for the analytic stem field, no gradient:
 * centres (B,N,3) -> camera space with w2c (B,16 row-major), pixel = round(cam.xy * f / cam.z + c) (half to even),
 * in-image test, nearest-depth test per pixel cell (iy*H + ix, the reference's hash) -> keep (B,N) uint8,
 * pix (B,N,2) int32, and xhat (B,N,C) = GroupNorm-normalised stem field at image[b, :, ix, iy] (the reference's own
 * index order), statistics from up3d_stem_group_stats' sums (B,G,2).  image (B,3,H,W), proj (C,3), shift (C). */
UP3D_API int up3d_fusion_project(int B, int N, int H, int W, int C, int G, float fx, float fy, float cx, float cy, float eps,
                                 const float *center, const float *w2c, const float *image, const float *proj,
                                 const float *shift, const double *sums, unsigned char *keep, int32_t *pix, float *xhat,
                                 up3d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Point-cloud serialization keys (first stage of PTv3: pointcept/models/utils/structure.py:47-107 `Point.serialization`,
 * serialization/default.py:8-25 `encode`, serialization/z_order.py:41-96).  grid_coord (n,3) int32 voxel coordinates,
 * batch (n) int64 or NULL -> code (n) int64: bit i of x, y, z at bits 3i+2, 3i+1, 3i (swap_xy != 0: the "z-trans" order,
 * x and y exchanged), batch index OR-ed in above bit 3*depth; depth in [1,16] as the reference asserts.
 * ---------------------------------------------------------------------------------------- */
UP3D_API int up3d_zorder_keys(int64_t n, int depth, int swap_xy, const int32_t *grid_coord, const int64_t *batch, int64_t *code,
                              up3d_stream_t stream);
/* same interface for the "hilbert" / "hilbert-trans" orders (serialization/hilbert.py:96-191: Skilling's transpose,
 * bit interleave with dimension 0 most significant, Gray decode). */
UP3D_API int up3d_hilbert_keys(int64_t n, int depth, int swap_xy, const int32_t *grid_coord, const int64_t *batch, int64_t *code,
                               up3d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Frozen image stem (stand-in for model/image_predictor.py:56-81, whose SD-VAE weights are not shipped):
 * the 128-channel field f[n,c,y,x] = sin(proj[c,:] . image[n,:,y,x] + shift[c]) feeds
 * image_conv = GroupNorm(G, C) + Conv1x1 (model/gaussian_predictor.py:61-66,139).  Writes the GroupNorm
 * statistics sums (n_images, G, 2) fp64 = [sum f, sum f^2] per (image, group) straight from the 3-channel
 * image (n,3,H,W); the dense field is never materialised.  proj (C,3), shift (C); C <= 256, C % G == 0.
 * ---------------------------------------------------------------------------------------- */
UP3D_API int up3d_stem_group_stats(int n_images, int H, int W, int C, int G, const float *image, const float *proj,
                                   const float *shift, double *sums, up3d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Dense Linear on the 5th-generation tensor cores (tcgen05.mma, TMEM accumulator, TMA tensor loads).
 * Replaces the nn.Linear / Conv1d(k=1) GEMMs of the point backbone
 * (/root/reference/openpoints/models/backbone/transformer.py:22-33 Mlp, :52-77 Attention qkv/proj,
 * :214-243 mini-PointNet), forward and the dX half of the backward.
 *   out (T,N) = A (T,K) bf16 row-major  x  op(B)  (+ bias (N) bf16)  -> epilogue
 *   b_major 0: B is (N,K) row-major (the weight as nn.Linear stores it: y = x W^T)
 *   b_major 1: B is (K,N) row-major (the same weight read for dx = dy W)
 *   b_major | 2: B is an activation written by a kernel that may still be in flight (by default B is treated as a
 *              weight and prefetched before the programmatic dependency on the previous kernel resolves)
 *   epilogue 0: out = acc (+ bias)
 *            1: aux_out (T,N) bf16 = acc + bias ; out = GELU(aux_out)          (fc1 + nn.GELU)
 *            2: out = acc * GELU'(aux_in (T,N) bf16)                          (dX of fc2 + GELU backward)
 *   out is bf16 (out_f32 = 0) or fp32 (out_f32 = 1); tile_n = 0 lets the library choose the CTA tile width
 *   (32/64/96/128; 64/128 for b_major 1) and the split-K factor; otherwise tile_n = width | (split_k << 16), split_k in
 *   {0 = auto, 1, 2, 4} (2 and 4: the K range is shared by a thread-block cluster; plain epilogue only).  K % 8 == 0, N % 32 == 0, all pointers 16-byte aligned.
 * ---------------------------------------------------------------------------------------- */
/* Programmatic dependent launch of the library's chained kernels (default on; UP3D_PDL=0 in the environment = off). */
UP3D_API int up3d_set_pdl(int enabled);
UP3D_API int up3d_tc_linear(int T, int N, int K, const void *A, const void *B, int b_major, const void *bias, int epilogue,
                            const void *aux_in, void *aux_out, void *out, int out_f32, int tile_n, up3d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Sparse 3-D convolution for the scene-level backbones: the arithmetic the reference takes from the un-vendored spconv
 * package (spconv.SubMConv3d / SparseConv3d / SparseInverseConv3d in
 * /root/reference/pointcept/models/sparse_unet/spconv_unet_v1m1_base.py:57-83,153-160,209-216,247-253 and
 * point_transformer_v3m1_base.py:281-287).  Output-stationary implicit GEMM over a rulebook
 * nbr (kernel_volume, n_out) int32: nbr[o][i] = input row read by kernel offset o for output row i, -1 = none.
 *   up3d_sparse_subm_rulebook: keys_sorted (n) int64 = batch<<48 | c0<<32 | c1<<16 | c2 ascending, coords (n,4) int32
 *     (batch, c0, c1, c2) in the same order -> nbr (k^3, n); offset o = (a*k + b)*k + c reads coord + (a,b,c) - k/2.
 *   up3d_sparse_conv: out (n_out, c_out) fp32 = sum_o in[nbr[o][i]] (fp32, rounded to bf16) x weight[o] (c_in, c_out) bf16,
 *     fp32 accumulation on the tensor cores; deterministic.  The input gradient is the same call with the transposed
 *     rulebook and per-offset transposed weights.  Channels are multiples of 16 (callers zero-pad).
 *   up3d_sparse_conv_wgrad: dweight (kernel_volume, c_in, c_out) fp32 += sum_i in[nbr[o][i]]^T dout[i] (fp32 atomics).
 * ---------------------------------------------------------------------------------------- */
UP3D_API int up3d_sparse_subm_rulebook(int n, int kernel_size, const int64_t *keys_sorted, const int32_t *coords,
                                       int32_t *nbr, up3d_stream_t stream);
UP3D_API int up3d_sparse_conv(int n_out, int c_in, int c_out, int kernel_volume, const int32_t *nbr, const float *in,
                              const void *weight_bf16, float *out, up3d_stream_t stream);
UP3D_API int up3d_sparse_conv_wgrad(int n_out, int c_in, int c_out, int kernel_volume, const int32_t *nbr, const float *in,
                                    const float *dout, float *dweight, up3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UP3D_H */
