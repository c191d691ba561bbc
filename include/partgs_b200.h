/*
 * partgs_b200 — C ABI of the B200-native (sm_100a) PartGS rendering hot path.
 *
 * Plain C: raw device pointers, sizes, a cudaStream_t passed as void*.  No torch
 * types.  Every entry point returns >= 0 on success and a negative pgs_status on
 * failure; pgs_last_error() returns a human-readable message for the calling thread.
 * All work is enqueued on `stream`; the only host synchronisation is the read-back
 * of the instance count inside pgs_dsr_forward / pgs_dsrp_forward (the reference
 * has the same one: rasterizer_impl.cu:282).
 *
 * Each declaration names the reference interface it replaces
 * (paths relative to zhirui-gao/PartGS; DSR = submodules/diff-surfel-rasterization,
 * DSRP = submodules/diff-surfel-rasterization_part, KNN = submodules/simple-knn).
 */
#ifndef PARTGS_B200_H_
#define PARTGS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  PGS_OK = 0,
  PGS_ERR_INVALID_ARG = -1,
  PGS_ERR_CUDA = -2,
  PGS_ERR_ALLOC = -3,
  PGS_ERR_UNSUPPORTED = -4
} pgs_status;

/* Resizable-buffer callback: must return a device pointer to at least `bytes`
 * bytes that stays valid until the matching backward call.  C equivalent of the
 * reference's std::function<char*(size_t)> (DSR/cuda_rasterizer/rasterizer.h:36-38,
 * produced by resizeFunctional in DSR/rasterize_points.cu:31-37). */
typedef char* (*pgs_alloc_fn)(size_t bytes, void* user);

const char* pgs_last_error(void);
int pgs_version(void);
/* Number of kernels launched by this library in the calling process so far. */
unsigned long long pgs_launch_count(void);

/* ---- per-stage device timing (CUDA events recorded on the caller's stream around each
 * kernel group); off by default.  pgs_timing_read synchronises the recorded events and
 * returns accumulated milliseconds / launch counts per stage. */
enum {
  PGS_STAGE_PREPROCESS_FWD = 0, PGS_STAGE_SCAN, PGS_STAGE_DUP_KEYS, PGS_STAGE_SORT, PGS_STAGE_TILE_RANGES,
  PGS_STAGE_RENDER_FWD, PGS_STAGE_RENDER_BWD, PGS_STAGE_PREPROCESS_BWD, PGS_STAGE_KNN, PGS_STAGE_SQ_FWD,
  PGS_STAGE_SQ_BWD, PGS_STAGE_SURFACE_FWD, PGS_STAGE_SURFACE_BWD, PGS_STAGE_PHOTO_FWD,
  PGS_STAGE_PHOTO_BWD, PGS_NUM_STAGES
};
void pgs_timing_enable(int on);
int pgs_timing_read(double* ms, unsigned long long* counts, int reset);

/* ---- base rasteriser ----------------------------------------------------------
 * Replaces CudaRasterizer::Rasterizer::forward (DSR/cuda_rasterizer/rasterizer.h:31-61,
 * rasterizer_impl.cu:198-342).  Same argument meaning; optional inputs are NULL
 * (shs / colors_precomp, scales+rotations / transMat_precomp).  Returns num_rendered.
 * out_color [3,H,W], out_others [7,H,W], radii [P] (may be NULL). */
int pgs_dsr_forward(pgs_alloc_fn geometry_buffer, void* geometry_user, pgs_alloc_fn binning_buffer,
                    void* binning_user, pgs_alloc_fn image_buffer, void* image_user, int P, int D, int M,
                    const float* background, int width, int height, const float* means3D, const float* shs,
                    const float* colors_precomp, const float* opacities, const float* scales, float scale_modifier,
                    const float* rotations, const float* transMat_precomp, const float* viewmatrix,
                    const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered,
                    float* out_color, float* out_others, int* radii, int debug, void* stream);

/* Lazy instance count.  `debug` of the forward entry points is a flag word: bit 0 = the reference's debug switch, bit 1 =
 * PGS_FWD_LAZY_COUNT: do not wait for the frame's instance count (the reference blocks on it, rasterizer_impl.cu:282;
 * by default this library waits AFTER queueing the frame).  The frame is queued for the instance capacity remembered
 * from earlier frames and the call returns PGS_COUNT_PENDING, which pgs_dsr_backward accepts as R; the first frames of a
 * scene (no capacity remembered yet) and debug calls are not lazy and return the count.  A forward call issued while its
 * stream is being captured into a CUDA graph is always lazy (a capture cannot contain the host wait).
 * pgs_dsr_resolve_count() waits for the counts of all lazy frames of this host thread on the current device (or, with
 * none pending, reports the slot a replayed graph refreshed — synchronise the stream first) and returns the latest;
 * *overflow = 1 if a frame needed more instances than it was queued for: its outputs are invalid (its kernels did
 * nothing) and it must be rendered again — the remembered capacity has been raised.  At most 32 count slots exist:
 * outstanding lazy frames plus captured graphs (a captured forward keeps its slot). */
#define PGS_FWD_LAZY_COUNT 2
#define PGS_COUNT_PENDING 0x7fffffff
int pgs_dsr_resolve_count(int* overflow);
/* The instance capacity speculative / lazy frames on the current device are queued for (x 1.25, rounded up to 512 Ki);
 * it grows with the frames seen and decays slowly.  0 forgets it (the next frame counts first, like the reference) —
 * e.g. when switching to a much smaller scene. */
void pgs_dsr_set_capacity_hint(size_t instances);

/* Bytes of scratch pgs_dsr_backward needs (per-surfel gradient accumulators). */
size_t pgs_dsr_backward_scratch_bytes(int P);

/* Replaces CudaRasterizer::Rasterizer::backward (DSR/cuda_rasterizer/rasterizer.h:63-94,
 * rasterizer_impl.cu:346-448).  R = num_rendered returned by the forward call that
 * filled the three buffers; binning_bytes = the size the binning callback was asked for by
 * that call (the arena is laid out for a capacity >= R that backward recovers from it).  All nine gradient arrays are fully written (no
 * pre-zeroing required): dL_dmean2D [P,3], dL_dopacity [P], dL_dcolor [P,3],
 * dL_dmean3D [P,3], dL_dtransMat [P,9], dL_dsh [P,M,3], dL_dscale [P,2], dL_drot [P,4].
 * `scratch` replaces the reference's internal dL_dnormal [P,3] tensor.
 * `debug` is a flag word: bit 0 = the reference's debug switch (synchronise and check after every stage); bit 1 =
 * PGS_BWD_ACCUMULATE: dL_dmean3D, dL_dsh, dL_dopacity, dL_dscale and dL_drot are ADDED to (gradient accumulation over
 * the views of a data-parallel batch: one all-reduce per batch); the other arrays are overwritten. */
#define PGS_BWD_ACCUMULATE 2
int pgs_dsr_backward(int P, int D, int M, int R, const float* background, int width, int height,
                     const float* means3D, const float* shs, const float* colors_precomp, const float* scales,
                     float scale_modifier, const float* rotations, const float* transMat_precomp,
                     const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                     float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer, size_t binning_bytes,
                     char* image_buffer, const float* dL_dpix, const float* dL_dothers, float* dL_dmean2D, float* scratch,
                     float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dtransMat, float* dL_dsh,
                     float* dL_dscale, float* dL_drot, int debug, void* stream);

/* ---- `_part` rasteriser (part / semantic maps) ------------------------------------
 * Replaces CudaRasterizer::Rasterizer::forward / backward of the fork
 * (DSRP/cuda_rasterizer/rasterizer.h, rasterizer_impl.cu:198-350 / :352-460): same as the base
 * pair plus `semantics` [P,S] in, `out_semantic` [S,H,W] out, an 8-channel aux map
 * (out_others [8,H,W]) and `dL_dsemantics` [P,S] (fully written).  S = semantic_types must be
 * <= 16 (the reference overflows a fixed 16-entry array silently; here it is an error).
 * transMat_precomp must be NULL (that path is unusable in the reference fork). */
int pgs_dsrp_forward(pgs_alloc_fn geometry_buffer, void* geometry_user, pgs_alloc_fn binning_buffer,
                     void* binning_user, pgs_alloc_fn image_buffer, void* image_user, int P, int D, int M,
                     const float* background, int width, int height, int semantic_types, const float* means3D,
                     const float* shs, const float* colors_precomp, const float* semantics, const float* opacities,
                     const float* scales, float scale_modifier, const float* rotations, const float* transMat_precomp,
                     const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                     float tan_fovy, int prefiltered, float* out_color, float* out_semantic, float* out_others,
                     int* radii, int debug, void* stream);
int pgs_dsrp_backward(int P, int D, int M, int R, const float* background, int width, int height, int semantic_types,
                      const float* means3D, const float* shs, const float* colors_precomp, const float* semantics,
                      const float* scales, float scale_modifier, const float* rotations,
                      const float* transMat_precomp, const float* viewmatrix, const float* projmatrix,
                      const float* campos, float tan_fovx, float tan_fovy, const int* radii, char* geom_buffer,
                      char* binning_buffer, size_t binning_bytes, char* image_buffer, const float* dL_dpix,
                      const float* dL_dsemantic_pix,
                      const float* dL_dothers, float* dL_dmean2D, float* scratch, float* dL_dopacity, float* dL_dcolor,
                      float* dL_dsemantics, float* dL_dmean3D, float* dL_dtransMat, float* dL_dsh, float* dL_dscale,
                      float* dL_drot, int debug, void* stream);

/* ---- block-level (superquadric) rasteriser --------------------------------------------
 * north_star (1): preprocess fused with the superquadric -> surfel placement.  Same pipeline as
 * pgs_dsr_forward / pgs_dsr_backward, but surfel i = (block, face, sample) is GENERATED inside the
 * preprocess kernels from the 13 parameters per superquadric (BlockGaussianModel.prepare_scaling_rot /
 * get_xyz / get_scaling / get_rotation / get_opacity, games/block_mesh_splatting/scene/
 * block_gaussian_model.py:98-109,189-256) instead of being read from means3D / scales / rotations /
 * opacities; P = B*F*K.  `vertices` [B,Vt,3] is written by forward and read by backward.  out_xyz [P,3],
 * out_scaling [P,2] (log), out_rotation [P,4], out_opacity [P] may be NULL (materialise only if a caller
 * reads get_xyz etc.).  Backward returns the gradients of the block parameters (d_sq_* fully written;
 * d_alpha [B*F,K,3] and d_scale_raw [B,F*K] optional) plus dL_dsh / dL_dcolor / dL_dmean2D as usual. */
int pgs_dsr_forward_blocks(pgs_alloc_fn geometry_buffer, void* geometry_user, pgs_alloc_fn binning_buffer,
                           void* binning_user, pgs_alloc_fn image_buffer, void* image_user, int B, int Vt, int F, int K,
                           const float* sq_r, const float* sq_s, const float* sq_t, const float* sq_eps,
                           const float* sq_occ, const float* eta, const float* omega, const int* faces,
                           const float* alpha, const float* scale_raw, float ratio, float scale_min, int D, int M,
                           const float* background, int width, int height, const float* shs,
                           const float* colors_precomp, float scale_modifier, const float* viewmatrix,
                           const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
                           float* vertices, float* out_xyz, float* out_scaling, float* out_rotation,
                           float* out_opacity, float* out_color, float* out_others, int* radii, int debug,
                           void* stream);
size_t pgs_dsr_backward_blocks_scratch_bytes(int B, int Vt, int F, int K);
int pgs_dsr_backward_blocks(int B, int Vt, int F, int K, const float* sq_r, const float* sq_s, const float* sq_t,
                            const float* sq_eps, const float* sq_occ, const float* eta, const float* omega,
                            const int* faces, const float* alpha, const float* scale_raw, float ratio, float scale_min,
                            const float* vertices, int D, int M, int R, const float* background, int width, int height,
                            const float* shs, const float* colors_precomp, float scale_modifier,
                            const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                            float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer,
                            size_t binning_bytes, char* image_buffer, const float* dL_dpix, const float* dL_dothers,
                            const float* dL_dvertices, /* [B,Vt,3] or NULL: upstream gradient of the returned vertices */
                            float* dL_dmean2D, void* scratch, float* dL_dcolor, float* dL_dsh, float* d_sq_r,
                            float* d_sq_s, float* d_sq_t, float* d_sq_eps, float* d_sq_occ, float* d_alpha,
                            float* d_scale_raw, int debug, void* stream);

/* ---- renderer post-processing: surface maps -----------------------------------------------
 * SURVEY.md section 8(f) rank 1.  What render() does with the rasteriser's allmap right after the call
 * (renderer/gaussian_renderer/__init__.py:110-149 + utils/point_utils.py:4-33: normal view->world,
 * nan_to_num, expected depth = depth/alpha, surf_depth blend, depth_to_normal stencil, x alpha), ~30 ATen
 * kernels in the reference, as one kernel forward and two backward.
 *   allmap [7,H,W]; view3x3 = world_view_transform[:3,:3] (row-major, 9 floats, DEVICE pointer);
 *   rays_m1 = intrins.inverse().T, rays_m2 = c2w[:3,:3].T (9 floats each), rays_o = c2w[:3,3] (3 floats),
 *   all device pointers prepared by the caller exactly like point_utils.depths_to_points does;
 *   out: rend_normal [3,H,W], surf_depth [1,H,W], surf_normal [3,H,W].
 * Backward: any of the three incoming gradients may be NULL (= zero); g_allmap [7,H,W] is fully written
 * (channel 6 and the direct alpha view gradient are left to the caller's autograd, they are plain slices). */
int pgs_surface_maps_forward(int width, int height, const float* allmap, const float* view3x3, const float* rays_m1,
                             const float* rays_m2, const float* rays_o, float depth_ratio, float* rend_normal,
                             float* surf_depth, float* surf_normal, void* stream);
size_t pgs_surface_maps_backward_scratch_bytes(int width, int height);
int pgs_surface_maps_backward(int width, int height, const float* allmap, const float* view3x3, const float* rays_m1,
                              const float* rays_m2, const float* rays_o, float depth_ratio, const float* g_rend_normal,
                              const float* g_surf_depth, const float* g_surf_normal, void* scratch, float* g_allmap,
                              void* stream);

/* ---- photometric loss: L1 + SSIM ----------------------------------------------------------
 * SURVEY.md section 8(f) rank 2.  loss = (1 - lambda) * mean|image - gt| + lambda * (1 - ssim(image, gt))
 * (train.py:230-231; utils/loss_utils.py:6-7 l1_loss, :12-54 ssim with the 11x11 sigma-1.5 Gaussian window,
 * zero padding).  Forward leaves {sum of the SSIM map, sum |image - gt|} in `sums` (2 doubles, device) and three
 * derivative maps in `dmaps` ([3][C][H][W] floats) for backward; the caller forms the scalar from the sums
 * (N = C*H*W).  Backward reads the upstream gradient of the scalar from DEVICE memory (g_loss, 1 float) and writes
 * d loss / d image [C][H][W]. */
int pgs_photometric_forward(int channels, int height, int width, const float* image, const float* gt, double* sums,
                            float* dmaps, void* stream);
int pgs_photometric_backward(int channels, int height, int width, const float* image, const float* gt,
                             const float* dmaps, const float* g_loss, float lambda_dssim, float* g_image,
                             void* stream);

/* ---- per-pixel regularisers of the training step -------------------------------------------------
 * SURVEY.md section 8(f) rank 2 (second half).  train.py:234-251 on the maps render() returns:
 *   mask entropy   -(mask * log(a) + (1 - mask) * log(1 - a)).mean(),  a = rend_alpha.clamp(1e-6, 1 - 1e-6)
 *   normal error   (1 - (rend_normal * surf_normal).sum(0)).mean()
 *   distortion     rend_dist.mean()
 * Forward leaves the three SUMS in `sums` (3 doubles, device; the caller divides by width*height and applies the
 * lambdas); a NULL mask / NULL normals / NULL dist skips that term.  Backward reads the upstream gradient of the
 * weighted scalar from DEVICE memory (g_loss, 1 float) and writes d/d rend_alpha [H,W], d/d rend_dist [H,W],
 * d/d rend_normal [3,H,W], d/d surf_normal [3,H,W] (any output may be NULL). */
int pgs_regularizers_forward(int width, int height, const float* rend_alpha, const float* gt_mask,
                             const float* rend_dist, const float* rend_normal, const float* surf_normal, double* sums,
                             void* stream);
int pgs_regularizers_backward(int width, int height, const float* rend_alpha, const float* gt_mask,
                              const float* rend_normal, const float* surf_normal, const float* g_loss,
                              float lambda_mask_entropy, float lambda_normal, float lambda_dist, float* g_rend_alpha,
                              float* g_rend_dist, float* g_rend_normal, float* g_surf_normal, void* stream);

/* ---- optimiser step / densification statistics ---------------------------------------------
 * SURVEY.md section 8(f) rank 3 (partial).  pgs_adam_step: torch.optim.Adam's update (no amsgrad / weight decay;
 * scene/gaussian_model.py:266 uses eps = 1e-15) for up to 16 tensors in ONE launch.  The tables are HOST arrays of
 * device pointers; step_size[i] = lr_i / (1 - beta1^step) and bias_correction2_sqrt = sqrt(1 - beta2^step) are
 * computed by the caller in double precision exactly like torch/optim/adam.py; the betas and eps travel as doubles so
 * that (1 - beta) is formed in double too (1.f - 0.999f is off by 5e-5).
 * pgs_densify_stats: for surfels with radii > 0: max_radii2D = max(max_radii2D, radii) (optional, may be NULL),
 * grad_accum += |grad_means2D.xy|, denom += 1  (scene/gaussian_model.py:515-517, train.py:295-297). */
int pgs_adam_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const size_t* numel, const float* step_size, double beta1, double beta2,
                  double eps, double bias_correction2_sqrt, void* stream);
int pgs_densify_stats(int P, const int* radii, const float* grad_means2D, float* max_radii2D, float* grad_accum,
                      float* denom, void* stream);

/* ---- densification (clone / split / prune) as one planned compaction ---------------------------
 * SURVEY.md section 8(f) rank 3.  Replaces TwoGaussianModel.densify_and_prune
 * (games/block_mesh_splatting/scene/two_gaussian_model.py:341-423 prune_points / densification_postfix /
 * densify_and_split / densify_and_clone; scene/gaussian_model.py:384-436 _prune_optimizer / cat_tensors_to_optimizer,
 * :495-509 densify_and_prune).  Row order of the result = the reference's:
 *     [surviving originals | surviving clones | surviving split children, replica-major].
 * 1. pgs_densify_plan classifies the P surfels.  Thresholds are the reference's Python scalars: max_grad,
 *    dense_threshold = percent_dense * extent, min_opacity, world_size_threshold = 0.1 * extent (tested only when
 *    use_world_size_test != 0, i.e. when the reference's max_screen_size is truthy; its screen-size test itself can
 *    never fire because densification_postfix zeroes max_radii2D first), split_divisor = 0.8 * N.
 *    scaling [P,2] and opacity [P] are the stored (pre-activation) parameters.  Outputs (device): code [P] bytes,
 *    block_offsets [4 * pgs_densify_blocks(P)] u32, counts [8] u32 = {surviving originals, surviving clones, rows
 *    selected for splitting (Ns: the caller draws z [N*Ns,3] standard normals, the reference's
 *    torch.normal(0, stds)), surviving children per replica, rows selected for cloning, 0, 0, 0}.  The caller
 *    reads counts back to size the outputs: n_out = counts[0] + counts[1] + N * counts[3].
 * 2. pgs_densify_map writes src_row [n_out] (output row -> source row) and sample_row [N * counts[3]].
 * 3. pgs_densify_gather moves up to 24 row-major float tensors in one launch: dst[t][r, :] = src[t][src_row[r], :],
 *    or zeros for r >= n_keep when zero_new[t] != 0 (Adam moments of new surfels).  Tables are HOST arrays.
 * 4. pgs_densify_children overwrites xyz / scaling of the split children:
 *    xyz = R(rotation) @ (z * [exp(scaling), 0]) + xyz_parent, scaling = log(exp(scaling_parent) / split_divisor). */
int pgs_densify_blocks(int P);
int pgs_densify_plan(int P, const float* grad_accum, const float* denom, const float* scaling, const float* opacity,
                     double max_grad, double dense_threshold, double min_opacity, int use_world_size_test,
                     double world_size_threshold, double split_divisor, unsigned char* code,
                     unsigned int* block_offsets, unsigned int* counts, void* stream);
int pgs_densify_map(int P, const unsigned char* code, const unsigned int* block_offsets, const unsigned int* counts,
                    int n_split, int* src_row, int* sample_row, void* stream);
int pgs_densify_gather(int n_tensors, const float* const* src, float* const* dst, const int* widths,
                       const int* zero_new, int n_out, int n_keep, const int* src_row, void* stream);
int pgs_densify_children(int n_children, const unsigned int* counts, const int* src_row, const int* sample_row,
                         const float* z, const float* xyz_in, const float* scaling_in, const float* rotation_in,
                         double split_divisor, float* xyz_out, float* scaling_out, void* stream);

/* ---- per-view epilogue of the multi-view extraction loop ---------------------------------------
 * SURVEY.md section 8(f) rank 4.  GaussianExtractor.reconstruction (utils/mesh_utils.py:102-129) per rendered view:
 * part_rgb = partmap_to_rgbmap(render_semantic) (:77-89: clamp to [0,1], argmax over the S part channels -> palette
 * colour, pixels whose clamped channels sum to < 0.1 -> white) and normal_unit = F.normalize(rend_normal, dim=0)
 * (:113).  semantic [S,H,W]; palette: (S+1) rows of palette_stride >= 3 floats (get_fancy_color(S+1), DEVICE
 * memory); either output (and its input) may be NULL to skip that half. */
int pgs_extract_maps(int width, int height, int n_parts, const float* semantic, const float* palette,
                     int palette_stride, const float* rend_normal, float* part_rgb, float* normal_unit, void* stream);

/* Replaces CudaRasterizer::Rasterizer::markVisible (rasterizer.h:24-29, rasterizer_impl.cu:141-153).
 * present: one byte per point (bool). */
int pgs_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     unsigned char* present, void* stream);

/* ---- superquadric -> surfel parameterisation ------------------------------------------
 * Replaces BlockGaussianModel.get_verts / prepare_scaling_rot / get_opacity
 * (games/block_mesh_splatting/scene/block_gaussian_model.py:189-256, 106-109) and the autograd
 * graph PyTorch builds for them.  B superquadrics, Vt icosphere vertices, F faces, K surfels per
 * face; surfel index = (b*F + f)*K + k.  All arrays float32 device memory except faces (int32).
 *   in : sq_r[B,4] sq_s[B,3] sq_t[B,3] sq_eps[B,2] sq_occ[B] (raw, pre-activation),
 *        eta[B,Vt] omega[B,Vt], faces[B,F,3], alpha[B*F,K,3] (normalised barycentrics, BGM:178-186),
 *        scale_raw[B,F*K] (BGM `_scale`), ratio (ratio_block_scene), scale_min (scale_block_min)
 *   out: vertices[B,Vt,3], xyz[P,3], scaling[P,2] (log space, BGM `_scaling`), rotation[P,4]
 *        (w,x,y,z, real part >= 0, BGM `_rotation`), opacity[P] = sigmoid(sq_occ) per block. */
int pgs_sq2surfel_forward(int B, int Vt, int F, int K, const float* sq_r, const float* sq_s, const float* sq_t,
                          const float* sq_eps, const float* sq_occ, const float* eta, const float* omega,
                          const int* faces, const float* alpha, const float* scale_raw, float ratio, float scale_min,
                          float* vertices, float* xyz, float* scaling, float* rotation, float* opacity, void* stream);
size_t pgs_sq2surfel_backward_scratch_bytes(int B, int Vt);
/* Gradients of the five block parameters (always) and of alpha / scale_raw (when non-NULL) given
 * gradients of the forward outputs; d_opacity and d_vertices_in may be NULL. */
int pgs_sq2surfel_backward(int B, int Vt, int F, int K, const float* sq_r, const float* sq_s, const float* sq_t,
                           const float* sq_eps, const float* sq_occ, const float* eta, const float* omega,
                           const int* faces, const float* alpha, const float* scale_raw, float ratio, float scale_min,
                           const float* vertices, const float* d_xyz, const float* d_scaling, const float* d_rotation,
                           const float* d_opacity, const float* d_vertices_in, float* d_sq_r, float* d_sq_s,
                           float* d_sq_t, float* d_sq_eps, float* d_sq_occ, float* d_alpha, float* d_scale_raw,
                           void* scratch, void* stream);

/* ---- gradient all-reduce over NVLink peer memory (no counterpart in the reference, which is
 * single-process) -------------------------------------------------------------------------------------------
 * `buckets[q]` = address of rank q's gradient bucket as mapped into THIS process (symmetric / peer memory), q <
 * world <= 8.  The caller owns slice [offset_floats, offset_floats + n_floats) (multiples of 4): the kernel loads
 * it from all `world` buckets, adds in rank order and stores the sum back into all of them.  The caller brackets the
 * call with two cross-rank barriers on `stream` (gradients complete / slices landed), see partgs_b200/dist.py.
 * max_ctas bounds the SMs the collective takes from concurrently running kernels (0 = default 32). */
int pgs_peer_allreduce_slice(int world, float* const* buckets, size_t offset_floats, size_t n_floats, int max_ctas,
                             void* stream);

/* ---- simple-knn ------------------------------------------------------------------
 * Replaces distCUDA2 / SimpleKNN::knn (KNN/spatial.cu:15-25, KNN/simple_knn.cu:185-221):
 * mean squared distance of every point to its 3 nearest neighbours (exact).  Two grid levels over the bounding
 * box: a ring search with pruning in the fine one, a bottom-up octree walk (Morton-numbered coarse cells) for the
 * queries the rings leave open.
 * points [P,3] float32 device, mean_dist2 [P] float32 device, temp >= pgs_knn_temp_bytes(P).
 * Synchronises the stream once (bounding-box read-back). */
size_t pgs_knn_temp_bytes(int P);
int pgs_knn_dist2(int P, const float* points, float* mean_dist2, void* temp, void* stream);

/* ---- binning stages, individually callable (parity tests compare each stage) ---- */
/* cub::DeviceScan::InclusiveSum, rasterizer_impl.cu:278 */
size_t pgs_scan_temp_bytes(int n);
int pgs_inclusive_scan_u32(const uint32_t* in, uint32_t* out, int n, void* temp, void* stream);
/* cub::DeviceRadixSort::SortPairs(u64,u32, bits [0,end_bit)), rasterizer_impl.cu:304-309.
 * Returns 0 if the sorted output is in (keys_a, vals_a), 1 if in (keys_b, vals_b). */
size_t pgs_sort_temp_bytes(int n, int end_bit);
int pgs_sort_pairs_u64(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, int n, int end_bit,
                       void* temp, void* stream);
/* duplicateWithKeys, rasterizer_impl.cu:70-111, run on a geometry buffer filled by
 * pgs_dsr_forward; keys/values must hold num_rendered entries. */
int pgs_dsr_duplicate_with_keys(int P, const char* geom_buffer, int width, int height, const int* radii,
                                uint64_t* keys, uint32_t* values, void* stream);
/* BinningState::point_list_keys (rasterizer_impl.cu:187-194): the (tile << 32 | depth bits) keys of the sorted
 * instance list of the frame pgs_dsr_forward last rendered into these buffers.  The production path sorts the
 * surfels by depth and the instances by tile id only and never materialises the 64-bit keys; this entry rebuilds
 * them (keys [num_rendered]) for stage-wise parity checks against the reference's sorted keys. */
int pgs_dsr_sorted_keys(int P, int width, int height, const char* geom_buffer, const char* binning_buffer,
                        size_t binning_bytes, int num_rendered, uint64_t* keys, void* stream);
/* identifyTileRanges, rasterizer_impl.cu:116-138 (ranges [ntiles] uint2, zeroed here). */
int pgs_identify_tile_ranges(int L, const uint64_t* sorted_keys, uint32_t* ranges, int ntiles, void* stream);
/* getHigherMsb, rasterizer_impl.cu:35-50 */
uint32_t pgs_higher_msb(uint32_t n);

/* Layout of the opaque buffers (byte offsets from the buffer base, which must be
 * 256-byte aligned), so that tests can inspect intermediate state the way the
 * reference's fromChunk does (rasterizer_impl.cu:155-194). */
typedef struct {
  size_t geom_bytes, geom_rec, geom_bbox, geom_radii, geom_tiles_touched, geom_point_offsets;
  size_t image_bytes, image_final_T, image_n_contrib, image_ranges;
  size_t binning_bytes, binning_keys_sorted, binning_point_list;
  size_t binning_frag_mask; /* u32 [8 warps][binning_mask_stride]: pixels of each warp footprint that blended instance i */
  size_t binning_mask_stride;
  int rec_floats;  /* floats per surfel record */
  int tile_pixels; /* per-pixel state is tile-major [tile][256] */
} pgs_dsr_layout;
int pgs_dsr_get_layout(int P, int width, int height, size_t binning_bytes, pgs_dsr_layout* out);

#ifdef __cplusplus
}
#endif
#endif /* PARTGS_B200_H_ */
