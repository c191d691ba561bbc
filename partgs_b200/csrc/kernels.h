// Internal launch interface between the C-ABI layer (api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pgs {

// bookkeeping for pgs_launch_count(): every kernel launch of this library is counted
void count_launch(int n = 1);

// ---- superquadric -> surfel parameterisation (games/block_mesh_splatting) ------------
struct SqArgs {
  int B, Vt, F, K;
  const float* sq_r;       // [B,4] raw quaternion (w,x,y,z)
  const float* sq_s;       // [B,3] raw log-scale
  const float* sq_t;       // [B,3]
  const float* sq_eps;     // [B,2] raw
  const float* sq_occ;     // [B,1] raw
  const float* eta;        // [B,Vt]
  const float* omega;      // [B,Vt]
  const int* faces;        // [B,F,3] int32
  const float* alpha;      // [B*F,K,3] normalised barycentrics
  const float* scale_raw;  // [B,F*K]
  float ratio, scale_min;
};

struct PreprocessFwdArgs {
  int P, D, M;
  const float* means3D;
  const float* scales;
  float scale_modifier;
  const float* rotations;
  const float* opacities;
  const float* shs;
  const float* transMat_precomp;
  const float* colors_precomp;
  const float* viewmatrix;
  const float* projmatrix;
  const float* cam_pos;
  int W, H;
  int grid_x, grid_y;
  float focal_x, focal_y;  // `_part` fork only
  // outputs
  int* radii;
  float4* rec;    // [P][REC_QUADS]
  float4* bbox;   // [P][CULL_QUADS] cull records (box + conic)
  uint32_t* tiles_touched;
  uint32_t* depth_key;  // [P] bits of the view depth; 0xffffffff for culled surfels (they sort to the end)
  uint32_t* depth_hist; // [4][256] digit histograms of depth_key, accumulated here (zeroed by the caller); or nullptr
  uint2* rect;          // [P] tile rectangle {x0 | x1 << 16, y0 | y1 << 16}; empty for culled surfels
  // block-level mode: surfel i is generated in the kernel from the superquadric parameters (means3D, scales,
  // rotations, opacities are then ignored); sq_out_* optionally materialise what was generated
  bool use_sq;
  SqArgs sq;
  const float* sq_vertices;  // [B,Vt,3] from launch_sq_vertices
  float* sq_out_xyz;         // [P,3] or nullptr
  float* sq_out_scaling;     // [P,2] log-scales or nullptr
  float* sq_out_rotation;    // [P,4] or nullptr
  float* sq_out_opacity;     // [P] or nullptr
};
void launch_preprocess_fwd(const PreprocessFwdArgs& a, cudaStream_t s);
void launch_preprocess_fwd_part(const PreprocessFwdArgs& a, cudaStream_t s);
void launch_check_frustum(int P, const float* means3D, const float* viewmatrix, unsigned char* present,
                          cudaStream_t s);

// ---- binning ----------------------------------------------------------------
size_t scan_temp_bytes(int n);
// inclusive prefix sum of uint32 (reference: cub::DeviceScan::InclusiveSum, rasterizer_impl.cu:278)
void launch_inclusive_scan_u32(const uint32_t* in, uint32_t* out, int n, void* temp, cudaStream_t s);

// reference duplicateWithKeys (rasterizer_impl.cu:70-111)
// `capacity`: number of instances keys/values can hold; if offsets[P-1] exceeds it the launch does nothing
void launch_duplicate_with_keys(int P, const float4* rec, const uint32_t* offsets, uint64_t* keys, uint32_t* values,
                                const int* radii, int grid_x, int grid_y, cudaStream_t s,
                                uint32_t capacity = 0xffffffffu);

// Stable LSD radix sort of (u64 key, u32 value) pairs on key bits [0, end_bit)
// (reference: cub::DeviceRadixSort::SortPairs, rasterizer_impl.cu:304-309).
// Buffers a/b ping-pong; returns 0 if the sorted result is in (keys_a, vals_a), 1 if in (keys_b, vals_b).
size_t radix_sort_temp_bytes(int n, int end_bit);
// n_dev != nullptr: `n` is only the capacity the launch is sized for, the element count is read on the device
// (speculative launch before the host knows the count; does nothing if *n_dev > n)
int launch_radix_sort_pairs(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, int n, int end_bit,
                            void* temp, cudaStream_t s, const uint32_t* n_dev = nullptr);
// 32-bit key variant (distCUDA2 Morton/cell sort)
size_t radix_sort32_temp_bytes(int n, int end_bit);
int launch_radix_sort_pairs32(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, int n,
                              int end_bit, void* temp, cudaStream_t s);

// reference identifyTileRanges (rasterizer_impl.cu:116-138); ranges must be zeroed first.
void launch_identify_tile_ranges(int L, const uint64_t* keys, uint2* ranges, cudaStream_t s,
                                 const uint32_t* n_dev = nullptr);
void launch_identify_tile_ranges32(int L, const uint32_t* tile_keys, uint2* ranges, cudaStream_t s,
                                   const uint32_t* n_dev = nullptr);

// ---- production binning path: depth-ordered emission + tile-id sort (binning.cu) ------------------------------
#define PGS_RS_MAX_PASSES 8
struct RsPlan {  // digit passes of a radix sort
  int passes;
  int shift[PGS_RS_MAX_PASSES];
  int bits[PGS_RS_MAX_PASSES];
};
RsPlan rs_plan_even(int end_bit);
// index sort: (keys_a, positions) sorted by keys_a on bits [0, end_bit); vals_a is scratch.  Returns where (0: a, 1: b)
// hist_ready: the four digit histograms were accumulated into `temp` by the producer of the keys (preprocess,
// after radix_sort32_prepare zeroed it)
int launch_radix_sort_index32(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, int n,
                              int end_bit, void* temp, cudaStream_t s, bool hist_ready = false);
void radix_sort32_prepare(int n, int end_bit, void* temp, cudaStream_t s);
size_t radix_sort_plan_temp_bytes(int n, const RsPlan& pl);
void radix_sort_plan_prepare(int n, const RsPlan& pl, void* temp, cudaStream_t s);  // zero histograms + look-back
// sort whose digit histograms were accumulated into `temp` by launch_emit_instances (after radix_sort_plan_prepare)
int launch_radix_sort_plan32(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, int n,
                             const RsPlan& pl, void* temp, cudaStream_t s, const uint32_t* n_dev,
                             uint2* ranges = nullptr);  // ranges: identifyTileRanges fused into the last pass
struct EmitArgs {
  int P;
  const uint32_t* sorted_ids;  // [P] surfel indices in depth order (culled surfels last)
  const uint2* rect;           // [P] tile rectangles written by preprocess
  unsigned gx;
  uint32_t* keys;              // [capacity] out: tile id of every instance
  uint32_t* vals;              // [capacity] out: surfel index of every instance
  uint32_t capacity;
  uint32_t* total;             // out: number of instances of the frame
  uint32_t* hist;              // digit histograms of the tile sort (head of its temp storage)
  uint2* ranges;               // [ntiles] initialised here to {0xffffffff, 0}
  int ntiles;
  RsPlan plan;
  uint32_t* counter;           // set by the launcher
  unsigned long long* state;   // set by the launcher
};
size_t emit_state_bytes(int P);
void launch_emit_instances(const EmitArgs& a, void* state_mem, cudaStream_t s);
void launch_rebuild_sorted_keys(int L, const uint32_t* tile_keys, const uint32_t* point_list, const float4* rec,
                                uint64_t* keys, cudaStream_t s);

// ---- render -----------------------------------------------------------------
// tile ids sorted longest-list-first (launch order of the render kernels)
// also rewrites tile ranges that were never touched ({0xffffffff, 0}, see EmitArgs::ranges) as {0, 0}
void launch_tile_order(uint2* ranges, int ntiles, uint32_t* order, cudaStream_t s);

struct RenderFwdArgs {
  const uint2* ranges;
  const uint32_t* tile_order;  // [ntiles] or nullptr (identity)
  const uint32_t* point_list;
  int W, H;
  int grid_x, grid_y;
  const float4* rec;
  const float4* bbox;
  const float* bg_color;
  // per-pixel state kept for backward, tile-major [tile][256]
  float* final_T;     // [3][ntile*256]: T, M1, M2
  uint32_t* n_contrib;  // [2][ntile*256]: last, median
  float* out_color;   // [3][H][W]
  float* out_others;  // [7][H][W] (base) / [8][H][W] (`_part`)
  // per (warp, list position): the pixels of the warp's 8x4 footprint that blended the surfel;
  // written here, consumed by the backward pass.  [8][mask_stride], indexed [warp][range.x + pos]
  uint32_t* frag_mask;
  size_t mask_stride;
  // `_part` fork only
  int S;                   // semantic channels (<= MAX_SEMANTIC)
  const float* semantics;  // [P][S]
  float* out_semantic;     // [S][H][W]
};
constexpr int MAX_SEMANTIC = 16;  // DSRP/cuda_rasterizer/forward.cu:317 (fixed register array in the reference)
void launch_render_fwd(const RenderFwdArgs& a, cudaStream_t s);
void launch_render_fwd_part(const RenderFwdArgs& a, cudaStream_t s);

struct RenderBwdArgs {
  const uint2* ranges;
  const uint32_t* tile_order;  // [ntiles] or nullptr (identity)
  const uint32_t* point_list;
  int W, H;
  int grid_x, grid_y;
  const float4* rec;
  const float4* bbox;
  const float* bg_color;
  const float* final_T;
  const uint32_t* n_contrib;
  const float* dL_dpixels;  // [3][H][W]
  const float* dL_dothers;  // [7][H][W]
  float* grad;              // [P][GRAD_FLOATS], zeroed
  const uint32_t* frag_mask;  // forward's blend masks, [8][mask_stride]
  size_t mask_stride;
  // `_part` fork only
  int S;
  const float* semantics;       // [P][S]
  const float* dL_dsemantic;    // [S][H][W]
  float* grad_semantics;        // [P][S], zeroed
};
void launch_render_bwd(const RenderBwdArgs& a, cudaStream_t s);
void launch_render_bwd_part(const RenderBwdArgs& a, cudaStream_t s);

struct PreprocessBwdArgs {
  int P, D, M;
  const float* means3D;
  const int* radii;
  const float* shs;
  const float* scales;
  const float* rotations;
  float scale_modifier;
  const float* transMat_precomp;
  const float* viewmatrix;
  const float* projmatrix;
  float focal_x, focal_y, tan_fovx, tan_fovy;
  const float* cam_pos;
  const float4* rec;
  const float* grad;  // [P][GRAD_FLOATS] from render bwd
  // outputs (all written for every surfel; no pre-zeroing needed)
  float* dL_dmean2D;    // [P][3]
  float* dL_dcolors;    // [P][3]
  float* dL_dopacity;   // [P]
  float* dL_dmean3D;    // [P][3]
  float* dL_dtransMat;  // [P][9]
  float* dL_dsh;        // [P][M][3]
  float* dL_dscales;    // [P][2]  (block-level mode: gradient w.r.t. the LOG scales sq_surfels would store)
  float* dL_drots;      // [P][4]
  // base fork: add the five parameter gradients (mean3D, SH, opacity, scales, rotations) to what is already there
  // instead of overwriting (gradient accumulation over the views of a data-parallel batch)
  int accumulate;
  // block-level mode (see PreprocessFwdArgs)
  bool use_sq;
  SqArgs sq;
  const float* sq_vertices;
};
void launch_preprocess_bwd(const PreprocessBwdArgs& a, cudaStream_t s);
void launch_preprocess_bwd_part(const PreprocessBwdArgs& a, cudaStream_t s);

// ---- superquadric -> surfel parameterisation: launchers (SqArgs is declared at the top) ----
void launch_sq_vertices(const SqArgs& a, float* vertices, cudaStream_t s);
void launch_sq_forward(const SqArgs& a, float* vertices, float* xyz, float* scaling, float* rotation, float* opacity,
                       cudaStream_t s);
void launch_sq_backward(const SqArgs& a, const float* vertices, const float* d_xyz, const float* d_scaling,
                        const float* d_rotation, const float* d_opacity, float* d_vertices, float* d_occ_acc,
                        float* d_alpha, float* d_scale_raw, float* d_sq_r, float* d_sq_s, float* d_sq_t,
                        float* d_sq_eps, float* d_sq_occ, cudaStream_t s);

// ---- renderer post-processing: surface maps (surface_maps.cu) -----------------------------------
void launch_surface_maps_fwd(int W, int H, const float* allmap, const float* A, const float* M1, const float* M2,
                             const float* o, float ratio, float* rend_normal, float* surf_depth, float* surf_normal,
                             cudaStream_t s);
// scratch: 6*W*H floats (only touched when g_surf_normal != nullptr); g_allmap [7,H,W] is fully written
void launch_surface_maps_bwd(int W, int H, const float* allmap, const float* A, const float* M1, const float* M2,
                             const float* o, float ratio, const float* g_rend_normal, const float* g_surf_depth,
                             const float* g_surf_normal, float* scratch, float* g_allmap, cudaStream_t s);

// ---- photometric loss: L1 + SSIM (photometric.cu) ------------------------------------------------
// sums: 2 doubles {sum of the SSIM map, sum |img - gt|} (zeroed inside); dmaps: [3][C][H][W] derivative maps
void launch_photometric_fwd(int C, int H, int W, const float* img, const float* gt, double* sums, float* dmaps,
                            cudaStream_t s);
void launch_photometric_bwd(int C, int H, int W, const float* img, const float* gt, const float* dmaps,
                            const float* g_loss, float lambda, float* g_img, cudaStream_t s);

// ---- per-pixel regularisers: mask entropy, normal consistency, distortion (regularizers.cu) ---------
// sums: 3 doubles (zeroed inside); any of mask / rend_normal+surf_normal / dist may be NULL (term skipped)
void launch_regularizers_fwd(int npix, const float* alpha, const float* mask, const float* dist,
                             const float* rend_normal, const float* surf_normal, double* sums, cudaStream_t s);
void launch_regularizers_bwd(int npix, const float* alpha, const float* mask, const float* rend_normal,
                             const float* surf_normal, const float* g_loss, float lambda_entropy, float lambda_normal,
                             float lambda_dist, float* g_alpha, float* g_dist, float* g_rend_normal,
                             float* g_surf_normal, cudaStream_t s);

// ---- optimiser step / densification statistics (optim.cu) -----------------------------------------
#define PGS_ADAM_MAX_TENSORS 16
struct AdamTable {
  int n;
  float* param[PGS_ADAM_MAX_TENSORS];
  const float* grad[PGS_ADAM_MAX_TENSORS];
  float* exp_avg[PGS_ADAM_MAX_TENSORS];
  float* exp_avg_sq[PGS_ADAM_MAX_TENSORS];
  size_t numel[PGS_ADAM_MAX_TENSORS];
  float step_size[PGS_ADAM_MAX_TENSORS];
  int block_start[PGS_ADAM_MAX_TENSORS];
};
int launch_adam_multi(AdamTable& t, double beta1, double beta2, double eps, double bias_correction2_sqrt,
                      cudaStream_t s);
void launch_densify_stats(int P, const int* radii, const float* grad_means2D, float* max_radii2D, float* accum,
                          float* denom, cudaStream_t s);

// ---- densification as one planned compaction (densify.cu) -------------------------------------------
#define PGS_GATHER_MAX_TENSORS 24
struct GatherTable {
  int n;
  const float* src[PGS_GATHER_MAX_TENSORS];
  float* dst[PGS_GATHER_MAX_TENSORS];
  size_t numel[PGS_GATHER_MAX_TENSORS];  // output elements = output rows x width
  int width[PGS_GATHER_MAX_TENSORS];     // floats per row
  int zero_new[PGS_GATHER_MAX_TENSORS];  // rows >= n_keep are written as zeros (Adam moments of new surfels)
  int block_start[PGS_GATHER_MAX_TENSORS];
};
int densify_blocks(int P);
// counts (device, 8 x u32): [0] surviving originals, [1] surviving clones, [2] rows selected for splitting,
// [3] surviving split children per replica, [4] rows selected for cloning
void launch_densify_plan(int P, const float* accum, const float* denom, const float* scaling, const float* opacity,
                         float max_grad, float dense_thr, float min_opacity, int use_ws, float ws_thr,
                         float inv_divisor, unsigned char* code, uint32_t* block_off, uint32_t* counts,
                         cudaStream_t s);
void launch_densify_map(int P, const unsigned char* code, const uint32_t* block_off, const uint32_t* counts,
                        int n_split, int* src_row, int* sample_row, cudaStream_t s);
void launch_densify_gather(GatherTable& t, int n_keep, const int* src_row, cudaStream_t s);
void launch_densify_children(int n_children, const uint32_t* counts, const int* src_row, const int* sample_row,
                             const float* z, const float* xyz_in, const float* scaling_in, const float* rotation_in,
                             float inv_divisor, float* xyz_out, float* scaling_out, cudaStream_t s);

// ---- per-view epilogue of the extraction loop (extract.cu) -----------------------------------------
void launch_extract_maps(int npix, int S, const float* semantic, const float* palette, int palette_stride,
                         const float* rend_normal, float* part_rgb, float* normal_unit, cudaStream_t s);

// ---- gradient all-reduce over NVLink peer memory (collective.cu) ------------------------------------
#define PGS_PEER_MAX_WORLD 8
struct PeerBuckets { float* p[PGS_PEER_MAX_WORLD]; };  // the same bucket on every rank (peer-mapped addresses)
// returns 0, -1 (slice not 16-byte aligned), -2 (world size not supported)
int launch_peer_allreduce_slice(const PeerBuckets& b, int world, size_t offset_floats, size_t n_floats, int ctas,
                                cudaStream_t s);

// ---- distCUDA2 (simple-knn) ---------------------------------------------------
size_t knn_temp_bytes(int P);
// returns 0, or <0 with a message in err (does one stream sync for the bounding box)
int launch_knn_dist2(int P, const float* points, float* out, void* temp, cudaStream_t s, char* err, size_t errlen);

}  // namespace pgs
