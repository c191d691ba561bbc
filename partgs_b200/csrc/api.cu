// C-ABI layer of partgs_b200 (see include/partgs_b200.h).  Host orchestration of the
// base rasteriser: buffer carving, kernel sequencing on the caller's stream.
// Mirrors the call order of reference CudaRasterizer::Rasterizer::forward/backward
// (cuda_rasterizer/rasterizer_impl.cu:198-342, 346-448).
#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "../../include/partgs_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace pgs {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int check_cuda(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(PGS_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}
int check_sync(cudaStream_t s, const char* what) {
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return set_error(PGS_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

// ---- optional per-stage device timing (CUDA events on the launching stream) ------
// bench.py uses it to time the dominant kernel live inside its timed region.
struct StageSpan { int stage; int dev; cudaEvent_t a, b; };
static std::atomic<int> g_timing{0};
static std::mutex g_timing_mu;
static std::vector<StageSpan> g_spans;
// CUDA events belong to the device they were created on: one pool per device (a process may drive several GPUs)
constexpr int PGS_MAX_DEVICES = 64;
static std::vector<cudaEvent_t> g_event_pool[PGS_MAX_DEVICES];
static int current_device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < PGS_MAX_DEVICES) ? dev : -1;
}
static cudaEvent_t get_event(int dev) {
  auto& pool = g_event_pool[dev];
  if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct StageTimer {
  bool on; int stage; int dev; cudaStream_t s; cudaEvent_t a, b;
  StageTimer(int stage_, cudaStream_t s_) : on(g_timing.load() != 0), stage(stage_), dev(-1), s(s_) {
    if (on) {
      dev = current_device_slot();
      on = dev >= 0;
    }
    if (on) { std::lock_guard<std::mutex> lk(g_timing_mu); a = get_event(dev); b = get_event(dev); cudaEventRecord(a, s); }
  }
  ~StageTimer() {
    if (on) { cudaEventRecord(b, s); std::lock_guard<std::mutex> lk(g_timing_mu); g_spans.push_back({stage, dev, a, b}); }
  }
};

// pinned host word + event through which forward reads the instance count of the frame in flight
struct CountFetch {
  // a ring of pinned words: a lazy forward (PGS_FWD_LAZY_COUNT) leaves its count in the next slot and returns
  // without waiting; pgs_dsr_resolve_count() later checks every slot still pending
  static constexpr int SLOTS = 32;
  int* host = nullptr;            // [SLOTS] pinned
  cudaEvent_t ev[SLOTS] = {};
  size_t cap[SLOTS] = {};         // capacity the frame was launched with
  bool pending[SLOTS] = {};
  bool captured[SLOTS] = {};      // the slot belongs to a captured CUDA graph (every replay refreshes it): never reused
  int next = 0, last = 0;
  bool init() {
    if (host) return true;
    if (cudaHostAlloc((void**)&host, SLOTS * sizeof(int), cudaHostAllocDefault) != cudaSuccess) { host = nullptr; return false; }
    for (int i = 0; i < SLOTS; i++) {
      host[i] = 0;
      if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) return false;
    }
    return true;
  }
};
// One ring per device, shared by all host threads (PyTorch runs the backward pass on its autograd thread: the count
// a forward left pending on the caller's thread is resolved from there).  Calls on one device are expected to be
// ordered by the caller (they share a stream); the mutex only keeps the bookkeeping consistent.
static std::mutex g_count_mu;
static CountFetch* count_fetch() {
  static CountFetch cf[PGS_MAX_DEVICES];
  const int dev = current_device_slot();
  return dev < 0 ? nullptr : &cf[dev];
}

// instance capacity remembered per device (grow-only with slow decay): what speculative / lazy frames are sized for
static std::atomic<size_t> g_capacity_hint[PGS_MAX_DEVICES];

// reference getHigherMsb (rasterizer_impl.cu:35-50)
static uint32_t higher_msb(uint32_t n) {
  uint32_t msb = sizeof(n) * 4;
  uint32_t step = msb;
  while (step > 1) {
    step /= 2;
    if (n >> msb)
      msb += step;
    else
      msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

struct GeomState {
  float4* rec;
  float4* bbox;
  int* radii;
  uint32_t* tiles_touched;
  uint32_t* point_offsets;  // stage-wise entry points only (pgs_dsr_duplicate_with_keys)
  char* scan_temp;
  // production binning path (binning.cu, "instance emission in depth order")
  uint2* rect;              // tile rectangle of every surfel
  uint32_t* dkey_a, *dkey_b;  // depth bits, ping-pong of the depth sort
  uint32_t* dval_a, *dval_b;  // surfel indices, ditto; the depth order ends up in dval_a
  uint32_t* total;          // number of instances of the frame (device copy)
  char* dsort_temp;
  char* emit_state;
  static GeomState from(char*& p, size_t P) {
    GeomState g;
    carve(p, g.rec, P * REC_QUADS);
    carve(p, g.bbox, P * CULL_QUADS);
    carve(p, g.radii, P);
    carve(p, g.tiles_touched, P);
    carve(p, g.point_offsets, P);
    carve(p, g.scan_temp, scan_temp_bytes((int)P));
    carve(p, g.rect, P);
    carve(p, g.dkey_a, P);
    carve(p, g.dkey_b, P);
    carve(p, g.dval_a, P);
    carve(p, g.dval_b, P);
    carve(p, g.total, 64);
    carve(p, g.dsort_temp, radix_sort32_temp_bytes((int)P, 32));
    carve(p, g.emit_state, emit_state_bytes((int)P));
    return g;
  }
};
struct ImageState {
  float* final_T;
  uint32_t* n_contrib;
  uint2* ranges;
  uint32_t* tile_order;
  static ImageState from(char*& p, size_t ntiles) {
    ImageState s;
    carve(p, s.final_T, 3 * ntiles * TILE_PIX);
    carve(p, s.n_contrib, 2 * ntiles * TILE_PIX);
    carve(p, s.ranges, ntiles);
    carve(p, s.tile_order, ntiles);
    return s;
  }
};
struct BinningState {
  uint64_t* keys_a;  // production path: two u32[R] halves = ping-pong of the tile-id sort
  uint64_t* keys_b;  // the reference's sorted 64-bit keys, rebuilt on request (pgs_dsr_sorted_keys)
  uint32_t* vals_a;
  uint32_t* vals_b;
  uint32_t* frag_mask;  // [8 warps][mask_stride]: forward's per-warp blend masks (render.cu)
  size_t mask_stride;
  char* sort_temp;
  // The arena is laid out for a CAPACITY (a multiple of 512 Ki instances), not for the frame's exact
  // instance count: forward launches binning + render before the host has read the count (see
  // forward_impl), and a capacity that only ever grows makes the request the caller's allocator sees
  // constant from frame to frame.  Backward recovers the capacity from the buffer size.
  static constexpr size_t GRAIN = (size_t)1 << 19;
  static size_t capacity_for(size_t instances) { return align_up(instances > 0 ? instances : 1, GRAIN); }
  static BinningState from(char*& p, size_t capacity, int end_bit) {
    const size_t R = capacity;
    BinningState b;
    carve(p, b.vals_a, R);  // the final point list always ends up here
    b.mask_stride = R;
    carve(p, b.frag_mask, (size_t)(TILE_PIX / 32) * R);
    carve(p, b.vals_b, R);
    carve(p, b.keys_a, R);
    carve(p, b.keys_b, R);
    carve(p, b.sort_temp, radix_sort_temp_bytes((int)R, end_bit));  // >= the two-pass tile sort's need
    return b;
  }
  uint32_t* tile_keys(int which) const { return reinterpret_cast<uint32_t*>(keys_a) + (which ? mask_stride : 0); }
  static size_t bytes(size_t capacity, int end_bit) {
    char* p = nullptr;
    from(p, capacity, end_bit);
    return (size_t)p + 256;
  }
  // inverse of bytes(): 0 if `nbytes` is not the size of any capacity
  static size_t capacity_from_bytes(size_t nbytes, int end_bit) {
    for (size_t c = GRAIN; c <= ((size_t)1 << 31); c += GRAIN) {
      const size_t b = bytes(c, end_bit);
      if (b == nbytes) return c;
      if (b > nbytes) break;
    }
    return 0;
  }
};
template <typename F> static size_t required(F f) {
  char* p = nullptr;
  f(p);
  return (size_t)p + 256;
}

}  // namespace pgs

using namespace pgs;

extern "C" {

const char* pgs_last_error(void) { return g_err; }
int pgs_version(void) { return 100; }
unsigned long long pgs_launch_count(void) { return g_launches.load(); }
uint32_t pgs_higher_msb(uint32_t n) { return higher_msb(n); }

void pgs_timing_enable(int on) { g_timing.store(on ? 1 : 0); }
int pgs_timing_read(double* ms, unsigned long long* counts, int reset) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  static double acc_ms[PGS_NUM_STAGES];
  static unsigned long long acc_n[PGS_NUM_STAGES];
  for (auto& sp : g_spans) {
    cudaError_t e = cudaEventSynchronize(sp.b);
    float t = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, sp.a, sp.b);
    if (e != cudaSuccess) return set_error(PGS_ERR_CUDA, "timing: %s", cudaGetErrorString(e));
    if (sp.stage >= 0 && sp.stage < PGS_NUM_STAGES) { acc_ms[sp.stage] += t; acc_n[sp.stage]++; }
    g_event_pool[sp.dev].push_back(sp.a);
    g_event_pool[sp.dev].push_back(sp.b);
  }
  g_spans.clear();
  for (int i = 0; i < PGS_NUM_STAGES; i++) {
    if (ms) ms[i] = acc_ms[i];
    if (counts) counts[i] = acc_n[i];
    if (reset) { acc_ms[i] = 0; acc_n[i] = 0; }
  }
  return 0;
}

int pgs_dsr_get_layout(int P, int width, int height, size_t binning_bytes, pgs_dsr_layout* out) {
  if (!out || P < 0 || width <= 0 || height <= 0) return set_error(PGS_ERR_INVALID_ARG, "bad layout query");
  const int gx = (width + TILE_X - 1) / TILE_X, gy = (height + TILE_Y - 1) / TILE_Y;
  const size_t ntiles = (size_t)gx * gy;
  const int end_bit = 32 + (int)higher_msb(gx * gy);
  memset(out, 0, sizeof(*out));
  {
    char* p = nullptr;
    GeomState g = GeomState::from(p, P);
    out->geom_bytes = (size_t)p + 256;
    out->geom_rec = (size_t)g.rec;
    out->geom_bbox = (size_t)g.bbox;
    out->geom_radii = (size_t)g.radii;
    out->geom_tiles_touched = (size_t)g.tiles_touched;
    out->geom_point_offsets = (size_t)g.point_offsets;
  }
  {
    char* p = nullptr;
    ImageState s = ImageState::from(p, ntiles);
    out->image_bytes = (size_t)p + 256;
    out->image_final_T = (size_t)s.final_T;
    out->image_n_contrib = (size_t)s.n_contrib;
    out->image_ranges = (size_t)s.ranges;
  }
  {
    size_t cap = binning_bytes ? BinningState::capacity_from_bytes(binning_bytes, end_bit) : BinningState::GRAIN;
    if (cap == 0) return set_error(PGS_ERR_INVALID_ARG, "binning buffer size %zu matches no capacity", binning_bytes);
    char* p = nullptr;
    BinningState b = BinningState::from(p, cap, end_bit);
    out->binning_bytes = (size_t)p + 256;
    out->binning_keys_sorted = (size_t)b.keys_b;  // filled by pgs_dsr_sorted_keys
    out->binning_point_list = (size_t)b.vals_a;
    out->binning_frag_mask = (size_t)b.frag_mask;
    out->binning_mask_stride = b.mask_stride;
  }
  out->rec_floats = REC_FLOATS;
  out->tile_pixels = TILE_PIX;
  return 0;
}

}  // extern "C"

// Shared host sequence of both forks (part == true: diff-surfel-rasterization_part).
struct SqForward {        // block-level mode of forward_impl
  SqArgs sq;
  float* vertices;        // [B,Vt,3] out
  float* out_xyz, *out_scaling, *out_rotation, *out_opacity;  // optional materialisation
};

static int forward_impl(const SqForward* sqf, bool part, int S, const float* semantics, float* out_semantic,
                        pgs_alloc_fn geometry_buffer, void* geometry_user, pgs_alloc_fn binning_buffer,
                        void* binning_user, pgs_alloc_fn image_buffer, void* image_user, int P, int D, int M,
                        const float* background, int width, int height, const float* means3D, const float* shs,
                        const float* colors_precomp, const float* opacities, const float* scales, float scale_modifier,
                        const float* rotations, const float* transMat_precomp, const float* viewmatrix,
                        const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered,
                        float* out_color, float* out_others, int* radii, int flags, void* stream) {
  (void)prefiltered;
  // flags: bit 0 = the reference's `debug` (synchronise and check after every stage), bit 1 = PGS_FWD_LAZY_COUNT
  const int debug = flags & 1;
  if (part) {
    if (S < 0 || S > MAX_SEMANTIC)
      return set_error(PGS_ERR_UNSUPPORTED, "semantic channels must be in [0, %d] (got %d)", MAX_SEMANTIC, S);
    if (S > 0 && (!semantics || !out_semantic)) return set_error(PGS_ERR_INVALID_ARG, "null semantics pointer");
    if (transMat_precomp)
      return set_error(PGS_ERR_UNSUPPORTED, "transMat_precomp is not usable in the _part fork (reference reads "
                                            "scales unconditionally and leaves the normal uninitialised)");
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (P <= 0 || width <= 0 || height <= 0) return set_error(PGS_ERR_INVALID_ARG, "P, width, height must be positive");
  if (!geometry_buffer || !binning_buffer || !image_buffer) return set_error(PGS_ERR_INVALID_ARG, "null allocator");
  if (!background || !viewmatrix || !projmatrix || !out_color || !out_others || (!sqf && (!means3D || !opacities)))
    return set_error(PGS_ERR_INVALID_ARG, "null required pointer");
  if (sqf && (part || transMat_precomp || !sqf->vertices))
    return set_error(PGS_ERR_INVALID_ARG, "block-level mode: base fork only, no transMat_precomp, vertices required");
  if (!shs && !colors_precomp)
    return set_error(PGS_ERR_INVALID_ARG, "provide SHs or precomputed colours");  // rasterizer_impl.cu:243-246
  if (shs && (!cam_pos || M <= 0)) return set_error(PGS_ERR_INVALID_ARG, "SH path needs campos and M > 0");
  if (!sqf && !transMat_precomp && (!scales || !rotations))
    return set_error(PGS_ERR_INVALID_ARG, "provide scales+rotations or transMat_precomp");
  if (D < 0 || D > 3 || (shs && (D + 1) * (D + 1) > M)) return set_error(PGS_ERR_INVALID_ARG, "bad SH degree");

  const int gx = (width + TILE_X - 1) / TILE_X, gy = (height + TILE_Y - 1) / TILE_Y;
  const size_t ntiles = (size_t)gx * gy;
  if (gx > 0xffff || gy > 0xffff) return set_error(PGS_ERR_UNSUPPORTED, "image too large (more than 65535 tiles per axis)");

  size_t geom_bytes = required([&](char*& p) { GeomState::from(p, P); });
  char* gptr = geometry_buffer(geom_bytes, geometry_user);
  if (!gptr) return set_error(PGS_ERR_ALLOC, "geometry buffer allocation failed (%zu B)", geom_bytes);
  GeomState geom = GeomState::from(gptr, P);
  if (radii == nullptr) radii = geom.radii;

  size_t img_bytes = required([&](char*& p) { ImageState::from(p, ntiles); });
  char* iptr = image_buffer(img_bytes, image_user);
  if (!iptr) return set_error(PGS_ERR_ALLOC, "image buffer allocation failed (%zu B)", img_bytes);
  ImageState img = ImageState::from(iptr, ntiles);

  PreprocessFwdArgs pa;
  pa.P = P; pa.D = D; pa.M = M;
  pa.means3D = means3D; pa.scales = scales; pa.scale_modifier = scale_modifier; pa.rotations = rotations;
  pa.opacities = opacities; pa.shs = shs; pa.transMat_precomp = transMat_precomp;
  pa.colors_precomp = colors_precomp; pa.viewmatrix = viewmatrix; pa.projmatrix = projmatrix; pa.cam_pos = cam_pos;
  pa.W = width; pa.H = height; pa.grid_x = gx; pa.grid_y = gy;
  pa.radii = radii; pa.rec = geom.rec; pa.bbox = geom.bbox; pa.tiles_touched = geom.tiles_touched;
  pa.depth_key = geom.dkey_a; pa.rect = geom.rect;
  // base fork: preprocess accumulates the digit histograms of the depth sort (head of its temp storage)
  pa.depth_hist = part ? nullptr : reinterpret_cast<uint32_t*>(geom.dsort_temp);
  radix_sort32_prepare(P, 32, geom.dsort_temp, s);
  pa.focal_y = height / (2.0f * tan_fovy);
  pa.focal_x = width / (2.0f * tan_fovx);
  pa.use_sq = sqf != nullptr;
  if (sqf) {
    pa.sq = sqf->sq; pa.sq_vertices = sqf->vertices;
    pa.sq_out_xyz = sqf->out_xyz; pa.sq_out_scaling = sqf->out_scaling; pa.sq_out_rotation = sqf->out_rotation;
    pa.sq_out_opacity = sqf->out_opacity;
    StageTimer t(PGS_STAGE_SQ_FWD, s);
    launch_sq_vertices(sqf->sq, sqf->vertices, s);  // B*Vt threads; the per-surfel part runs inside preprocess
  }
  {
    StageTimer t(PGS_STAGE_PREPROCESS_FWD, s);
    if (part) launch_preprocess_fwd_part(pa, s); else launch_preprocess_fwd(pa, s);
  }
  if (int e = check_cuda("preprocess_fwd")) return e;
  if (debug) if (int e = check_sync(s, "preprocess_fwd")) return e;

  // Depth order of the surfels (stable sort of the depth bits; culled surfels carry 0xffffffff and end up last).
  // Reference: the low 32 key bits of its one big sort, rasterizer_impl.cu:301-309.
  {
    StageTimer t(PGS_STAGE_SCAN, s);
    const int where = launch_radix_sort_index32(geom.dkey_a, geom.dval_a, geom.dkey_b, geom.dval_b, P, 32,
                                                geom.dsort_temp, s, /*hist_ready=*/pa.depth_hist != nullptr);
    if (where) return set_error(PGS_ERR_CUDA, "depth sort ended in the wrong buffer");  // 4 passes: never
  }
  if (int e = check_cuda("depth_sort")) return e;

  // Number of surfel x tile instances R (reference: blocking cudaMemcpy, rasterizer_impl.cu:282).  Binning +
  // render are launched SPECULATIVELY for a capacity remembered from earlier frames, the kernels reading the
  // count on the device (the emission kernel computes it); the count travels to pinned host memory
  // asynchronously and only then does the host wait for it: the GPU already has the rest of the frame queued, so
  // the round trip costs no device idle time, and the exact count is still returned to the caller.  If the count
  // exceeds the capacity the speculative kernels did nothing and everything is launched again with a larger
  // arena (first frames of a scene only).
  int dev = 0;
  cudaGetDevice(&dev);
  CountFetch* cfp = count_fetch();
  if (!cfp) return set_error(PGS_ERR_UNSUPPORTED, "device index %d out of range (max %d devices per process)", dev, PGS_MAX_DEVICES);
  CountFetch& cf = *cfp;
  std::unique_lock<std::mutex> count_lock(g_count_mu);
  if (!cf.init()) return set_error(PGS_ERR_CUDA, "pinned count buffer: %s", cudaGetErrorString(cudaGetLastError()));
  // Lazy count (flag bit 1, or implied by stream capture — a capture cannot contain the host wait): see below
  cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s, &cap_status);
  const bool capturing = cap_status != cudaStreamCaptureStatusNone;
  const bool want_lazy = !debug && (capturing || (flags & PGS_FWD_LAZY_COUNT));
  int slot = cf.next;
  for (int k = 0; k < CountFetch::SLOTS && cf.captured[slot]; k++) slot = (slot + 1) % CountFetch::SLOTS;
  if (cf.captured[slot])
    return set_error(PGS_ERR_UNSUPPORTED, "all %d count slots belong to captured graphs", CountFetch::SLOTS);
  if (cf.pending[slot])
    return set_error(PGS_ERR_UNSUPPORTED, "more than %d frames rendered with a lazy instance count without "
                                          "pgs_dsr_resolve_count()", CountFetch::SLOTS);
  const uint32_t* n_dev = geom.total;

  const int end_bit = 32 + (int)higher_msb(gx * gy);
  auto launch_rest = [&](size_t capacity, const uint32_t* count_dev, int count_host) -> int {
    const size_t bin_bytes = BinningState::bytes(capacity, end_bit);
    char* bptr = binning_buffer(bin_bytes, binning_user);
    if (!bptr) return set_error(PGS_ERR_ALLOC, "binning buffer allocation failed (%zu B)", bin_bytes);
    BinningState bin = BinningState::from(bptr, capacity, end_bit);
    const int n = count_dev ? (int)capacity : count_host;  // launch size

    // Emit the (tile, surfel) instances in depth order, then stable-sort them by tile id: the reference's
    // (tile | depth) order (duplicateWithKeys + SortPairs, rasterizer_impl.cu:70-111,301-309) with two digit
    // passes over R instead of six.  The sort must end in (tile_keys(0), vals_a): start in b for an odd pass count.
    const RsPlan plan = rs_plan_even(end_bit - 32);
    const int first = plan.passes & 1;
    uint32_t* k0 = bin.tile_keys(first), *k1 = bin.tile_keys(first ^ 1);
    uint32_t* v0 = first ? bin.vals_b : bin.vals_a, *v1 = first ? bin.vals_a : bin.vals_b;
    {
      StageTimer t(PGS_STAGE_DUP_KEYS, s);
      radix_sort_plan_prepare(n, plan, bin.sort_temp, s);
      EmitArgs ea;
      ea.P = P; ea.sorted_ids = geom.dval_a; ea.rect = geom.rect; ea.gx = (unsigned)gx;
      ea.keys = k0; ea.vals = v0; ea.capacity = (uint32_t)capacity; ea.total = geom.total;
      ea.hist = reinterpret_cast<uint32_t*>(bin.sort_temp); ea.plan = plan; ea.counter = nullptr; ea.state = nullptr;
      ea.ranges = img.ranges; ea.ntiles = (int)ntiles;
      launch_emit_instances(ea, geom.emit_state, s);
    }
    if (int e = check_cuda("emit_instances")) return e;
    if (count_dev) {  // speculative launch: the count goes to the host while the rest of the frame is queued
      cudaError_t ce = cudaMemcpyAsync(cf.host + slot, geom.total, sizeof(int), cudaMemcpyDeviceToHost, s);
      if (ce != cudaSuccess) return set_error(PGS_ERR_CUDA, "memcpy num_rendered: %s", cudaGetErrorString(ce));
      if (!capturing) cudaEventRecord(cf.ev[slot], s);
    }
    if (n > 0) {
      { StageTimer t(PGS_STAGE_SORT, s);
        // identifyTileRanges (rasterizer_impl.cu:116-138) is fused into the last pass
        launch_radix_sort_plan32(k0, v0, k1, v1, n, plan, bin.sort_temp, s, count_dev, img.ranges); }
      if (int e = check_cuda("radix_sort")) return e;
      if (debug) if (int e = check_sync(s, "binning")) return e;
    }
    launch_tile_order(img.ranges, (int)ntiles, img.tile_order, s);
    if (int e = check_cuda("tile_order")) return e;

    RenderFwdArgs ra;
    ra.ranges = img.ranges; ra.tile_order = img.tile_order; ra.point_list = bin.vals_a; ra.W = width; ra.H = height;
    ra.grid_x = gx; ra.grid_y = gy;
    ra.rec = geom.rec; ra.bbox = geom.bbox; ra.bg_color = background;
    ra.final_T = img.final_T; ra.n_contrib = img.n_contrib; ra.out_color = out_color; ra.out_others = out_others;
    ra.S = S; ra.semantics = semantics; ra.out_semantic = out_semantic;
    ra.frag_mask = bin.frag_mask; ra.mask_stride = bin.mask_stride;
    {
      StageTimer t(PGS_STAGE_RENDER_FWD, s);
      if (part) launch_render_fwd_part(ra, s); else launch_render_fwd(ra, s);
    }
    if (int e = check_cuda("render_fwd")) return e;
    return 0;
  };

  static const bool no_spec = getenv("PGS_NO_SPECULATE") != nullptr;  // diagnostic switch
  const bool speculate = !debug && !no_spec && g_capacity_hint[dev].load() > 0;
  if (want_lazy && !speculate)
    if (capturing)
      return set_error(PGS_ERR_UNSUPPORTED, "stream capture needs a remembered instance capacity: render one frame of "
                                            "the scene eagerly before capturing");
  const bool lazy = want_lazy && speculate;
  size_t capacity = 0;
  if (speculate) {
    capacity = BinningState::capacity_for(g_capacity_hint[dev].load());
    if (int e = launch_rest(capacity, n_dev, 0)) return e;
  } else {
    // no capacity to speculate with (first frame, debug mode): count first, as the reference does
    launch_inclusive_scan_u32(geom.tiles_touched, geom.point_offsets, P, geom.scan_temp, s);
    if (int e = check_cuda("scan")) return e;
    cudaError_t ce0 = cudaMemcpyAsync(cf.host + slot, geom.point_offsets + P - 1, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (ce0 != cudaSuccess) return set_error(PGS_ERR_CUDA, "memcpy num_rendered: %s", cudaGetErrorString(ce0));
    cudaEventRecord(cf.ev[slot], s);
  }
  if (lazy) {
    // The caller does not want to wait: the frame is queued for `capacity`; whether it fitted is established by
    // pgs_dsr_resolve_count() (at the backward's entry, or after a CUDA-graph replay).
    cf.pending[slot] = !capturing;   // a captured frame has no host-side bookkeeping: replays just refresh the slot
    cf.captured[slot] = capturing;
    cf.cap[slot] = capacity;
    cf.last = slot;
    cf.next = (slot + 1) % CountFetch::SLOTS;
    return PGS_COUNT_PENDING;
  }
  cudaError_t ce = cudaEventSynchronize(cf.ev[slot]);
  if (ce != cudaSuccess) return set_error(PGS_ERR_CUDA, "scan/num_rendered: %s", cudaGetErrorString(ce));
  const int num_rendered = cf.host[slot];
  cf.cap[slot] = capacity;
  cf.last = slot;
  if (num_rendered < 0) return set_error(PGS_ERR_UNSUPPORTED, "more than 2^31 surfel-tile instances");
  if (!speculate || (size_t)num_rendered > capacity) {
    // first frame / debug mode / the speculative capacity was too small
    size_t want = (size_t)num_rendered + (size_t)num_rendered / 4;
    if (!debug) {
      size_t cur = g_capacity_hint[dev].load();
      while (want > cur && !g_capacity_hint[dev].compare_exchange_weak(cur, want)) {}
      want = std::max(want, cur);
    }
    capacity = BinningState::capacity_for(debug ? (size_t)num_rendered : want);
    if (int e = launch_rest(capacity, nullptr, num_rendered)) return e;
  } else if ((size_t)num_rendered * 2 < g_capacity_hint[dev].load()) {
    // the arena is more than twice what this frame needed: let the remembered capacity decay (3 % per such
    // frame, never below 1.25 x this frame), so that one exceptional frame does not size every later arena
    size_t cur = g_capacity_hint[dev].load();
    const size_t floor_ = (size_t)num_rendered + (size_t)num_rendered / 4;
    g_capacity_hint[dev].store(std::max(floor_, cur - cur / 32));
  }
  if (debug) if (int e = check_sync(s, "render_fwd")) return e;
  return num_rendered;
}

extern "C" {

int pgs_dsr_forward(pgs_alloc_fn geometry_buffer, void* geometry_user, pgs_alloc_fn binning_buffer,
                    void* binning_user, pgs_alloc_fn image_buffer, void* image_user, int P, int D, int M,
                    const float* background, int width, int height, const float* means3D, const float* shs,
                    const float* colors_precomp, const float* opacities, const float* scales, float scale_modifier,
                    const float* rotations, const float* transMat_precomp, const float* viewmatrix,
                    const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered,
                    float* out_color, float* out_others, int* radii, int debug, void* stream) {
  return forward_impl(nullptr, false, 0, nullptr, nullptr, geometry_buffer, geometry_user, binning_buffer, binning_user,
                      image_buffer, image_user, P, D, M, background, width, height, means3D, shs, colors_precomp,
                      opacities, scales, scale_modifier, rotations, transMat_precomp, viewmatrix, projmatrix, cam_pos,
                      tan_fovx, tan_fovy, prefiltered, out_color, out_others, radii, debug, stream);
}

int pgs_dsrp_forward(pgs_alloc_fn geometry_buffer, void* geometry_user, pgs_alloc_fn binning_buffer,
                     void* binning_user, pgs_alloc_fn image_buffer, void* image_user, int P, int D, int M,
                     const float* background, int width, int height, int semantic_types, const float* means3D,
                     const float* shs, const float* colors_precomp, const float* semantics, const float* opacities,
                     const float* scales, float scale_modifier, const float* rotations, const float* transMat_precomp,
                     const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                     float tan_fovy, int prefiltered, float* out_color, float* out_semantic, float* out_others,
                     int* radii, int debug, void* stream) {
  return forward_impl(nullptr, true, semantic_types, semantics, out_semantic, geometry_buffer, geometry_user, binning_buffer,
                      binning_user, image_buffer, image_user, P, D, M, background, width, height, means3D, shs,
                      colors_precomp, opacities, scales, scale_modifier, rotations, transMat_precomp, viewmatrix,
                      projmatrix, cam_pos, tan_fovx, tan_fovy, prefiltered, out_color, out_others, radii, debug, stream);
}

void pgs_dsr_set_capacity_hint(size_t instances) {
  const int dev = current_device_slot();
  if (dev >= 0) g_capacity_hint[dev].store(instances);
}

int pgs_dsr_resolve_count(int* overflow) {
  if (overflow) *overflow = 0;
  CountFetch* cf = count_fetch();
  std::unique_lock<std::mutex> count_lock(g_count_mu);
  if (!cf || !cf->host) return 0;
  const int dev = current_device_slot();
  int last_count = cf->host[cf->last];
  bool any = false;
  for (int k = 0; k < CountFetch::SLOTS; k++) {
    const int i = (cf->next + k) % CountFetch::SLOTS;   // oldest first
    if (!cf->pending[i]) continue;
    any = true;
    cudaError_t ce = cudaEventSynchronize(cf->ev[i]);
    if (ce != cudaSuccess) return set_error(PGS_ERR_CUDA, "resolve_count: %s", cudaGetErrorString(ce));
    cf->pending[i] = false;
    const int n = cf->host[i];
    last_count = n;
    if (n < 0 || (size_t)n > cf->cap[i]) {
      if (overflow) *overflow = 1;
      // remember the larger need, so that the re-rendered frame (and the following ones) fit
      size_t want = (size_t)(n < 0 ? 0x7fffffff : n);
      want += want / 4;
      size_t cur = g_capacity_hint[dev].load();
      while (want > cur && !g_capacity_hint[dev].compare_exchange_weak(cur, want)) {}
    }
  }
  if (!any) {
    // nothing pending: a captured frame replayed by the caller (who synchronised the stream): report its slot
    const int n = cf->host[cf->last];
    if (overflow && (n < 0 || (size_t)n > cf->cap[cf->last])) *overflow = 1;
    for (int i = 0; i < CountFetch::SLOTS; i++)   // every captured graph's latest replay
      if (overflow && cf->captured[i] && (cf->host[i] < 0 || (size_t)cf->host[i] > cf->cap[i])) *overflow = 1;
    return n;
  }
  return last_count;
}

size_t pgs_dsr_backward_scratch_bytes(int P) { return (size_t)(P > 0 ? P : 0) * GRAD_FLOATS * sizeof(float) + 256; }

}  // extern "C"

static int backward_impl(const SqArgs* sq, const float* sq_vertices, bool part, int S, const float* semantics,
                         const float* dL_dsemantic_pix, float* dL_dsemantics,
                         int P, int D, int M, int R, const float* background, int width, int height,
                         const float* means3D, const float* shs, const float* colors_precomp, const float* scales,
                         float scale_modifier, const float* rotations, const float* transMat_precomp,
                         const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                         float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer,
                         size_t binning_bytes, char* image_buffer,
                         const float* dL_dpix, const float* dL_dothers, float* dL_dmean2D, float* scratch,
                         float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dtransMat, float* dL_dsh,
                         float* dL_dscale, float* dL_drot, int flags, void* stream) {
  (void)colors_precomp;
  // flags: bit 0 = debug (synchronise and check after every stage, like the reference's `debug`), bit 1 =
  // PGS_BWD_ACCUMULATE (add the five parameter gradients to the arrays instead of overwriting them)
  const int debug = flags & 1;
  const int accumulate = (flags & PGS_BWD_ACCUMULATE) ? 1 : 0;
  if (accumulate && part)
    return set_error(PGS_ERR_UNSUPPORTED, "gradient accumulation is not implemented for the _part fork");
  // (block-level mode: only dL_dsh accumulates in the kernel — the per-surfel geometry gradients are scratch of the
  //  superquadric backward; the caller adds the five small block gradients itself)
  if (part) {
    if (S < 0 || S > MAX_SEMANTIC)
      return set_error(PGS_ERR_UNSUPPORTED, "semantic channels must be in [0, %d] (got %d)", MAX_SEMANTIC, S);
    if (S > 0 && (!semantics || !dL_dsemantic_pix || !dL_dsemantics))
      return set_error(PGS_ERR_INVALID_ARG, "null semantics pointer");
    if (transMat_precomp) return set_error(PGS_ERR_UNSUPPORTED, "transMat_precomp is not usable in the _part fork");
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (P <= 0 || width <= 0 || height <= 0 || R < 0) return set_error(PGS_ERR_INVALID_ARG, "bad sizes");
  const bool r_pending = R == PGS_COUNT_PENDING;   // lazy forward: the count is not known here; the arena's capacity is
  if (!geom_buffer || !image_buffer || (R > 0 && !binning_buffer))
    return set_error(PGS_ERR_INVALID_ARG, "null state buffer");
  if (!dL_dpix || !dL_dothers || !scratch || !dL_dmean2D || !dL_dopacity || !dL_dcolor || !dL_dmean3D ||
      !dL_dtransMat)
    return set_error(PGS_ERR_INVALID_ARG, "null gradient pointer");
  if (!transMat_precomp && ((!sq && (!scales || !rotations)) || !dL_dscale || !dL_drot))
    return set_error(PGS_ERR_INVALID_ARG, "scale/rotation path needs scales, rotations and their gradient arrays");
  if (shs && !dL_dsh) return set_error(PGS_ERR_INVALID_ARG, "SH path needs dL_dsh");

  const int gx = (width + TILE_X - 1) / TILE_X, gy = (height + TILE_Y - 1) / TILE_Y;
  const size_t ntiles = (size_t)gx * gy;
  const int end_bit = 32 + (int)higher_msb(gx * gy);

  GeomState geom = GeomState::from(geom_buffer, P);
  ImageState img = ImageState::from(image_buffer, ntiles);
  if (radii == nullptr) radii = geom.radii;
  const uint32_t* point_list = nullptr;
  const uint32_t* frag_mask = nullptr;
  size_t mask_stride = 0;
  if (R > 0) {
    const size_t capacity = BinningState::capacity_from_bytes(binning_bytes, end_bit);
    if (capacity == 0 || (!r_pending && capacity < (size_t)R))
      return set_error(PGS_ERR_INVALID_ARG, "binning buffer of %zu bytes was not produced by the forward pass of "
                                            "this frame (%d instances)", binning_bytes, R);
    BinningState bin = BinningState::from(binning_buffer, capacity, end_bit);
    point_list = bin.vals_a;
    frag_mask = bin.frag_mask;
    mask_stride = bin.mask_stride;
  }

  float* grad = reinterpret_cast<float*>(align_up(reinterpret_cast<size_t>(scratch), 256));
  cudaMemsetAsync(grad, 0, (size_t)P * GRAD_FLOATS * sizeof(float), s);

  const float focal_y = height / (2.0f * tan_fovy);
  const float focal_x = width / (2.0f * tan_fovx);

  RenderBwdArgs rb;
  rb.ranges = img.ranges; rb.tile_order = img.tile_order; rb.point_list = point_list; rb.W = width; rb.H = height; rb.grid_x = gx; rb.grid_y = gy;
  rb.rec = geom.rec; rb.bbox = geom.bbox; rb.bg_color = background; rb.final_T = img.final_T;
  rb.n_contrib = img.n_contrib; rb.dL_dpixels = dL_dpix; rb.dL_dothers = dL_dothers; rb.grad = grad;
  rb.S = S; rb.semantics = semantics; rb.dL_dsemantic = dL_dsemantic_pix; rb.grad_semantics = dL_dsemantics;
  rb.frag_mask = frag_mask; rb.mask_stride = mask_stride;
  if (part && S > 0) cudaMemsetAsync(dL_dsemantics, 0, (size_t)P * S * sizeof(float), s);
  {
    StageTimer t(PGS_STAGE_RENDER_BWD, s);
    if (part) launch_render_bwd_part(rb, s); else launch_render_bwd(rb, s);
  }
  if (int e = check_cuda("render_bwd")) return e;
  if (debug) if (int e = check_sync(s, "render_bwd")) return e;

  PreprocessBwdArgs pb;
  pb.P = P; pb.D = D; pb.M = M; pb.means3D = means3D; pb.radii = radii; pb.shs = shs;
  pb.scales = transMat_precomp ? nullptr : scales; pb.rotations = rotations; pb.scale_modifier = scale_modifier;
  pb.use_sq = sq != nullptr;
  pb.accumulate = accumulate;
  if (sq) { pb.sq = *sq; pb.sq_vertices = sq_vertices; }
  pb.transMat_precomp = transMat_precomp; pb.viewmatrix = viewmatrix; pb.projmatrix = projmatrix;
  pb.focal_x = focal_x; pb.focal_y = focal_y; pb.tan_fovx = tan_fovx; pb.tan_fovy = tan_fovy; pb.cam_pos = campos;
  pb.rec = geom.rec; pb.grad = grad;
  pb.dL_dmean2D = dL_dmean2D; pb.dL_dcolors = dL_dcolor; pb.dL_dopacity = dL_dopacity; pb.dL_dmean3D = dL_dmean3D;
  pb.dL_dtransMat = dL_dtransMat; pb.dL_dsh = dL_dsh; pb.dL_dscales = dL_dscale; pb.dL_drots = dL_drot;
  {
    StageTimer t(PGS_STAGE_PREPROCESS_BWD, s);
    if (part) launch_preprocess_bwd_part(pb, s); else launch_preprocess_bwd(pb, s);
  }
  if (int e = check_cuda("preprocess_bwd")) return e;
  if (debug) if (int e = check_sync(s, "preprocess_bwd")) return e;
  return 0;
}

extern "C" {

int pgs_dsr_backward(int P, int D, int M, int R, const float* background, int width, int height,
                     const float* means3D, const float* shs, const float* colors_precomp, const float* scales,
                     float scale_modifier, const float* rotations, const float* transMat_precomp,
                     const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                     float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer, size_t binning_bytes,
                     char* image_buffer, const float* dL_dpix, const float* dL_dothers, float* dL_dmean2D, float* scratch,
                     float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dtransMat, float* dL_dsh,
                     float* dL_dscale, float* dL_drot, int debug, void* stream) {
  return backward_impl(nullptr, nullptr, false, 0, nullptr, nullptr, nullptr, P, D, M, R, background, width, height, means3D, shs,
                       colors_precomp, scales, scale_modifier, rotations, transMat_precomp, viewmatrix, projmatrix,
                       campos, tan_fovx, tan_fovy, radii, geom_buffer, binning_buffer, binning_bytes, image_buffer,
                       dL_dpix, dL_dothers, dL_dmean2D, scratch, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dtransMat, dL_dsh,
                       dL_dscale, dL_drot, debug, stream);
}

int pgs_dsrp_backward(int P, int D, int M, int R, const float* background, int width, int height, int semantic_types,
                      const float* means3D, const float* shs, const float* colors_precomp, const float* semantics,
                      const float* scales, float scale_modifier, const float* rotations,
                      const float* transMat_precomp, const float* viewmatrix, const float* projmatrix,
                      const float* campos, float tan_fovx, float tan_fovy, const int* radii, char* geom_buffer,
                      char* binning_buffer, size_t binning_bytes, char* image_buffer, const float* dL_dpix,
                      const float* dL_dsemantic_pix,
                      const float* dL_dothers, float* dL_dmean2D, float* scratch, float* dL_dopacity, float* dL_dcolor,
                      float* dL_dsemantics, float* dL_dmean3D, float* dL_dtransMat, float* dL_dsh, float* dL_dscale,
                      float* dL_drot, int debug, void* stream) {
  return backward_impl(nullptr, nullptr, true, semantic_types, semantics, dL_dsemantic_pix, dL_dsemantics, P, D, M, R, background,
                       width, height, means3D, shs, colors_precomp, scales, scale_modifier, rotations,
                       transMat_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, radii, geom_buffer,
                       binning_buffer, binning_bytes, image_buffer, dL_dpix, dL_dothers, dL_dmean2D, scratch, dL_dopacity, dL_dcolor,
                       dL_dmean3D, dL_dtransMat, dL_dsh, dL_dscale, dL_drot, debug, stream);
}

int pgs_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     unsigned char* present, void* stream) {
  (void)projmatrix;
  if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return set_error(PGS_ERR_INVALID_ARG, "bad args");
  launch_check_frustum(P, means3D, viewmatrix, present, (cudaStream_t)stream);
  return check_cuda("check_frustum");
}

static int sq_fill(SqArgs& a, int B, int Vt, int F, int K, const float* sq_r, const float* sq_s, const float* sq_t,
                   const float* sq_eps, const float* sq_occ, const float* eta, const float* omega, const int* faces,
                   const float* alpha, const float* scale_raw, float ratio, float scale_min) {
  if (B <= 0 || Vt <= 0 || F <= 0 || K <= 0) return set_error(PGS_ERR_INVALID_ARG, "B, Vt, F, K must be positive");
  if ((long long)B * F * K > 0x7fffffffLL) return set_error(PGS_ERR_UNSUPPORTED, "too many surfels");
  if (!sq_r || !sq_s || !sq_t || !sq_eps || !sq_occ || !eta || !omega || !faces || !alpha || !scale_raw)
    return set_error(PGS_ERR_INVALID_ARG, "null superquadric input");
  a.B = B; a.Vt = Vt; a.F = F; a.K = K; a.sq_r = sq_r; a.sq_s = sq_s; a.sq_t = sq_t; a.sq_eps = sq_eps;
  a.sq_occ = sq_occ; a.eta = eta; a.omega = omega; a.faces = faces; a.alpha = alpha; a.scale_raw = scale_raw;
  a.ratio = ratio; a.scale_min = scale_min;
  return 0;
}

int pgs_sq2surfel_forward(int B, int Vt, int F, int K, const float* sq_r, const float* sq_s, const float* sq_t,
                          const float* sq_eps, const float* sq_occ, const float* eta, const float* omega,
                          const int* faces, const float* alpha, const float* scale_raw, float ratio, float scale_min,
                          float* vertices, float* xyz, float* scaling, float* rotation, float* opacity, void* stream) {
  SqArgs a;
  if (int e = sq_fill(a, B, Vt, F, K, sq_r, sq_s, sq_t, sq_eps, sq_occ, eta, omega, faces, alpha, scale_raw, ratio,
                      scale_min))
    return e;
  if (!vertices || !xyz || !scaling || !rotation || !opacity) return set_error(PGS_ERR_INVALID_ARG, "null output");
  cudaStream_t s = (cudaStream_t)stream;
  {
    StageTimer t(PGS_STAGE_SQ_FWD, s);
    launch_sq_forward(a, vertices, xyz, scaling, rotation, opacity, s);
  }
  return check_cuda("sq2surfel_forward");
}

size_t pgs_sq2surfel_backward_scratch_bytes(int B, int Vt) {
  return (size_t)(B > 0 ? B : 0) * ((size_t)(Vt > 0 ? Vt : 0) * 3 + 1) * sizeof(float) + 512;
}

int pgs_sq2surfel_backward(int B, int Vt, int F, int K, const float* sq_r, const float* sq_s, const float* sq_t,
                           const float* sq_eps, const float* sq_occ, const float* eta, const float* omega,
                           const int* faces, const float* alpha, const float* scale_raw, float ratio, float scale_min,
                           const float* vertices, const float* d_xyz, const float* d_scaling, const float* d_rotation,
                           const float* d_opacity, const float* d_vertices_in, float* d_sq_r, float* d_sq_s,
                           float* d_sq_t, float* d_sq_eps, float* d_sq_occ, float* d_alpha, float* d_scale_raw,
                           void* scratch, void* stream) {
  SqArgs a;
  if (int e = sq_fill(a, B, Vt, F, K, sq_r, sq_s, sq_t, sq_eps, sq_occ, eta, omega, faces, alpha, scale_raw, ratio,
                      scale_min))
    return e;
  if (!vertices || !d_xyz || !d_scaling || !d_rotation || !d_sq_r || !d_sq_s || !d_sq_t || !d_sq_eps || !d_sq_occ ||
      !scratch)
    return set_error(PGS_ERR_INVALID_ARG, "null gradient pointer");
  cudaStream_t s = (cudaStream_t)stream;
  float* d_vertices = reinterpret_cast<float*>(align_up(reinterpret_cast<size_t>(scratch), 256));
  float* d_occ_acc = d_vertices + (size_t)B * Vt * 3;
  const size_t nv = (size_t)B * Vt * 3 * sizeof(float);
  if (d_vertices_in)
    cudaMemcpyAsync(d_vertices, d_vertices_in, nv, cudaMemcpyDeviceToDevice, s);
  else
    cudaMemsetAsync(d_vertices, 0, nv, s);
  cudaMemsetAsync(d_occ_acc, 0, (size_t)B * sizeof(float), s);
  {
    StageTimer t(PGS_STAGE_SQ_BWD, s);
    launch_sq_backward(a, vertices, d_xyz, d_scaling, d_rotation, d_opacity, d_vertices, d_occ_acc, d_alpha,
                       d_scale_raw, d_sq_r, d_sq_s, d_sq_t, d_sq_eps, d_sq_occ, s);
  }
  return check_cuda("sq2surfel_backward");
}

// ---- block-level (superquadric) rasteriser: surfels are generated inside preprocess -------------------
size_t pgs_dsr_backward_blocks_scratch_bytes(int B, int Vt, int F, int K) {
  const size_t P = (size_t)(B > 0 ? B : 0) * (F > 0 ? F : 0) * (K > 0 ? K : 0);
  // render gradient records + per-surfel (xyz 3, log-scale 2, rotation 4, opacity 1, transMat 9) + sq scratch
  return P * (GRAD_FLOATS + 3 + 2 + 4 + 1 + 9) * sizeof(float) + pgs_sq2surfel_backward_scratch_bytes(B, Vt) + 2048;
}

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
                           void* stream) {
  SqForward f;
  if (int e = sq_fill(f.sq, B, Vt, F, K, sq_r, sq_s, sq_t, sq_eps, sq_occ, eta, omega, faces, alpha, scale_raw, ratio,
                      scale_min))
    return e;
  f.vertices = vertices; f.out_xyz = out_xyz; f.out_scaling = out_scaling; f.out_rotation = out_rotation;
  f.out_opacity = out_opacity;
  return forward_impl(&f, false, 0, nullptr, nullptr, geometry_buffer, geometry_user, binning_buffer, binning_user,
                      image_buffer, image_user, B * F * K, D, M, background, width, height, nullptr, shs, colors_precomp,
                      nullptr, nullptr, scale_modifier, nullptr, nullptr, viewmatrix, projmatrix, cam_pos, tan_fovx,
                      tan_fovy, 0, out_color, out_others, radii, debug, stream);
}

int pgs_dsr_backward_blocks(int B, int Vt, int F, int K, const float* sq_r, const float* sq_s, const float* sq_t,
                            const float* sq_eps, const float* sq_occ, const float* eta, const float* omega,
                            const int* faces, const float* alpha, const float* scale_raw, float ratio, float scale_min,
                            const float* vertices, int D, int M, int R, const float* background, int width, int height,
                            const float* shs, const float* colors_precomp, float scale_modifier,
                            const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                            float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer,
                            size_t binning_bytes, char* image_buffer, const float* dL_dpix, const float* dL_dothers,
                            const float* dL_dvertices, float* dL_dmean2D, void* scratch, float* dL_dcolor,
                            float* dL_dsh, float* d_sq_r,
                            float* d_sq_s, float* d_sq_t, float* d_sq_eps, float* d_sq_occ, float* d_alpha,
                            float* d_scale_raw, int debug, void* stream) {
  SqArgs a;
  if (int e = sq_fill(a, B, Vt, F, K, sq_r, sq_s, sq_t, sq_eps, sq_occ, eta, omega, faces, alpha, scale_raw, ratio,
                      scale_min))
    return e;
  if (!vertices || !scratch || !d_sq_r || !d_sq_s || !d_sq_t || !d_sq_eps || !d_sq_occ)
    return set_error(PGS_ERR_INVALID_ARG, "null block gradient pointer");
  const size_t P = (size_t)B * F * K;
  cudaStream_t s = (cudaStream_t)stream;
  float* base = reinterpret_cast<float*>(align_up(reinterpret_cast<size_t>(scratch), 256));
  float* grad = base;                       base += align_up(P * GRAD_FLOATS, 64);
  float* d_xyz = base;                      base += align_up(P * 3, 64);
  float* d_scaling = base;                  base += align_up(P * 2, 64);
  float* d_rotation = base;                 base += align_up(P * 4, 64);
  float* d_opacity = base;                  base += align_up(P, 64);
  float* d_transMat = base;                 base += align_up(P * 9, 64);
  float* d_vertices = base;                 base += align_up((size_t)B * Vt * 3, 64);
  float* d_occ_acc = base;
  // render backward + preprocess backward (regenerating the surfels), per-surfel gradients into scratch
  if (int e = backward_impl(&a, vertices, false, 0, nullptr, nullptr, nullptr, (int)P, D, M, R, background, width, height,
                            nullptr, shs, colors_precomp, nullptr, scale_modifier, nullptr, nullptr, viewmatrix,
                            projmatrix, campos, tan_fovx, tan_fovy, radii, geom_buffer, binning_buffer, binning_bytes,
                            image_buffer, dL_dpix, dL_dothers, dL_dmean2D, grad, d_opacity, dL_dcolor, d_xyz, d_transMat,
                            dL_dsh, d_scaling, d_rotation, debug, stream))
    return e;
  // per-surfel -> per-face -> per-vertex -> 13 parameters per block
  if (dL_dvertices)   // a loss on the returned mesh vertices: its gradient joins the face -> vertex sums
    cudaMemcpyAsync(d_vertices, dL_dvertices, (size_t)B * Vt * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s);
  else
    cudaMemsetAsync(d_vertices, 0, (size_t)B * Vt * 3 * sizeof(float), s);
  cudaMemsetAsync(d_occ_acc, 0, (size_t)B * sizeof(float), s);
  {
    StageTimer t(PGS_STAGE_SQ_BWD, s);
    launch_sq_backward(a, vertices, d_xyz, d_scaling, d_rotation, d_opacity, d_vertices, d_occ_acc, d_alpha,
                       d_scale_raw, d_sq_r, d_sq_s, d_sq_t, d_sq_eps, d_sq_occ, s);
  }
  return check_cuda("sq2surfel_backward");
}

// ---- renderer post-processing (surface maps) ----------------------------------------------------------
int pgs_surface_maps_forward(int width, int height, const float* allmap, const float* view3x3, const float* rays_m1,
                             const float* rays_m2, const float* rays_o, float depth_ratio, float* rend_normal,
                             float* surf_depth, float* surf_normal, void* stream) {
  if (width <= 0 || height <= 0 || !allmap || !view3x3 || !rays_m1 || !rays_m2 || !rays_o || !rend_normal ||
      !surf_depth || !surf_normal)
    return set_error(PGS_ERR_INVALID_ARG, "surface_maps_forward: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  {
    StageTimer t(PGS_STAGE_SURFACE_FWD, s);
    launch_surface_maps_fwd(width, height, allmap, view3x3, rays_m1, rays_m2, rays_o, depth_ratio, rend_normal,
                            surf_depth, surf_normal, s);
  }
  return check_cuda("surface_maps_forward");
}
size_t pgs_surface_maps_backward_scratch_bytes(int width, int height) {
  return (size_t)6 * (width > 0 ? width : 0) * (height > 0 ? height : 0) * sizeof(float) + 256;
}
int pgs_surface_maps_backward(int width, int height, const float* allmap, const float* view3x3, const float* rays_m1,
                              const float* rays_m2, const float* rays_o, float depth_ratio, const float* g_rend_normal,
                              const float* g_surf_depth, const float* g_surf_normal, void* scratch, float* g_allmap,
                              void* stream) {
  if (width <= 0 || height <= 0 || !allmap || !view3x3 || !rays_m1 || !rays_m2 || !rays_o || !g_allmap ||
      (g_surf_normal && !scratch))
    return set_error(PGS_ERR_INVALID_ARG, "surface_maps_backward: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  float* sc = scratch ? reinterpret_cast<float*>(align_up(reinterpret_cast<size_t>(scratch), 256)) : nullptr;
  {
    StageTimer t(PGS_STAGE_SURFACE_BWD, s);
    launch_surface_maps_bwd(width, height, allmap, view3x3, rays_m1, rays_m2, rays_o, depth_ratio, g_rend_normal,
                            g_surf_depth, g_surf_normal, sc, g_allmap, s);
  }
  return check_cuda("surface_maps_backward");
}

// ---- photometric loss (L1 + SSIM) -----------------------------------------------------------------------
int pgs_photometric_forward(int channels, int height, int width, const float* image, const float* gt, double* sums,
                            float* dmaps, void* stream) {
  if (channels <= 0 || height <= 0 || width <= 0 || !image || !gt || !sums || !dmaps)
    return set_error(PGS_ERR_INVALID_ARG, "photometric_forward: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  {
    StageTimer t(PGS_STAGE_PHOTO_FWD, s);
    launch_photometric_fwd(channels, height, width, image, gt, sums, dmaps, s);
  }
  return check_cuda("photometric_forward");
}
int pgs_photometric_backward(int channels, int height, int width, const float* image, const float* gt,
                             const float* dmaps, const float* g_loss, float lambda_dssim, float* g_image,
                             void* stream) {
  if (channels <= 0 || height <= 0 || width <= 0 || !image || !gt || !dmaps || !g_loss || !g_image)
    return set_error(PGS_ERR_INVALID_ARG, "photometric_backward: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  {
    StageTimer t(PGS_STAGE_PHOTO_BWD, s);
    launch_photometric_bwd(channels, height, width, image, gt, dmaps, g_loss, lambda_dssim, g_image, s);
  }
  return check_cuda("photometric_backward");
}

// ---- per-pixel regularisers (mask entropy, normal consistency, distortion) ---------------------------------
int pgs_regularizers_forward(int width, int height, const float* rend_alpha, const float* gt_mask,
                             const float* rend_dist, const float* rend_normal, const float* surf_normal, double* sums,
                             void* stream) {
  if (width < 0 || height < 0 || (long long)width * height > 0x7fffffffLL || !sums)
    return set_error(PGS_ERR_INVALID_ARG, "regularizers_forward: bad arguments");
  if ((gt_mask && !rend_alpha) || ((rend_normal == nullptr) != (surf_normal == nullptr)))
    return set_error(PGS_ERR_INVALID_ARG, "regularizers_forward: mask needs rend_alpha; normals come in pairs");
  launch_regularizers_fwd(width * height, rend_alpha, gt_mask, rend_dist, rend_normal, surf_normal, sums,
                          (cudaStream_t)stream);
  return check_cuda("regularizers_forward");
}
int pgs_regularizers_backward(int width, int height, const float* rend_alpha, const float* gt_mask,
                              const float* rend_normal, const float* surf_normal, const float* g_loss,
                              float lambda_mask_entropy, float lambda_normal, float lambda_dist, float* g_rend_alpha,
                              float* g_rend_dist, float* g_rend_normal, float* g_surf_normal, void* stream) {
  if (width < 0 || height < 0 || (long long)width * height > 0x7fffffffLL || !g_loss)
    return set_error(PGS_ERR_INVALID_ARG, "regularizers_backward: bad arguments");
  if ((gt_mask && g_rend_alpha && !rend_alpha) || ((g_rend_normal == nullptr) != (g_surf_normal == nullptr)) ||
      (g_rend_normal && (!rend_normal || !surf_normal)))
    return set_error(PGS_ERR_INVALID_ARG, "regularizers_backward: inconsistent pointers");
  launch_regularizers_bwd(width * height, rend_alpha, gt_mask, rend_normal, surf_normal, g_loss, lambda_mask_entropy,
                          lambda_normal, lambda_dist, g_rend_alpha, g_rend_dist, g_rend_normal, g_surf_normal,
                          (cudaStream_t)stream);
  return check_cuda("regularizers_backward");
}

// ---- optimiser step / densification statistics -----------------------------------------------------------
int pgs_adam_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const size_t* numel, const float* step_size, double beta1, double beta2,
                  double eps, double bias_correction2_sqrt, void* stream) {
  if (n_tensors < 0 || n_tensors > PGS_ADAM_MAX_TENSORS)
    return set_error(PGS_ERR_INVALID_ARG, "adam_step: at most %d tensors per call", PGS_ADAM_MAX_TENSORS);
  if (n_tensors > 0 && (!params || !grads || !exp_avg || !exp_avg_sq || !numel || !step_size))
    return set_error(PGS_ERR_INVALID_ARG, "adam_step: null table");
  AdamTable t;
  t.n = n_tensors;
  for (int i = 0; i < n_tensors; i++) {
    if (numel[i] > 0 && (!params[i] || !grads[i] || !exp_avg[i] || !exp_avg_sq[i]))
      return set_error(PGS_ERR_INVALID_ARG, "adam_step: null tensor %d", i);
    t.param[i] = params[i]; t.grad[i] = grads[i]; t.exp_avg[i] = exp_avg[i]; t.exp_avg_sq[i] = exp_avg_sq[i];
    t.numel[i] = numel[i]; t.step_size[i] = step_size[i];
  }
  launch_adam_multi(t, beta1, beta2, eps, bias_correction2_sqrt, (cudaStream_t)stream);
  return check_cuda("adam_step");
}
int pgs_densify_stats(int P, const int* radii, const float* grad_means2D, float* max_radii2D, float* grad_accum,
                      float* denom, void* stream) {
  if (P < 0 || (P > 0 && (!radii || !grad_means2D || !grad_accum || !denom)))
    return set_error(PGS_ERR_INVALID_ARG, "densify_stats: bad arguments");
  launch_densify_stats(P, radii, grad_means2D, max_radii2D, grad_accum, denom, (cudaStream_t)stream);
  return check_cuda("densify_stats");
}

// ---- per-view epilogue of the extraction loop --------------------------------------------------------------
int pgs_extract_maps(int width, int height, int n_parts, const float* semantic, const float* palette,
                     int palette_stride, const float* rend_normal, float* part_rgb, float* normal_unit, void* stream) {
  if (width < 0 || height < 0 || n_parts < 0 || (long long)width * height > 0x7fffffffLL)
    return set_error(PGS_ERR_INVALID_ARG, "extract_maps: bad image size");
  if (part_rgb && n_parts > 0 && (!semantic || !palette || palette_stride < 3))
    return set_error(PGS_ERR_INVALID_ARG, "extract_maps: part map needs semantic, palette and palette_stride >= 3");
  if (normal_unit && !rend_normal) return set_error(PGS_ERR_INVALID_ARG, "extract_maps: rend_normal is null");
  launch_extract_maps(width * height, n_parts, semantic, palette, palette_stride, rend_normal, part_rgb, normal_unit,
                      (cudaStream_t)stream);
  return check_cuda("extract_maps");
}

// ---- densification as one planned compaction -----------------------------------------------------------------
int pgs_densify_blocks(int P) { return densify_blocks(P); }

int pgs_densify_plan(int P, const float* grad_accum, const float* denom, const float* scaling, const float* opacity,
                     double max_grad, double dense_threshold, double min_opacity, int use_world_size_test,
                     double world_size_threshold, double split_divisor, unsigned char* code,
                     unsigned int* block_offsets, unsigned int* counts, void* stream) {
  if (P < 0 || !counts || (P > 0 && (!grad_accum || !denom || !scaling || !opacity || !code || !block_offsets)))
    return set_error(PGS_ERR_INVALID_ARG, "densify_plan: bad arguments");
  if (!(split_divisor > 0.0)) return set_error(PGS_ERR_INVALID_ARG, "densify_plan: split_divisor must be positive");
  // ATen compares float tensors with Python scalars in float, and divides by a scalar as x * (1.f / scalar)
  launch_densify_plan(P, grad_accum, denom, scaling, opacity, (float)max_grad, (float)dense_threshold,
                      (float)min_opacity, use_world_size_test, (float)world_size_threshold,
                      1.0f / (float)split_divisor, code, block_offsets, counts, (cudaStream_t)stream);
  return check_cuda("densify_plan");
}

int pgs_densify_map(int P, const unsigned char* code, const unsigned int* block_offsets, const unsigned int* counts,
                    int n_split, int* src_row, int* sample_row, void* stream) {
  if (P < 0 || n_split < 1 || !counts || (P > 0 && (!code || !block_offsets)))
    return set_error(PGS_ERR_INVALID_ARG, "densify_map: bad arguments");
  launch_densify_map(P, code, block_offsets, counts, n_split, src_row, sample_row, (cudaStream_t)stream);
  return check_cuda("densify_map");
}

int pgs_densify_gather(int n_tensors, const float* const* src, float* const* dst, const int* widths,
                       const int* zero_new, int n_out, int n_keep, const int* src_row, void* stream) {
  if (n_tensors < 0 || n_tensors > PGS_GATHER_MAX_TENSORS)
    return set_error(PGS_ERR_INVALID_ARG, "densify_gather: at most %d tensors per call", PGS_GATHER_MAX_TENSORS);
  if (n_out < 0 || n_keep < 0 || n_keep > n_out || (n_tensors > 0 && (!src || !dst || !widths || !zero_new)))
    return set_error(PGS_ERR_INVALID_ARG, "densify_gather: bad arguments");
  if (n_out == 0 || n_tensors == 0) return 0;
  if (!src_row) return set_error(PGS_ERR_INVALID_ARG, "densify_gather: null row map");
  GatherTable t;
  t.n = n_tensors;
  for (int i = 0; i < n_tensors; i++) {
    if (!src[i] || !dst[i] || widths[i] < 1)
      return set_error(PGS_ERR_INVALID_ARG, "densify_gather: null tensor or bad width at %d", i);
    t.src[i] = src[i]; t.dst[i] = dst[i]; t.width[i] = widths[i]; t.zero_new[i] = zero_new[i];
    t.numel[i] = (size_t)n_out * (size_t)widths[i];
  }
  launch_densify_gather(t, n_keep, src_row, (cudaStream_t)stream);
  return check_cuda("densify_gather");
}

int pgs_densify_children(int n_children, const unsigned int* counts, const int* src_row, const int* sample_row,
                         const float* z, const float* xyz_in, const float* scaling_in, const float* rotation_in,
                         double split_divisor, float* xyz_out, float* scaling_out, void* stream) {
  if (n_children < 0) return set_error(PGS_ERR_INVALID_ARG, "densify_children: bad arguments");
  if (n_children == 0) return 0;
  if (!counts || !src_row || !sample_row || !z || !xyz_in || !scaling_in || !rotation_in || !xyz_out || !scaling_out ||
      !(split_divisor > 0.0))
    return set_error(PGS_ERR_INVALID_ARG, "densify_children: bad arguments");
  launch_densify_children(n_children, counts, src_row, sample_row, z, xyz_in, scaling_in, rotation_in,
                          1.0f / (float)split_divisor, xyz_out, scaling_out, (cudaStream_t)stream);
  return check_cuda("densify_children");
}

int pgs_peer_allreduce_slice(int world, float* const* buckets, size_t offset_floats, size_t n_floats, int max_ctas,
                             void* stream) {
  if (world < 1 || world > PGS_PEER_MAX_WORLD || !buckets)
    return set_error(PGS_ERR_INVALID_ARG, "peer all-reduce: world size must be in [1, %d]", PGS_PEER_MAX_WORLD);
  PeerBuckets b;
  for (int i = 0; i < PGS_PEER_MAX_WORLD; i++) b.p[i] = i < world ? buckets[i] : nullptr;
  for (int i = 0; i < world; i++)
    if (!b.p[i] || (reinterpret_cast<size_t>(b.p[i]) & 15)) return set_error(PGS_ERR_INVALID_ARG, "peer all-reduce: bad bucket pointer");
  const int rc = launch_peer_allreduce_slice(b, world, offset_floats, n_floats, max_ctas, (cudaStream_t)stream);
  if (rc) return set_error(PGS_ERR_INVALID_ARG, "peer all-reduce: slice must be a multiple of 4 floats");
  return check_cuda("peer_allreduce_slice");
}

size_t pgs_knn_temp_bytes(int P) { return knn_temp_bytes(P > 0 ? P : 0) + 256; }
int pgs_knn_dist2(int P, const float* points, float* mean_dist2, void* temp, void* stream) {
  if (P < 0 || (P > 0 && (!points || !mean_dist2 || !temp))) return set_error(PGS_ERR_INVALID_ARG, "bad args");
  if (P == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  void* t = reinterpret_cast<void*>(align_up(reinterpret_cast<size_t>(temp), 256));
  char msg[256] = "";
  int rc;
  {
    StageTimer tm(PGS_STAGE_KNN, s);
    rc = launch_knn_dist2(P, points, mean_dist2, t, s, msg, sizeof(msg));
  }
  if (rc < 0) return set_error(rc == -1 ? PGS_ERR_ALLOC : PGS_ERR_CUDA, "%s", msg);
  return check_cuda("knn");
}

size_t pgs_scan_temp_bytes(int n) { return scan_temp_bytes(n > 0 ? n : 0) + 256; }
int pgs_inclusive_scan_u32(const uint32_t* in, uint32_t* out, int n, void* temp, void* stream) {
  if (n < 0 || (n > 0 && (!in || !out || !temp))) return set_error(PGS_ERR_INVALID_ARG, "bad args");
  void* t = reinterpret_cast<void*>(align_up(reinterpret_cast<size_t>(temp), 256));
  launch_inclusive_scan_u32(in, out, n, t, (cudaStream_t)stream);
  return check_cuda("scan");
}
size_t pgs_sort_temp_bytes(int n, int end_bit) { return radix_sort_temp_bytes(n > 0 ? n : 0, end_bit) + 256; }
int pgs_sort_pairs_u64(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, int n, int end_bit,
                       void* temp, void* stream) {
  if (n < 0 || end_bit <= 0 || end_bit > 64 || (n > 0 && (!keys_a || !vals_a || !keys_b || !vals_b || !temp)))
    return set_error(PGS_ERR_INVALID_ARG, "bad args");
  void* t = reinterpret_cast<void*>(align_up(reinterpret_cast<size_t>(temp), 256));
  int where = launch_radix_sort_pairs(keys_a, vals_a, keys_b, vals_b, n, end_bit, t, (cudaStream_t)stream);
  if (int e = check_cuda("radix_sort")) return e;
  return where;
}
int pgs_dsr_duplicate_with_keys(int P, const char* geom_buffer, int width, int height, const int* radii,
                                uint64_t* keys, uint32_t* values, void* stream) {
  if (P <= 0 || !geom_buffer || !keys || !values) return set_error(PGS_ERR_INVALID_ARG, "bad args");
  const int gx = (width + TILE_X - 1) / TILE_X, gy = (height + TILE_Y - 1) / TILE_Y;
  char* p = const_cast<char*>(geom_buffer);
  GeomState geom = GeomState::from(p, P);
  if (!radii) radii = geom.radii;
  launch_inclusive_scan_u32(geom.tiles_touched, geom.point_offsets, P, geom.scan_temp, (cudaStream_t)stream);
  launch_duplicate_with_keys(P, geom.rec, geom.point_offsets, keys, values, radii, gx, gy, (cudaStream_t)stream);
  return check_cuda("duplicate_with_keys");
}
int pgs_dsr_sorted_keys(int P, int width, int height, const char* geom_buffer, const char* binning_buffer,
                        size_t binning_bytes, int num_rendered, uint64_t* keys, void* stream) {
  if (P <= 0 || width <= 0 || height <= 0 || !geom_buffer || !binning_buffer || num_rendered < 0 ||
      (num_rendered > 0 && !keys))
    return set_error(PGS_ERR_INVALID_ARG, "bad args");
  const int gx = (width + TILE_X - 1) / TILE_X, gy = (height + TILE_Y - 1) / TILE_Y;
  const int end_bit = 32 + (int)higher_msb(gx * gy);
  const size_t cap = BinningState::capacity_from_bytes(binning_bytes, end_bit);
  if (cap == 0 || (size_t)num_rendered > cap)
    return set_error(PGS_ERR_INVALID_ARG, "binning buffer size %zu matches no capacity", binning_bytes);
  char* gp = const_cast<char*>(geom_buffer);
  GeomState geom = GeomState::from(gp, P);
  char* bp = const_cast<char*>(binning_buffer);
  BinningState bin = BinningState::from(bp, cap, end_bit);
  launch_rebuild_sorted_keys(num_rendered, bin.tile_keys(0), bin.vals_a, geom.rec, keys, (cudaStream_t)stream);
  return check_cuda("rebuild_sorted_keys");
}
int pgs_identify_tile_ranges(int L, const uint64_t* sorted_keys, uint32_t* ranges, int ntiles, void* stream) {
  if (L < 0 || ntiles <= 0 || !ranges || (L > 0 && !sorted_keys)) return set_error(PGS_ERR_INVALID_ARG, "bad args");
  cudaMemsetAsync(ranges, 0, (size_t)ntiles * sizeof(uint2), (cudaStream_t)stream);
  launch_identify_tile_ranges(L, sorted_keys, reinterpret_cast<uint2*>(ranges), (cudaStream_t)stream);
  return check_cuda("identify_tile_ranges");
}

}  // extern "C"
