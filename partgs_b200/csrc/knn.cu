// distCUDA2: mean squared distance to the 3 nearest neighbours of every point (sm_100a).
//
// Replaces reference SimpleKNN::knn (submodules/simple-knn/simple_knn.cu:185-221): Morton
// sort + 1024-point boxes + box-pruned brute force, i.e. O(P * P/1024) box tests per call,
// 8 raw cudaMalloc/cudaFree and two host syncs.
//
// B200 design: uniform grid ("grid hash" with a dense cell table):
//   1. bounding box by block reduction + ordered-int atomics            (1 pass over points)
//   2. cell histogram, exclusive offsets (our decoupled look-back scan),
//      counting-sort scatter of float4{x,y,z,index}                     (2 passes)
//   3. one thread per point in cell order: visit the 27-cell neighbourhood, then expand
//      ring by ring until the 3rd-best distance is provably inside the searched cube.
// Two grid levels (the fine one has half the cell size): real clouds are concentrated, and a grid sized for one
// point per cell of the BOUNDING BOX puts tens of points into the cells where most of the points live (a 27-cell
// neighbourhood of ~800 candidates for a 3-NN query).  Every query first searches the fine grid (at most
// KNN_FINE_RINGS rings); the minority that is still open there — the sparse fringe, where a fine grid would walk
// ring after empty ring — is appended to a work list and searched in the coarse grid by a second kernel.
// Rows and cells that cannot hold a point nearer than the current third-best distance are skipped.
// The result is the exact 3-NN (same definition as simple_knn.cu:131-183: self excluded by
// index, duplicates count with distance 0, FLT_MAX for missing neighbours when P < 4).
// One host sync (bounding box read-back) instead of the reference's two + mallocs; all
// scratch comes from the caller.
#include <cfloat>
#include <climits>
#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace pgs {

constexpr int KNN_MAX_RES = 256;

__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float ordered_to_float(int i) {
  int j = i >= 0 ? i : i ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
  return __int_as_float(j);
#else
  float f;
  memcpy(&f, &j, 4);
  return f;
#endif
}

// bbox[0..2] = min (ordered ints), bbox[3..5] = max
__global__ void __launch_bounds__(256) knn_bbox_kernel(int P, const float* __restrict__ pts, int* bbox) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      float v = pts[3 * i + a];
      mn[a] = fminf(mn[a], v);
      mx[a] = fmaxf(mx[a], v);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      atomicMin(&bbox[a], float_to_ordered(mn[a]));
      atomicMax(&bbox[3 + a], float_to_ordered(mx[a]));
    }
  }
}

struct KnnGrid {
  float ox, oy, oz;  // origin
  float inv_h, h;
  float eps;         // slack on every cell-boundary distance: covers the rounding of the cell assignment
  int nx, ny, nz;
};

#ifndef KNN_FINE_RINGS
#define KNN_FINE_RINGS 3
#endif

__device__ __forceinline__ int3 cell_of(const KnnGrid& g, float x, float y, float z) {
  int cx = (int)((x - g.ox) * g.inv_h), cy = (int)((y - g.oy) * g.inv_h), cz = (int)((z - g.oz) * g.inv_h);
  cx = min(max(cx, 0), g.nx - 1);
  cy = min(max(cy, 0), g.ny - 1);
  cz = min(max(cz, 0), g.nz - 1);
  return make_int3(cx, cy, cz);
}
// the coarse cell of a point is its fine cell halved: both levels agree on every cell boundary
__device__ __forceinline__ int3 coarse_of(int3 f) { return make_int3(f.x >> 1, f.y >> 1, f.z >> 1); }
// coarse cells are numbered in Morton order: the cells of an octree node are one contiguous range of the table
__device__ __forceinline__ uint32_t spread3(uint32_t v) {
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__device__ __forceinline__ uint32_t compact3(uint32_t v) {
  v &= 0x09249249u;
  v = (v | (v >> 2)) & 0x030c30c3u;
  v = (v | (v >> 4)) & 0x0300f00fu;
  v = (v | (v >> 8)) & 0x030000ffu;
  v = (v | (v >> 16)) & 0x3ffu;
  return v;
}
__device__ __forceinline__ uint32_t morton3(int3 c) {
  return spread3((uint32_t)c.x) | (spread3((uint32_t)c.y) << 1) | (spread3((uint32_t)c.z) << 2);
}
__device__ __forceinline__ uint32_t linear_cell(const KnnGrid& g, int3 c) {
  return (uint32_t)c.x + (uint32_t)g.nx * ((uint32_t)c.y + (uint32_t)g.ny * (uint32_t)c.z);
}

__global__ void __launch_bounds__(256) knn_count_kernel(int P, const float* __restrict__ pts, KnnGrid gf, KnnGrid gc,
                                                        uint32_t* __restrict__ cell_f, uint32_t* __restrict__ cell_c,
                                                        uint32_t* __restrict__ counts_f,
                                                        uint32_t* __restrict__ counts_c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int3 f = cell_of(gf, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
  const uint32_t lf = linear_cell(gf, f), lc = morton3(coarse_of(f));
  cell_f[i] = lf;
  cell_c[i] = lc;
  atomicAdd(&counts_f[lf], 1u);
  atomicAdd(&counts_c[lc], 1u);
}

__global__ void __launch_bounds__(256) knn_scatter_kernel(int P, const float* __restrict__ pts,
                                                          const uint32_t* __restrict__ cell_f,
                                                          const uint32_t* __restrict__ cell_c,
                                                          const uint32_t* __restrict__ incl_f,
                                                          const uint32_t* __restrict__ incl_c,
                                                          uint32_t* __restrict__ fill_f, uint32_t* __restrict__ fill_c,
                                                          float4* __restrict__ sorted_f,
                                                          float4* __restrict__ sorted_c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float4 v = make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], __int_as_float(i));
  const uint32_t lf = cell_f[i], lc = cell_c[i];
  sorted_f[(lf == 0 ? 0u : incl_f[lf - 1]) + atomicAdd(&fill_f[lf], 1u)] = v;
  sorted_c[(lc == 0 ? 0u : incl_c[lc - 1]) + atomicAdd(&fill_c[lc], 1u)] = v;
}

// insert into the ascending triple; most candidates fail the first test
__device__ __forceinline__ void update3(float dist, float (&best)[3]) {
  if (dist < best[2]) {
    const float b1 = fminf(best[1], dist), t2 = fmaxf(best[1], dist);
    best[2] = t2;
    best[1] = fmaxf(best[0], b1);
    best[0] = fminf(best[0], b1);
  }
}

__device__ __forceinline__ void scan_cells(const float4* __restrict__ sorted, const uint32_t* __restrict__ incl,
                                           uint32_t ca, uint32_t cb, const float4& q, float (&best)[3]) {
  const uint32_t s = ca == 0 ? 0u : incl[ca - 1], e = incl[cb];
  for (uint32_t k = s; k < e; k++) {
    const float4 p = sorted[k];
    if (__float_as_int(p.w) == __float_as_int(q.w)) continue;   // the query itself (by index: duplicates count)
    const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
    update3(dx * dx + dy * dy + dz * dz, best);
  }
}

// distance from coordinate v to the slab of cells `cell` along one axis (0 inside), minus the slack
__device__ __forceinline__ float slab_dist(const KnnGrid& g, float o, float v, int cell, int own) {
  if (cell == own) return 0.f;
  const float d = cell < own ? v - (o + (float)(cell + 1) * g.h) : (o + (float)cell * g.h) - v;
  return fmaxf(d - g.eps, 0.f);
}

// Visit the shell of Chebyshev radius r around cell c; rows and cells that cannot hold a point nearer than the
// current third-best distance are skipped (conservatively: the slack covers the rounding of the cell assignment).
__device__ __forceinline__ void visit_shell(const KnnGrid& g, const float4* __restrict__ sorted,
                                            const uint32_t* __restrict__ incl, const int3 c, const int r,
                                            const float4& q, float (&best)[3]) {
  const int z0 = max(c.z - r, 0), z1 = min(c.z + r, g.nz - 1);
  const int y0 = max(c.y - r, 0), y1 = min(c.y + r, g.ny - 1);
  const int x0 = max(c.x - r, 0), x1 = min(c.x + r, g.nx - 1);
  for (int z = z0; z <= z1; z++) {
    const bool zface = (z == c.z - r) || (z == c.z + r);
    const float dz = slab_dist(g, g.oz, q.z, z, c.z);
    const float dz2 = dz * dz;
    if (dz2 >= best[2]) continue;
    for (int y = y0; y <= y1; y++) {
      const bool yface = (y == c.y - r) || (y == c.y + r);
      const float dy = slab_dist(g, g.oy, q.y, y, c.y);
      const float dyz2 = dz2 + dy * dy;
      if (dyz2 >= best[2]) continue;
      const uint32_t row = (uint32_t)g.nx * ((uint32_t)y + (uint32_t)g.ny * (uint32_t)z);
      if (zface || yface) {
        // the whole x-run belongs to the shell: cells are contiguous in memory -> one range, clipped to the
        // cells within reach of the third-best distance
        int xa = x0, xb = x1;
        if (best[2] < FLT_MAX) {
          const float rad = sqrtf(best[2] - dyz2) + g.eps;
          xa = max(xa, (int)floorf((q.x - rad - g.ox) * g.inv_h - 1e-3f));
          xb = min(xb, (int)floorf((q.x + rad - g.ox) * g.inv_h + 1e-3f));
          if (xa > xb) continue;
        }
        scan_cells(sorted, incl, row + xa, row + xb, q, best);
      } else {
        // only the two x end caps
#pragma unroll
        for (int side = 0; side < 2; side++) {
          const int x = side == 0 ? c.x - r : c.x + r;
          if (x < 0 || x >= g.nx || (side == 1 && r == 0)) continue;
          const float dx = slab_dist(g, g.ox, q.x, x, c.x);
          if (dyz2 + dx * dx >= best[2]) continue;
          scan_cells(sorted, incl, row + x, row + x, q, best);
        }
      }
    }
  }
}

// true when every point outside the cube of radius r around cell c is provably farther than the third best
__device__ __forceinline__ bool shell_closes(const KnnGrid& g, const int3 c, const int r, const float4& q,
                                             const float (&best)[3]) {
  const float lox = g.ox + (float)(c.x - r) * g.h, hix = g.ox + (float)(c.x + r + 1) * g.h;
  const float loy = g.oy + (float)(c.y - r) * g.h, hiy = g.oy + (float)(c.y + r + 1) * g.h;
  const float loz = g.oz + (float)(c.z - r) * g.h, hiz = g.oz + (float)(c.z + r + 1) * g.h;
  float reach = fminf(fminf(fminf(q.x - lox, hix - q.x), fminf(q.y - loy, hiy - q.y)), fminf(q.z - loz, hiz - q.z));
  reach = fmaxf(reach - g.eps, 0.f) * 0.9999f;  // guard against rounding of the cell assignment
  const bool covers_all = (c.x - r <= 0) && (c.y - r <= 0) && (c.z - r <= 0) && (c.x + r >= g.nx - 1) &&
                          (c.y + r >= g.ny - 1) && (c.z + r >= g.nz - 1);
  return covers_all || best[2] <= reach * reach;
}

// fine level: every point, in fine-cell order; open queries go to the work list (their point index)
__global__ void __launch_bounds__(128) knn_search_fine_kernel(int P, KnnGrid g, const float4* __restrict__ sorted,
                                                              const uint32_t* __restrict__ incl,
                                                              float* __restrict__ out, uint32_t* __restrict__ work,
                                                              uint32_t* __restrict__ n_work) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  bool open = false;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (slot < P) {
    q = sorted[slot];
    const int3 c = cell_of(g, q.x, q.y, q.z);
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    open = true;
    for (int r = 0; r <= KNN_FINE_RINGS; r++) {
      visit_shell(g, sorted, incl, c, r, q, best);
      if (shell_closes(g, c, r, q, best)) {
        open = false;
        break;
      }
    }
    if (!open) out[__float_as_int(q.w)] = (best[0] + best[1] + best[2]) / 3.0f;
  }
  // warp-aggregated append (all 32 lanes take part)
  const unsigned m = __ballot_sync(0xffffffffu, open);
  if (m == 0) return;
  const unsigned lane = threadIdx.x & 31u;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(n_work, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (open) work[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)__float_as_int(q.w);
}

// Octree level: the queries the fine level left open (the sparse fringe, outliers, clouds much sparser than the
// fine grid).  The coarse cells are the leaves of an implicit octree (Morton numbering: node m of level k is the
// cell range [m << 3k, (m + 1) << 3k) of the table).  Bottom-up exact search with a stack: the query's own leaf
// first, then the seven siblings of its ancestor at every level, nearest levels first; a node farther than the
// third-best distance is dropped before its range is even read, a node with few points is scanned, any other node
// is replaced by its eight children (nearest octant on top).
constexpr int KNN_TREE_STACK = 128;
constexpr uint32_t KNN_TREE_SCAN = 16;   // scan a node with at most this many points instead of descending

__device__ __forceinline__ void scan_points(const float4* __restrict__ sorted, uint32_t s, uint32_t e,
                                            const float4& q, float (&best)[3]) {
  for (uint32_t j = s; j < e; j++) {
    const float4 p = sorted[j];
    if (__float_as_int(p.w) == __float_as_int(q.w)) continue;
    const float ex = p.x - q.x, ey = p.y - q.y, ez = p.z - q.z;
    update3(ex * ex + ey * ey + ez * ez, best);
  }
}

__global__ void __launch_bounds__(128) knn_search_tree_kernel(const uint32_t* __restrict__ n_work,
                                                              const uint32_t* __restrict__ work,
                                                              const float* __restrict__ pts, KnnGrid gf, KnnGrid gc,
                                                              int levels, const float4* __restrict__ sorted,
                                                              const uint32_t* __restrict__ incl,
                                                              float* __restrict__ out) {
  // the work list's length is only known on the device: a fixed grid strides over it
  uint32_t stack[KNN_TREE_STACK];
  const uint32_t n = *n_work;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const uint32_t i = work[t];
    const float4 q = make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], __int_as_float((int)i));
    const uint32_t leaf = morton3(coarse_of(cell_of(gf, q.x, q.y, q.z)));
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    int sp = 0;
    for (int k = levels - 1; k >= 0; k--) {      // siblings of the ancestor at level k; lower levels end up on top
      const uint32_t a = leaf >> (3 * k);
      for (int j = 7; j >= 1; j--) stack[sp++] = ((uint32_t)k << 24) | (a ^ (uint32_t)j);
    }
    stack[sp++] = leaf;                           // level 0
    while (sp > 0) {
      const uint32_t e = stack[--sp];
      const int k = (int)(e >> 24);
      const uint32_t m = e & 0xffffffu;
      const float hk = gc.h * (float)(1 << k);
      const float bx = gc.ox + (float)compact3(m) * hk, by = gc.oy + (float)compact3(m >> 1) * hk,
                  bz = gc.oz + (float)compact3(m >> 2) * hk;
      const float dx = fmaxf(fmaxf(fmaxf(bx - q.x, q.x - (bx + hk)), 0.f) - gc.eps, 0.f);
      const float dy = fmaxf(fmaxf(fmaxf(by - q.y, q.y - (by + hk)), 0.f) - gc.eps, 0.f);
      const float dz = fmaxf(fmaxf(fmaxf(bz - q.z, q.z - (bz + hk)), 0.f) - gc.eps, 0.f);
      if (dx * dx + dy * dy + dz * dz >= best[2]) continue;
      const uint32_t lo = m << (3 * k), hi = ((m + 1u) << (3 * k)) - 1u;
      const uint32_t s = lo == 0 ? 0u : incl[lo - 1], en = incl[hi];
      if (s == en) continue;
      // a leaf, a node with few points, or (cannot happen: <= 7 levels x 7 siblings + 7 per level descended) no room
      if (k == 0 || en - s <= KNN_TREE_SCAN || sp + 8 > KNN_TREE_STACK) {
        scan_points(sorted, s, en, q, best);
        continue;
      }
      // eight children, the octant nearest to the query on top
      const float half = 0.5f * hk;
      const uint32_t qo = (q.x >= bx + half ? 1u : 0u) | (q.y >= by + half ? 2u : 0u) | (q.z >= bz + half ? 4u : 0u);
      for (int j = 7; j >= 0; j--) stack[sp++] = ((uint32_t)(k - 1) << 24) | ((m << 3) | ((uint32_t)j ^ qo));
    }
    out[i] = (best[0] + best[1] + best[2]) / 3.0f;
  }
}

static int knn_coarse_res(int P) {
  int res = 8;  // ~1 point per cell of the bounding box
  while (res < KNN_MAX_RES / 2 && (double)res * res * res < (double)P) res *= 2;
  return res;
}

size_t knn_temp_bytes(int P) {
  const int rc = knn_coarse_res(P), rf = 2 * rc;
  const size_t nc = (size_t)rc * rc * rc, nf = (size_t)rf * rf * rf;
  return 256 + 3 * align_up((size_t)P * 4, 256) + 2 * align_up(nc * 4, 256) + 2 * align_up(nf * 4, 256) +
         2 * align_up((size_t)P * 16, 256) + scan_temp_bytes((int)nf) + 2048;
}

int launch_knn_dist2(int P, const float* points, float* out, void* temp, cudaStream_t s, char* err, size_t errlen) {
  char* p = (char*)temp;
  int* bbox;
  carve(p, bbox, 8);
  int init[8] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN, 0, 0};   // [6] = work-list length
  cudaMemcpyAsync(bbox, init, sizeof(init), cudaMemcpyHostToDevice, s);
  int blocks = min((P + 255) / 256, 148 * 8);
  knn_bbox_kernel<<<blocks, 256, 0, s>>>(P, points, bbox);
  count_launch();
  int hb[6];
  cudaMemcpyAsync(hb, bbox, sizeof(hb), cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) {
    snprintf(err, errlen, "knn bbox: %s", cudaGetErrorString(e));
    return -2;
  }
  float mn[3], mx[3];
  float amax = 0.f;
  for (int a = 0; a < 3; a++) {
    mn[a] = ordered_to_float(hb[a]);
    mx[a] = ordered_to_float(hb[3 + a]);
    amax = fmaxf(amax, fmaxf(fabsf(mn[a]), fabsf(mx[a])));
  }
  float ext = fmaxf(fmaxf(mx[0] - mn[0], mx[1] - mn[1]), mx[2] - mn[2]);
  if (!(ext > 0.f) || !isfinite(ext)) ext = 1.f;
  if (!isfinite(amax)) amax = 0.f;
  const int res_c = knn_coarse_res(P), res_f = 2 * res_c;
  KnnGrid gf, gc;
  gf.h = ext / (float)res_f * 1.0001f;
  gf.inv_h = 1.0f / gf.h;
  gf.ox = mn[0]; gf.oy = mn[1]; gf.oz = mn[2];
  // slack: 1e-3 of a cell + a few ulps of the largest coordinate (a cloud far from the origin)
  gf.eps = 1e-3f * gf.h + 8.f * FLT_EPSILON * amax;
  gf.nx = max(1, min(res_f, (int)ceilf((mx[0] - mn[0]) * gf.inv_h + 1e-3f)));
  gf.ny = max(1, min(res_f, (int)ceilf((mx[1] - mn[1]) * gf.inv_h + 1e-3f)));
  gf.nz = max(1, min(res_f, (int)ceilf((mx[2] - mn[2]) * gf.inv_h + 1e-3f)));
  gc = gf;
  gc.h = 2.f * gf.h;          // exact: coarse boundaries are fine boundaries
  gc.inv_h = 0.5f * gf.inv_h;
  gc.nx = (gf.nx + 1) / 2; gc.ny = (gf.ny + 1) / 2; gc.nz = (gf.nz + 1) / 2;
  const size_t nf = (size_t)gf.nx * gf.ny * gf.nz;
  int levels = 0;                // the Morton cube that holds the coarse grid: 2^levels cells per axis
  while ((1 << levels) < max(max(gc.nx, gc.ny), gc.nz)) levels++;
  const size_t nc = (size_t)1 << (3 * levels);

  uint32_t *cell_f, *cell_c, *work, *counts_f, *incl_f, *counts_c, *incl_c;
  float4 *sorted_f, *sorted_c;
  char* scan_tmp;
  carve(p, cell_f, (size_t)P);
  carve(p, cell_c, (size_t)P);
  carve(p, work, (size_t)P);
  carve(p, counts_f, nf);
  carve(p, incl_f, nf);
  carve(p, counts_c, nc);
  carve(p, incl_c, nc);
  carve(p, sorted_f, (size_t)P);
  carve(p, sorted_c, (size_t)P);
  carve(p, scan_tmp, scan_temp_bytes((int)nf));
  if ((size_t)(p - (char*)temp) > knn_temp_bytes(P)) {
    snprintf(err, errlen, "knn scratch too small");
    return -1;
  }
  uint32_t* n_work = reinterpret_cast<uint32_t*>(bbox + 6);
  cudaMemsetAsync(counts_f, 0, nf * 4, s);
  cudaMemsetAsync(counts_c, 0, nc * 4, s);
  knn_count_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, points, gf, gc, cell_f, cell_c, counts_f, counts_c);
  count_launch();
  launch_inclusive_scan_u32(counts_f, incl_f, (int)nf, scan_tmp, s);
  launch_inclusive_scan_u32(counts_c, incl_c, (int)nc, scan_tmp, s);
  cudaMemsetAsync(counts_f, 0, nf * 4, s);
  cudaMemsetAsync(counts_c, 0, nc * 4, s);
  knn_scatter_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, points, cell_f, cell_c, incl_f, incl_c, counts_f, counts_c,
                                                     sorted_f, sorted_c);
  count_launch();
  knn_search_fine_kernel<<<(P + 127) / 128, 128, 0, s>>>(P, gf, sorted_f, incl_f, out, work, n_work);
  count_launch();
  knn_search_tree_kernel<<<min((P + 127) / 128, 148 * 16), 128, 0, s>>>(n_work, work, points, gf, gc, levels, sorted_c,
                                                                         incl_c, out);
  count_launch();
  return 0;
}

}  // namespace pgs
