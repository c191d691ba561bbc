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
  int nx, ny, nz;
};

__device__ __forceinline__ int3 cell_of(const KnnGrid& g, float x, float y, float z) {
  int cx = (int)((x - g.ox) * g.inv_h), cy = (int)((y - g.oy) * g.inv_h), cz = (int)((z - g.oz) * g.inv_h);
  cx = min(max(cx, 0), g.nx - 1);
  cy = min(max(cy, 0), g.ny - 1);
  cz = min(max(cz, 0), g.nz - 1);
  return make_int3(cx, cy, cz);
}

__global__ void __launch_bounds__(256) knn_count_kernel(int P, const float* __restrict__ pts, KnnGrid g,
                                                        uint32_t* __restrict__ cell_of_point,
                                                        uint32_t* __restrict__ counts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  int3 c = cell_of(g, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
  uint32_t cell = (uint32_t)c.x + (uint32_t)g.nx * ((uint32_t)c.y + (uint32_t)g.ny * (uint32_t)c.z);
  cell_of_point[i] = cell;
  atomicAdd(&counts[cell], 1u);
}

__global__ void __launch_bounds__(256) knn_scatter_kernel(int P, const float* __restrict__ pts,
                                                          const uint32_t* __restrict__ cell_of_point,
                                                          const uint32_t* __restrict__ incl,
                                                          uint32_t* __restrict__ fill, float4* __restrict__ sorted) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  uint32_t cell = cell_of_point[i];
  uint32_t start = cell == 0 ? 0u : incl[cell - 1];
  uint32_t slot = start + atomicAdd(&fill[cell], 1u);
  sorted[slot] = make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], __int_as_float(i));
}

__device__ __forceinline__ void update3(float dist, float (&best)[3]) {
#pragma unroll
  for (int j = 0; j < 3; j++) {
    if (best[j] > dist) {
      float t = best[j];
      best[j] = dist;
      dist = t;
    }
  }
}

__global__ void __launch_bounds__(128) knn_search_kernel(int P, KnnGrid g, const float4* __restrict__ sorted,
                                                         const uint32_t* __restrict__ incl,
                                                         float* __restrict__ out) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= P) return;
  const float4 q = sorted[slot];
  const int3 c = cell_of(g, q.x, q.y, q.z);
  float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
  const int rmax = max(max(g.nx, g.ny), g.nz);
  for (int r = 0; r <= rmax; r++) {
    // visit the shell of Chebyshev radius r around the query's cell
    const int z0 = max(c.z - r, 0), z1 = min(c.z + r, g.nz - 1);
    const int y0 = max(c.y - r, 0), y1 = min(c.y + r, g.ny - 1);
    const int x0 = max(c.x - r, 0), x1 = min(c.x + r, g.nx - 1);
    for (int z = z0; z <= z1; z++) {
      const bool zface = (z == c.z - r) || (z == c.z + r);
      for (int y = y0; y <= y1; y++) {
        const bool yface = (y == c.y - r) || (y == c.y + r);
        const uint32_t row = (uint32_t)g.nx * ((uint32_t)y + (uint32_t)g.ny * (uint32_t)z);
        if (zface || yface) {
          // whole x-run belongs to the shell: cells are contiguous in memory -> one range
          const uint32_t ca = row + x0, cb = row + x1;
          const uint32_t s = ca == 0 ? 0u : incl[ca - 1], e = incl[cb];
          for (uint32_t k = s; k < e; k++) {
            if (k == (uint32_t)slot) continue;
            const float4 p = sorted[k];
            const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
            update3(dx * dx + dy * dy + dz * dz, best);
          }
        } else {
          // only the two x end caps
#pragma unroll
          for (int side = 0; side < 2; side++) {
            const int x = side == 0 ? c.x - r : c.x + r;
            if (x < 0 || x >= g.nx || (side == 1 && r == 0)) continue;
            const uint32_t cc = row + x;
            const uint32_t s = cc == 0 ? 0u : incl[cc - 1], e = incl[cc];
            for (uint32_t k = s; k < e; k++) {
              if (k == (uint32_t)slot) continue;
              const float4 p = sorted[k];
              const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
              update3(dx * dx + dy * dy + dz * dz, best);
            }
          }
        }
      }
    }
    // everything outside the cube of radius r is at least `reach` away
    const float lox = g.ox + (float)(c.x - r) * g.h, hix = g.ox + (float)(c.x + r + 1) * g.h;
    const float loy = g.oy + (float)(c.y - r) * g.h, hiy = g.oy + (float)(c.y + r + 1) * g.h;
    const float loz = g.oz + (float)(c.z - r) * g.h, hiz = g.oz + (float)(c.z + r + 1) * g.h;
    float reach = fminf(fminf(fminf(q.x - lox, hix - q.x), fminf(q.y - loy, hiy - q.y)),
                        fminf(q.z - loz, hiz - q.z));
    reach = fmaxf(reach, 0.f) * 0.9999f;  // guard against rounding of the cell assignment
    const bool covers_all = (c.x - r <= 0) && (c.y - r <= 0) && (c.z - r <= 0) && (c.x + r >= g.nx - 1) &&
                            (c.y + r >= g.ny - 1) && (c.z + r >= g.nz - 1);
    if (covers_all || best[2] <= reach * reach) break;
  }
  out[__float_as_int(q.w)] = (best[0] + best[1] + best[2]) / 3.0f;
}

size_t knn_temp_bytes(int P) {
  int res = 8;  // same rule as launch_knn_dist2
  while (res < KNN_MAX_RES && (double)res * res * res < (double)P) res *= 2;
  size_t ncell = (size_t)res * res * res;
  return 256 + align_up((size_t)P * 4, 256) + 2 * align_up(ncell * 4, 256) + align_up((size_t)P * 16, 256) +
         scan_temp_bytes((int)ncell) + 1024;
}

int launch_knn_dist2(int P, const float* points, float* out, void* temp, cudaStream_t s, char* err, size_t errlen) {
  char* p = (char*)temp;
  int* bbox;
  carve(p, bbox, 8);
  int init[8] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN, 0, 0};
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
  for (int a = 0; a < 3; a++) {
    mn[a] = ordered_to_float(hb[a]);
    mx[a] = ordered_to_float(hb[3 + a]);
  }
  float ext = fmaxf(fmaxf(mx[0] - mn[0], mx[1] - mn[1]), mx[2] - mn[2]);
  if (!(ext > 0.f) || !isfinite(ext)) ext = 1.f;
  // resolution: ~1 point per cell for a volumetric cloud, a few per occupied cell for a surface
  int res = 8;
  while (res < KNN_MAX_RES && (double)res * res * res < (double)P) res *= 2;
  KnnGrid g;
  g.h = ext / (float)res * 1.0001f;
  g.inv_h = 1.0f / g.h;
  g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2];
  g.nx = max(1, min(res, (int)ceilf((mx[0] - mn[0]) * g.inv_h + 1e-3f)));
  g.ny = max(1, min(res, (int)ceilf((mx[1] - mn[1]) * g.inv_h + 1e-3f)));
  g.nz = max(1, min(res, (int)ceilf((mx[2] - mn[2]) * g.inv_h + 1e-3f)));
  const size_t ncell = (size_t)g.nx * g.ny * g.nz;

  uint32_t *cell_of_point, *counts, *incl;
  float4* sorted;
  char* scan_tmp;
  carve(p, cell_of_point, (size_t)P);
  carve(p, counts, ncell);
  carve(p, incl, ncell);
  carve(p, sorted, (size_t)P);
  carve(p, scan_tmp, scan_temp_bytes((int)ncell));
  if ((size_t)(p - (char*)temp) > knn_temp_bytes(P)) {
    snprintf(err, errlen, "knn scratch too small");
    return -1;
  }
  cudaMemsetAsync(counts, 0, ncell * 4, s);
  knn_count_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, points, g, cell_of_point, counts);
  count_launch();
  launch_inclusive_scan_u32(counts, incl, (int)ncell, scan_tmp, s);
  cudaMemsetAsync(counts, 0, ncell * 4, s);
  knn_scatter_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, points, cell_of_point, incl, counts, sorted);
  count_launch();
  knn_search_kernel<<<(P + 127) / 128, 128, 0, s>>>(P, g, sorted, incl, out);
  count_launch();
  return 0;
}

}  // namespace pgs
