// Surface maps of the 2DGS renderer, forward and backward, as fused per-pixel kernels (sm_100a).
//
// Replaces the ~30 ATen kernels (and their full-image temporaries: meshgrid, an [H*W,3] point map, two
// matmuls, slices / pads) that PartGS runs right after the rasteriser every iteration:
//   renderer/gaussian_renderer/__init__.py:110-149   allmap slicing, normal view->world, nan_to_num,
//                                                    expected depth = depth / alpha, surf_depth blend
//   utils/point_utils.py:4-33                        depths_to_points + depth_to_normal (central-difference
//                                                    cross product, normalised, zero border)
// Inputs per pixel: allmap[7] (rasteriser output: depth, alpha, normal xyz, median depth, distortion).
// Outputs: rend_normal[3] (world space), surf_depth[1], surf_normal[3] (= pseudo normal * alpha).
// rend_alpha / rend_dist are plain channel views of allmap and stay on the Python side.
//
// Per-camera constants (computed by the caller with the reference's own tiny torch expressions and passed
// as device pointers, so no host synchronisation is needed):
//   A[9]   = world_view_transform[:3,:3]                      rend_normal = n_view @ A^T
//   M1[9]  = intrins.inverse().T,  M2[9] = c2w[:3,:3].T       rays_d(x,y) = ([x,y,1] @ M1) @ M2
//   o[3]   = c2w[:3,3]                                          point = depth * rays_d + o
// HBM-bound: 28 B read + 28 B written per pixel forward (neighbour reads hit L1/L2).
#include "common.cuh"
#include "kernels.h"

namespace pgs {

struct CamConst {
  float A[9], M1[9], M2[9], o[3];
};
__device__ __forceinline__ CamConst load_cam(const float* __restrict__ A, const float* __restrict__ M1,
                                             const float* __restrict__ M2, const float* __restrict__ o) {
  CamConst c;
#pragma unroll
  for (int i = 0; i < 9; i++) { c.A[i] = __ldg(A + i); c.M1[i] = __ldg(M1 + i); c.M2[i] = __ldg(M2 + i); }
#pragma unroll
  for (int i = 0; i < 3; i++) c.o[i] = __ldg(o + i);
  return c;
}

// torch.nan_to_num(x, 0, 0) as the reference calls it (positional: nan=0, posinf=0, neginf left at its default,
// the most negative finite float)
__device__ __forceinline__ float finite_or_zero(float x) {
  if (isfinite(x)) return x;
  return (x < 0.f) ? -3.402823466e+38f : 0.f;  // NaN compares false -> 0
}

__device__ __forceinline__ float surf_depth_at(const float* __restrict__ allmap, size_t HW, size_t pid, float ratio) {
  const float D = __ldg(allmap + pid), alpha = __ldg(allmap + HW + pid), med = __ldg(allmap + 5 * HW + pid);
  const float expd = finite_or_zero(D / alpha);
  return expd * (1.f - ratio) + ratio * finite_or_zero(med);
}

__device__ __forceinline__ float3 ray_dir(const CamConst& c, int x, int y) {
  const float px = (float)x, py = (float)y;
  // [x,y,1] @ M1 (row vector times row-major matrix), then @ M2
  const float t0 = px * c.M1[0] + py * c.M1[3] + c.M1[6];
  const float t1 = px * c.M1[1] + py * c.M1[4] + c.M1[7];
  const float t2 = px * c.M1[2] + py * c.M1[5] + c.M1[8];
  return make_float3(t0 * c.M2[0] + t1 * c.M2[3] + t2 * c.M2[6], t0 * c.M2[1] + t1 * c.M2[4] + t2 * c.M2[7],
                     t0 * c.M2[2] + t1 * c.M2[5] + t2 * c.M2[8]);
}
__device__ __forceinline__ float3 point_at(const CamConst& c, const float* __restrict__ allmap, int W, size_t HW, int x,
                                           int y, float ratio) {
  const float d = surf_depth_at(allmap, HW, (size_t)y * W + x, ratio);
  const float3 r = ray_dir(c, x, y);
  return make_float3(d * r.x + c.o[0], d * r.y + c.o[1], d * r.z + c.o[2]);
}
__device__ __forceinline__ float3 cross3f(float3 a, float3 b) {
  return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// dx / dy of depth_to_normal (utils/point_utils.py:30-31): differences along image rows / columns
__device__ __forceinline__ void stencil(const CamConst& c, const float* __restrict__ allmap, int W, size_t HW, int x,
                                        int y, float ratio, float3& dx, float3& dy) {
  const float3 pu = point_at(c, allmap, W, HW, x, y - 1, ratio), pd = point_at(c, allmap, W, HW, x, y + 1, ratio);
  const float3 pl = point_at(c, allmap, W, HW, x - 1, y, ratio), pr = point_at(c, allmap, W, HW, x + 1, y, ratio);
  dx = make_float3(pd.x - pu.x, pd.y - pu.y, pd.z - pu.z);
  dy = make_float3(pr.x - pl.x, pr.y - pl.y, pr.z - pl.z);
}

__global__ void __launch_bounds__(256) surface_maps_fwd_kernel(int W, int H, const float* __restrict__ allmap,
                                                               const float* __restrict__ A, const float* __restrict__ M1,
                                                               const float* __restrict__ M2, const float* __restrict__ o,
                                                               float ratio, float* __restrict__ rend_normal,
                                                               float* __restrict__ surf_depth,
                                                               float* __restrict__ surf_normal) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const CamConst c = load_cam(A, M1, M2, o);
  const size_t HW = (size_t)W * H, pid = (size_t)y * W + x;
  const float n0 = allmap[2 * HW + pid], n1 = allmap[3 * HW + pid], n2 = allmap[4 * HW + pid];
  // (n @ A^T)[c] = sum_k n[k] * A[c][k]
  rend_normal[pid] = n0 * c.A[0] + n1 * c.A[1] + n2 * c.A[2];
  rend_normal[HW + pid] = n0 * c.A[3] + n1 * c.A[4] + n2 * c.A[5];
  rend_normal[2 * HW + pid] = n0 * c.A[6] + n1 * c.A[7] + n2 * c.A[8];
  surf_depth[pid] = surf_depth_at(allmap, HW, pid, ratio);
  float3 n = make_float3(0.f, 0.f, 0.f);
  if (x > 0 && y > 0 && x < W - 1 && y < H - 1) {
    float3 dx, dy;
    stencil(c, allmap, W, HW, x, y, ratio, dx, dy);
    const float3 cr = cross3f(dx, dy);
    const float inv = 1.f / fmaxf(sqrtf(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z), 1e-12f);  // F.normalize eps
    const float alpha = allmap[HW + pid];
    n = make_float3(cr.x * inv * alpha, cr.y * inv * alpha, cr.z * inv * alpha);
  }
  surf_normal[pid] = n.x;
  surf_normal[HW + pid] = n.y;
  surf_normal[2 * HW + pid] = n.z;
}

// backward pass 1: per interior pixel, gradient w.r.t. its two point differences (6 floats into scratch)
__global__ void __launch_bounds__(256) surface_maps_bwd_diff_kernel(int W, int H, const float* __restrict__ allmap,
                                                                    const float* __restrict__ A,
                                                                    const float* __restrict__ M1,
                                                                    const float* __restrict__ M2,
                                                                    const float* __restrict__ o, float ratio,
                                                                    const float* __restrict__ g_surf_normal,
                                                                    float* __restrict__ gdiff /* [6][H][W] */) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const size_t HW = (size_t)W * H, pid = (size_t)y * W + x;
  float3 gdx = make_float3(0.f, 0.f, 0.f), gdy = gdx;
  if (g_surf_normal != nullptr && x > 0 && y > 0 && x < W - 1 && y < H - 1) {
    const CamConst c = load_cam(A, M1, M2, o);
    float3 dx, dy;
    stencil(c, allmap, W, HW, x, y, ratio, dx, dy);
    const float3 cr = cross3f(dx, dy);
    const float len = sqrtf(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z);
    const float alpha = allmap[HW + pid];  // detached in the reference
    const float3 gn = make_float3(g_surf_normal[pid] * alpha, g_surf_normal[HW + pid] * alpha,
                                  g_surf_normal[2 * HW + pid] * alpha);
    float3 gc;
    if (len > 1e-12f) {
      const float inv = 1.f / len;
      const float3 nh = make_float3(cr.x * inv, cr.y * inv, cr.z * inv);
      const float d = nh.x * gn.x + nh.y * gn.y + nh.z * gn.z;
      gc = make_float3((gn.x - nh.x * d) * inv, (gn.y - nh.y * d) * inv, (gn.z - nh.z * d) * inv);
    } else {  // clamped denominator: v / eps
      gc = make_float3(gn.x * 1e12f, gn.y * 1e12f, gn.z * 1e12f);
    }
    gdx = cross3f(dy, gc);  // c = dx x dy
    gdy = cross3f(gc, dx);
  }
  gdiff[pid] = gdx.x; gdiff[HW + pid] = gdx.y; gdiff[2 * HW + pid] = gdx.z;
  gdiff[3 * HW + pid] = gdy.x; gdiff[4 * HW + pid] = gdy.y; gdiff[5 * HW + pid] = gdy.z;
}

// backward pass 2: gather the point gradient from the four pixels whose stencil touches this one, chain into
// surf_depth and on into the rasteriser's channels.  g_allmap is fully written.
__global__ void __launch_bounds__(256) surface_maps_bwd_kernel(int W, int H, const float* __restrict__ allmap,
                                                               const float* __restrict__ A, const float* __restrict__ M1,
                                                               const float* __restrict__ M2, const float* __restrict__ o,
                                                               float ratio, const float* __restrict__ g_rend_normal,
                                                               const float* __restrict__ g_surf_depth,
                                                               const float* __restrict__ gdiff,
                                                               float* __restrict__ g_allmap) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const CamConst c = load_cam(A, M1, M2, o);
  const size_t HW = (size_t)W * H, pid = (size_t)y * W + x;
  float gd = g_surf_depth ? g_surf_depth[pid] : 0.f;
  if (gdiff != nullptr) {
    float3 gp = make_float3(0.f, 0.f, 0.f);
    auto add = [&](int qx, int qy, int ch0, float sgn) {
      if (qx < 0 || qy < 0 || qx >= W || qy >= H) return;
      const size_t q = (size_t)qy * W + qx;
      gp.x += sgn * __ldg(gdiff + (size_t)ch0 * HW + q);
      gp.y += sgn * __ldg(gdiff + (size_t)(ch0 + 1) * HW + q);
      gp.z += sgn * __ldg(gdiff + (size_t)(ch0 + 2) * HW + q);
    };
    add(x, y - 1, 0, 1.f);   // this point is the "+" end of dx of the pixel above
    add(x, y + 1, 0, -1.f);  // and the "-" end of dx of the pixel below
    add(x - 1, y, 3, 1.f);
    add(x + 1, y, 3, -1.f);
    const float3 r = ray_dir(c, x, y);
    gd += gp.x * r.x + gp.y * r.y + gp.z * r.z;
  }
  const float D = allmap[pid], alpha = allmap[HW + pid], med = allmap[5 * HW + pid];
  const float e = D / alpha;
  float gD = 0.f, ga = 0.f;
  if (isfinite(e)) {  // nan_to_num passes the gradient only where its input is finite
    const float ge = gd * (1.f - ratio);
    gD = ge / alpha;
    ga = -ge * D / (alpha * alpha);
    // (alpha == 0 gives a non-finite e: the reference produces 0/0 = NaN gradients there, which no surfel ever
    //  receives because nothing was blended into such a pixel; here they are 0)
  }
  g_allmap[pid] = gD;
  g_allmap[HW + pid] = ga;
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
  if (g_rend_normal) { g0 = g_rend_normal[pid]; g1 = g_rend_normal[HW + pid]; g2 = g_rend_normal[2 * HW + pid]; }
  g_allmap[2 * HW + pid] = g0 * c.A[0] + g1 * c.A[3] + g2 * c.A[6];
  g_allmap[3 * HW + pid] = g0 * c.A[1] + g1 * c.A[4] + g2 * c.A[7];
  g_allmap[4 * HW + pid] = g0 * c.A[2] + g1 * c.A[5] + g2 * c.A[8];
  g_allmap[5 * HW + pid] = isfinite(med) ? gd * ratio : 0.f;
  g_allmap[6 * HW + pid] = 0.f;
}

void launch_surface_maps_fwd(int W, int H, const float* allmap, const float* A, const float* M1, const float* M2,
                             const float* o, float ratio, float* rend_normal, float* surf_depth, float* surf_normal,
                             cudaStream_t s) {
  if (W <= 0 || H <= 0) return;
  dim3 grid((W + 31) / 32, (H + 7) / 8);
  surface_maps_fwd_kernel<<<grid, 256, 0, s>>>(W, H, allmap, A, M1, M2, o, ratio, rend_normal, surf_depth, surf_normal);
  count_launch();
}

void launch_surface_maps_bwd(int W, int H, const float* allmap, const float* A, const float* M1, const float* M2,
                             const float* o, float ratio, const float* g_rend_normal, const float* g_surf_depth,
                             const float* g_surf_normal, float* scratch, float* g_allmap, cudaStream_t s) {
  if (W <= 0 || H <= 0) return;
  dim3 grid((W + 31) / 32, (H + 7) / 8);
  if (g_surf_normal) {
    surface_maps_bwd_diff_kernel<<<grid, 256, 0, s>>>(W, H, allmap, A, M1, M2, o, ratio, g_surf_normal, scratch);
    count_launch();
  }
  surface_maps_bwd_kernel<<<grid, 256, 0, s>>>(W, H, allmap, A, M1, M2, o, ratio, g_rend_normal, g_surf_depth,
                                               g_surf_normal ? scratch : nullptr, g_allmap);
  count_launch();
}

}  // namespace pgs
