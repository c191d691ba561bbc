// Forward preprocess for the base surfel rasteriser (sm_100a).
//
// One thread per surfel: near-cull, ray-splat transform T, facing normal, 3-sigma
// AABB -> radius / tile rectangle, SH -> RGB, packed 80-byte render record and a
// conservative cull box for the render kernels.
//
// Replaces reference preprocessCUDA (cuda_rasterizer/forward.cu:148-251) and its
// helpers compute_transmat (:75-115), compute_aabb (:119-145), computeColorFromSH
// (:20-71), in_frustum / quat_to_rotmat / scale_to_mat / getRect
// (cuda_rasterizer/auxiliary.h:67-77,185-235,285-292).
//
// Bit-exactness: radii, tile rectangles and sort keys must equal the reference's,
// so every fp32 expression below keeps the reference's association order (GLM sums
// are left-to-right; see linalg.cuh) and is compiled with the same nvcc defaults
// (-fmad=true, no fast-math).
#include "common.cuh"
#include "linalg.cuh"
#include "kernels.h"
#include "sq_device.cuh"
#include <cstdlib>

namespace pgs {

__device__ const float kSH_C0 = 0.28209479177387814f;
__device__ const float kSH_C1 = 0.4886025119029199f;
__device__ const float kSH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                   -1.0925484305920792f, 0.5462742152960396f};
__device__ const float kSH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                                   -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

__device__ __forceinline__ float3 xform_point4x3(const float3& p, const float* m) {
  float3 t = {
      m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
      m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
      m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
  };
  return t;
}
__device__ __forceinline__ float3 xform_vec4x3(const float3& p, const float* m) {
  float3 t = {
      m[0] * p.x + m[4] * p.y + m[8] * p.z,
      m[1] * p.x + m[5] * p.y + m[9] * p.z,
      m[2] * p.x + m[6] * p.y + m[10] * p.z,
  };
  return t;
}

// (w,x,y,z) quaternion stored in fields (x,y,z,w); normalised here with rsqrtf.
__device__ __forceinline__ m3 quat_to_rotmat(const v4 quat) {
  float s = rsqrtf(quat.w * quat.w + quat.x * quat.x + quat.y * quat.y + quat.z * quat.z);
  float w = quat.x * s;
  float x = quat.y * s;
  float y = quat.z * s;
  float z = quat.w * s;
  m3 R;
  R[0] = v3(1.f - 2.f * (y * y + z * z), 2.f * (x * y + w * z), 2.f * (x * z - w * y));
  R[1] = v3(2.f * (x * y - w * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z + w * x));
  R[2] = v3(2.f * (x * z + w * y), 2.f * (y * z - w * x), 1.f - 2.f * (x * x + y * y));
  return R;
}

__device__ __forceinline__ m3 scale_to_mat(const v2 scale, const float glob_scale) {
  m3 S = diag3(1.f);
  S[0][0] = glob_scale * scale.x;
  S[1][1] = glob_scale * scale.y;
  return S;
}

// T = transpose(splat2world) * world2ndc * ndc2pix, normal = view rotation of the
// third rotation column.
__device__ __forceinline__ void compute_transmat(const float3& p_orig, const v2 scale, float mod, const v4 rot,
                                                 const float* projmatrix, const float* viewmatrix, const int W,
                                                 const int H, m3& T, float3& normal) {
  m3 R = quat_to_rotmat(rot);
  m3 S = scale_to_mat(scale, mod);
  m3 L = R * S;

  m3x4 splat2world = make_m3x4(v4(L[0], 0.0f), v4(L[1], 0.0f), v4(p_orig.x, p_orig.y, p_orig.z, 1.f));

  m4 world2ndc;
  world2ndc[0] = v4(projmatrix[0], projmatrix[4], projmatrix[8], projmatrix[12]);
  world2ndc[1] = v4(projmatrix[1], projmatrix[5], projmatrix[9], projmatrix[13]);
  world2ndc[2] = v4(projmatrix[2], projmatrix[6], projmatrix[10], projmatrix[14]);
  world2ndc[3] = v4(projmatrix[3], projmatrix[7], projmatrix[11], projmatrix[15]);

  m3x4 ndc2pix = make_m3x4(v4((float)(float(W) / 2.0), 0.0f, 0.0f, (float)(float(W - 1) / 2.0)),
                           v4(0.0f, (float)(float(H) / 2.0), 0.0f, (float)(float(H - 1) / 2.0)),
                           v4(0.0f, 0.0f, 0.0f, 1.0f));

  T = transpose(splat2world) * world2ndc * ndc2pix;
  normal = xform_vec4x3({L[2].x, L[2].y, L[2].z}, viewmatrix);
}

// Bounding box of the cutoff-sigma level set in pixel space.
__device__ __forceinline__ bool compute_aabb(m3 T, float cutoff, float2& point_image, float2& extent) {
  v3 t = v3(cutoff * cutoff, cutoff * cutoff, -1.0f);
  float d = dot(t, T[2] * T[2]);
  if (d == 0.0) return false;
  v3 f = (1 / d) * t;

  v2 p = v2(dot(f, T[0] * T[2]), dot(f, T[1] * T[2]));
  v2 h0 = p * p - v2(dot(f, T[0] * T[0]), dot(f, T[1] * T[1]));
  v2 h = vsqrt(vmax(v2(1e-4, 1e-4), h0));
  point_image = {p.x, p.y};
  extent = {h.x, h.y};
  return true;
}

__device__ __forceinline__ void tile_rect(const float2 p, int max_radius, uint2& rect_min, uint2& rect_max,
                                          dim3 grid) {
  rect_min = {min(grid.x, max((int)0, (int)((p.x - max_radius) / TILE_X))),
              min(grid.y, max((int)0, (int)((p.y - max_radius) / TILE_Y)))};
  rect_max = {min(grid.x, max((int)0, (int)((p.x + max_radius + TILE_X - 1) / TILE_X))),
              min(grid.y, max((int)0, (int)((p.y + max_radius + TILE_Y - 1) / TILE_Y)))};
}

// NOTE: keep this function exactly in this shape — its FMA contraction (decided by nvcc from the
// expression tree and its surroundings) is what makes `rgb` bit-identical to the reference build;
// a variant that first copied the coefficients to registers with 128-bit loads changed it.
template <bool DIRECT = false, bool STAGED = false>
__device__ __forceinline__ v3 color_from_sh(int idx, int deg, int max_coeffs, const v3* means, v3 campos,
                                            const float* shs, unsigned& clamped_mask, v3 pos_direct = v3()) {
  v3 pos = DIRECT ? pos_direct : means[idx];
  v3 dir = pos - campos;
  dir = dir / length(dir);

  const v3* sh = STAGED ? (const v3*)shs : ((const v3*)shs) + (size_t)idx * max_coeffs;  // STAGED: shs = this surfel's row
  v3 result = kSH_C0 * sh[0];

  if (deg > 0) {
    float x = dir.x;
    float y = dir.y;
    float z = dir.z;
    result = result - kSH_C1 * y * sh[1] + kSH_C1 * z * sh[2] - kSH_C1 * x * sh[3];

    if (deg > 1) {
      float xx = x * x, yy = y * y, zz = z * z;
      float xy = x * y, yz = y * z, xz = x * z;
      result = result + kSH_C2[0] * xy * sh[4] + kSH_C2[1] * yz * sh[5] +
               kSH_C2[2] * (2.0f * zz - xx - yy) * sh[6] + kSH_C2[3] * xz * sh[7] + kSH_C2[4] * (xx - yy) * sh[8];

      if (deg > 2) {
        result = result + kSH_C3[0] * y * (3.0f * xx - yy) * sh[9] + kSH_C3[1] * xy * z * sh[10] +
                 kSH_C3[2] * y * (4.0f * zz - xx - yy) * sh[11] +
                 kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[12] +
                 kSH_C3[4] * x * (4.0f * zz - xx - yy) * sh[13] + kSH_C3[5] * z * (xx - yy) * sh[14] +
                 kSH_C3[6] * x * (xx - 3.0f * yy) * sh[15];
      }
    }
  }
  result += 0.5f;
  clamped_mask = (result.x < 0 ? 1u : 0u) | (result.y < 0 ? 2u : 0u) | (result.z < 0 ? 4u : 0u);
  return vmax(result, 0.0f);
}

// Conservative screen-space box outside of which the reference is guaranteed to
// skip the fragment with `alpha < 1/255`: alpha = opa*exp(-rho/2) and
// rho = min(rho3d, rho2d), so a pixel can only pass if rho3d <= c2 or rho2d <= c2
// with c2 = 2 ln(255 opa).  {rho3d <= c2} projects inside the c-sigma AABB of the
// splat when the whole c-sigma disk lies in front of the camera plane;
// {rho2d <= c2} is a disc of radius sqrt(c2/2) around the low-pass centre.
// Evaluated in fp64 with margins; falls back to "everything" when the projected
// conic is not an ellipse.  This is new relative to the reference (which visits
// every pixel of every binned tile); it does not change any result.
__device__ __forceinline__ void cull_record(const m3& T, float2 xy, float opa, float4* out) {
  const float BIG = 3.0e38f;
  const float4 none = make_float4(BIG, BIG, -BIG, -BIG);
  const float4 all = make_float4(-BIG, -BIG, BIG, BIG);
  // conic "always passes" (g = -BIG) unless computed below
  out[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  out[2] = make_float4(0.f, -BIG, xy.x, xy.y);
  if (!(opa > 0.f)) { out[0] = none; return; }  // alpha <= 0 < 1/255 always
  double c2 = 2.0 * log(255.0 * (double)opa);
  c2 = c2 * 1.02 + 0.05;
  if (c2 <= 0.0) { out[0] = none; return; }
  const double ux = T[0].x, uy = T[0].y, uz = T[0].z;
  const double vx = T[1].x, vy = T[1].y, vz = T[1].z;
  const double wx = T[2].x, wy = T[2].y, wz = T[2].z;

  // ---- conic: p(x,y) = x*(Tv x Tw) + y*(Tw x Tu) + (Tu x Tv);  rho3d <= c2  <=>  p.x^2 + p.y^2 - c2 p.z^2 <= 0
  // (holds for every conic type, no front-facing assumption).  Centred on the low-pass centre and
  // normalised; only used when opa <= 1 so that the fixed low-pass radius bound is valid.
  if (opa <= 1.0f) {
    const double a1x = vy * wz - vz * wy, a1y = vz * wx - vx * wz, a1z = vx * wy - vy * wx;
    const double a2x = wy * uz - wz * uy, a2y = wz * ux - wx * uz, a2z = wx * uy - wy * ux;
    const double a0x = uy * vz - uz * vy, a0y = uz * vx - ux * vz, a0z = ux * vy - uy * vx;
    const double cx = xy.x, cy = xy.y;
    // p at the centre
    const double p0x = a1x * cx + a2x * cy + a0x, p0y = a1y * cx + a2y * cy + a0y, p0z = a1z * cx + a2z * cy + a0z;
    double qa = a1x * a1x + a1y * a1y - c2 * a1z * a1z;
    double qb = a1x * a2x + a1y * a2y - c2 * a1z * a2z;
    double qc = a2x * a2x + a2y * a2y - c2 * a2z * a2z;
    double qd = a1x * p0x + a1y * p0y - c2 * a1z * p0z;
    double qe = a2x * p0x + a2y * p0y - c2 * a2z * p0z;
    double qg = p0x * p0x + p0y * p0y - c2 * p0z * p0z;
    const double sc = fmax(fmax(fabs(qa), fabs(qc)), fabs(qb));
    if (sc > 0.0 && isfinite(sc) && isfinite(qd) && isfinite(qe) && isfinite(qg)) {
      const double inv = 1.0 / sc;
      out[1] = make_float4((float)(qa * inv), (float)(qb * inv), (float)(qc * inv), (float)(qd * inv));
      const float ge = (float)(qe * inv);
      const float gg = __double2float_rd(qg * inv);
      if (isfinite(out[1].w) && isfinite(ge) && isfinite(gg)) out[2] = make_float4(ge, gg, xy.x, xy.y);
      else out[2] = make_float4(0.f, -BIG, xy.x, xy.y);
    }
  }

  // ---- box of the c-sigma ellipse (only when the whole c-sigma disk is in front of the camera plane)
  const double wn = c2 * (wx * wx + wy * wy);
  if (!(wz > 0.0) || !(wz * wz > wn * 1.0201)) { out[0] = all; return; }
  const double d = wn - wz * wz;  // < 0
  const double inv = 1.0 / d;
  const double cx = (c2 * (ux * wx + uy * wy) - uz * wz) * inv;
  const double cy = (c2 * (vx * wx + vy * wy) - vz * wz) * inv;
  const double hx2 = cx * cx - (c2 * (ux * ux + uy * uy) - uz * uz) * inv;
  const double hy2 = cy * cy - (c2 * (vx * vx + vy * vy) - vz * vz) * inv;
  if (!(hx2 >= 0.0) || !(hy2 >= 0.0) || !isfinite(hx2) || !isfinite(hy2)) { out[0] = all; return; }
  const double hx = sqrt(hx2) * 1.01 + 0.5;
  const double hy = sqrt(hy2) * 1.01 + 0.5;
  const double r2 = sqrt(0.5 * c2) + 0.5;
  const double x0 = fmin(cx - hx, (double)xy.x - r2), x1 = fmax(cx + hx, (double)xy.x + r2);
  const double y0 = fmin(cy - hy, (double)xy.y - r2), y1 = fmax(cy + hy, (double)xy.y + r2);
  if (!isfinite(x0) || !isfinite(x1) || !isfinite(y0) || !isfinite(y1)) { out[0] = all; return; }
  out[0] = make_float4(__double2float_rd(x0), __double2float_rd(y0), __double2float_ru(x1), __double2float_ru(y1));
}

// SQ = block-level mode: the surfel is generated here from the superquadric parameters
// (sq_device.cuh: sq_generate) instead of being read from means3D / scales / rotations / opacities.
//
// The SH coefficients (192 B per surfel, the bulk of this kernel's traffic) are loaded warp-cooperatively: read per
// thread they are 48 strided 4-byte loads (32 sectors per warp request), which kept L1TEX 91 % busy at 32 % of DRAM
// peak (profiles/r1_ncu_full_C3_step.txt).  A warp therefore first runs the geometry of its 32 surfels (no early
// return: culled lanes only clear `vis`), compacts the visible ones, copies their coefficient rows to shared memory
// with fully coalesced loads (lane k reads float k of a row: 4 + 2 sectors per row), four rows in flight per step, and
// every visible thread then evaluates the SH polynomial from its shared-memory row (row stride 49 floats:
// conflict-free).  The arithmetic — the expression trees of the geometry and of color_from_sh — is unchanged, so the
// outputs stay bit-identical to the reference build (checked stage by stage in tests/test_gpu_base_raster.py).
constexpr int SH_ROW = 49;  // floats per staged row (48 coefficients + 1 pad)
constexpr size_t SH_COOP_SMEM = (size_t)8 * 32 * SH_ROW * sizeof(float) + 8 * 32 * sizeof(uint32_t);

template <bool SQ>
__global__ void __launch_bounds__(256) preprocess_fwd_kernel(PreprocessFwdArgs a) {
  extern __shared__ __align__(16) unsigned char coop_smem[];
  float* s_rows = reinterpret_cast<float*>(coop_smem);
  uint32_t* s_ids = reinterpret_cast<uint32_t*>(coop_smem + (size_t)8 * 32 * SH_ROW * sizeof(float));
  const unsigned lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int W = a.W, H = a.H;
  const float* orig_points = a.means3D;

  bool vis = false;
  SqSurfel sf;
  float3 p_orig = make_float3(0.f, 0.f, 0.f), p_view = make_float3(0.f, 0.f, 0.f), normal = make_float3(0.f, 0.f, 0.f);
  m3 T;
  float2 point_image = make_float2(0.f, 0.f);
  float radius = 0.f;
  uint2 rect_min = {0u, 0u}, rect_max = {0u, 0u};
  if (idx < a.P) do {
    a.radii[idx] = 0;
    a.tiles_touched[idx] = 0;
    a.depth_key[idx] = 0xffffffffu;
    a.rect[idx] = make_uint2(0u, 0u);
    if (SQ) {
      sf = sq_generate(a.sq, a.sq_vertices, idx);
      if (a.sq_out_xyz) { a.sq_out_xyz[3 * idx] = sf.mean.x; a.sq_out_xyz[3 * idx + 1] = sf.mean.y; a.sq_out_xyz[3 * idx + 2] = sf.mean.z; }
      if (a.sq_out_scaling) { a.sq_out_scaling[2 * idx] = sf.log_scale.x; a.sq_out_scaling[2 * idx + 1] = sf.log_scale.y; }
      if (a.sq_out_rotation) reinterpret_cast<float4*>(a.sq_out_rotation)[idx] = sf.quat;
      if (a.sq_out_opacity) a.sq_out_opacity[idx] = sf.opacity;
    }
    p_orig = SQ ? sf.mean : make_float3(orig_points[3 * idx], orig_points[3 * idx + 1], orig_points[3 * idx + 2]);
    p_view = xform_point4x3(p_orig, a.viewmatrix);
    if (p_view.z <= 0.2f) break;
    if (SQ || a.transMat_precomp == nullptr) {
      const v2 scale_in = SQ ? v2(sf.scale.x, sf.scale.y) : ((const v2*)a.scales)[idx];
      const v4 rot_in = SQ ? v4(sf.quat.x, sf.quat.y, sf.quat.z, sf.quat.w) : ((const v4*)a.rotations)[idx];
      compute_transmat(p_orig, scale_in, a.scale_modifier, rot_in, a.projmatrix, a.viewmatrix, W, H, T, normal);
    } else {
      const v3* T_ptr = (const v3*)a.transMat_precomp;
      T = make_m3(T_ptr[idx * 3 + 0], T_ptr[idx * 3 + 1], T_ptr[idx * 3 + 2]);
      normal = make_float3(0.0, 0.0, 1.0);
    }
    float cosv = -(p_view.x * normal.x + p_view.y * normal.y + p_view.z * normal.z);
    if (cosv == 0) break;
    float multiplier = cosv > 0 ? 1 : -1;
    normal = make_float3(multiplier * normal.x, multiplier * normal.y, multiplier * normal.z);
    float cutoff = 3.0f;
    {
      float2 extent;
      bool ok = compute_aabb(T, cutoff, point_image, extent);
      if (!ok) break;
      radius = ceil(max(max(extent.x, extent.y), cutoff * PGS_FILTER_SIZE));
    }
    dim3 grid(a.grid_x, a.grid_y, 1);
    tile_rect(point_image, radius, rect_min, rect_max, grid);
    if ((rect_max.x - rect_min.x) * (rect_max.y - rect_min.y) == 0) break;
    vis = true;
  } while (0);

  // Digit histograms of the depth keys for the depth sort that follows (binning.cu), accumulated here so that the
  // sort needs no histogram pass of its own.  Culled surfels carry the key 0xffffffff.
  if (a.depth_hist != nullptr) {
    __shared__ uint32_t s_dh[4 * 256];
    for (int i = threadIdx.x; i < 4 * 256; i += 256) s_dh[i] = 0;
    __syncthreads();
    const bool live = idx < a.P;
    const uint32_t key = vis ? __float_as_uint(p_view.z) : 0xffffffffu;
    if (live) {
      atomicAdd(&s_dh[key & 0xffu], 1u);  // the low digits are spread out ...
      atomicAdd(&s_dh[256 + ((key >> 8) & 0xffu)], 1u);
    }
#pragma unroll
    for (int p = 2; p < 4; p++) {  // ... the high ones take a handful of values: one atomic per value and warp
      const uint32_t d = live ? ((key >> (8 * p)) & 0xffu) : (0x100u + lane);
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      if (live && (int)lane == __ffs(peers) - 1) atomicAdd(&s_dh[p * 256 + d], (uint32_t)__popc(peers));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * 256; i += 256) {
      const uint32_t c = s_dh[i];
      if (c) atomicAdd(&a.depth_hist[i], c);
    }
  }

  float r, g, b;
  unsigned clamped = 0;
  if (a.colors_precomp == nullptr) {
    // ---- warp-cooperative copy of the visible surfels' SH rows to shared memory ----
    const unsigned vmask = __ballot_sync(0xffffffffu, vis);
    const int nv = __popc(vmask), slot = __popc(vmask & ((1u << lane) - 1u));
    float* rows = s_rows + (size_t)wrp * 32 * SH_ROW;
    uint32_t* ids = s_ids + wrp * 32;
    if (vis) ids[slot] = (uint32_t)idx;
    __syncwarp();
    const int n = 3 * (a.D + 1) * (a.D + 1);  // floats the active degree reads (<= 48: degrees 0..3, as in the reference)
    const size_t row_floats = (size_t)a.M * 3;
    for (int s0 = 0; s0 < nv; s0 += 4) {
      float v0[4], v1[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        v0[u] = v1[u] = 0.f;
        if (s0 + u < nv) {
          const float* src = a.shs + (size_t)ids[s0 + u] * row_floats;
          if ((int)lane < n) v0[u] = __ldg(src + lane);
          if ((int)lane + 32 < n) v1[u] = __ldg(src + 32 + lane);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (s0 + u < nv) {
          rows[(s0 + u) * SH_ROW + lane] = v0[u];
          if (lane + 32 < (unsigned)SH_ROW) rows[(s0 + u) * SH_ROW + 32 + lane] = v1[u];
        }
      }
    }
    __syncwarp();
    if (!vis) return;
    v3 c = color_from_sh<SQ, true>(idx, a.D, a.M, (const v3*)orig_points, *(const v3*)a.cam_pos, rows + slot * SH_ROW,
                                   clamped, v3(p_orig.x, p_orig.y, p_orig.z));
    r = c.x; g = c.y; b = c.z;
  } else {
    if (!vis) return;
    r = a.colors_precomp[idx * 3 + 0];
    g = a.colors_precomp[idx * 3 + 1];
    b = a.colors_precomp[idx * 3 + 2];
  }

  const float opa = SQ ? sf.opacity : a.opacities[idx];
  float4* rec = a.rec + (size_t)idx * REC_QUADS;
  rec[0] = make_float4(T[0].x, T[0].y, T[0].z, point_image.x);
  rec[1] = make_float4(T[1].x, T[1].y, T[1].z, point_image.y);
  rec[2] = make_float4(T[2].x, T[2].y, T[2].z, opa);
  rec[3] = make_float4(normal.x, normal.y, normal.z, p_view.z);
  rec[4] = make_float4(r, g, b, __uint_as_float(clamped));
  cull_record(T, point_image, opa, a.bbox + (size_t)idx * CULL_QUADS);

  a.radii[idx] = (int)radius;
  a.tiles_touched[idx] = (rect_max.y - rect_min.y) * (rect_max.x - rect_min.x);
  a.depth_key[idx] = __float_as_uint(p_view.z);
  a.rect[idx] = make_uint2(rect_min.x | (rect_max.x << 16), rect_min.y | (rect_max.y << 16));
}

// =============================================================================
// `_part` fork (submodules/diff-surfel-rasterization_part): older 2DGS math.
//   computeTransMat   DSRP/cuda_rasterizer/forward.cu:75-128  (view matrix + intrinsics
//                     {fx, fy, W/2, H/2}; scale_modifier ignored; normal flipped inside)
//   computeAABB       :133-163 (1-sigma box, h = sqrt(max(0,h0)))
//   preprocessCUDA    :166-260 (radius = ceil(3 * max(ext, 0.7071067811865476)) in double)
// =============================================================================
__device__ __forceinline__ bool part_transmat(const v3& p_world, const v4& quat, const v2& scale,
                                              const float* viewmat, const float4& intrins, m3& Tout,
                                              float3& normal) {
  m3 W;
  W[0] = v3(viewmat[0], viewmat[1], viewmat[2]);
  W[1] = v3(viewmat[4], viewmat[5], viewmat[6]);
  W[2] = v3(viewmat[8], viewmat[9], viewmat[10]);
  const v3 cam_pos = v3(viewmat[12], viewmat[13], viewmat[14]);
  m4 Pm;
  Pm[0] = v4(intrins.x, 0.0f, 0.0f, 0.0f);
  Pm[1] = v4(0.0f, intrins.y, 0.0f, 0.0f);
  Pm[2] = v4(intrins.z, intrins.w, 1.0f, 1.0f);
  Pm[3] = v4(0.0f, 0.0f, 0.0f, 0.0f);

  v3 p_view = W * p_world + cam_pos;
  m3 S = diag3(1.f);
  S[0][0] = 1.0f * scale.x;
  S[1][1] = 1.0f * scale.y;
  S[2][2] = 1.0f * 1.0f;
  m3 R = quat_to_rotmat(quat) * S;
  m3 M = make_m3(W * R[0], W * R[1], p_view);
  v3 tn = W * R[2];
  float cosv = dot(-tn, p_view);
  if (cosv == 0.0f) return false;
  float multiplier = cosv > 0 ? 1 : -1;
  tn *= multiplier;
  m4x3 T = transpose(Pm * make_m3x4(v4(M[0], 0.0f), v4(M[1], 0.0f), v4(M[2], 1.0f)));
  Tout[0] = T[0];
  Tout[1] = T[1];
  Tout[2] = T[2];
  normal = {tn.x, tn.y, tn.z};
  return true;
}

__device__ __forceinline__ bool part_aabb(const m3& T3, float2& center, float2& extent) {
  m4x3 T;
  T[0] = T3[0]; T[1] = T3[1]; T[2] = T3[2]; T[3] = T3[2];
  float d = dot(v3(1.0f, 1.0f, -1.0f), T[3] * T[3]);
  if (d == 0.0f) return false;
  v3 f = v3(1.0f, 1.0f, -1.0f) * (1.0f / d);
  v3 p = v3(dot(f, T[0] * T[3]), dot(f, T[1] * T[3]), dot(f, T[2] * T[3]));
  v3 h0 = p * p - v3(dot(f, T[0] * T[0]), dot(f, T[1] * T[1]), dot(f, T[2] * T[2]));
  v3 h = vsqrt(vmax(v3(0.0f, 0.0f, 0.0f), h0)) + v3(0.0f, 0.0f, (float)1e-2);
  center = {p.x, p.y};
  extent = {h.x, h.y};
  return true;
}

__global__ void __launch_bounds__(256) preprocess_fwd_part_kernel(PreprocessFwdArgs a) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.P) return;
  a.radii[idx] = 0;
  a.tiles_touched[idx] = 0;
  a.depth_key[idx] = 0xffffffffu;
  a.rect[idx] = make_uint2(0u, 0u);

  const int W = a.W, H = a.H;
  const float* orig_points = a.means3D;
  v3 p_world = v3(orig_points[3 * idx], orig_points[3 * idx + 1], orig_points[3 * idx + 2]);
  float3 p_orig = {p_world.x, p_world.y, p_world.z};
  float3 p_view = xform_point4x3(p_orig, a.viewmatrix);
  if (p_view.z <= 0.2f) return;

  float4 intrins = {a.focal_x, a.focal_y, (float)(float(W) / 2.0), (float)(float(H) / 2.0)};
  v2 scale = ((const v2*)a.scales)[idx];
  v4 quat = ((const v4*)a.rotations)[idx];
  m3 T;
  float3 normal;
  if (!part_transmat(p_world, quat, scale, a.viewmatrix, intrins, T, normal)) return;

  float2 center, extent;
  if (!part_aabb(T, center, extent)) return;
  float truncated_R = 3.f;
  float radius = ceil(truncated_R * max(max(extent.x, extent.y), 0.7071067811865476));

  dim3 grid(a.grid_x, a.grid_y, 1);
  uint2 rect_min, rect_max;
  tile_rect(center, radius, rect_min, rect_max, grid);
  if ((rect_max.x - rect_min.x) * (rect_max.y - rect_min.y) == 0) return;

  float r, g, b;
  unsigned clamped = 0;
  if (a.colors_precomp == nullptr) {
    v3 c = color_from_sh(idx, a.D, a.M, (const v3*)orig_points, *(const v3*)a.cam_pos, a.shs, clamped);
    r = c.x; g = c.y; b = c.z;
  } else {
    r = a.colors_precomp[idx * 3 + 0];
    g = a.colors_precomp[idx * 3 + 1];
    b = a.colors_precomp[idx * 3 + 2];
  }
  const float opa = a.opacities[idx];
  float4* rec = a.rec + (size_t)idx * REC_QUADS;
  rec[0] = make_float4(T[0].x, T[0].y, T[0].z, center.x);
  rec[1] = make_float4(T[1].x, T[1].y, T[1].z, center.y);
  rec[2] = make_float4(T[2].x, T[2].y, T[2].z, opa);
  rec[3] = make_float4(normal.x, normal.y, normal.z, p_view.z);
  rec[4] = make_float4(r, g, b, __uint_as_float(clamped));
  cull_record(T, center, opa, a.bbox + (size_t)idx * CULL_QUADS);
  a.radii[idx] = (int)radius;
  a.tiles_touched[idx] = (rect_max.y - rect_min.y) * (rect_max.x - rect_min.x);
  a.depth_key[idx] = __float_as_uint(p_view.z);
  a.rect[idx] = make_uint2(rect_min.x | (rect_max.x << 16), rect_min.y | (rect_max.y << 16));
}

void launch_preprocess_fwd_part(const PreprocessFwdArgs& a, cudaStream_t s) {
  if (a.P <= 0) return;
  preprocess_fwd_part_kernel<<<(a.P + 255) / 256, 256, 0, s>>>(a);
  count_launch();
}

// mark_visible (reference checkFrustum, rasterizer_impl.cu:54-66)
__global__ void __launch_bounds__(256) check_frustum_kernel(int P, const float* means3D, const float* viewmatrix,
                                                            unsigned char* present) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  float3 p = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
  float3 pv = xform_point4x3(p, viewmatrix);
  present[idx] = (pv.z <= 0.2f) ? 0 : 1;
}

void launch_preprocess_fwd(const PreprocessFwdArgs& a, cudaStream_t s) {
  if (a.P <= 0) return;
  static bool attr_set[2][64] = {};
  if (first_use_on_device(attr_set[a.use_sq ? 1 : 0])) {
    if (a.use_sq) cudaFuncSetAttribute(preprocess_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SH_COOP_SMEM);
    else cudaFuncSetAttribute(preprocess_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SH_COOP_SMEM);
  }
  if (a.use_sq) preprocess_fwd_kernel<true><<<(a.P + 255) / 256, 256, SH_COOP_SMEM, s>>>(a);
  else preprocess_fwd_kernel<false><<<(a.P + 255) / 256, 256, SH_COOP_SMEM, s>>>(a);
  count_launch();
}

void launch_check_frustum(int P, const float* means3D, const float* viewmatrix, unsigned char* present,
                          cudaStream_t s) {
  if (P <= 0) return;
  check_frustum_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, viewmatrix, present);
  count_launch();
}

}  // namespace pgs
