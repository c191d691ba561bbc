// Device-side superquadric -> surfel arithmetic shared by the stand-alone parameterisation kernels
// (sq2surfel.cu) and the preprocess kernels that generate block-level surfels in place
// (preprocess_fwd.cu / preprocess_bwd.cu, `SQ` variants).  Reference:
// games/block_mesh_splatting/scene/block_gaussian_model.py:189-256, utils/superquadric.py:10-14,93-101,
// utils/general_utils.py:34-87.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace pgs {

constexpr unsigned SQ_FULL = 0xffffffffu;
#define SQ_EPS 1e-8f

__device__ __forceinline__ float spow(float t, float e) {
  float s = (t > 0.f) ? 1.f : ((t < 0.f) ? -1.f : 0.f);
  return s * powf(fabsf(t), e);
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

struct BlockPose {
  float S[3];
  float R[3][3];  // row-major, quaternion_to_rotation_matrix (utils/superquadric.py:93-101)
  float t[3];
  float e1, e2;
  float q[4];     // normalised (w,x,y,z)
  float rn;       // |sq_r| clamped like F.normalize
};

__device__ __forceinline__ BlockPose load_pose(const SqArgs& a, int b) {
  BlockPose p;
  p.e1 = sigmoidf(a.sq_eps[2 * b]) * 1.8f + 0.1f;
  p.e2 = sigmoidf(a.sq_eps[2 * b + 1]) * 1.8f + 0.1f;
  for (int i = 0; i < 3; i++) {
    p.S[i] = expf(a.sq_s[3 * b + i]) + a.scale_min;
    p.t[i] = a.sq_t[3 * b + i];
  }
  float r0 = a.sq_r[4 * b], r1 = a.sq_r[4 * b + 1], r2 = a.sq_r[4 * b + 2], r3 = a.sq_r[4 * b + 3];
  p.rn = fmaxf(sqrtf(r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3), 1e-12f);
  float w = r0 / p.rn, x = r1 / p.rn, y = r2 / p.rn, z = r3 / p.rn;
  p.q[0] = w; p.q[1] = x; p.q[2] = y; p.q[3] = z;
  p.R[0][0] = 1 - 2 * y * y - 2 * z * z; p.R[0][1] = 2 * x * y - 2 * z * w; p.R[0][2] = 2 * x * z + 2 * y * w;
  p.R[1][0] = 2 * x * y + 2 * z * w; p.R[1][1] = 1 - 2 * x * x - 2 * z * z; p.R[1][2] = 2 * y * z - 2 * x * w;
  p.R[2][0] = 2 * x * z - 2 * y * w; p.R[2][1] = 2 * y * z + 2 * x * w; p.R[2][2] = 1 - 2 * x * x - 2 * y * y;
  return p;
}

struct Frame {
  float t0[3], t1[3], t2[3];
  float n[3], ln, v0[3];
  float a[3], la, v1[3];
  float b[3], c0, c1, w[3], lw, v2[3];
  float s1, s2;
};

__device__ __forceinline__ float dot3(const float* x, const float* y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; }
__device__ __forceinline__ void cross3(const float* x, const float* y, float* o) {
  o[0] = x[1] * y[2] - x[2] * y[1];
  o[1] = x[2] * y[0] - x[0] * y[2];
  o[2] = x[0] * y[1] - x[1] * y[0];
}

__device__ __forceinline__ Frame make_frame(const float* verts_b, const int* face) {
  Frame f;
  for (int j = 0; j < 3; j++) {
    f.t0[j] = verts_b[3 * face[0] + j];
    f.t1[j] = verts_b[3 * face[1] + j];
    f.t2[j] = verts_b[3 * face[2] + j];
  }
  float e1[3], e2[3], m[3];
  for (int j = 0; j < 3; j++) {
    e1[j] = f.t1[j] - f.t0[j];
    e2[j] = f.t2[j] - f.t0[j];
    m[j] = (f.t0[j] + f.t1[j] + f.t2[j]) / 3.f;
  }
  cross3(e1, e2, f.n);
  f.ln = sqrtf(dot3(f.n, f.n));
  for (int j = 0; j < 3; j++) {
    f.v0[j] = f.n[j] / (f.ln + SQ_EPS);
    f.a[j] = f.t1[j] - m[j];
    f.b[j] = f.t2[j] - m[j];
  }
  f.la = sqrtf(dot3(f.a, f.a));
  for (int j = 0; j < 3; j++) f.v1[j] = f.a[j] / (f.la + SQ_EPS);
  f.c0 = dot3(f.b, f.v0);
  f.c1 = dot3(f.b, f.v1);
  for (int j = 0; j < 3; j++) f.w[j] = f.b[j] - f.c0 * f.v0[j] - f.c1 * f.v1[j];
  f.lw = sqrtf(dot3(f.w, f.w));
  for (int j = 0; j < 3; j++) f.v2[j] = f.w[j] / (f.lw + SQ_EPS);
  f.s1 = (f.la + SQ_EPS) / 2.f;
  f.s2 = dot3(f.b, f.v2) / 2.f;
  return f;
}

// pytorch3d-style matrix_to_quaternion (utils/general_utils.py:34-87); m[i][j] row-major.
// Returns the selected candidate index and sign so that backward can replay it.
__device__ __forceinline__ void mat_to_quat(const float m[3][3], float q[4], int& sel, float& sign, float qa[4]) {
  const float x[4] = {1.f + m[0][0] + m[1][1] + m[2][2], 1.f + m[0][0] - m[1][1] - m[2][2],
                      1.f - m[0][0] + m[1][1] - m[2][2], 1.f - m[0][0] - m[1][1] + m[2][2]};
  sel = 0;
  for (int i = 0; i < 4; i++) qa[i] = x[i] > 0.f ? sqrtf(x[i]) : 0.f;
  for (int i = 1; i < 4; i++)
    if (qa[i] > qa[sel]) sel = i;
  float row[4];
  switch (sel) {
    case 0: row[0] = qa[0] * qa[0]; row[1] = m[2][1] - m[1][2]; row[2] = m[0][2] - m[2][0]; row[3] = m[1][0] - m[0][1]; break;
    case 1: row[0] = m[2][1] - m[1][2]; row[1] = qa[1] * qa[1]; row[2] = m[1][0] + m[0][1]; row[3] = m[0][2] + m[2][0]; break;
    case 2: row[0] = m[0][2] - m[2][0]; row[1] = m[1][0] + m[0][1]; row[2] = qa[2] * qa[2]; row[3] = m[1][2] + m[2][1]; break;
    default: row[0] = m[1][0] - m[0][1]; row[1] = m[2][0] + m[0][2]; row[2] = m[2][1] + m[1][2]; row[3] = qa[3] * qa[3]; break;
  }
  const float D = 2.0f * fmaxf(qa[sel], 0.1f);
  for (int i = 0; i < 4; i++) q[i] = row[i] / D;
  sign = q[0] < 0.f ? -1.f : 1.f;
  for (int i = 0; i < 4; i++) q[i] *= sign;
}

// One surfel of the (block, face, sample) row-major list, exactly as sq_surfels_kernel materialises it and as
// the accessors of BlockGaussianModel hand it to the rasteriser: centre, ACTIVATED scale exp(log(relu(s)+eps)),
// quaternion (w,x,y,z) from the tangent frame, opacity sigmoid(occ).
struct SqSurfel {
  float3 mean;
  float2 scale;      // activated
  float2 log_scale;  // what sq_surfels_kernel stores
  float4 quat;
  float opacity;
};
__device__ __forceinline__ SqSurfel sq_generate(const SqArgs& a, const float* __restrict__ vertices, long long idx) {
  const int FK = a.F * a.K;
  const int b = (int)(idx / FK);
  const int fk = (int)(idx - (long long)b * FK);
  const int f = fk / a.K;
  const int* face = a.faces + ((size_t)b * a.F + f) * 3;
  const Frame fr = make_frame(vertices + (size_t)b * a.Vt * 3, face);
  SqSurfel s;
  const float* al = a.alpha + idx * 3;
  s.mean = make_float3(al[0] * fr.t0[0] + al[1] * fr.t1[0] + al[2] * fr.t2[0],
                       al[0] * fr.t0[1] + al[1] * fr.t1[1] + al[2] * fr.t2[1],
                       al[0] * fr.t0[2] + al[1] * fr.t1[2] + al[2] * fr.t2[2]);
  const float sc = a.scale_raw[idx];
  s.log_scale = make_float2(logf(fmaxf(sc * fr.s1, 0.f) + SQ_EPS), logf(fmaxf(sc * fr.s2, 0.f) + SQ_EPS));
  s.scale = make_float2(expf(s.log_scale.x), expf(s.log_scale.y));
  float m[3][3];
  for (int i = 0; i < 3; i++) { m[i][0] = fr.v1[i]; m[i][1] = fr.v2[i]; m[i][2] = fr.v0[i]; }
  float q[4], qa[4], sign;
  int sel;
  mat_to_quat(m, q, sel, sign, qa);
  s.quat = make_float4(q[0], q[1], q[2], q[3]);
  s.opacity = sigmoidf(a.sq_occ[b]);
  return s;
}

}  // namespace pgs
