// Superquadric -> surfel parameterisation, forward and backward (sm_100a).
//
// Replaces the ~45 forward / ~90 backward ATen kernels PartGS launches every iteration in
// BlockGaussianModel.prepare_scaling_rot / get_verts / get_opacity
// (games/block_mesh_splatting/scene/block_gaussian_model.py:189-256, 106-109) with
//   forward : 1 kernel for the B x Vt superquadric vertices (parametric_sq,
//             utils/superquadric.py:10-14; signed_pow utils/pytorch.py:28-29; block S,R,t)
//             + 1 kernel with one thread per surfel (barycentric centre, tangent frame,
//             log-scales, pytorch3d-style matrix->quaternion, utils/general_utils.py:34-87)
//   backward: 1 kernel with one warp per triangle (the K surfels of a face share the frame:
//             per-surfel gradients are reduced across the warp with the transposed butterfly,
//             then one lane differentiates the frame and scatters 9 atomics into the
//             vertex gradient) + 1 kernel with one CTA per superquadric reducing the vertex
//             gradients to the 13 block parameters.
// Surfel order is (block, face, sample) row-major like the reference (BGM:256).
#include "common.cuh"
#include "kernels.h"
#include "sq_device.cuh"

namespace pgs {

// ---- forward: vertices ------------------------------------------------------------
__global__ void __launch_bounds__(128) sq_vertices_kernel(SqArgs a, float* __restrict__ vertices) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.B * a.Vt) return;
  const int b = idx / a.Vt;
  const BlockPose p = load_pose(a, b);
  const float eta = a.eta[idx], omega = a.omega[idx];
  const float ce = spow(cosf(eta), p.e1), se = spow(sinf(eta), p.e1);
  const float co = spow(cosf(omega), p.e2), so = spow(sinf(omega), p.e2);
  const float v[3] = {ce * so * a.ratio, se * a.ratio, ce * co * a.ratio};
  const float u[3] = {v[0] * p.S[0], v[1] * p.S[1], v[2] * p.S[2]};
  for (int j = 0; j < 3; j++)
    vertices[3 * idx + j] = u[0] * p.R[0][j] + u[1] * p.R[1][j] + u[2] * p.R[2][j] + p.t[j];
}

// ---- forward: surfels ---------------------------------------------------------------
__global__ void __launch_bounds__(256) sq_surfels_kernel(SqArgs a, const float* __restrict__ vertices,
                                                        float* __restrict__ xyz, float* __restrict__ scaling,
                                                        float* __restrict__ rotation, float* __restrict__ opacity) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long P = (long long)a.B * a.F * a.K;
  if (idx >= P) return;
  const SqSurfel sf = sq_generate(a, vertices, idx);
  xyz[idx * 3 + 0] = sf.mean.x; xyz[idx * 3 + 1] = sf.mean.y; xyz[idx * 3 + 2] = sf.mean.z;
  scaling[idx * 2 + 0] = sf.log_scale.x;
  scaling[idx * 2 + 1] = sf.log_scale.y;
  rotation[idx * 4 + 0] = sf.quat.x; rotation[idx * 4 + 1] = sf.quat.y;
  rotation[idx * 4 + 2] = sf.quat.z; rotation[idx * 4 + 3] = sf.quat.w;
  opacity[idx] = sf.opacity;
}

// Sum v[0..15] over the warp; component (lane >> 1) is returned on every lane.
__device__ __forceinline__ float sq_warp_reduce16(float (&v)[16], unsigned lane) {
#pragma unroll
  for (int step = 0; step < 4; step++) {
    const int half = 8 >> step;
    const unsigned bit = 16u >> step;
    const bool hi = lane & bit;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (i < half) {
        float keep = hi ? v[i + half] : v[i];
        float send = hi ? v[i] : v[i + half];
        v[i] = keep + __shfl_xor_sync(SQ_FULL, send, bit);
      }
    }
  }
  v[0] += __shfl_xor_sync(SQ_FULL, v[0], 1);
  return v[0];
}

// ---- backward: one warp per face ------------------------------------------------------
// d_vertices [B,Vt,3] must be zeroed (or hold an upstream vertex gradient); d_occ_acc [B] zeroed.
__global__ void __launch_bounds__(256) sq_faces_bwd_kernel(SqArgs a, const float* __restrict__ vertices,
                                                          const float* __restrict__ d_xyz,
                                                          const float* __restrict__ d_scaling,
                                                          const float* __restrict__ d_rotation,
                                                          const float* __restrict__ d_opacity,
                                                          float* __restrict__ d_vertices, float* __restrict__ d_occ_acc,
                                                          float* __restrict__ d_alpha, float* __restrict__ d_scale_raw) {
  const unsigned lane = threadIdx.x & 31;
  const int face_id = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (face_id >= a.B * a.F) return;
  const int b = face_id / a.F;
  const int* face = a.faces + (size_t)face_id * 3;
  const Frame fr = make_frame(vertices + (size_t)b * a.Vt * 3, face);

  // acc: [0..8] dt0,dt1,dt2 from xyz ; [9] ds1 ; [10] ds2 ; [11..14] dq ; [15] d opacity
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = 0.f;
  const long long base = (long long)face_id * a.K;
  for (int k = lane; k < a.K; k += 32) {
    const long long idx = base + k;
    const float* al = a.alpha + idx * 3;
    const float gx[3] = {d_xyz[idx * 3], d_xyz[idx * 3 + 1], d_xyz[idx * 3 + 2]};
    for (int j = 0; j < 3; j++) {
      acc[j] += al[0] * gx[j];
      acc[3 + j] += al[1] * gx[j];
      acc[6 + j] += al[2] * gx[j];
    }
    if (d_alpha) {
      d_alpha[idx * 3 + 0] = dot3(gx, fr.t0);
      d_alpha[idx * 3 + 1] = dot3(gx, fr.t1);
      d_alpha[idx * 3 + 2] = dot3(gx, fr.t2);
    }
    const float sc = a.scale_raw[idx];
    const float z1 = sc * fr.s1, z2 = sc * fr.s2;
    const float dz1 = z1 > 0.f ? d_scaling[idx * 2] / (z1 + SQ_EPS) : 0.f;
    const float dz2 = z2 > 0.f ? d_scaling[idx * 2 + 1] / (z2 + SQ_EPS) : 0.f;
    acc[9] += dz1 * sc;
    acc[10] += dz2 * sc;
    if (d_scale_raw) d_scale_raw[idx] = dz1 * fr.s1 + dz2 * fr.s2;
    for (int i = 0; i < 4; i++) acc[11 + i] += d_rotation[idx * 4 + i];
    acc[15] += d_opacity ? d_opacity[idx] : 0.f;
  }
  const float red = sq_warp_reduce16(acc, lane);
  // gather the 16 sums on every lane (component c lives on lanes 2c, 2c+1)
  float g[16];
#pragma unroll
  for (int c = 0; c < 16; c++) g[c] = __shfl_sync(SQ_FULL, red, 2 * c);
  if (lane != 0) return;

  float dt0[3] = {g[0], g[1], g[2]}, dt1[3] = {g[3], g[4], g[5]}, dt2[3] = {g[6], g[7], g[8]};
  const float ds1 = g[9], ds2 = g[10];
  const float dq_in[4] = {g[11], g[12], g[13], g[14]};
  atomicAdd(&d_occ_acc[b], g[15]);

  // ---- quaternion -> rotation matrix gradient (replay of mat_to_quat) ----
  float m[3][3];
  for (int i = 0; i < 3; i++) { m[i][0] = fr.v1[i]; m[i][1] = fr.v2[i]; m[i][2] = fr.v0[i]; }
  float q[4], qa[4], sign;
  int sel;
  mat_to_quat(m, q, sel, sign, qa);
  const float Dq = 2.0f * fmaxf(qa[sel], 0.1f);
  float row[4], drow[4];
  for (int i = 0; i < 4; i++) { row[i] = sign * q[i] * Dq; drow[i] = sign * dq_in[i] / Dq; }
  float dD = 0.f;
  for (int i = 0; i < 4; i++) dD -= sign * dq_in[i] * row[i] / (Dq * Dq);
  float dqa = (qa[sel] > 0.1f) ? 2.f * dD : 0.f;   // D = 2*max(qa,0.1)
  dqa += 2.f * qa[sel] * drow[sel];                 // row[sel] = qa^2
  const float dx = qa[sel] > 0.f ? dqa / (2.f * qa[sel]) : 0.f;
  float dm[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  const float sg[4][3] = {{1, 1, 1}, {1, -1, -1}, {-1, 1, -1}, {-1, -1, 1}};
  dm[0][0] += sg[sel][0] * dx; dm[1][1] += sg[sel][1] * dx; dm[2][2] += sg[sel][2] * dx;
  switch (sel) {
    case 0: dm[2][1] += drow[1]; dm[1][2] -= drow[1]; dm[0][2] += drow[2]; dm[2][0] -= drow[2]; dm[1][0] += drow[3]; dm[0][1] -= drow[3]; break;
    case 1: dm[2][1] += drow[0]; dm[1][2] -= drow[0]; dm[1][0] += drow[2]; dm[0][1] += drow[2]; dm[0][2] += drow[3]; dm[2][0] += drow[3]; break;
    case 2: dm[0][2] += drow[0]; dm[2][0] -= drow[0]; dm[1][0] += drow[1]; dm[0][1] += drow[1]; dm[1][2] += drow[3]; dm[2][1] += drow[3]; break;
    default: dm[1][0] += drow[0]; dm[0][1] -= drow[0]; dm[2][0] += drow[1]; dm[0][2] += drow[1]; dm[2][1] += drow[2]; dm[1][2] += drow[2]; break;
  }
  float dv1[3], dv2[3], dv0[3];
  for (int i = 0; i < 3; i++) { dv1[i] = dm[i][0]; dv2[i] = dm[i][1]; dv0[i] = dm[i][2]; }

  // ---- frame gradient ----
  float db[3], dw[3], da[3];
  for (int j = 0; j < 3; j++) { db[j] = 0.5f * ds2 * fr.v2[j]; dv2[j] += 0.5f * ds2 * fr.b[j]; }
  {
    const float inv = 1.f / (fr.lw + SQ_EPS);
    const float dlw = -dot3(dv2, fr.w) * inv * inv;
    for (int j = 0; j < 3; j++) dw[j] = dv2[j] * inv + (fr.lw > 0.f ? dlw * fr.w[j] / fr.lw : 0.f);
  }
  const float dc0 = -dot3(dw, fr.v0), dc1 = -dot3(dw, fr.v1);
  for (int j = 0; j < 3; j++) {
    db[j] += dw[j] + dc0 * fr.v0[j] + dc1 * fr.v1[j];
    dv0[j] += -fr.c0 * dw[j] + dc0 * fr.b[j];
    dv1[j] += -fr.c1 * dw[j] + dc1 * fr.b[j];
  }
  {
    const float inv = 1.f / (fr.la + SQ_EPS);
    const float dla = 0.5f * ds1 - dot3(dv1, fr.a) * inv * inv;
    for (int j = 0; j < 3; j++) da[j] = dv1[j] * inv + (fr.la > 0.f ? dla * fr.a[j] / fr.la : 0.f);
  }
  float dn[3];
  {
    const float inv = 1.f / (fr.ln + SQ_EPS);
    const float dln = -dot3(dv0, fr.n) * inv * inv;
    for (int j = 0; j < 3; j++) dn[j] = dv0[j] * inv + (fr.ln > 0.f ? dln * fr.n[j] / fr.ln : 0.f);
  }
  float e1[3], e2[3], de1[3], de2[3];
  for (int j = 0; j < 3; j++) { e1[j] = fr.t1[j] - fr.t0[j]; e2[j] = fr.t2[j] - fr.t0[j]; }
  cross3(e2, dn, de1);
  cross3(dn, e1, de2);
  for (int j = 0; j < 3; j++) {
    const float dmj = -(da[j] + db[j]) / 3.f;
    dt0[j] += dmj - de1[j] - de2[j];
    dt1[j] += dmj + da[j] + de1[j];
    dt2[j] += dmj + db[j] + de2[j];
  }
  float* dvb = d_vertices + (size_t)b * a.Vt * 3;
  for (int j = 0; j < 3; j++) {
    atomicAdd(&dvb[3 * face[0] + j], dt0[j]);
    atomicAdd(&dvb[3 * face[1] + j], dt1[j]);
    atomicAdd(&dvb[3 * face[2] + j], dt2[j]);
  }
}

// ---- backward: one CTA per superquadric ---------------------------------------------
__global__ void __launch_bounds__(256) sq_blocks_bwd_kernel(SqArgs a, const float* __restrict__ d_vertices,
                                                           const float* __restrict__ d_occ_acc,
                                                           float* __restrict__ d_sq_r, float* __restrict__ d_sq_s,
                                                           float* __restrict__ d_sq_t, float* __restrict__ d_sq_eps,
                                                           float* __restrict__ d_sq_occ) {
  __shared__ float s_red[8][17];
  const int b = blockIdx.x;
  const BlockPose p = load_pose(a, b);
  // per-thread partials: [0..2] dt, [3..5] dS, [6..14] dR (row-major), [15] de1, [16] de2
  float part[17];
#pragma unroll
  for (int i = 0; i < 17; i++) part[i] = 0.f;
  for (int i = threadIdx.x; i < a.Vt; i += blockDim.x) {
    const int idx = b * a.Vt + i;
    const float eta = a.eta[idx], omega = a.omega[idx];
    const float ce0 = cosf(eta), se0 = sinf(eta), co0 = cosf(omega), so0 = sinf(omega);
    const float A = spow(ce0, p.e1), Bq = spow(se0, p.e1), Cc = spow(co0, p.e2), Dd = spow(so0, p.e2);
    const float v[3] = {A * Dd * a.ratio, Bq * a.ratio, A * Cc * a.ratio};
    const float dv[3] = {d_vertices[3 * idx], d_vertices[3 * idx + 1], d_vertices[3 * idx + 2]};
    float du[3];
    for (int r = 0; r < 3; r++) {
      const float u = v[r] * p.S[r];
      part[r] += dv[r];
      du[r] = p.R[r][0] * dv[0] + p.R[r][1] * dv[1] + p.R[r][2] * dv[2];
      for (int j = 0; j < 3; j++) part[6 + 3 * r + j] += u * dv[j];
      part[3 + r] += du[r] * v[r];
    }
    const float dvl[3] = {du[0] * p.S[0] * a.ratio, du[1] * p.S[1] * a.ratio, du[2] * p.S[2] * a.ratio};
    const float dA = dvl[0] * Dd + dvl[2] * Cc, dB = dvl[1], dC = dvl[2] * A, dDd = dvl[0] * A;
    // d/de sign(t)|t|^e = spow * ln|t|, 0 at t == 0 (torch.pow's exponent-gradient mask)
    const float lce = ce0 != 0.f ? logf(fabsf(ce0)) : 0.f, lse = se0 != 0.f ? logf(fabsf(se0)) : 0.f;
    const float lco = co0 != 0.f ? logf(fabsf(co0)) : 0.f, lso = so0 != 0.f ? logf(fabsf(so0)) : 0.f;
    part[15] += dA * A * lce + dB * Bq * lse;
    part[16] += dC * Cc * lco + dDd * Dd * lso;
  }
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 17; i++) {
    float v = part[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SQ_FULL, v, o);
    if (lane == 0) s_red[wid][i] = v;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  float tot[17];
  for (int i = 0; i < 17; i++) {
    tot[i] = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot[i] += s_red[w][i];
  }
  for (int j = 0; j < 3; j++) {
    d_sq_t[3 * b + j] = tot[j];
    d_sq_s[3 * b + j] = tot[3 + j] * expf(a.sq_s[3 * b + j]);  // S = exp(s) + min
  }
  const float sg1 = sigmoidf(a.sq_eps[2 * b]), sg2 = sigmoidf(a.sq_eps[2 * b + 1]);
  d_sq_eps[2 * b] = tot[15] * 1.8f * sg1 * (1.f - sg1);
  d_sq_eps[2 * b + 1] = tot[16] * 1.8f * sg2 * (1.f - sg2);
  const float so_ = sigmoidf(a.sq_occ[b]);
  d_sq_occ[b] = d_occ_acc[b] * so_ * (1.f - so_);
  // rotation matrix -> unit quaternion -> raw quaternion (F.normalize)
  const float w = p.q[0], x = p.q[1], y = p.q[2], z = p.q[3];
  const float* dR = &tot[6];
#define DR(i, j) dR[3 * (i) + (j)]
  float dq[4];
  dq[0] = 2.f * (-z * DR(0, 1) + y * DR(0, 2) + z * DR(1, 0) - x * DR(1, 2) - y * DR(2, 0) + x * DR(2, 1));
  dq[1] = 2.f * (y * DR(0, 1) + z * DR(0, 2) + y * DR(1, 0) - 2.f * x * DR(1, 1) - w * DR(1, 2) + z * DR(2, 0) +
                 w * DR(2, 1) - 2.f * x * DR(2, 2));
  dq[2] = 2.f * (-2.f * y * DR(0, 0) + x * DR(0, 1) + w * DR(0, 2) + x * DR(1, 0) + z * DR(1, 2) - w * DR(2, 0) +
                 z * DR(2, 1) - 2.f * y * DR(2, 2));
  dq[3] = 2.f * (-2.f * z * DR(0, 0) - w * DR(0, 1) + x * DR(0, 2) + w * DR(1, 0) - 2.f * z * DR(1, 1) + y * DR(1, 2) +
                 x * DR(2, 0) + y * DR(2, 1));
#undef DR
  const float qd = p.q[0] * dq[0] + p.q[1] * dq[1] + p.q[2] * dq[2] + p.q[3] * dq[3];
  for (int i = 0; i < 4; i++) d_sq_r[4 * b + i] = (dq[i] - p.q[i] * qd) / p.rn;
}

void launch_sq_vertices(const SqArgs& a, float* vertices, cudaStream_t s) {
  const int nv = a.B * a.Vt;
  if (nv <= 0) return;
  sq_vertices_kernel<<<(nv + 127) / 128, 128, 0, s>>>(a, vertices);
  count_launch();
}

void launch_sq_forward(const SqArgs& a, float* vertices, float* xyz, float* scaling, float* rotation, float* opacity,
                       cudaStream_t s) {
  const int nv = a.B * a.Vt;
  if (nv <= 0) return;
  launch_sq_vertices(a, vertices, s);
  const long long P = (long long)a.B * a.F * a.K;
  if (P <= 0) return;
  sq_surfels_kernel<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(a, vertices, xyz, scaling, rotation, opacity);
  count_launch();
}

void launch_sq_backward(const SqArgs& a, const float* vertices, const float* d_xyz, const float* d_scaling,
                        const float* d_rotation, const float* d_opacity, float* d_vertices, float* d_occ_acc,
                        float* d_alpha, float* d_scale_raw, float* d_sq_r, float* d_sq_s, float* d_sq_t,
                        float* d_sq_eps, float* d_sq_occ, cudaStream_t s) {
  const int nf = a.B * a.F;
  if (nf > 0 && a.K > 0) {
    sq_faces_bwd_kernel<<<(nf + 7) / 8, 256, 0, s>>>(a, vertices, d_xyz, d_scaling, d_rotation, d_opacity, d_vertices,
                                                      d_occ_acc, d_alpha, d_scale_raw);
    count_launch();
  }
  if (a.B > 0) {
    sq_blocks_bwd_kernel<<<a.B, 256, 0, s>>>(a, d_vertices, d_occ_acc, d_sq_r, d_sq_s, d_sq_t, d_sq_eps, d_sq_occ);
    count_launch();
  }
}

}  // namespace pgs
