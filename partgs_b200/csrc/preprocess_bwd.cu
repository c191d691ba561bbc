// Backward preprocess for the base surfel rasteriser (sm_100a).
//
// One thread per surfel: takes the per-surfel gradient record accumulated by the
// backward render kernel (dL/dT, dL/dmean2D, dL/dnormal, dL/dopacity, dL/dcolour)
// and produces gradients for mean3D, scale, rotation, SH coefficients plus the
// "densification" mean2D gradient.  Every output row is written (zeros for
// culled surfels), so the caller never has to zero-fill 304 B/surfel of gradient
// tensors as the reference glue does (rasterize_points.cu:187-195).
//
// Replaces reference BACKWARD::preprocessCUDA (cuda_rasterizer/backward.cu:575-630),
// compute_transmat_aabb (:443-573), SH backward (:20-139) and quat_to_rotmat_vjp
// (auxiliary.h:238-282).  Quirks kept: scale_modifier ignored (:481), W/H re-derived
// from focal*tan (:607-608), mean2D gradient overwritten by the hack (:626-629).
#include "common.cuh"
#include "linalg.cuh"
#include "kernels.h"
#include "sq_device.cuh"

namespace pgs {

__device__ const float bSH_C0 = 0.28209479177387814f;
__device__ const float bSH_C1 = 0.4886025119029199f;
__device__ const float bSH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                   -1.0925484305920792f, 0.5462742152960396f};
__device__ const float bSH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                                   -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

__device__ __forceinline__ float3 dnormvdv3(float3 v, float3 dv) {
  float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
  float invsum32 = 1.0f / sqrt(sum2 * sum2 * sum2);
  float3 r;
  r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
  r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
  r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
  return r;
}

__device__ __forceinline__ m3 quat_to_rotmat_b(const v4 quat) {
  float s = rsqrtf(quat.w * quat.w + quat.x * quat.x + quat.y * quat.y + quat.z * quat.z);
  float w = quat.x * s, x = quat.y * s, y = quat.z * s, z = quat.w * s;
  m3 R;
  R[0] = v3(1.f - 2.f * (y * y + z * z), 2.f * (x * y + w * z), 2.f * (x * z - w * y));
  R[1] = v3(2.f * (x * y - w * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z + w * x));
  R[2] = v3(2.f * (x * z + w * y), 2.f * (y * z - w * x), 1.f - 2.f * (x * x + y * y));
  return R;
}

// VJP of quat_to_rotmat with respect to the *normalised* quaternion (the reference
// does not differentiate through the in-kernel normalisation).
__device__ __forceinline__ v4 quat_to_rotmat_vjp(const v4 quat, const m3 v_R) {
  float s = rsqrtf(quat.w * quat.w + quat.x * quat.x + quat.y * quat.y + quat.z * quat.z);
  float w = quat.x * s, x = quat.y * s, y = quat.z * s, z = quat.w * s;
  v4 v_quat;
  v_quat.x = 2.f * (x * (v_R[1][2] - v_R[2][1]) + y * (v_R[2][0] - v_R[0][2]) + z * (v_R[0][1] - v_R[1][0]));
  v_quat.y = 2.f * (-2.f * x * (v_R[1][1] + v_R[2][2]) + y * (v_R[0][1] + v_R[1][0]) + z * (v_R[0][2] + v_R[2][0]) +
                    w * (v_R[1][2] - v_R[2][1]));
  v_quat.z = 2.f * (x * (v_R[0][1] + v_R[1][0]) - 2.f * y * (v_R[0][0] + v_R[2][2]) + z * (v_R[1][2] + v_R[2][1]) +
                    w * (v_R[2][0] - v_R[0][2]));
  v_quat.w = 2.f * (x * (v_R[0][2] + v_R[2][0]) + y * (v_R[1][2] + v_R[2][1]) - 2.f * z * (v_R[0][0] + v_R[1][1]) +
                    w * (v_R[0][1] - v_R[1][0]));
  return v_quat;
}

__device__ __forceinline__ void zero_sh_row(float* o_sh, int M) {
  if (M == 16) {
    float4* d4 = reinterpret_cast<float4*>(o_sh);
#pragma unroll
    for (int i = 0; i < 12; i++) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (int i = 0; i < M * 3; i++) o_sh[i] = 0.f;
  }
}

// SH -> RGB backward (reference backward.cu:20-139): writes dL/dsh[M] and returns the mean3D
// gradient that flows through the view direction.
template <bool DIRECT = false>
__device__ __forceinline__ v3 sh_backward(int idx, int deg, int M, const v3* means, v3 campos, const float* shs,
                                          unsigned clamped, v3 dL_dcolor, v3* dL_dsh_out, v3 pos_direct = v3(),
                                          bool acc = false) {
  v3 pos = DIRECT ? pos_direct : means[idx];
  v3 dir_orig = pos - campos;
  v3 dir = dir_orig / length(dir_orig);
  // SH coefficients through registers: 12 x 128-bit loads / stores per surfel when M == 16
  // (192-byte rows are 16-byte aligned) instead of 48 + 48 scalar accesses.
  v3 sh[16];
  v3 dL_dsh[16];
  const float* sh_g = shs + (size_t)idx * M * 3;
  if (M == 16) {
    const float4* s4 = reinterpret_cast<const float4*>(sh_g);
    float* dst = reinterpret_cast<float*>(sh);
#pragma unroll
    for (int i = 0; i < 12; i++) {
      const float4 v = __ldg(s4 + i);
      dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (i < M) sh[i] = v3(sh_g[3 * i], sh_g[3 * i + 1], sh_g[3 * i + 2]);
      else sh[i] = v3(0.f, 0.f, 0.f);
    }
  }

  v3 dL_dRGB = dL_dcolor;
  dL_dRGB.x *= (clamped & 1u) ? 0 : 1;
  dL_dRGB.y *= (clamped & 2u) ? 0 : 1;
  dL_dRGB.z *= (clamped & 4u) ? 0 : 1;

  v3 dRGBdx(0, 0, 0), dRGBdy(0, 0, 0), dRGBdz(0, 0, 0);
  float x = dir.x, y = dir.y, z = dir.z;
#pragma unroll
  for (int i = 0; i < 16; i++) dL_dsh[i] = v3(0.f, 0.f, 0.f);

  float dRGBdsh0 = bSH_C0;
  dL_dsh[0] = dRGBdsh0 * dL_dRGB;
  if (deg > 0) {
    float dRGBdsh1 = -bSH_C1 * y;
    float dRGBdsh2 = bSH_C1 * z;
    float dRGBdsh3 = -bSH_C1 * x;
    dL_dsh[1] = dRGBdsh1 * dL_dRGB;
    dL_dsh[2] = dRGBdsh2 * dL_dRGB;
    dL_dsh[3] = dRGBdsh3 * dL_dRGB;
    dRGBdx = -bSH_C1 * sh[3];
    dRGBdy = -bSH_C1 * sh[1];
    dRGBdz = bSH_C1 * sh[2];
    if (deg > 1) {
      float xx = x * x, yy = y * y, zz = z * z;
      float xy = x * y, yz = y * z, xz = x * z;
      float dRGBdsh4 = bSH_C2[0] * xy;
      float dRGBdsh5 = bSH_C2[1] * yz;
      float dRGBdsh6 = bSH_C2[2] * (2.f * zz - xx - yy);
      float dRGBdsh7 = bSH_C2[3] * xz;
      float dRGBdsh8 = bSH_C2[4] * (xx - yy);
      dL_dsh[4] = dRGBdsh4 * dL_dRGB;
      dL_dsh[5] = dRGBdsh5 * dL_dRGB;
      dL_dsh[6] = dRGBdsh6 * dL_dRGB;
      dL_dsh[7] = dRGBdsh7 * dL_dRGB;
      dL_dsh[8] = dRGBdsh8 * dL_dRGB;
      dRGBdx += bSH_C2[0] * y * sh[4] + bSH_C2[2] * 2.f * -x * sh[6] + bSH_C2[3] * z * sh[7] +
                bSH_C2[4] * 2.f * x * sh[8];
      dRGBdy += bSH_C2[0] * x * sh[4] + bSH_C2[1] * z * sh[5] + bSH_C2[2] * 2.f * -y * sh[6] +
                bSH_C2[4] * 2.f * -y * sh[8];
      dRGBdz += bSH_C2[1] * y * sh[5] + bSH_C2[2] * 2.f * 2.f * z * sh[6] + bSH_C2[3] * x * sh[7];
      if (deg > 2) {
        float dRGBdsh9 = bSH_C3[0] * y * (3.f * xx - yy);
        float dRGBdsh10 = bSH_C3[1] * xy * z;
        float dRGBdsh11 = bSH_C3[2] * y * (4.f * zz - xx - yy);
        float dRGBdsh12 = bSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
        float dRGBdsh13 = bSH_C3[4] * x * (4.f * zz - xx - yy);
        float dRGBdsh14 = bSH_C3[5] * z * (xx - yy);
        float dRGBdsh15 = bSH_C3[6] * x * (xx - 3.f * yy);
        dL_dsh[9] = dRGBdsh9 * dL_dRGB;
        dL_dsh[10] = dRGBdsh10 * dL_dRGB;
        dL_dsh[11] = dRGBdsh11 * dL_dRGB;
        dL_dsh[12] = dRGBdsh12 * dL_dRGB;
        dL_dsh[13] = dRGBdsh13 * dL_dRGB;
        dL_dsh[14] = dRGBdsh14 * dL_dRGB;
        dL_dsh[15] = dRGBdsh15 * dL_dRGB;
        dRGBdx += (bSH_C3[0] * sh[9] * 3.f * 2.f * xy + bSH_C3[1] * sh[10] * yz + bSH_C3[2] * sh[11] * -2.f * xy +
                   bSH_C3[3] * sh[12] * -3.f * 2.f * xz + bSH_C3[4] * sh[13] * (-3.f * xx + 4.f * zz - yy) +
                   bSH_C3[5] * sh[14] * 2.f * xz + bSH_C3[6] * sh[15] * 3.f * (xx - yy));
        dRGBdy += (bSH_C3[0] * sh[9] * 3.f * (xx - yy) + bSH_C3[1] * sh[10] * xz +
                   bSH_C3[2] * sh[11] * (-3.f * yy + 4.f * zz - xx) + bSH_C3[3] * sh[12] * -3.f * 2.f * yz +
                   bSH_C3[4] * sh[13] * -2.f * xy + bSH_C3[5] * sh[14] * -2.f * yz +
                   bSH_C3[6] * sh[15] * -3.f * 2.f * xy);
        dRGBdz += (bSH_C3[1] * sh[10] * xy + bSH_C3[2] * sh[11] * 4.f * 2.f * yz +
                   bSH_C3[3] * sh[12] * 3.f * (2.f * zz - xx - yy) + bSH_C3[4] * sh[13] * 4.f * 2.f * xz +
                   bSH_C3[5] * sh[14] * (xx - yy));
      }
    }
  }
  // acc: the row already holds the sum over the earlier views of a batch (gradient accumulation in place)
  if (M == 16) {
    float4* d4 = reinterpret_cast<float4*>(dL_dsh_out);
    const float* src = reinterpret_cast<const float*>(dL_dsh);
    if (acc) {
#pragma unroll
      for (int i = 0; i < 12; i++) {
        const float4 o = d4[i];
        d4[i] = make_float4(o.x + src[4 * i], o.y + src[4 * i + 1], o.z + src[4 * i + 2], o.w + src[4 * i + 3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 12; i++) d4[i] = make_float4(src[4 * i], src[4 * i + 1], src[4 * i + 2], src[4 * i + 3]);
    }
  } else {
    float* dg = reinterpret_cast<float*>(dL_dsh_out);
#pragma unroll
    for (int i = 0; i < 16; i++)
      if (i < M) {
        if (acc) { dg[3 * i] += dL_dsh[i].x; dg[3 * i + 1] += dL_dsh[i].y; dg[3 * i + 2] += dL_dsh[i].z; }
        else { dg[3 * i] = dL_dsh[i].x; dg[3 * i + 1] = dL_dsh[i].y; dg[3 * i + 2] = dL_dsh[i].z; }
      }
    if (!acc)
      for (int i = 16; i < M; i++) { dg[3 * i] = 0.f; dg[3 * i + 1] = 0.f; dg[3 * i + 2] = 0.f; }
  }
  v3 dL_ddir(dot(dRGBdx, dL_dRGB), dot(dRGBdy, dL_dRGB), dot(dRGBdz, dL_dRGB));
  float3 dL_dmean = dnormvdv3(float3{dir_orig.x, dir_orig.y, dir_orig.z}, float3{dL_ddir.x, dL_ddir.y, dL_ddir.z});
  return v3(dL_dmean.x, dL_dmean.y, dL_dmean.z);
}

// SQ = block-level mode: the surfel is regenerated from the superquadric parameters (as in the forward
// kernel); dL_dscales then receives the gradient w.r.t. the LOG scale (= dL/dscale * scale), which is what the
// superquadric backward kernels consume together with dL_dmean3D, dL_drots and dL_dopacity.
template <bool SQ>
__global__ void __launch_bounds__(256) preprocess_bwd_kernel(PreprocessBwdArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.P) return;
  const int M = a.M;
  const bool visible = a.radii[idx] > 0;

  float* o_m2d = a.dL_dmean2D + (size_t)idx * 3;
  float* o_col = a.dL_dcolors + (size_t)idx * 3;
  float* o_m3d = a.dL_dmean3D + (size_t)idx * 3;
  float* o_T = a.dL_dtransMat + (size_t)idx * 9;
  float* o_sh = a.dL_dsh ? a.dL_dsh + (size_t)idx * M * 3 : nullptr;

  // Accumulate mode (a data-parallel batch: a rank renders several views and all-reduces ONE gradient bucket): the
  // five parameter gradients (mean3D, SH, opacity, scales, rotations) are added to what the earlier views of the
  // batch left in place; the per-view outputs (mean2D, colours, transMat) are still overwritten.
  // Block-level mode: only the SH rows are parameters (and accumulate); mean3D / scales / rotations / opacity are
  // per-view scratch that the face -> vertex -> block reduction consumes right after this kernel: always overwritten.
  const bool acc_sh = a.accumulate != 0;
  const bool acc = acc_sh && !SQ;
  if (!visible) {
    o_m2d[0] = o_m2d[1] = o_m2d[2] = 0.f;
    o_col[0] = o_col[1] = o_col[2] = 0.f;
    for (int i = 0; i < 9; i++) o_T[i] = 0.f;
    if (o_sh && !acc_sh) zero_sh_row(o_sh, M);
    if (acc) return;   // nothing to add
    a.dL_dopacity[idx] = 0.f;
    o_m3d[0] = o_m3d[1] = o_m3d[2] = 0.f;
    if (a.dL_dscales) { a.dL_dscales[idx * 2] = 0.f; a.dL_dscales[idx * 2 + 1] = 0.f; }
    if (a.dL_drots) { for (int i = 0; i < 4; i++) a.dL_drots[idx * 4 + i] = 0.f; }
    return;
  }
  auto put = [acc](float* dst, float v) { *dst = acc ? *dst + v : v; };

  // accumulated per-surfel gradients from the render pass
  const float4* gq = reinterpret_cast<const float4*>(a.grad + (size_t)idx * GRAD_FLOATS);
  const float4 g0 = gq[0], g1 = gq[1], g2 = gq[2], g3 = gq[3], g4 = gq[4];
  float gT[9] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w, g2.x};
  float3 dL_dmean2D = {g2.y, g2.z, 0.f};
  const float g_opacity = g2.w;
  const float3 dL_dnormal = {g3.x, g3.y, g3.z};
  v3 dL_dcolor = v3(g4.x, g4.y, g4.z);

  const int W = int(a.focal_x * a.tan_fovx * 2);
  const int H = int(a.focal_y * a.tan_fovy * 2);
  const bool precomp = !SQ && (a.scales == nullptr);

  const float4* rec = a.rec + (size_t)idx * REC_QUADS;
  const float4 q0 = rec[0], q1 = rec[1], q2 = rec[2], q4 = rec[4];

  m3 T;
  float3 normal;
  m3x4 Pm;
  m3 R;
  float3 p_orig;
  v4 rot;
  v2 scale;
  if (precomp) {
    T[0] = v3(q0.x, q0.y, q0.z);
    T[1] = v3(q1.x, q1.y, q1.z);
    T[2] = v3(q2.x, q2.y, q2.z);
    normal = {0.0f, 0.0f, 0.0f};
  } else {
    if (SQ) {
      const SqSurfel sf = sq_generate(a.sq, a.sq_vertices, idx);
      p_orig = sf.mean;
      rot = v4(sf.quat.x, sf.quat.y, sf.quat.z, sf.quat.w);
      scale = v2(sf.scale.x, sf.scale.y);
    } else {
      p_orig = {a.means3D[3 * idx], a.means3D[3 * idx + 1], a.means3D[3 * idx + 2]};
      rot = ((const v4*)a.rotations)[idx];
      scale = ((const v2*)a.scales)[idx];
    }
    R = quat_to_rotmat_b(rot);
    m3 S = diag3(1.f);
    S[0][0] = 1.0f * scale.x;  // scale_modifier deliberately ignored (backward.cu:481)
    S[1][1] = 1.0f * scale.y;
    m3 L = R * S;
    m3x4 Mm = make_m3x4(v4(L[0], 0.0f), v4(L[1], 0.0f), v4(p_orig.x, p_orig.y, p_orig.z, 1.f));
    m4 world2ndc;
    world2ndc[0] = v4(a.projmatrix[0], a.projmatrix[4], a.projmatrix[8], a.projmatrix[12]);
    world2ndc[1] = v4(a.projmatrix[1], a.projmatrix[5], a.projmatrix[9], a.projmatrix[13]);
    world2ndc[2] = v4(a.projmatrix[2], a.projmatrix[6], a.projmatrix[10], a.projmatrix[14]);
    world2ndc[3] = v4(a.projmatrix[3], a.projmatrix[7], a.projmatrix[11], a.projmatrix[15]);
    m3x4 ndc2pix = make_m3x4(v4((float)(float(W) / 2.0), 0.0f, 0.0f, (float)(float(W - 1) / 2.0)),
                             v4(0.0f, (float)(float(H) / 2.0), 0.0f, (float)(float(H - 1) / 2.0)),
                             v4(0.0f, 0.0f, 0.0f, 1.0f));
    Pm = world2ndc * ndc2pix;
    T = transpose(Mm) * Pm;
    const float* vm = a.viewmatrix;
    normal = {vm[0] * L[2].x + vm[4] * L[2].y + vm[8] * L[2].z, vm[1] * L[2].x + vm[5] * L[2].y + vm[9] * L[2].z,
              vm[2] * L[2].x + vm[6] * L[2].y + vm[10] * L[2].z};
  }

  m3 dL_dT;
  dL_dT[0] = v3(gT[0], gT[1], gT[2]);
  dL_dT[1] = v3(gT[3], gT[4], gT[5]);
  dL_dT[2] = v3(gT[6], gT[7], gT[8]);

  bool T_written = false;
  if (dL_dmean2D.x != 0 || dL_dmean2D.y != 0) {
    v3 t_vec = v3(9.0f, 9.0f, -1.0f);
    float d = dot(t_vec, T[2] * T[2]);
    v3 f_vec = t_vec * (1.0f / d);
    v3 dL_dT0 = dL_dmean2D.x * f_vec * T[2];
    v3 dL_dT1 = dL_dmean2D.y * f_vec * T[2];
    v3 dL_dT3 = dL_dmean2D.x * f_vec * T[0] + dL_dmean2D.y * f_vec * T[1];
    v3 dL_df = dL_dmean2D.x * T[0] * T[2] + dL_dmean2D.y * T[1] * T[2];
    float dL_dd = dot(dL_df, f_vec) * (-1.0 / d);
    v3 dd_dT3 = t_vec * T[2] * 2.0f;
    dL_dT3 += dL_dd * dd_dT3;
    dL_dT[0] += dL_dT0;
    dL_dT[1] += dL_dT1;
    dL_dT[2] += dL_dT3;
    T_written = precomp;
  }

  // dL_dtransMat output: the reference only writes the AABB-augmented value back on
  // the precomputed-T path (backward.cu:533-544); otherwise the caller sees the
  // render-accumulated one.
  if (T_written) {
    o_T[0] = dL_dT[0].x; o_T[1] = dL_dT[0].y; o_T[2] = dL_dT[0].z;
    o_T[3] = dL_dT[1].x; o_T[4] = dL_dT[1].y; o_T[5] = dL_dT[1].z;
    o_T[6] = dL_dT[2].x; o_T[7] = dL_dT[2].y; o_T[8] = dL_dT[2].z;
  } else {
    for (int i = 0; i < 9; i++) o_T[i] = gT[i];
  }
  // densification hack (backward.cu:626-629) reads dL_dtransMat as it stands in
  // memory: render-accumulated, AABB-augmented only on the precomputed-T path.
  const float depth = q2.z;
  const float hack_x = o_T[2] * depth * 0.5 * float(W);
  const float hack_y = o_T[5] * depth * 0.5 * float(H);

  v3 dmean = v3(0.f, 0.f, 0.f);
  if (!precomp) {
    m3x4 dL_dM = Pm * transpose(dL_dT);
    const float* vm = a.viewmatrix;
    float3 dL_dtn = {vm[0] * dL_dnormal.x + vm[1] * dL_dnormal.y + vm[2] * dL_dnormal.z,
                     vm[4] * dL_dnormal.x + vm[5] * dL_dnormal.y + vm[6] * dL_dnormal.z,
                     vm[8] * dL_dnormal.x + vm[9] * dL_dnormal.y + vm[10] * dL_dnormal.z};
    float3 p_view = {vm[0] * p_orig.x + vm[4] * p_orig.y + vm[8] * p_orig.z + vm[12],
                     vm[1] * p_orig.x + vm[5] * p_orig.y + vm[9] * p_orig.z + vm[13],
                     vm[2] * p_orig.x + vm[6] * p_orig.y + vm[10] * p_orig.z + vm[14]};
    float cosv = -(p_view.x * normal.x + p_view.y * normal.y + p_view.z * normal.z);
    float multiplier = cosv > 0 ? 1 : -1;
    dL_dtn = {multiplier * dL_dtn.x, multiplier * dL_dtn.y, multiplier * dL_dtn.z};

    m3 dL_dRS = make_m3(xyz(dL_dM[0]), xyz(dL_dM[1]), v3(dL_dtn.x, dL_dtn.y, dL_dtn.z));
    m3 dL_dR = make_m3(dL_dRS[0] * v3(scale.x, scale.x, scale.x), dL_dRS[1] * v3(scale.y, scale.y, scale.y),
                       dL_dRS[2]);
    v4 dq = quat_to_rotmat_vjp(rot, dL_dR);
    put(&a.dL_drots[idx * 4 + 0], dq.x);
    put(&a.dL_drots[idx * 4 + 1], dq.y);
    put(&a.dL_drots[idx * 4 + 2], dq.z);
    put(&a.dL_drots[idx * 4 + 3], dq.w);
    put(&a.dL_dscales[idx * 2 + 0], (float)dot(dL_dRS[0], R[0]) * (SQ ? scale.x : 1.0f));
    put(&a.dL_dscales[idx * 2 + 1], (float)dot(dL_dRS[1], R[1]) * (SQ ? scale.y : 1.0f));
    dmean = xyz(dL_dM[2]);
  } else if (!acc) {
    if (a.dL_dscales) { a.dL_dscales[idx * 2] = 0.f; a.dL_dscales[idx * 2 + 1] = 0.f; }
    if (a.dL_drots) { for (int i = 0; i < 4; i++) a.dL_drots[idx * 4 + i] = 0.f; }
  }

  // SH backward (backward.cu:20-139), incl. view-direction -> mean3D path
  if (a.shs) {
    dmean += sh_backward<SQ>(idx, a.D, M, (const v3*)a.means3D, *(const v3*)a.cam_pos, a.shs, __float_as_uint(q4.w),
                             dL_dcolor, (v3*)o_sh, v3(p_orig.x, p_orig.y, p_orig.z), acc_sh);
  } else if (o_sh && !acc_sh) {
    for (int i = 0; i < M * 3; i++) o_sh[i] = 0.f;
  }

  put(&o_m3d[0], dmean.x); put(&o_m3d[1], dmean.y); put(&o_m3d[2], dmean.z);
  o_col[0] = dL_dcolor.x; o_col[1] = dL_dcolor.y; o_col[2] = dL_dcolor.z;
  put(&a.dL_dopacity[idx], g_opacity);
  o_m2d[0] = hack_x;
  o_m2d[1] = hack_y;
  o_m2d[2] = 0.f;
}

// =============================================================================
// `_part` fork: computeAABB backward (DSRP/cuda_rasterizer/backward.cu:621-671) fused with
// preprocessCUDA backward (:555-619) and the computeTransMat VJP (:473-551).  Quirks kept:
// the AABB-centre term is always added (no dL_dmean2D != 0 test), dL_dtransMat returned to
// the caller includes it, mean2D hack = dL_dT[2|5] * z * {fx*tanx | fy*tany}.
// =============================================================================
__global__ void __launch_bounds__(256) preprocess_bwd_part_kernel(PreprocessBwdArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.P) return;
  const int M = a.M;
  float* o_m2d = a.dL_dmean2D + (size_t)idx * 3;
  float* o_col = a.dL_dcolors + (size_t)idx * 3;
  float* o_m3d = a.dL_dmean3D + (size_t)idx * 3;
  float* o_T = a.dL_dtransMat + (size_t)idx * 9;
  float* o_sh = a.dL_dsh ? a.dL_dsh + (size_t)idx * M * 3 : nullptr;
  if (!(a.radii[idx] > 0)) {
    o_m2d[0] = o_m2d[1] = o_m2d[2] = 0.f;
    o_col[0] = o_col[1] = o_col[2] = 0.f;
    a.dL_dopacity[idx] = 0.f;
    o_m3d[0] = o_m3d[1] = o_m3d[2] = 0.f;
    for (int i = 0; i < 9; i++) o_T[i] = 0.f;
    if (o_sh) zero_sh_row(o_sh, M);
    a.dL_dscales[idx * 2] = 0.f; a.dL_dscales[idx * 2 + 1] = 0.f;
    for (int i = 0; i < 4; i++) a.dL_drots[idx * 4 + i] = 0.f;
    return;
  }
  const float4* gq = reinterpret_cast<const float4*>(a.grad + (size_t)idx * GRAD_FLOATS);
  const float4 g0 = gq[0], g1 = gq[1], g2 = gq[2], g3 = gq[3], g4 = gq[4];
  float gT[9] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w, g2.x};
  const float3 dL_dmean2D = {g2.y, g2.z, 0.f};
  const float g_opacity = g2.w;
  const float dL_dnormal3D[3] = {g3.x, g3.y, g3.z};
  v3 dL_dcolor = v3(g4.x, g4.y, g4.z);

  const float4* rec = a.rec + (size_t)idx * REC_QUADS;
  const float4 q0 = rec[0], q1 = rec[1], q2 = rec[2], q4 = rec[4];

  // ---- computeAABB backward ----
  {
    m4x3 T;
    T[0] = v3(q0.x, q0.y, q0.z); T[1] = v3(q1.x, q1.y, q1.z); T[2] = v3(q2.x, q2.y, q2.z); T[3] = T[2];
    float d = dot(v3(1.0f, 1.0f, -1.0f), T[3] * T[3]);
    v3 f = v3(1.0f, 1.0f, -1.0f) * (1.0f / d);
    v3 dL_dT0 = dL_dmean2D.x * f * T[3];
    v3 dL_dT1 = dL_dmean2D.y * f * T[3];
    v3 dL_dT3 = dL_dmean2D.x * f * T[0] + dL_dmean2D.y * f * T[1];
    v3 dL_df = (dL_dmean2D.x * T[0] * T[3]) + (dL_dmean2D.y * T[1] * T[3]);
    float dL_dd = dot(dL_df, f) * (-1.0 / d);
    v3 dd_dT3 = v3(1.0f, 1.0f, -1.0f) * T[3] * 2.0f;
    dL_dT3 += dL_dd * dd_dT3;
    gT[0] += dL_dT0.x; gT[1] += dL_dT0.y; gT[2] += dL_dT0.z;
    gT[3] += dL_dT1.x; gT[4] += dL_dT1.y; gT[5] += dL_dT1.z;
    gT[6] += dL_dT3.x; gT[7] += dL_dT3.y; gT[8] += dL_dT3.z;
  }
  for (int i = 0; i < 9; i++) o_T[i] = gT[i];
  const float Wh = a.focal_x * a.tan_fovx, Hh = a.focal_y * a.tan_fovy;
  const float z = q2.z;
  o_m2d[0] = gT[2] * z * Wh;
  o_m2d[1] = gT[5] * z * Hh;
  o_m2d[2] = 0.f;

  // ---- computeTransMat VJP ----
  const float* viewmat = a.viewmatrix;
  const float4 intrins = {a.focal_x, a.focal_y, a.focal_x * a.tan_fovx, a.focal_y * a.tan_fovy};
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
  const v3 p_world = v3(a.means3D[3 * idx], a.means3D[3 * idx + 1], a.means3D[3 * idx + 2]);
  const v4 quat = ((const v4*)a.rotations)[idx];
  const v2 scale = ((const v2*)a.scales)[idx];
  m3 S = diag3(1.f);
  S[0][0] = 1.0f * scale.x; S[1][1] = 1.0f * scale.y; S[2][2] = 1.0f;
  m3 R = quat_to_rotmat_b(quat);
  m3 RS = R * S;
  v3 p_view = W * p_world + cam_pos;
  m3 Mm = make_m3(W * RS[0], W * RS[1], p_view);

  m4x3 dL_dT;
  dL_dT[0] = v3(gT[0], gT[1], gT[2]);
  dL_dT[1] = v3(gT[3], gT[4], gT[5]);
  dL_dT[2] = v3(gT[6], gT[7], gT[8]);
  dL_dT[3] = v3(0.0f, 0.0f, 0.0f);
  m3x4 dL_dM_aug = transpose(Pm) * transpose(dL_dT);
  m3 dL_dM = make_m3(xyz(dL_dM_aug[0]), xyz(dL_dM_aug[1]), xyz(dL_dM_aug[2]));
  m3 W_t = transpose(W);
  m3 dL_dRS = W_t * dL_dM;
  v3 dL_dRS0 = dL_dRS[0];
  v3 dL_dRS1 = dL_dRS[1];
  v3 dL_dpw = dL_dRS[2];
  v3 dL_dtn = W_t * v3(dL_dnormal3D[0], dL_dnormal3D[1], dL_dnormal3D[2]);
  v3 tn = W * R[2];
  float cosv = dot(-tn, Mm[2]);
  float multiplier = cosv > 0 ? 1 : -1;
  dL_dtn *= multiplier;
  m3 dL_dR = make_m3(dL_dRS0 * v3(scale.x, scale.x, scale.x), dL_dRS1 * v3(scale.y, scale.y, scale.y), dL_dtn);
  v4 dq = quat_to_rotmat_vjp(quat, dL_dR);
  a.dL_drots[idx * 4 + 0] = dq.x; a.dL_drots[idx * 4 + 1] = dq.y;
  a.dL_drots[idx * 4 + 2] = dq.z; a.dL_drots[idx * 4 + 3] = dq.w;
  a.dL_dscales[idx * 2 + 0] = (float)dot(dL_dRS0, R[0]);
  a.dL_dscales[idx * 2 + 1] = (float)dot(dL_dRS1, R[1]);
  v3 dmean = dL_dpw;

  if (a.shs) {
    v3 extra = sh_backward(idx, a.D, M, (const v3*)a.means3D, *(const v3*)a.cam_pos, a.shs, __float_as_uint(q4.w),
                           dL_dcolor, (v3*)o_sh);
    dmean += extra;
  } else if (o_sh) {
    for (int i = 0; i < M * 3; i++) o_sh[i] = 0.f;
  }
  o_m3d[0] = dmean.x; o_m3d[1] = dmean.y; o_m3d[2] = dmean.z;
  o_col[0] = dL_dcolor.x; o_col[1] = dL_dcolor.y; o_col[2] = dL_dcolor.z;
  a.dL_dopacity[idx] = g_opacity;
}

void launch_preprocess_bwd_part(const PreprocessBwdArgs& a, cudaStream_t s) {
  if (a.P <= 0) return;
  preprocess_bwd_part_kernel<<<(a.P + 255) / 256, 256, 0, s>>>(a);
  count_launch();
}

void launch_preprocess_bwd(const PreprocessBwdArgs& a, cudaStream_t s) {
  if (a.P <= 0) return;
  if (a.use_sq) preprocess_bwd_kernel<true><<<(a.P + 255) / 256, 256, 0, s>>>(a);
  else preprocess_bwd_kernel<false><<<(a.P + 255) / 256, 256, 0, s>>>(a);
  count_launch();
}

}  // namespace pgs
