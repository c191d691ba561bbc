// Per-fragment (pixel x surfel) arithmetic shared by the four render kernels, with every
// rounding step pinned.
//
// The reference writes this math as plain C (cuda_rasterizer/forward.cu:344-387,
// backward.cu:282-313 and the `_part` twins) and lets nvcc/ptxas contract a*b+c into FMAs.
// Which products get fused decides single bits of rho/alpha, and a fragment sitting on the
// `alpha < 1/255` or `T(1-alpha) < 1e-4` threshold then flips between blended and skipped.
// To stay bit-identical to the reference build *independently of how the surrounding kernel
// is restructured*, the contraction pattern observed in the reference SASS (sm_100a, nvcc
// 12.9) is written out here with explicit __fmaf_rn / __fmul_rn / __fadd_rn:
//   k.c   = fma(px, Tw.c, -Tu.c)                 l.c = fma(py, Tw.c, -Tv.c)
//   p     = k x l with  a*b - c*d -> fma(a, b, -(c*d))
//   s     = p.xy / p.z (IEEE division)
//   rho3d = fma(s.x, s.x, s.y*s.y)
//   rho2d = 2 * fma(d.y, d.y, d.x*d.x)                        (base fork)
//         = float(double(fma(d.x, d.x, d.y*d.y)) * (1/0.7071067811865476^2))   (`_part` fork)
//   depth = Tw.z + fma(Tw.x, s.x, Tw.y*s.y)   when rho3d <= rho2d, else Tw.z
// Forward and backward use the same routine, so the backward replay takes exactly the
// forward's decisions.
#pragma once
#include <cuda_runtime.h>

namespace pgs {

// MUFU approximations (~1 ulp) for the backward pass, where no decision depends on the values.
#ifndef PGS_EMU
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#else  // CPU lock-step emulator (tests/cuda_emu): exact stand-ins for the MUFU approximations
inline float rcp_approx(float x) { return 1.0f / x; }
inline float ex2_approx(float x) { return exp2f(x); }
#endif

struct FragGeom {
  float3 k, l, p;
  float2 s, d;
  float rho3d, rho2d, depth;
};

// Returns false when the two pixel planes do not intersect the splat plane in a point
// (p.z == 0; the reference `continue`s).
template <bool PART>
__device__ __forceinline__ bool frag_geometry(const float2 pixf, const float3 Tu, const float3 Tv, const float3 Tw,
                                              const float2 xy, FragGeom& f) {
  f.k = make_float3(__fmaf_rn(pixf.x, Tw.x, -Tu.x), __fmaf_rn(pixf.x, Tw.y, -Tu.y), __fmaf_rn(pixf.x, Tw.z, -Tu.z));
  f.l = make_float3(__fmaf_rn(pixf.y, Tw.x, -Tv.x), __fmaf_rn(pixf.y, Tw.y, -Tv.y), __fmaf_rn(pixf.y, Tw.z, -Tv.z));
  f.p.x = __fmaf_rn(f.k.y, f.l.z, -__fmul_rn(f.k.z, f.l.y));
  f.p.y = __fmaf_rn(f.k.z, f.l.x, -__fmul_rn(f.k.x, f.l.z));
  f.p.z = __fmaf_rn(f.k.x, f.l.y, -__fmul_rn(f.k.y, f.l.x));
  const bool ok = !(f.p.z == 0.0f);  // branch-free: a zero p.z only produces inf/nan that the caller discards
  f.s = make_float2(__fdiv_rn(f.p.x, f.p.z), __fdiv_rn(f.p.y, f.p.z));
  f.rho3d = __fmaf_rn(f.s.x, f.s.x, __fmul_rn(f.s.y, f.s.y));
  f.d = make_float2(__fsub_rn(xy.x, pixf.x), __fsub_rn(xy.y, pixf.y));
  if (PART) {
    const float dd = __fmaf_rn(f.d.x, f.d.x, __fmul_rn(f.d.y, f.d.y));
    // reference: (float)((double)dd * (1 / 0.7071067811865476^2)).  The double constant is 2 (1 +- 2.2e-16), 2 dd is
    // a float, and the double product lies 2.2e-16 (relative) from it — eight orders of magnitude inside the
    // rounding interval of that float: the conversion returns 2 dd, always.  No double arithmetic needed.
    f.rho2d = __fadd_rn(dd, dd);
  } else {
    const float dd = __fmaf_rn(f.d.y, f.d.y, __fmul_rn(f.d.x, f.d.x));
    f.rho2d = __fadd_rn(dd, dd);  // FilterInvSquare = 2
  }
  f.depth = (f.rho3d <= f.rho2d) ? __fadd_rn(Tw.z, __fmaf_rn(Tw.x, f.s.x, __fmul_rn(Tw.y, f.s.y))) : Tw.z;
  return ok;
}

// alpha of a fragment: min(0.99, opacity * exp(-rho/2)).  `G` returns the Gaussian weight.
__device__ __forceinline__ float frag_alpha(float rho3d, float rho2d, float opa, float& power, float& G) {
  const float rho = fminf(rho3d, rho2d);
  power = __fmul_rn(-0.5f, rho);
  G = expf(power);
  return fminf(0.99f, __fmul_rn(opa, G));
}

// Forward-side evaluation of one fragment up to the `alpha < 1/255` test: true when the
// reference would go on to the transmittance test / blending (forward.cu:350-383).  Does not
// depend on the pixel's running state, so several fragments can be evaluated ahead of the blend.
template <bool PART>
__device__ __forceinline__ bool frag_eval(const float2 pixf, const float4 q0, const float4 q1, const float4 q2,
                                          float& alpha, float& depth) {
  FragGeom f;
  bool ok = frag_geometry<PART>(pixf, make_float3(q0.x, q0.y, q0.z), make_float3(q1.x, q1.y, q1.z),
                                make_float3(q2.x, q2.y, q2.z), make_float2(q0.w, q1.w), f);
  depth = f.depth;
  // `_part` compares in double: (double)depth < 0.2.  The largest float below 0.2 (double) is the predecessor of
  // 0.2f, and 0.2f itself lies above 0.2 (double): as a float comparison this is depth < 0.2f, exactly.
  if (PART) ok = ok && !(depth < 0.2f);
  else ok = ok && !(depth < 0.2f);
  float power, G;
  alpha = frag_alpha(f.rho3d, f.rho2d, q2.w, power, G);
  return ok && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
}

}  // namespace pgs
