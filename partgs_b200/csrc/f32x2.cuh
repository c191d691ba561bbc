// Packed FP32 pairs for sm_100a: two IEEE-rn single-precision lanes in one 64-bit register pair, computed by the
// Blackwell FFMA2 / FMUL2 / FADD2 instructions (PTX fma.rn.f32x2 / mul.rn.f32x2 / add.rn.f32x2).  One issue slot does
// the work of two scalar instructions, which is the lever for the issue-bound render kernels: they evaluate TWO
// staged surfels (A = low half, B = high half) per loop iteration for the same pixel.
//
// Each half is rounded exactly like the scalar __fmaf_rn / __fmul_rn / __fadd_rn, so pinned rounding sequences
// (frag_math.cuh) stay bit-identical when written with these.  ptxas folds `pk(-x, -y)` into an operand negation
// and `bc(s)` (both halves the same scalar) into the FFMA2 scalar-broadcast operand form, so neither costs an
// instruction; lo()/hi() are free (a pair is two ordinary registers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pgs {

#ifndef PGS_EMU
struct P2 {
  unsigned long long v;
};
__device__ __forceinline__ P2 pk(float a, float b) {
  P2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float lo(P2 p) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v));
  (void)b;
  return a;
}
__device__ __forceinline__ float hi(P2 p) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v));
  (void)a;
  return b;
}
__device__ __forceinline__ P2 fma2(P2 a, P2 b, P2 c) {
  P2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ P2 mul2(P2 a, P2 b) {
  P2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ P2 add2(P2 a, P2 b) {
  P2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
#else  // CPU lock-step emulator (tests/cuda_emu): the same operations on two plain floats, each rounded once
struct P2 {
  float a, b;
};
inline P2 pk(float a, float b) { return P2{a, b}; }
inline float lo(P2 p) { return p.a; }
inline float hi(P2 p) { return p.b; }
inline P2 fma2(P2 a, P2 b, P2 c) { return P2{fmaf(a.a, b.a, c.a), fmaf(a.b, b.b, c.b)}; }
inline P2 mul2(P2 a, P2 b) { return P2{__fmul_rn(a.a, b.a), __fmul_rn(a.b, b.b)}; }
inline P2 add2(P2 a, P2 b) { return P2{__fadd_rn(a.a, b.a), __fadd_rn(a.b, b.b)}; }
#endif

__device__ __forceinline__ P2 bc(float s) { return pk(s, s); }                    // broadcast
__device__ __forceinline__ P2 neg2(P2 a) { return pk(-lo(a), -hi(a)); }            // folded into the consumer
__device__ __forceinline__ P2 sub2(P2 a, P2 b) { return add2(a, neg2(b)); }        // a - b
__device__ __forceinline__ P2 fms2(P2 a, P2 b, P2 c) { return fma2(a, b, neg2(c)); }  // a*b - c
// 64-bit shuffle of a pair (two SHFL)
__device__ __forceinline__ P2 shfl_xor2(P2 a, int m) {
  return pk(__shfl_xor_sync(0xffffffffu, lo(a), m), __shfl_xor_sync(0xffffffffu, hi(a), m));
}

// A 16-byte shared-memory quad holding two pairs: {f0.A, f0.B, f1.A, f1.B}
struct __align__(16) Q2 {
  P2 x, y;
};

// ---- explicit shared-memory accesses for the hot loops ----------------------------------------------------------
// Per-lane shared-memory addresses are loop invariants; written as C++ pointers, ptxas re-derives them from
// %tid inside the loop (a dozen integer instructions per iteration) to save registers.  The hot loops therefore
// keep them as opaque 32-bit shared-window addresses and access memory with a compile-time byte offset.
#ifndef PGS_EMU
typedef uint32_t saddr_t;
__device__ __forceinline__ saddr_t smem_addr(const void* p) {
  saddr_t a = (saddr_t)__cvta_generic_to_shared(p);
  asm volatile("" : "+r"(a));  // opaque: keep it in a register
  return a;
}
__device__ __forceinline__ uint32_t opaque_u32(uint32_t v) {
  asm volatile("" : "+r"(v));
  return v;
}
template <int OFF> __device__ __forceinline__ Q2 lds_q2(saddr_t a) {
  Q2 r;
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2+%3];" : "=l"(r.x.v), "=l"(r.y.v) : "r"(a), "n"(OFF) : "memory");
  return r;
}
template <int OFF> __device__ __forceinline__ uint4 lds_u4(saddr_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4+%5];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "r"(a), "n"(OFF)
               : "memory");
  return r;
}
template <int OFF> __device__ __forceinline__ float4 lds_f4(saddr_t a) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "r"(a), "n"(OFF)
               : "memory");
  return r;
}
template <int OFF> __device__ __forceinline__ uint2 lds_u2(saddr_t a) {
  uint2 r;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2+%3];" : "=r"(r.x), "=r"(r.y) : "r"(a), "n"(OFF) : "memory");
  return r;
}
template <int OFF> __device__ __forceinline__ uint32_t lds_u32(saddr_t a) {
  uint32_t r;
  asm volatile("ld.shared.b32 %0, [%1+%2];" : "=r"(r) : "r"(a), "n"(OFF) : "memory");
  return r;
}
template <typename T> __device__ __forceinline__ T* opaque_ptr(T* p) {
  asm volatile("" : "+l"(p));
  return p;
}
template <int OFF> __device__ __forceinline__ void sts_p2(saddr_t a, P2 v) {
  asm volatile("st.shared.v2.f32 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "f"(lo(v)), "f"(hi(v)) : "memory");
}
template <int OFF> __device__ __forceinline__ void sts_f32(saddr_t a, float v) {
  asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(a), "n"(OFF), "f"(v) : "memory");
}
#else
typedef unsigned char* saddr_t;
inline saddr_t smem_addr(const void* p) { return (saddr_t)p; }
inline uint32_t opaque_u32(uint32_t v) { return v; }
template <int OFF> inline Q2 lds_q2(saddr_t a) { return *reinterpret_cast<const Q2*>(a + OFF); }
template <int OFF> inline uint4 lds_u4(saddr_t a) { return *reinterpret_cast<const uint4*>(a + OFF); }
template <int OFF> inline float4 lds_f4(saddr_t a) { return *reinterpret_cast<const float4*>(a + OFF); }
template <int OFF> inline uint2 lds_u2(saddr_t a) { return *reinterpret_cast<const uint2*>(a + OFF); }
template <int OFF> inline uint32_t lds_u32(saddr_t a) { return *reinterpret_cast<const uint32_t*>(a + OFF); }
template <typename T> inline T* opaque_ptr(T* p) { return p; }
template <int OFF> inline void sts_p2(saddr_t a, P2 v) { *reinterpret_cast<P2*>(a + OFF) = v; }
template <int OFF> inline void sts_f32(saddr_t a, float v) { *reinterpret_cast<float*>(a + OFF) = v; }
#endif

}  // namespace pgs
