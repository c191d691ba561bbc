// Shared constants, HBM data layouts and small device helpers for the surfel
// rasteriser kernels (sm_100a).  Product code: no torch, no CPU fallback.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace pgs {

// ---- tile geometry (reference: cuda_rasterizer/config.h:15-17) --------------
constexpr int TILE_X = 16;
constexpr int TILE_Y = 16;
constexpr int TILE_PIX = TILE_X * TILE_Y;  // 256 threads = 8 warps per tile
// Each warp owns an 8x4 pixel footprint inside the 16x16 tile (compact footprints
// make the per-warp surfel culling effective).
constexpr int WARP_FX = 8;
constexpr int WARP_FY = 4;

// ---- reference constants (base fork: cuda_rasterizer/auxiliary.h:37-40) -----
#define PGS_NEAR_N 0.2f
#define PGS_FAR_N 100.0f
#define PGS_FILTER_SIZE 0.707106f
#define PGS_FILTER_INV_SQUARE 2.0f

// aux-map channel offsets (auxiliary.h:23-27)
constexpr int DEPTH_OFFSET = 0;
constexpr int ALPHA_OFFSET = 1;
constexpr int NORMAL_OFFSET = 2;
constexpr int MIDDEPTH_OFFSET = 5;
constexpr int DISTORTION_OFFSET = 6;
constexpr int MEDIAN_WEIGHT_OFFSET = 7;  // _part fork only

// ---- per-surfel record written by preprocess, gathered by render ------------
// 5 x float4 = 80 B, 16-B aligned so that every field moves as one 128-bit load.
//   q0 = {Tu.x, Tu.y, Tu.z, xy.x}
//   q1 = {Tv.x, Tv.y, Tv.z, xy.y}
//   q2 = {Tw.x, Tw.y, Tw.z, opacity}
//   q3 = {n.x,  n.y,  n.z,  depth(view z)}
//   q4 = {r, g, b, bitcast(clamped mask)}
constexpr int REC_QUADS = 5;
constexpr int REC_FLOATS = REC_QUADS * 4;

// Per-surfel cull record (3 x float4 = 48 B), written by preprocess, read by the render kernels:
//   c0 = conservative screen box {x0, y0, x1, y1} outside of which no pixel can pass alpha >= 1/255
//   c1 = {a, b, c, d}, c2 = {e, g, cx, cy}: the conic  f(X,Y) = aX^2 + 2bXY + cY^2 + 2dX + 2eY + g
//        (X = x - cx, Y = y - cy) whose non-positive set is exactly {rho3d <= c^2}; a pixel can
//        only be blended if f <= 0 there or it lies within LOWPASS_RADIUS of (cx, cy).
constexpr int CULL_QUADS = 3;
#define PGS_LOWPASS_RADIUS 2.9f

// Per-surfel gradient accumulator filled by the backward render kernel.
//   [0..8] dL/dT (Tu,Tv,Tw)  [9,10] dL/dmean2D.xy  [11] dL/dopacity(G*dL_dalpha)
//   [12..14] dL/dnormal(view) [15] pad  [16..18] dL/dcolor [19] pad
constexpr int GRAD_FLOATS = 20;

// ---- helpers ----------------------------------------------------------------
// Function attributes (opt-in dynamic shared memory) are per device: `flags` is a per-kernel array of 64
// "already set on device d" markers.  Returns true exactly once per device.
inline bool first_use_on_device(bool (&flags)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <typename T>
inline void carve(char*& p, T*& out, size_t count) {
  size_t off = align_up(reinterpret_cast<size_t>(p), 256);
  out = reinterpret_cast<T*>(off);
  p = reinterpret_cast<char*>(out + count);
}

#ifndef PGS_EMU
__device__ __forceinline__ unsigned lane_id() {
  unsigned r;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(r));
  return r;
}

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// fire-and-forget float add into global memory (RED.ADD.F32: no return value, no generic-address check)
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

// 16-byte async global->shared copy (LDGSTS).
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
#else
// CPU lock-step emulator (tests/cuda_emu): the same helpers without PTX.  Test builds only — nvcc never defines
// PGS_EMU.  cp.async completes immediately, so the emulator checks data flow, not the wait_group placement.
inline unsigned lane_id() { return emu::t_lane; }
inline float4 ldg4(const float4* p) { return *p; }
inline void red_add_f32(float* addr, float v) { atomicAdd(addr, v); }
inline void cp_async16(void* smem, const void* gmem) { memcpy(smem, gmem, 16); }
inline void cp_async_commit() {}
template <int N> inline void cp_async_wait() {}
#endif

}  // namespace pgs
