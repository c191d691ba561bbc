// Gradient all-reduce over NVLink peer memory as ONE kernel (sm_100a).
//
// The five parameter gradients of a rank live in one flat bucket in symmetric (peer-mapped) memory; every rank can
// load from and store to every peer's bucket through NVLink / NVSwitch.  Rank r owns slice r of the bucket: after a
// barrier ("all gradients are complete") it loads that slice from all W buckets, adds the W values in rank order
// (the same order on every rank and in every run: deterministic, and bitwise equal to ((g0 + g1) + g2) + ...) and
// stores the sum into slice r of all W buckets — reduce-scatter and all-gather fused, the transfers overlapping the
// additions element by element; a second barrier ("all slices have landed") ends the collective.  2 (W-1)/W of the
// bucket crosses NVLink per rank, like a ring all-reduce, in one pass and without staging buffers.
//
// The kernel runs beside the render kernels of the next views (which want every SM): it is launched with a small,
// fixed number of CTAs — enough loads in flight to cover the NVLink round trip (W-1 independent 16-byte loads per
// thread), not more.
#include "common.cuh"
#include "kernels.h"

namespace pgs {

#ifndef PGS_EMU
__device__ __forceinline__ float4 ld_peer(const float4* p) { return __ldcg(p); }   // L2 / fabric, never a stale L1 line
__device__ __forceinline__ void st_peer(float4* p, float4 v) { __stcg(p, v); }
#else
inline float4 ld_peer(const float4* p) { return *p; }
inline void st_peer(float4* p, float4 v) { *p = v; }
#endif

template <int W>
__global__ void __launch_bounds__(512) peer_allreduce_slice_kernel(PeerBuckets b, size_t off4, size_t n4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v[W];
#pragma unroll
    for (int q = 0; q < W; q++) v[q] = ld_peer(reinterpret_cast<const float4*>(b.p[q]) + off4 + i);
    float4 acc = v[0];
#pragma unroll
    for (int q = 1; q < W; q++) {
      acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w;
    }
#pragma unroll
    for (int q = 0; q < W; q++) st_peer(reinterpret_cast<float4*>(b.p[q]) + off4 + i, acc);
  }
}

int launch_peer_allreduce_slice(const PeerBuckets& b, int world, size_t offset_floats, size_t n_floats, int ctas,
                                cudaStream_t s) {
  if (n_floats == 0) return 0;
  if ((offset_floats | n_floats) & 3) return -1;  // 16-byte granularity
  if (ctas <= 0) ctas = 32;
  const size_t off4 = offset_floats / 4, n4 = n_floats / 4;
  const int grid = (int)((n4 + 511) / 512 < (size_t)ctas ? (n4 + 511) / 512 : (size_t)ctas);
  switch (world) {
#define PGS_CASE(W) case W: peer_allreduce_slice_kernel<W><<<grid, 512, 0, s>>>(b, off4, n4); break;
    PGS_CASE(1) PGS_CASE(2) PGS_CASE(3) PGS_CASE(4) PGS_CASE(5) PGS_CASE(6) PGS_CASE(7) PGS_CASE(8)
#undef PGS_CASE
    default: return -2;
  }
  count_launch();
  return 0;
}

}  // namespace pgs
