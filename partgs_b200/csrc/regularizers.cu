// The per-pixel regularisers of the PartGS training step, fused (sm_100a) — SURVEY.md 8(f) rank 2, the part of
// train.py:234-251 that follows the photometric loss:
//   opacity = rend_alpha.clamp(1e-6, 1 - 1e-6)
//   loss_mask_entropy = -(mask * log(opacity) + (1 - mask) * log(1 - opacity)).mean()          (:235-237)
//   normal_loss = lambda_normal * (1 - (rend_normal * surf_normal).sum(0)).mean()                (:248-249)
//   dist_loss   = lambda_dist * rend_dist.mean()                                                 (:250)
// The reference spends ~14 elementwise ATen kernels and three reductions forward and as many backward on these four
// maps (9 floats per pixel).  Here: one kernel forward (three double sums) and one pointwise kernel backward.
// HBM-bound: 36 B read per pixel forward (32 without a mask); backward 36 B read + 32 B written.
#include "common.cuh"
#include "kernels.h"

namespace pgs {

constexpr float REG_LO = 1e-6f;
constexpr float REG_HI = (float)(1.0 - 1e-6);

__device__ __forceinline__ float reg_block_sum(float v, float* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[wid] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid on thread 0
}

// sums[0] += sum of the entropy terms, sums[1] += sum (1 - <rend_normal, surf_normal>), sums[2] += sum rend_dist
__global__ void __launch_bounds__(256) regularizers_fwd_kernel(int npix, const float* __restrict__ alpha,
                                                               const float* __restrict__ mask,
                                                               const float* __restrict__ dist,
                                                               const float* __restrict__ rend_normal,
                                                               const float* __restrict__ surf_normal,
                                                               double* __restrict__ sums) {
  __shared__ float s_red[8];
  float e = 0.f, n = 0.f, d = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
    if (mask) {
      const float a = fminf(fmaxf(alpha[i], REG_LO), REG_HI);
      const float m = mask[i];
      e -= m * logf(a) + (1.f - m) * logf(1.f - a);
    }
    if (rend_normal) {
      const size_t p = (size_t)i;
      n += 1.f - (rend_normal[p] * surf_normal[p] + rend_normal[p + npix] * surf_normal[p + npix] +
                  rend_normal[p + 2 * (size_t)npix] * surf_normal[p + 2 * (size_t)npix]);
    }
    if (dist) d += dist[i];
  }
  const float te = reg_block_sum(e, s_red), tn = reg_block_sum(n, s_red), td = reg_block_sum(d, s_red);
  if (threadIdx.x == 0) {
    if (mask) atomicAdd(&sums[0], (double)te);
    if (rend_normal) atomicAdd(&sums[1], (double)tn);
    if (dist) atomicAdd(&sums[2], (double)td);
  }
}

// g = upstream gradient of the scalar (device memory); k_x = g * lambda_x / npix
__global__ void __launch_bounds__(256) regularizers_bwd_kernel(int npix, const float* __restrict__ alpha,
                                                               const float* __restrict__ mask,
                                                               const float* __restrict__ rend_normal,
                                                               const float* __restrict__ surf_normal,
                                                               const float* __restrict__ g_loss, float lambda_entropy,
                                                               float lambda_normal, float lambda_dist,
                                                               float* __restrict__ g_alpha, float* __restrict__ g_dist,
                                                               float* __restrict__ g_rend_normal,
                                                               float* __restrict__ g_surf_normal) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const float g = g_loss[0] / (float)npix;
  if (g_alpha) {
    float ga = 0.f;
    if (mask) {
      const float x = alpha[i];
      if (x >= REG_LO && x <= REG_HI) {  // clamp passes the gradient inside [min, max] only
        const float m = mask[i];
        ga = -(m / x - (1.f - m) / (1.f - x)) * (g * lambda_entropy);
      }
    }
    g_alpha[i] = ga;
  }
  if (g_dist) g_dist[i] = g * lambda_dist;
  if (g_rend_normal) {
    const float k = -(g * lambda_normal);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const size_t p = (size_t)i + (size_t)c * npix;
      g_rend_normal[p] = k * surf_normal[p];
      g_surf_normal[p] = k * rend_normal[p];
    }
  }
}

void launch_regularizers_fwd(int npix, const float* alpha, const float* mask, const float* dist,
                             const float* rend_normal, const float* surf_normal, double* sums, cudaStream_t s) {
  cudaMemsetAsync(sums, 0, 3 * sizeof(double), s);
  if (npix <= 0) return;
  int blocks = (npix + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;  // grid-stride: a few double atomics per SM instead of one per 256 pixels
  regularizers_fwd_kernel<<<blocks, 256, 0, s>>>(npix, alpha, mask, dist, rend_normal, surf_normal, sums);
  count_launch();
}

void launch_regularizers_bwd(int npix, const float* alpha, const float* mask, const float* rend_normal,
                             const float* surf_normal, const float* g_loss, float lambda_entropy, float lambda_normal,
                             float lambda_dist, float* g_alpha, float* g_dist, float* g_rend_normal,
                             float* g_surf_normal, cudaStream_t s) {
  if (npix <= 0) return;
  regularizers_bwd_kernel<<<(npix + 255) / 256, 256, 0, s>>>(npix, alpha, mask, rend_normal, surf_normal, g_loss,
                                                              lambda_entropy, lambda_normal, lambda_dist, g_alpha,
                                                              g_dist, g_rend_normal, g_surf_normal);
  count_launch();
}

}  // namespace pgs
