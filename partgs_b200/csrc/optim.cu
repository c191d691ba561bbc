// Optimiser step and densification bookkeeping of the PartGS training loop, fused (sm_100a).
//   * torch.optim.Adam(l, lr=0.0, eps=1e-15) over the six per-surfel parameter groups
//     (scene/gaussian_model.py:256-266; train.py: gaussians.optimizer.step()): ONE launch for all groups instead of
//     ~10 ATen kernels per group.  Arithmetic follows torch's single-tensor Adam (lerp form of exp_avg, bias
//     corrections folded into step_size / denom exactly as torch/optim/adam.py does).
//   * add_densification_stats + the max_radii2D update (scene/gaussian_model.py:515-517, train.py:295-297):
//     one kernel over the surfels instead of four boolean-mask gathers / scatters.
// HBM-bound: 28 B per parameter element (p, g, m, v read; p, m, v written).
#include "common.cuh"
#include "kernels.h"

namespace pgs {

__global__ void __launch_bounds__(256) adam_multi_kernel(AdamTable t, float one_minus_beta1, float beta2,
                                                         float one_minus_beta2, float eps,
                                                         float inv_bias_correction2_sqrt) {
  // which tensor does this CTA work on?  (block_start is an exclusive prefix of the CTAs per tensor)
  int k = 0;
#pragma unroll
  for (int i = 1; i < PGS_ADAM_MAX_TENSORS; i++)
    if (i < t.n && (int)blockIdx.x >= t.block_start[i]) k = i;
  const size_t base = ((size_t)blockIdx.x - t.block_start[k]) * (256 * 4);
  float* __restrict__ p = t.param[k];
  const float* __restrict__ g = t.grad[k];
  float* __restrict__ m = t.exp_avg[k];
  float* __restrict__ v = t.exp_avg_sq[k];
  const float step_size = t.step_size[k];  // lr / (1 - beta1^step), computed in double on the host like torch
  const size_t n = t.numel[k];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const size_t i = base + (size_t)j * 256 + threadIdx.x;
    if (i < n) {
      const float gi = g[i];
      float mi = m[i], vi = v[i];
      // (1 - beta) is formed in double on the host like torch's Python does: 1.f - 0.999f would be off by 5e-5
      mi = mi + (gi - mi) * one_minus_beta1;               // exp_avg.lerp_(grad, 1 - beta1)
      vi = vi * beta2 + one_minus_beta2 * gi * gi;         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      // ATen divides a tensor by a host scalar as a multiplication by its float reciprocal
      const float denom = sqrtf(vi) * inv_bias_correction2_sqrt + eps;
      p[i] = p[i] - step_size * (mi / denom);              // param.addcdiv_(exp_avg, denom, value=-step_size)
      m[i] = mi;
      v[i] = vi;
    }
  }
}

int launch_adam_multi(AdamTable& t, double beta1, double beta2, double eps, double bias_correction2_sqrt,
                      cudaStream_t s) {
  int blocks = 0;
  for (int i = 0; i < t.n; i++) {
    t.block_start[i] = blocks;
    blocks += (int)((t.numel[i] + 1023) / 1024);
  }
  if (blocks == 0) return 0;
  adam_multi_kernel<<<blocks, 256, 0, s>>>(t, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps,
                                           1.0f / (float)bias_correction2_sqrt);
  count_launch();
  return 0;
}

// visibility_filter = radii > 0:  max_radii2D = max(max_radii2D, radii);  xyz_gradient_accum += |grad_means2D.xy|;  denom += 1
__global__ void __launch_bounds__(256) densify_stats_kernel(int P, const int* __restrict__ radii,
                                                            const float* __restrict__ grad_means2D /* [P,3] */,
                                                            float* __restrict__ max_radii2D, float* __restrict__ accum,
                                                            float* __restrict__ denom) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int r = radii[i];
  if (r <= 0) return;
  if (max_radii2D) max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
  const float gx = grad_means2D[3 * i], gy = grad_means2D[3 * i + 1];
  accum[i] += sqrtf(gx * gx + gy * gy);
  denom[i] += 1.f;
}

void launch_densify_stats(int P, const int* radii, const float* grad_means2D, float* max_radii2D, float* accum,
                          float* denom, cudaStream_t s) {
  if (P <= 0) return;
  densify_stats_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, radii, grad_means2D, max_radii2D, accum, denom);
  count_launch();
}

}  // namespace pgs
