// Photometric loss of the PartGS training step, forward and backward, fused (sm_100a):
//   Ll1  = mean |image - gt|                                   utils/loss_utils.py:6-7
//   ssim = mean of the 11x11 Gaussian-window SSIM map          utils/loss_utils.py:12-54
//   loss = (1 - lambda) * Ll1 + lambda * (1 - ssim)            train.py:230-231
// The reference runs five grouped 11x11 conv2d (mu1, mu2, E[x^2], E[y^2], E[xy]) plus ~15 pointwise
// kernels forward and as many backward, with full-size temporaries.  Here one kernel per direction: a 16x16
// pixel tile (+5 halo) of both images is staged in shared memory, the Gaussian window is applied separably
// (it is an outer product: create_window, loss_utils.py:16-20), the SSIM map is reduced to per-launch sums, and
// the three partial-derivative maps the backward convolution needs are written out (12 B per pixel x channel).
// Zero padding like F.conv2d(padding=5).  HBM-bound: 8 B read + 12 B written per pixel x channel forward.
#include "common.cuh"
#include "kernels.h"

namespace pgs {

constexpr int PH_T = 16;             // output tile edge
constexpr int PH_R = 5;              // window radius (window_size 11)
constexpr int PH_S = PH_T + 2 * PH_R;  // staged tile edge (26)
constexpr float PH_C1 = 0.01f * 0.01f;
constexpr float PH_C2 = 0.03f * 0.03f;

struct Gauss11 {
  float w[11];
};
// gaussian(11, 1.5) of the reference: float32 exp values divided by their float32 sum
static Gauss11 make_gauss() {
  Gauss11 g;
  float sum = 0.f;
  for (int x = 0; x < 11; x++) {
    g.w[x] = (float)exp(-(double)((x - 5) * (x - 5)) / (2.0 * 1.5 * 1.5));
    sum += g.w[x];
  }
  for (int x = 0; x < 11; x++) g.w[x] /= sum;
  return g;
}

__device__ __forceinline__ float block_sum(float v, float* s_red) {
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

// sums[0] += sum of the SSIM map, sums[1] += sum |img - gt|   (double accumulators, zeroed by the launcher)
__global__ void __launch_bounds__(PH_T * PH_T) photometric_fwd_kernel(int C, int H, int W, const float* __restrict__ img,
                                                                      const float* __restrict__ gt, Gauss11 gw,
                                                                      double* __restrict__ sums,
                                                                      float* __restrict__ dmaps /* [3][C][H][W] */) {
  __shared__ float s1[PH_S][PH_S + 1], s2[PH_S][PH_S + 1];
  __shared__ float hx[5][PH_S][PH_T];
  __shared__ float s_red[8];
  const int tx = threadIdx.x % PH_T, ty = threadIdx.x / PH_T;
  const int x0 = blockIdx.x * PH_T, y0 = blockIdx.y * PH_T;
  const size_t HW = (size_t)H * W, CHW = HW * C;
  float acc_ssim = 0.f, acc_l1 = 0.f;
  for (int c = 0; c < C; c++) {
    const float* a = img + c * HW;
    const float* b = gt + c * HW;
    for (int i = threadIdx.x; i < PH_S * PH_S; i += PH_T * PH_T) {
      const int ly = i / PH_S, lx = i % PH_S;
      const int gx = x0 + lx - PH_R, gy = y0 + ly - PH_R;
      const bool in = gx >= 0 && gy >= 0 && gx < W && gy < H;
      s1[ly][lx] = in ? __ldg(a + (size_t)gy * W + gx) : 0.f;
      s2[ly][lx] = in ? __ldg(b + (size_t)gy * W + gx) : 0.f;
    }
    __syncthreads();
    // horizontal pass: PH_S rows x PH_T columns, five moments
    for (int i = threadIdx.x; i < PH_S * PH_T; i += PH_T * PH_T) {
      const int ly = i / PH_T, lx = i % PH_T;
      float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
      for (int k = 0; k < 11; k++) {
        const float u = s1[ly][lx + k], v = s2[ly][lx + k], w = gw.w[k];
        m1 += w * u; m2 += w * v; e11 += w * u * u; e22 += w * v * v; e12 += w * u * v;
      }
      hx[0][ly][lx] = m1; hx[1][ly][lx] = m2; hx[2][ly][lx] = e11; hx[3][ly][lx] = e22; hx[4][ly][lx] = e12;
    }
    __syncthreads();
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; k++) {
      const float w = gw.w[k];
      mu1 += w * hx[0][ty + k][tx]; mu2 += w * hx[1][ty + k][tx]; e11 += w * hx[2][ty + k][tx];
      e22 += w * hx[3][ty + k][tx]; e12 += w * hx[4][ty + k][tx];
    }
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx < W && gy < H) {
      const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
      const float sg1 = e11 - mu1_sq, sg2 = e22 - mu2_sq, sg12 = e12 - mu12;
      const float A1 = 2.f * mu12 + PH_C1, A2 = 2.f * sg12 + PH_C2;
      const float B1 = mu1_sq + mu2_sq + PH_C1, B2 = sg1 + sg2 + PH_C2;
      const float inv = 1.f / (B1 * B2);
      const float m = A1 * A2 * inv;
      acc_ssim += m;
      acc_l1 += fabsf(s1[ty + PH_R][tx + PH_R] - s2[ty + PH_R][tx + PH_R]);
      // partial derivatives of the map value w.r.t. the three window statistics that depend on `img`
      const float d_s1 = -m / B2;                 // via sigma1_sq (in B2)
      const float d_s12 = 2.f * A1 * inv;         // via sigma12 (in A2)
      const float d_mu1 = 2.f * mu2 * A2 * inv - m * 2.f * mu1 / B1   // direct (A1, B1)
                          - 2.f * mu1 * d_s1 - mu2 * d_s12;           // sigma1_sq = E11 - mu1^2, sigma12 = E12 - mu1 mu2
      const size_t p = (size_t)c * HW + (size_t)gy * W + gx;
      dmaps[p] = d_mu1;
      dmaps[CHW + p] = d_s1;
      dmaps[2 * CHW + p] = d_s12;
    }
    __syncthreads();
  }
  const float t_ssim = block_sum(acc_ssim, s_red);
  const float t_l1 = block_sum(acc_l1, s_red);
  if (threadIdx.x == 0) {
    atomicAdd(&sums[0], (double)t_ssim);
    atomicAdd(&sums[1], (double)t_l1);
  }
}

// d loss / d img = k_l1 * sign(img - gt) + k_ssim * [ conv(d_mu1) + 2 img conv(d_s1) + gt conv(d_s12) ]
// with k_l1 = g * (1 - lambda) / N and k_ssim = -g * lambda / N (g = upstream gradient of the scalar loss, read
// from device memory so no host synchronisation is needed).
__global__ void __launch_bounds__(PH_T * PH_T) photometric_bwd_kernel(int C, int H, int W, const float* __restrict__ img,
                                                                      const float* __restrict__ gt, Gauss11 gw,
                                                                      const float* __restrict__ dmaps,
                                                                      const float* __restrict__ g_loss, float lambda,
                                                                      float* __restrict__ g_img) {
  __shared__ float sd[3][PH_S][PH_S + 1];
  __shared__ float hx[3][PH_S][PH_T];
  const int tx = threadIdx.x % PH_T, ty = threadIdx.x / PH_T;
  const int x0 = blockIdx.x * PH_T, y0 = blockIdx.y * PH_T;
  const size_t HW = (size_t)H * W, CHW = HW * C;
  const float g = __ldg(g_loss);
  const float inv_n = 1.f / (float)CHW;
  const float k_l1 = g * (1.f - lambda) * inv_n, k_ssim = -g * lambda * inv_n;
  for (int c = 0; c < C; c++) {
    for (int i = threadIdx.x; i < PH_S * PH_S; i += PH_T * PH_T) {
      const int ly = i / PH_S, lx = i % PH_S;
      const int gx = x0 + lx - PH_R, gy = y0 + ly - PH_R;
      const bool in = gx >= 0 && gy >= 0 && gx < W && gy < H;
      const size_t p = (size_t)c * HW + (size_t)gy * W + gx;
#pragma unroll
      for (int q = 0; q < 3; q++) sd[q][ly][lx] = in ? __ldg(dmaps + q * CHW + p) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PH_S * PH_T; i += PH_T * PH_T) {
      const int ly = i / PH_T, lx = i % PH_T;
      float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
      for (int k = 0; k < 11; k++) {
        const float w = gw.w[k];
        r0 += w * sd[0][ly][lx + k]; r1 += w * sd[1][ly][lx + k]; r2 += w * sd[2][ly][lx + k];
      }
      hx[0][ly][lx] = r0; hx[1][ly][lx] = r1; hx[2][ly][lx] = r2;
    }
    __syncthreads();
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; k++) {
      const float w = gw.w[k];
      c0 += w * hx[0][ty + k][tx]; c1 += w * hx[1][ty + k][tx]; c2 += w * hx[2][ty + k][tx];
    }
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx < W && gy < H) {
      const size_t p = (size_t)c * HW + (size_t)gy * W + gx;
      const float u = img[p], v = gt[p];
      const float d = u - v;
      const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      g_img[p] = k_l1 * sgn + k_ssim * (c0 + 2.f * u * c1 + v * c2);
    }
    __syncthreads();
  }
}

void launch_photometric_fwd(int C, int H, int W, const float* img, const float* gt, double* sums, float* dmaps,
                            cudaStream_t s) {
  static const Gauss11 gw = make_gauss();
  cudaMemsetAsync(sums, 0, 2 * sizeof(double), s);
  dim3 grid((W + PH_T - 1) / PH_T, (H + PH_T - 1) / PH_T);
  photometric_fwd_kernel<<<grid, PH_T * PH_T, 0, s>>>(C, H, W, img, gt, gw, sums, dmaps);
  count_launch();
}

void launch_photometric_bwd(int C, int H, int W, const float* img, const float* gt, const float* dmaps,
                            const float* g_loss, float lambda, float* g_img, cudaStream_t s) {
  static const Gauss11 gw = make_gauss();
  dim3 grid((W + PH_T - 1) / PH_T, (H + PH_T - 1) / PH_T);
  photometric_bwd_kernel<<<grid, PH_T * PH_T, 0, s>>>(C, H, W, img, gt, gw, dmaps, g_loss, lambda, g_img);
  count_launch();
}

}  // namespace pgs
