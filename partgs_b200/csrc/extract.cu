// Per-view epilogue of the multi-view extraction loop (render.py -> GaussianExtractor.reconstruction,
// utils/mesh_utils.py:77-116), fused (sm_100a).
//
// After every render the reference turns the S-channel part map into a colour image with
// partmap_to_rgbmap (clamp, sum over parts, argmax, one boolean-mask scatter per class, background fill:
// 2 S + 6 ATen kernels and a host-side tensor) and normalises the rendered normals
// (torch.nn.functional.normalize(dim=0): 4 kernels).  Here: one kernel, every input plane read once,
// every output plane written once — (4 S + 12) + 24 B per pixel, HBM-bound.
#include "common.cuh"
#include "kernels.h"

namespace pgs {

__global__ void __launch_bounds__(256) extract_maps_kernel(int npix, int S, const float* __restrict__ semantic,
                                                           const float* __restrict__ palette, int palette_stride,
                                                           const float* __restrict__ rend_normal,
                                                           float* __restrict__ part_rgb,
                                                           float* __restrict__ normal_unit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  if (part_rgb) {
    // part = clamp(part, 0, 1); background = part.sum(0); cls = part.argmax(0) (first maximum wins, NaN is a maximum)
    float sum = 0.f, best = 0.f;
    int cls = 0;
    for (int s = 0; s < S; s++) {
      float v = semantic[(size_t)s * npix + i];
      v = (v != v) ? v : fminf(fmaxf(v, 0.f), 1.f);  // torch.clamp keeps NaN
      sum += v;
      if (s == 0 || v > best || (v != v && best == best)) {
        best = v;
        cls = s;
      }
    }
    float r = 0.f, g = 0.f, b = 0.f;
    if (S > 0) {
      r = palette[cls * palette_stride];
      g = palette[cls * palette_stride + 1];
      b = palette[cls * palette_stride + 2];
    }
    if (sum < 1e-1f) r = g = b = 1.f;  // color_image_part[background < 1e-1, :] = 1.
    part_rgb[i] = r;
    part_rgb[(size_t)npix + i] = g;
    part_rgb[(size_t)2 * npix + i] = b;
  }
  if (normal_unit) {
    // F.normalize(x, dim=0): x / max(||x||_2, 1e-12)
    const float x = rend_normal[i], y = rend_normal[(size_t)npix + i], z = rend_normal[(size_t)2 * npix + i];
    const float n = fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
    normal_unit[i] = __fdiv_rn(x, n);
    normal_unit[(size_t)npix + i] = __fdiv_rn(y, n);
    normal_unit[(size_t)2 * npix + i] = __fdiv_rn(z, n);
  }
}

void launch_extract_maps(int npix, int S, const float* semantic, const float* palette, int palette_stride,
                         const float* rend_normal, float* part_rgb, float* normal_unit, cudaStream_t s) {
  if (npix <= 0) return;
  extract_maps_kernel<<<(npix + 255) / 256, 256, 0, s>>>(npix, S, semantic, palette, palette_stride, rend_normal,
                                                         part_rgb, normal_unit);
  count_launch();
}

}  // namespace pgs
