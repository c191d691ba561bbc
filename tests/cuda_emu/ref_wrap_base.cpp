// TEST INFRASTRUCTURE: extern "C" doors into the UNMODIFIED reference rasteriser (base fork,
// diff-surfel-rasterization) compiled for the CPU emulator (tests/cuda_emu/build.py::build_reference).
#include <functional>
#include <vector>

#include "rasterizer.h"

namespace {
std::vector<char> g_geom, g_bin, g_img;
std::function<char*(size_t)> resizer(std::vector<char>& v) {
  return [&v](size_t n) { v.assign(n + 256, 0); return (char*)(((size_t)v.data() + 127) & ~(size_t)127); };
}
char* base(std::vector<char>& v) { return (char*)(((size_t)v.data() + 127) & ~(size_t)127); }
}  // namespace

extern "C" {
int ref_base_forward(int P, int D, int M, const float* bg, int W, int H, const float* means3D, const float* shs,
                     const float* colors_precomp, const float* opacities, const float* scales, float scale_modifier,
                     const float* rotations, const float* transMat_precomp, const float* viewmatrix, const float* projmatrix,
                     const float* campos, float tan_fovx, float tan_fovy, float* out_color, float* out_others,
                     int* radii) {
  return CudaRasterizer::Rasterizer::forward(resizer(g_geom), resizer(g_bin), resizer(g_img), P, D, M, bg, W, H,
                                             means3D, shs, colors_precomp, opacities, scales, scale_modifier,
                                             rotations, transMat_precomp, viewmatrix, projmatrix, campos, tan_fovx,
                                             tan_fovy, false, out_color, out_others, radii, false);
}
void ref_base_backward(int P, int D, int M, int R, const float* bg, int W, int H, const float* means3D,
                       const float* shs, const float* colors_precomp, const float* scales, float scale_modifier,
                       const float* rotations, const float* transMat_precomp, const float* viewmatrix,
                       const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy, const int* radii,
                       const float* dL_dpix, const float* dL_dothers,
                       float* dL_dmean2D, float* dL_dnormal, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D,
                       float* dL_dtransMat, float* dL_dsh, float* dL_dscale, float* dL_drot) {
  CudaRasterizer::Rasterizer::backward(P, D, M, R, bg, W, H, means3D, shs, colors_precomp, scales, scale_modifier,
                                       rotations, transMat_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy,
                                       radii, base(g_geom), base(g_bin), base(g_img), dL_dpix, dL_dothers, dL_dmean2D,
                                       dL_dnormal, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dtransMat, dL_dsh, dL_dscale,
                                       dL_drot, false);
}
}
