/*
 * TEST INFRASTRUCTURE — CPU restatement of the reference surfel rasteriser.
 * Not product code: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load this library; the product path (partgs_b200/) never does.
 *
 * Plain C (fp32, one statement per reference statement, same association order) of
 * zhirui-gao/PartGS submodules/diff-surfel-rasterization:
 *   preprocess      cuda_rasterizer/forward.cu:148-251 (+ :20-71 SH, :75-115 T, :119-145 AABB,
 *                   auxiliary.h:67-77 getRect, :185-210 in_frustum, :213-235 quat_to_rotmat)
 *   binning         cuda_rasterizer/rasterizer_impl.cu:35-50 getHigherMsb, :70-111 duplicateWithKeys,
 *                   :301-309 stable sort on bits [0,32+msb), :116-138 identifyTileRanges
 *   render fwd      cuda_rasterizer/forward.cu:256-441
 *   render bwd      cuda_rasterizer/backward.cu:143-440
 *   preprocess bwd  cuda_rasterizer/backward.cu:443-630 (+ :20-139 SH backward,
 *                   auxiliary.h:128-138 dnormvdv, :238-282 quat_to_rotmat_vjp)
 *
 * Parity pin:
 *   (1) on the CPU, against golden vectors computed by the reference's OWN CUDA source (the unmodified
 *       cuda_rasterizer/{forward,backward,rasterizer_impl}.cu, executed by the lock-step emulator of
 *       tests/cuda_emu; tools/make_golden_ref_emu.py -> tests/golden/ref_emu_base_*.npz): radii and the
 *       instance count exactly, images <= 2e-5, gradients <= 2e-4, incl. scale_modifier != 1, SH degree 2 and
 *       a non-zero background (tests/test_emu_base_ref.py);
 *   (2) on the GPU, against the reference CUDA built for sm_100a (oracle/_ref) in the same process
 *       (tests/test_gpu_base_raster.py).
 * The GPU contracts a*b+c into FMAs; this file is compiled with -ffp-contract=off, so results agree with
 * the reference build to rounding (images ~1e-6, a handful of radii may differ by one on borderline surfels);
 * the bit-exact claims of the product are tested against oracle/_ref, not against this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BLOCK_X 16
#define BLOCK_Y 16
#define BLOCK_SIZE 256
static const float near_n = 0.2f, far_n = 100.0f, FilterSize = 0.707106f, FilterInvSquare = 2.0f;
static const float SH_C0 = 0.28209479177387814f, SH_C1 = 0.4886025119029199f;
static const float SH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                              0.5462742152960396f};
static const float SH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                              -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

typedef struct {
  int P, D, M, W, H, gx, gy, R;
  float *depths, *means2D, *transMat, *normal_opacity, *rgb;
  uint8_t* clamped;
  int* radii;
  uint32_t *tiles_touched, *point_offsets;
  uint64_t *keys_unsorted, *keys;
  uint32_t *vals_unsorted, *point_list;
  uint32_t* ranges;    /* [ntiles][2] */
  float* final_T;      /* [3][N]: T, M1, M2 */
  uint32_t* n_contrib; /* [2][N]: last, median */
} oracle_state;

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* reference getHigherMsb (rasterizer_impl.cu:35-50) */
uint32_t oracle_higher_msb(uint32_t n) {
  uint32_t msb = sizeof(n) * 4, step = msb;
  while (step > 1) {
    step /= 2;
    if (n >> msb) msb += step; else msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

static void quat_to_rotmat(const float* q /*w,x,y,z*/, float R[3][3] /* R[col][row] */) {
  float s = 1.0f / sqrtf(q[3] * q[3] + q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  float w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
  R[0][0] = 1.f - 2.f * (y * y + z * z); R[0][1] = 2.f * (x * y + w * z); R[0][2] = 2.f * (x * z - w * y);
  R[1][0] = 2.f * (x * y - w * z); R[1][1] = 1.f - 2.f * (x * x + z * z); R[1][2] = 2.f * (y * z + w * x);
  R[2][0] = 2.f * (x * z + w * y); R[2][1] = 2.f * (y * z - w * x); R[2][2] = 1.f - 2.f * (x * x + y * y);
}

/* P = world2ndc * ndc2pix as 3 columns of 4 rows (backward.cu:489-503); forward uses the
 * other association (A*w)*n (forward.cu:112) — both are restated where they are used. */
static void compute_T_forward(const float* p, const float* scale, float mod, const float* rot, const float* proj,
                              const float* view, int W, int H, float T[3][3], float normal[3]) {
  float R[3][3];
  quat_to_rotmat(rot, R);
  float L[3][3];
  for (int r = 0; r < 3; r++) {
    L[0][r] = R[0][r] * (mod * scale[0]);
    L[1][r] = R[1][r] * (mod * scale[1]);
    L[2][r] = R[2][r];
  }
  /* A = transpose(splat2world): 4 columns of 3 rows */
  float A[4][3] = {{L[0][0], L[1][0], p[0]}, {L[0][1], L[1][1], p[1]}, {L[0][2], L[1][2], p[2]}, {0.f, 0.f, 1.f}};
  float B[4][3];
  for (int i = 0; i < 4; i++)
    for (int r = 0; r < 3; r++) {
      float acc = A[0][r] * proj[i];
      acc += A[1][r] * proj[4 + i];
      acc += A[2][r] * proj[8 + i];
      acc += A[3][r] * proj[12 + i];
      B[i][r] = acc;
    }
  float n0[4] = {(float)((float)W / 2.0), 0.f, 0.f, (float)((float)(W - 1) / 2.0)};
  float n1[4] = {0.f, (float)((float)H / 2.0), 0.f, (float)((float)(H - 1) / 2.0)};
  float n2[4] = {0.f, 0.f, 0.f, 1.f};
  const float* n[3] = {n0, n1, n2};
  for (int i = 0; i < 3; i++)
    for (int r = 0; r < 3; r++) {
      float acc = B[0][r] * n[i][0];
      acc += B[1][r] * n[i][1];
      acc += B[2][r] * n[i][2];
      acc += B[3][r] * n[i][3];
      T[i][r] = acc;
    }
  normal[0] = view[0] * L[2][0] + view[4] * L[2][1] + view[8] * L[2][2];
  normal[1] = view[1] * L[2][0] + view[5] * L[2][1] + view[9] * L[2][2];
  normal[2] = view[2] * L[2][0] + view[6] * L[2][1] + view[10] * L[2][2];
}

static int compute_aabb(float T[3][3], float cutoff, float pim[2], float ext[2]) {
  float t[3] = {cutoff * cutoff, cutoff * cutoff, -1.0f};
  float d = t[0] * (T[2][0] * T[2][0]);
  d += t[1] * (T[2][1] * T[2][1]);
  d += t[2] * (T[2][2] * T[2][2]);
  if (d == 0.0) return 0;
  float f[3] = {(1 / d) * t[0], (1 / d) * t[1], (1 / d) * t[2]};
  float pp[2], q[2];
  for (int a = 0; a < 2; a++) {
    float s = f[0] * (T[a][0] * T[2][0]);
    s += f[1] * (T[a][1] * T[2][1]);
    s += f[2] * (T[a][2] * T[2][2]);
    pp[a] = s;
    float u = f[0] * (T[a][0] * T[a][0]);
    u += f[1] * (T[a][1] * T[a][1]);
    u += f[2] * (T[a][2] * T[a][2]);
    q[a] = u;
  }
  for (int a = 0; a < 2; a++) {
    float h0 = pp[a] * pp[a] - q[a];
    ext[a] = sqrtf(fmaxf((float)1e-4, h0));
    pim[a] = pp[a];
  }
  return 1;
}

static void get_rect(const float p[2], int max_radius, uint32_t rmin[2], uint32_t rmax[2], uint32_t gx, uint32_t gy) {
  int v;
  v = (int)((p[0] - max_radius) / BLOCK_X); v = v > 0 ? v : 0; rmin[0] = (uint32_t)v < gx ? (uint32_t)v : gx;
  v = (int)((p[1] - max_radius) / BLOCK_Y); v = v > 0 ? v : 0; rmin[1] = (uint32_t)v < gy ? (uint32_t)v : gy;
  v = (int)((p[0] + max_radius + BLOCK_X - 1) / BLOCK_X); v = v > 0 ? v : 0; rmax[0] = (uint32_t)v < gx ? (uint32_t)v : gx;
  v = (int)((p[1] + max_radius + BLOCK_Y - 1) / BLOCK_Y); v = v > 0 ? v : 0; rmax[1] = (uint32_t)v < gy ? (uint32_t)v : gy;
}

static void color_from_sh(int idx, int deg, int M, const float* means, const float* campos, const float* shs,
                          uint8_t* clamped, float out[3]) {
  const float* pos = means + 3 * idx;
  float dir[3] = {pos[0] - campos[0], pos[1] - campos[1], pos[2] - campos[2]};
  float len = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
  dir[0] /= len; dir[1] /= len; dir[2] /= len;
  const float* sh = shs + (size_t)idx * M * 3;
  float x = dir[0], y = dir[1], z = dir[2];
  for (int c = 0; c < 3; c++) {
#define SH(i) sh[(i) * 3 + c]
    float r = SH_C0 * SH(0);
    if (deg > 0) {
      r = r - SH_C1 * y * SH(1) + SH_C1 * z * SH(2) - SH_C1 * x * SH(3);
      if (deg > 1) {
        float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        r = r + SH_C2[0] * xy * SH(4) + SH_C2[1] * yz * SH(5) + SH_C2[2] * (2.0f * zz - xx - yy) * SH(6) +
            SH_C2[3] * xz * SH(7) + SH_C2[4] * (xx - yy) * SH(8);
        if (deg > 2) {
          r = r + SH_C3[0] * y * (3.0f * xx - yy) * SH(9) + SH_C3[1] * xy * z * SH(10) +
              SH_C3[2] * y * (4.0f * zz - xx - yy) * SH(11) + SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SH(12) +
              SH_C3[4] * x * (4.0f * zz - xx - yy) * SH(13) + SH_C3[5] * z * (xx - yy) * SH(14) +
              SH_C3[6] * x * (xx - 3.0f * yy) * SH(15);
        }
      }
    }
#undef SH
    r += 0.5f;
    clamped[3 * idx + c] = (r < 0);
    out[c] = fmaxf(r, 0.0f);
  }
}

static void radix_sort_pairs(uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp, size_t n,
                             int end_bit) {
  /* stable LSD radix sort on bits [0,end_bit); result in keys/vals */
  uint64_t *ki = keys, *ko = keys_tmp;
  uint32_t *vi = vals, *vo = vals_tmp;
  int swaps = 0;
  for (int shift = 0; shift < end_bit; shift += 8) {
    int bits = end_bit - shift < 8 ? end_bit - shift : 8;
    uint32_t mask = (1u << bits) - 1;
    size_t cnt[257];
    memset(cnt, 0, sizeof(cnt));
    for (size_t i = 0; i < n; i++) cnt[((ki[i] >> shift) & mask) + 1]++;
    for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
    for (size_t i = 0; i < n; i++) {
      size_t o = cnt[(ki[i] >> shift) & mask]++;
      ko[o] = ki[i];
      vo[o] = vi[i];
    }
    uint64_t* tk = ki; ki = ko; ko = tk;
    uint32_t* tv = vi; vi = vo; vo = tv;
    swaps++;
  }
  if (swaps & 1) {
    memcpy(keys, keys_tmp, n * sizeof(uint64_t));
    memcpy(vals, vals_tmp, n * sizeof(uint32_t));
  }
}

void oracle_free(oracle_state* s) {
  if (!s) return;
  free(s->depths); free(s->means2D); free(s->transMat); free(s->normal_opacity); free(s->rgb); free(s->clamped);
  free(s->radii); free(s->tiles_touched); free(s->point_offsets); free(s->keys_unsorted); free(s->keys);
  free(s->vals_unsorted); free(s->point_list); free(s->ranges); free(s->final_T); free(s->n_contrib);
  free(s);
}

/* Accessors (ctypes friendliness) */
#define ACC(type, name) type* oracle_##name(oracle_state* s) { return s->name; }
ACC(float, depths) ACC(float, means2D) ACC(float, transMat) ACC(float, normal_opacity) ACC(float, rgb)
ACC(uint8_t, clamped) ACC(int, radii) ACC(uint32_t, tiles_touched) ACC(uint32_t, point_offsets)
ACC(uint64_t, keys_unsorted) ACC(uint64_t, keys) ACC(uint32_t, vals_unsorted) ACC(uint32_t, point_list)
ACC(uint32_t, ranges) ACC(float, final_T) ACC(uint32_t, n_contrib)
int oracle_num_rendered(oracle_state* s) { return s->R; }

/* Forward: Rasterizer::forward (rasterizer_impl.cu:198-342). out_color [3,H,W], out_others [7,H,W]. */
oracle_state* oracle_forward(int P, int D, int M, const float* bg, int W, int H, const float* means3D,
                             const float* shs, const float* colors_precomp, const float* opacities,
                             const float* scales, float scale_modifier, const float* rotations,
                             const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                             float tan_fovy, float* out_color, float* out_others, int* radii_out) {
  (void)tan_fovx; (void)tan_fovy;
  oracle_state* s = (oracle_state*)calloc(1, sizeof(oracle_state));
  const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
  const size_t N = (size_t)W * H, ntiles = (size_t)gx * gy;
  s->P = P; s->D = D; s->M = M; s->W = W; s->H = H; s->gx = gx; s->gy = gy;
  s->depths = (float*)calloc(P, 4); s->means2D = (float*)calloc(P, 8); s->transMat = (float*)calloc(P, 36);
  s->normal_opacity = (float*)calloc(P, 16); s->rgb = (float*)calloc(P, 12); s->clamped = (uint8_t*)calloc(P, 3);
  s->radii = (int*)calloc(P, 4); s->tiles_touched = (uint32_t*)calloc(P, 4);
  s->point_offsets = (uint32_t*)calloc(P, 4);
  s->ranges = (uint32_t*)calloc(ntiles * 2, 4); s->final_T = (float*)calloc(3 * N, 4);
  s->n_contrib = (uint32_t*)calloc(2 * N, 4);

  /* ---- preprocessCUDA ---- */
#pragma omp parallel for schedule(static)
  for (int idx = 0; idx < P; idx++) {
    const float* p = means3D + 3 * idx;
    float pv[3];
    for (int r = 0; r < 3; r++)
      pv[r] = viewmatrix[r] * p[0] + viewmatrix[4 + r] * p[1] + viewmatrix[8 + r] * p[2] + viewmatrix[12 + r];
    if (pv[2] <= 0.2f) continue;
    float T[3][3], normal[3];
    compute_T_forward(p, scales + 2 * idx, scale_modifier, rotations + 4 * idx, projmatrix, viewmatrix, W, H, T, normal);
    for (int i = 0; i < 3; i++)
      for (int r = 0; r < 3; r++) s->transMat[9 * idx + 3 * i + r] = T[i][r];
    float cosv = -(pv[0] * normal[0] + pv[1] * normal[1] + pv[2] * normal[2]);
    if (cosv == 0) continue;
    float mult = cosv > 0 ? 1 : -1;
    normal[0] *= mult; normal[1] *= mult; normal[2] *= mult;
    float cutoff = 3.0f, pim[2], ext[2];
    if (!compute_aabb(T, cutoff, pim, ext)) continue;
    float radius = ceilf(fmaxf(fmaxf(ext[0], ext[1]), cutoff * FilterSize));
    uint32_t rmin[2], rmax[2];
    get_rect(pim, (int)radius, rmin, rmax, gx, gy);
    if ((rmax[0] - rmin[0]) * (rmax[1] - rmin[1]) == 0) continue;
    if (colors_precomp == NULL) {
      color_from_sh(idx, D, M, means3D, campos, shs, s->clamped, s->rgb + 3 * idx);
    } else {
      for (int c = 0; c < 3; c++) s->rgb[3 * idx + c] = colors_precomp[3 * idx + c];
    }
    s->depths[idx] = pv[2];
    s->radii[idx] = (int)radius;
    s->means2D[2 * idx] = pim[0]; s->means2D[2 * idx + 1] = pim[1];
    s->normal_opacity[4 * idx] = normal[0]; s->normal_opacity[4 * idx + 1] = normal[1];
    s->normal_opacity[4 * idx + 2] = normal[2]; s->normal_opacity[4 * idx + 3] = opacities[idx];
    s->tiles_touched[idx] = (rmax[1] - rmin[1]) * (rmax[0] - rmin[0]);
  }
  if (radii_out) memcpy(radii_out, s->radii, (size_t)P * 4);

  /* ---- InclusiveSum + duplicateWithKeys ---- */
  uint32_t run = 0;
  for (int i = 0; i < P; i++) { run += s->tiles_touched[i]; s->point_offsets[i] = run; }
  const size_t R = run;
  s->R = (int)R;
  s->keys_unsorted = (uint64_t*)malloc((R + 1) * 8); s->keys = (uint64_t*)malloc((R + 1) * 8);
  s->vals_unsorted = (uint32_t*)malloc((R + 1) * 4); s->point_list = (uint32_t*)malloc((R + 1) * 4);
#pragma omp parallel for schedule(static)
  for (int idx = 0; idx < P; idx++) {
    if (s->radii[idx] > 0) {
      uint32_t off = idx == 0 ? 0 : s->point_offsets[idx - 1];
      uint32_t rmin[2], rmax[2];
      get_rect(s->means2D + 2 * idx, s->radii[idx], rmin, rmax, gx, gy);
      uint32_t dbits;
      memcpy(&dbits, &s->depths[idx], 4);
      for (uint32_t y = rmin[1]; y < rmax[1]; y++)
        for (uint32_t x = rmin[0]; x < rmax[0]; x++) {
          uint64_t key = y * gx + x;
          key <<= 32; key |= dbits;
          s->keys_unsorted[off] = key; s->vals_unsorted[off] = idx; off++;
        }
    }
  }
  /* ---- SortPairs ---- */
  memcpy(s->keys, s->keys_unsorted, R * 8);
  memcpy(s->point_list, s->vals_unsorted, R * 4);
  {
    uint64_t* kt = (uint64_t*)malloc((R + 1) * 8);
    uint32_t* vt = (uint32_t*)malloc((R + 1) * 4);
    radix_sort_pairs(s->keys, s->point_list, kt, vt, R, 32 + (int)oracle_higher_msb(gx * gy));
    free(kt); free(vt);
  }
  /* ---- identifyTileRanges ---- */
  for (size_t i = 0; i < R; i++) {
    uint32_t cur = (uint32_t)(s->keys[i] >> 32);
    if (i == 0) s->ranges[2 * cur] = 0;
    else {
      uint32_t prev = (uint32_t)(s->keys[i - 1] >> 32);
      if (cur != prev) { s->ranges[2 * prev + 1] = (uint32_t)i; s->ranges[2 * cur] = (uint32_t)i; }
    }
    if (i == R - 1) s->ranges[2 * cur + 1] = (uint32_t)R;
  }

  /* ---- renderCUDA (forward.cu:256-441), one pixel at a time ---- */
#pragma omp parallel for schedule(dynamic, 1)
  for (int tile = 0; tile < (int)ntiles; tile++) {
    const int tx = tile % gx, ty = tile / gx;
    const uint32_t r0 = s->ranges[2 * tile], r1 = s->ranges[2 * tile + 1];
    for (int ly = 0; ly < BLOCK_Y; ly++)
      for (int lx = 0; lx < BLOCK_X; lx++) {
        const uint32_t px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
        if (px >= (uint32_t)W || py >= (uint32_t)H) continue;
        const size_t pix_id = (size_t)W * py + px;
        const float pixfx = (float)px, pixfy = (float)py;
        float T = 1.0f, C[3] = {0, 0, 0}, Nn[3] = {0, 0, 0}, Dd = 0, M1 = 0, M2 = 0, distortion = 0;
        float median_depth = 0, median_contributor = -1;
        uint32_t contributor = 0, last_contributor = 0;
        for (uint32_t i = r0; i < r1; i++) {
          contributor++;
          const uint32_t id = s->point_list[i];
          const float* Tm = s->transMat + 9 * id;
          const float Tu[3] = {Tm[0], Tm[1], Tm[2]}, Tv[3] = {Tm[3], Tm[4], Tm[5]}, Tw[3] = {Tm[6], Tm[7], Tm[8]};
          float k[3], l[3];
          for (int a = 0; a < 3; a++) { k[a] = pixfx * Tw[a] - Tu[a]; l[a] = pixfy * Tw[a] - Tv[a]; }
          float p[3] = {k[1] * l[2] - k[2] * l[1], k[2] * l[0] - k[0] * l[2], k[0] * l[1] - k[1] * l[0]};
          if (p[2] == 0.0) continue;
          float sx = p[0] / p[2], sy = p[1] / p[2];
          float rho3d = (sx * sx + sy * sy);
          float dx = s->means2D[2 * id] - pixfx, dy = s->means2D[2 * id + 1] - pixfy;
          float rho2d = FilterInvSquare * (dx * dx + dy * dy);
          float rho = fminf(rho3d, rho2d);
          float depth = (rho3d <= rho2d) ? (sx * Tw[0] + sy * Tw[1]) + Tw[2] : Tw[2];
          if (depth < near_n) continue;
          const float* no = s->normal_opacity + 4 * id;
          float opa = no[3];
          float power = -0.5f * rho;
          if (power > 0.0f) continue;
          float alpha = fminf(0.99f, opa * expf(power));
          if (alpha < 1.0f / 255.0f) continue;
          float test_T = T * (1 - alpha);
          if (test_T < 0.0001f) break; /* done = true */
          float w = alpha * T;
          float A = 1 - T;
          float m = far_n / (far_n - near_n) * (1 - near_n / depth);
          distortion += (m * m * A + M2 - 2 * m * M1) * w;
          Dd += depth * w;
          M1 += m * w;
          M2 += m * m * w;
          if (T > 0.5) { median_depth = depth; median_contributor = (float)contributor; }
          for (int ch = 0; ch < 3; ch++) Nn[ch] += no[ch] * w;
          for (int ch = 0; ch < 3; ch++) C[ch] += s->rgb[3 * id + ch] * w;
          T = test_T;
          last_contributor = contributor;
        }
        s->final_T[pix_id] = T;
        s->n_contrib[pix_id] = last_contributor;
        for (int ch = 0; ch < 3; ch++) out_color[ch * N + pix_id] = C[ch] + T * bg[ch];
        s->n_contrib[pix_id + N] = median_contributor < 0 ? 0u : (uint32_t)median_contributor;
        s->final_T[pix_id + N] = M1;
        s->final_T[pix_id + 2 * N] = M2;
        out_others[pix_id + 0 * N] = Dd;
        out_others[pix_id + 1 * N] = 1 - T;
        for (int ch = 0; ch < 3; ch++) out_others[pix_id + (2 + ch) * N] = Nn[ch];
        out_others[pix_id + 5 * N] = median_depth;
        out_others[pix_id + 6 * N] = distortion;
      }
  }
  return s;
}

static inline void atomic_addf(float* p, float v) {
#pragma omp atomic
  *p += v;
}

/* Backward: Rasterizer::backward (rasterizer_impl.cu:346-448).  All outputs must be zeroed by the caller
 * (the reference glue passes torch::zeros, rasterize_points.cu:187-195). */
void oracle_backward(oracle_state* s, const float* bg, const float* means3D, const float* shs, const float* scales,
                     const float* rotations, const float* viewmatrix, const float* projmatrix, const float* campos,
                     float tan_fovx, float tan_fovy, const float* dL_dpix, const float* dL_dothers,
                     float* dL_dmean2D /*[P,3]*/, float* dL_dnormal /*[P,3]*/, float* dL_dopacity /*[P]*/,
                     float* dL_dcolor /*[P,3]*/, float* dL_dmean3D /*[P,3]*/, float* dL_dtransMat /*[P,9]*/,
                     float* dL_dsh /*[P,M,3]*/, float* dL_dscale /*[P,2]*/, float* dL_drot /*[P,4]*/) {
  const int W = s->W, H = s->H, gx = s->gx, gy = s->gy, P = s->P, M = s->M, D = s->D;
  const size_t N = (size_t)W * H;
  const int ntiles = gx * gy;
  const float focal_y = H / (2.0f * tan_fovy), focal_x = W / (2.0f * tan_fovx);

  /* ---- BACKWARD::renderCUDA (backward.cu:143-440) ---- */
#pragma omp parallel for schedule(dynamic, 1)
  for (int tile = 0; tile < ntiles; tile++) {
    const int tx = tile % gx, ty = tile / gx;
    const uint32_t r0 = s->ranges[2 * tile], r1 = s->ranges[2 * tile + 1];
    for (int ly = 0; ly < BLOCK_Y; ly++)
      for (int lx = 0; lx < BLOCK_X; lx++) {
        const uint32_t px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
        if (px >= (uint32_t)W || py >= (uint32_t)H) continue;
        const size_t pix_id = (size_t)W * py + px;
        const float pixfx = (float)px, pixfy = (float)py;
        const float T_final = s->final_T[pix_id];
        float T = T_final;
        uint32_t contributor = r1 - r0;
        const uint32_t last_contributor = s->n_contrib[pix_id];
        const int median_contributor = (int)s->n_contrib[pix_id + N];
        float accum_rec[3] = {0, 0, 0}, dL_dpixel[3];
        const float dL_ddepth = dL_dothers[0 * N + pix_id], dL_daccum = dL_dothers[1 * N + pix_id];
        const float dL_dreg = dL_dothers[6 * N + pix_id];
        float dL_dnormal2D[3];
        for (int i = 0; i < 3; i++) dL_dnormal2D[i] = dL_dothers[(2 + i) * N + pix_id];
        const float dL_dmedian_depth = dL_dothers[5 * N + pix_id];
        float last_depth = 0, last_normal[3] = {0, 0, 0}, accum_depth_rec = 0, accum_alpha_rec = 0;
        float accum_normal_rec[3] = {0, 0, 0};
        const float final_D = s->final_T[pix_id + N], final_D2 = s->final_T[pix_id + 2 * N];
        const float final_A = 1 - T_final;
        float last_dL_dT = 0;
        for (int i = 0; i < 3; i++) dL_dpixel[i] = dL_dpix[i * N + pix_id];
        float last_alpha = 0, last_color[3] = {0, 0, 0};
        for (uint32_t ii = r1; ii > r0; ii--) {
          contributor--;
          if (contributor >= last_contributor) continue;
          const uint32_t id = s->point_list[ii - 1];
          const float* Tm = s->transMat + 9 * id;
          const float Tu[3] = {Tm[0], Tm[1], Tm[2]}, Tv[3] = {Tm[3], Tm[4], Tm[5]}, Tw[3] = {Tm[6], Tm[7], Tm[8]};
          float k[3], l[3];
          for (int a = 0; a < 3; a++) { k[a] = pixfx * Tw[a] - Tu[a]; l[a] = pixfy * Tw[a] - Tv[a]; }
          float p[3] = {k[1] * l[2] - k[2] * l[1], k[2] * l[0] - k[0] * l[2], k[0] * l[1] - k[1] * l[0]};
          if (p[2] == 0.0) continue;
          float sx = p[0] / p[2], sy = p[1] / p[2];
          float rho3d = (sx * sx + sy * sy);
          float dx = s->means2D[2 * id] - pixfx, dy = s->means2D[2 * id + 1] - pixfy;
          float rho2d = FilterInvSquare * (dx * dx + dy * dy);
          float rho = fminf(rho3d, rho2d);
          float c_d = (rho3d <= rho2d) ? (sx * Tw[0] + sy * Tw[1]) + Tw[2] : Tw[2];
          if (c_d < near_n) continue;
          const float* no = s->normal_opacity + 4 * id;
          float opa = no[3];
          float power = -0.5f * rho;
          if (power > 0.0f) continue;
          const float G = expf(power);
          const float alpha = fminf(0.99f, opa * G);
          if (alpha < 1.0f / 255.0f) continue;
          T = T / (1.f - alpha);
          const float dchannel_dcolor = alpha * T;
          float dL_dalpha = 0.0f;
          for (int ch = 0; ch < 3; ch++) {
            const float c = s->rgb[3 * id + ch];
            accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
            last_color[ch] = c;
            dL_dalpha += (c - accum_rec[ch]) * dL_dpixel[ch];
            atomic_addf(&dL_dcolor[3 * id + ch], dchannel_dcolor * dL_dpixel[ch]);
          }
          float dL_dz = 0.0f, dL_dweight = 0;
          const float m_d = far_n / (far_n - near_n) * (1 - near_n / c_d);
          const float dmd_dd = (far_n * near_n) / ((far_n - near_n) * c_d * c_d);
          if (contributor == (uint32_t)(median_contributor - 1)) dL_dz += dL_dmedian_depth;
          dL_dweight += (final_D2 + m_d * m_d * final_A - 2 * m_d * final_D) * dL_dreg;
          dL_dalpha += dL_dweight - last_dL_dT;
          last_dL_dT = dL_dweight * alpha + (1 - alpha) * last_dL_dT;
          const float dL_dmd = 2.0f * (T * alpha) * (m_d * final_A - final_D) * dL_dreg;
          dL_dz += dL_dmd * dmd_dd;
          accum_depth_rec = last_alpha * last_depth + (1.f - last_alpha) * accum_depth_rec;
          last_depth = c_d;
          dL_dalpha += (c_d - accum_depth_rec) * dL_ddepth;
          accum_alpha_rec = (float)(last_alpha * 1.0 + (1.f - last_alpha) * accum_alpha_rec);
          dL_dalpha += (1 - accum_alpha_rec) * dL_daccum;
          for (int ch = 0; ch < 3; ch++) {
            accum_normal_rec[ch] = last_alpha * last_normal[ch] + (1.f - last_alpha) * accum_normal_rec[ch];
            last_normal[ch] = no[ch];
            dL_dalpha += (no[ch] - accum_normal_rec[ch]) * dL_dnormal2D[ch];
            atomic_addf(&dL_dnormal[3 * id + ch], alpha * T * dL_dnormal2D[ch]);
          }
          dL_dalpha *= T;
          last_alpha = alpha;
          float bg_dot_dpixel = 0;
          for (int i = 0; i < 3; i++) bg_dot_dpixel += bg[i] * dL_dpixel[i];
          dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
          const float dL_dG = opa * dL_dalpha;
          dL_dz += alpha * T * dL_ddepth;
          if (rho3d <= rho2d) {
            const float dsx = dL_dG * -G * sx + dL_dz * Tw[0], dsy = dL_dG * -G * sy + dL_dz * Tw[1];
            const float dz_dTw[3] = {sx, sy, 1.0f};
            const float dsx_pz = dsx / p[2], dsy_pz = dsy / p[2];
            const float dp[3] = {dsx_pz, dsy_pz, -(dsx_pz * sx + dsy_pz * sy)};
            const float dk[3] = {l[1] * dp[2] - l[2] * dp[1], l[2] * dp[0] - l[0] * dp[2], l[0] * dp[1] - l[1] * dp[0]};
            const float dl[3] = {dp[1] * k[2] - dp[2] * k[1], dp[2] * k[0] - dp[0] * k[2], dp[0] * k[1] - dp[1] * k[0]};
            for (int a = 0; a < 3; a++) {
              atomic_addf(&dL_dtransMat[9 * id + a], -dk[a]);
              atomic_addf(&dL_dtransMat[9 * id + 3 + a], -dl[a]);
              atomic_addf(&dL_dtransMat[9 * id + 6 + a], pixfx * dk[a] + pixfy * dl[a] + dL_dz * dz_dTw[a]);
            }
          } else {
            const float dG_ddelx = -G * FilterInvSquare * dx, dG_ddely = -G * FilterInvSquare * dy;
            atomic_addf(&dL_dmean2D[3 * id + 0], dL_dG * dG_ddelx);
            atomic_addf(&dL_dmean2D[3 * id + 1], dL_dG * dG_ddely);
            atomic_addf(&dL_dtransMat[9 * id + 8], dL_dz);
          }
          atomic_addf(&dL_dopacity[id], G * dL_dalpha);
        }
      }
  }

  /* ---- BACKWARD::preprocessCUDA (backward.cu:575-630) ---- */
  const int Wd = (int)(focal_x * tan_fovx * 2), Hd = (int)(focal_y * tan_fovy * 2);
#pragma omp parallel for schedule(static)
  for (int idx = 0; idx < P; idx++) {
    if (!(s->radii[idx] > 0)) continue;
    const float* p = means3D + 3 * idx;
    const float* rot = rotations + 4 * idx;
    const float* sc = scales + 2 * idx;
    float Rm[3][3];
    quat_to_rotmat(rot, Rm);
    float L[3][3];
    for (int r = 0; r < 3; r++) { L[0][r] = Rm[0][r] * (1.0f * sc[0]); L[1][r] = Rm[1][r] * (1.0f * sc[1]); L[2][r] = Rm[2][r]; }
    /* M (3 cols x 4 rows), Pm = world2ndc * ndc2pix (3 cols x 4 rows) */
    float Mm[3][4] = {{L[0][0], L[0][1], L[0][2], 0.f}, {L[1][0], L[1][1], L[1][2], 0.f}, {p[0], p[1], p[2], 1.f}};
    float n0[4] = {(float)((float)Wd / 2.0), 0.f, 0.f, (float)((float)(Wd - 1) / 2.0)};
    float n1[4] = {0.f, (float)((float)Hd / 2.0), 0.f, (float)((float)(Hd - 1) / 2.0)};
    float n2[4] = {0.f, 0.f, 0.f, 1.f};
    const float* nn[3] = {n0, n1, n2};
    float Pm[3][4];
    for (int i = 0; i < 3; i++)
      for (int r = 0; r < 4; r++) {
        /* world2ndc column j = (proj[j], proj[4+j], proj[8+j], proj[12+j]); row r of column j = proj[4*r + j] */
        float acc = projmatrix[4 * r + 0] * nn[i][0];
        acc += projmatrix[4 * r + 1] * nn[i][1];
        acc += projmatrix[4 * r + 2] * nn[i][2];
        acc += projmatrix[4 * r + 3] * nn[i][3];
        Pm[i][r] = acc;
      }
    float T[3][3];
    for (int i = 0; i < 3; i++)
      for (int r = 0; r < 3; r++) {
        /* transpose(M) column j (j=0..3) = (M[0][j], M[1][j], M[2][j]) */
        float acc = Mm[r][0] * Pm[i][0];
        acc += Mm[r][1] * Pm[i][1];
        acc += Mm[r][2] * Pm[i][2];
        acc += Mm[r][3] * Pm[i][3];
        T[i][r] = acc;
      }
    float normal[3];
    for (int r = 0; r < 3; r++) normal[r] = viewmatrix[r] * L[2][0] + viewmatrix[4 + r] * L[2][1] + viewmatrix[8 + r] * L[2][2];

    float dT[3][3];
    for (int i = 0; i < 3; i++)
      for (int r = 0; r < 3; r++) dT[i][r] = dL_dtransMat[9 * idx + 3 * i + r];
    const float m2x = dL_dmean2D[3 * idx], m2y = dL_dmean2D[3 * idx + 1];
    if (m2x != 0 || m2y != 0) {
      const float t[3] = {9.0f, 9.0f, -1.0f};
      float d = t[0] * (T[2][0] * T[2][0]); d += t[1] * (T[2][1] * T[2][1]); d += t[2] * (T[2][2] * T[2][2]);
      float f[3], dT0[3], dT1[3], dT3[3], df[3];
      for (int a = 0; a < 3; a++) f[a] = t[a] * (1.0f / d);
      for (int a = 0; a < 3; a++) {
        dT0[a] = m2x * f[a] * T[2][a];
        dT1[a] = m2y * f[a] * T[2][a];
        dT3[a] = m2x * f[a] * T[0][a] + m2y * f[a] * T[1][a];
        df[a] = m2x * T[0][a] * T[2][a] + m2y * T[1][a] * T[2][a];
      }
      float dd = (float)((df[0] * f[0] + df[1] * f[1] + df[2] * f[2]) * (-1.0 / d));
      for (int a = 0; a < 3; a++) {
        dT3[a] += dd * (t[a] * T[2][a] * 2.0f);
        dT[0][a] += dT0[a]; dT[1][a] += dT1[a]; dT[2][a] += dT3[a];
      }
    }
    /* dL_dM = Pm * transpose(dT): 3 cols x 4 rows; column i = sum_j Pm[j] * dT[j][i] */
    float dM[3][4];
    for (int i = 0; i < 3; i++)
      for (int r = 0; r < 4; r++) {
        float acc = Pm[0][r] * dT[0][i];
        acc += Pm[1][r] * dT[1][i];
        acc += Pm[2][r] * dT[2][i];
        dM[i][r] = acc;
      }
    const float* dn = dL_dnormal + 3 * idx;
    float dtn[3] = {viewmatrix[0] * dn[0] + viewmatrix[1] * dn[1] + viewmatrix[2] * dn[2],
                    viewmatrix[4] * dn[0] + viewmatrix[5] * dn[1] + viewmatrix[6] * dn[2],
                    viewmatrix[8] * dn[0] + viewmatrix[9] * dn[1] + viewmatrix[10] * dn[2]};
    float pv[3];
    for (int r = 0; r < 3; r++)
      pv[r] = viewmatrix[r] * p[0] + viewmatrix[4 + r] * p[1] + viewmatrix[8 + r] * p[2] + viewmatrix[12 + r];
    float cosv = -(pv[0] * normal[0] + pv[1] * normal[1] + pv[2] * normal[2]);
    float mult = cosv > 0 ? 1 : -1;
    for (int a = 0; a < 3; a++) dtn[a] *= mult;
    float dRS[3][3] = {{dM[0][0], dM[0][1], dM[0][2]}, {dM[1][0], dM[1][1], dM[1][2]}, {dtn[0], dtn[1], dtn[2]}};
    float vR[3][3];
    for (int a = 0; a < 3; a++) { vR[0][a] = dRS[0][a] * sc[0]; vR[1][a] = dRS[1][a] * sc[1]; vR[2][a] = dRS[2][a]; }
    {
      float sn = 1.0f / sqrtf(rot[3] * rot[3] + rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2]);
      float w = rot[0] * sn, x = rot[1] * sn, y = rot[2] * sn, z = rot[3] * sn;
      float* q = dL_drot + 4 * idx;
      q[0] = 2.f * (x * (vR[1][2] - vR[2][1]) + y * (vR[2][0] - vR[0][2]) + z * (vR[0][1] - vR[1][0]));
      q[1] = 2.f * (-2.f * x * (vR[1][1] + vR[2][2]) + y * (vR[0][1] + vR[1][0]) + z * (vR[0][2] + vR[2][0]) + w * (vR[1][2] - vR[2][1]));
      q[2] = 2.f * (x * (vR[0][1] + vR[1][0]) - 2.f * y * (vR[0][0] + vR[2][2]) + z * (vR[1][2] + vR[2][1]) + w * (vR[2][0] - vR[0][2]));
      q[3] = 2.f * (x * (vR[0][2] + vR[2][0]) + y * (vR[1][2] + vR[2][1]) - 2.f * z * (vR[0][0] + vR[1][1]) + w * (vR[0][1] - vR[1][0]));
    }
    dL_dscale[2 * idx] = dRS[0][0] * Rm[0][0] + dRS[0][1] * Rm[0][1] + dRS[0][2] * Rm[0][2];
    dL_dscale[2 * idx + 1] = dRS[1][0] * Rm[1][0] + dRS[1][1] * Rm[1][1] + dRS[1][2] * Rm[1][2];
    for (int a = 0; a < 3; a++) dL_dmean3D[3 * idx + a] = dM[2][a];

    if (shs) {
      /* SH backward (backward.cu:20-139) */
      float dir_orig[3] = {p[0] - campos[0], p[1] - campos[1], p[2] - campos[2]};
      float len = sqrtf(dir_orig[0] * dir_orig[0] + dir_orig[1] * dir_orig[1] + dir_orig[2] * dir_orig[2]);
      float x = dir_orig[0] / len, y = dir_orig[1] / len, z = dir_orig[2] / len;
      const float* sh = shs + (size_t)idx * M * 3;
      float* dsh = dL_dsh + (size_t)idx * M * 3;
      float dRGB[3], ddir[3] = {0, 0, 0};
      for (int c = 0; c < 3; c++) dRGB[c] = dL_dcolor[3 * idx + c] * (s->clamped[3 * idx + c] ? 0.f : 1.f);
      for (int c = 0; c < 3; c++) {
#define SH(i) sh[(i) * 3 + c]
#define DSH(i) dsh[(i) * 3 + c]
        float dx_ = 0, dy_ = 0, dz_ = 0;
        DSH(0) = SH_C0 * dRGB[c];
        if (D > 0) {
          DSH(1) = (-SH_C1 * y) * dRGB[c]; DSH(2) = (SH_C1 * z) * dRGB[c]; DSH(3) = (-SH_C1 * x) * dRGB[c];
          dx_ = -SH_C1 * SH(3); dy_ = -SH_C1 * SH(1); dz_ = SH_C1 * SH(2);
          if (D > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            DSH(4) = (SH_C2[0] * xy) * dRGB[c]; DSH(5) = (SH_C2[1] * yz) * dRGB[c];
            DSH(6) = (SH_C2[2] * (2.f * zz - xx - yy)) * dRGB[c]; DSH(7) = (SH_C2[3] * xz) * dRGB[c];
            DSH(8) = (SH_C2[4] * (xx - yy)) * dRGB[c];
            dx_ += SH_C2[0] * y * SH(4) + SH_C2[2] * 2.f * -x * SH(6) + SH_C2[3] * z * SH(7) + SH_C2[4] * 2.f * x * SH(8);
            dy_ += SH_C2[0] * x * SH(4) + SH_C2[1] * z * SH(5) + SH_C2[2] * 2.f * -y * SH(6) + SH_C2[4] * 2.f * -y * SH(8);
            dz_ += SH_C2[1] * y * SH(5) + SH_C2[2] * 2.f * 2.f * z * SH(6) + SH_C2[3] * x * SH(7);
            if (D > 2) {
              DSH(9) = (SH_C3[0] * y * (3.f * xx - yy)) * dRGB[c]; DSH(10) = (SH_C3[1] * xy * z) * dRGB[c];
              DSH(11) = (SH_C3[2] * y * (4.f * zz - xx - yy)) * dRGB[c];
              DSH(12) = (SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * dRGB[c];
              DSH(13) = (SH_C3[4] * x * (4.f * zz - xx - yy)) * dRGB[c]; DSH(14) = (SH_C3[5] * z * (xx - yy)) * dRGB[c];
              DSH(15) = (SH_C3[6] * x * (xx - 3.f * yy)) * dRGB[c];
              dx_ += (SH_C3[0] * SH(9) * 3.f * 2.f * xy + SH_C3[1] * SH(10) * yz + SH_C3[2] * SH(11) * -2.f * xy +
                      SH_C3[3] * SH(12) * -3.f * 2.f * xz + SH_C3[4] * SH(13) * (-3.f * xx + 4.f * zz - yy) +
                      SH_C3[5] * SH(14) * 2.f * xz + SH_C3[6] * SH(15) * 3.f * (xx - yy));
              dy_ += (SH_C3[0] * SH(9) * 3.f * (xx - yy) + SH_C3[1] * SH(10) * xz +
                      SH_C3[2] * SH(11) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * SH(12) * -3.f * 2.f * yz +
                      SH_C3[4] * SH(13) * -2.f * xy + SH_C3[5] * SH(14) * -2.f * yz + SH_C3[6] * SH(15) * -3.f * 2.f * xy);
              dz_ += (SH_C3[1] * SH(10) * xy + SH_C3[2] * SH(11) * 4.f * 2.f * yz +
                      SH_C3[3] * SH(12) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * SH(13) * 4.f * 2.f * xz +
                      SH_C3[5] * SH(14) * (xx - yy));
            }
          }
        }
#undef SH
#undef DSH
        ddir[0] += dx_ * dRGB[c]; ddir[1] += dy_ * dRGB[c]; ddir[2] += dz_ * dRGB[c];
      }
      /* dnormvdv (auxiliary.h:128-138) */
      const float* v = dir_orig;
      float sum2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
      float inv = 1.0f / sqrtf(sum2 * sum2 * sum2);
      dL_dmean3D[3 * idx + 0] += ((+sum2 - v[0] * v[0]) * ddir[0] - v[1] * v[0] * ddir[1] - v[2] * v[0] * ddir[2]) * inv;
      dL_dmean3D[3 * idx + 1] += (-v[0] * v[1] * ddir[0] + (sum2 - v[1] * v[1]) * ddir[1] - v[2] * v[1] * ddir[2]) * inv;
      dL_dmean3D[3 * idx + 2] += (-v[0] * v[2] * ddir[0] - v[1] * v[2] * ddir[1] + (sum2 - v[2] * v[2]) * ddir[2]) * inv;
    }
    /* densification hack (backward.cu:626-629) */
    float depth = s->transMat[9 * idx + 8];
    dL_dmean2D[3 * idx + 0] = (float)(dL_dtransMat[9 * idx + 2] * depth * 0.5 * (float)Wd);
    dL_dmean2D[3 * idx + 1] = (float)(dL_dtransMat[9 * idx + 5] * depth * 0.5 * (float)Hd);
  }
}
