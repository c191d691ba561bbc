// Forward per-tile alpha blending of 2D-Gaussian surfels (sm_100a).
//
// Replaces reference renderCUDA (cuda_rasterizer/forward.cu:256-441): ray-splat
// intersection, low-pass filter, alpha compositing of colour / depth / normal /
// distortion / median depth, per-pixel state for the backward pass.
//
// B200 design (not in the reference):
//   * one CTA (256 threads, 8 warps) per 16x16 tile, each warp owns an 8x4 pixel
//     footprint;
//   * the tile's depth-sorted surfel list is consumed in rounds of 256 candidates:
//     each thread fetches one candidate's cull box (16 B), candidates whose box
//     misses the tile are dropped by a ballot compaction, the survivors' 80-byte
//     records are copied to shared memory with 128-bit cp.async (LDGSTS);
//   * each warp then tests 32 survivors at a time against its own 8x4 footprint
//     and only walks the set bits of the ballot, so a surfel costs ALU work only
//     in the warps it can actually touch;
//   * warp-level and CTA-level early termination votes.
// Per-fragment arithmetic keeps the reference's expression order exactly, so the
// same fragments are blended and images match bit-for-bit up to FMA contraction.
#include "common.cuh"
#include "kernels.h"

namespace pgs {

constexpr int RF_ROUND = TILE_PIX;  // candidates fetched per round (one per thread)

struct __align__(16) StagedSurfel {
  float4 q[REC_QUADS];
};

__global__ void __launch_bounds__(TILE_PIX) render_fwd_kernel(RenderFwdArgs a) {
  __shared__ StagedSurfel s_rec[RF_ROUND];
  __shared__ float4 s_box[RF_ROUND];
  __shared__ uint32_t s_pos[RF_ROUND];  // 1-based position in the tile list (the reference's `contributor`)
  __shared__ int s_warp_cnt[TILE_PIX / 32];

  const int tid = threadIdx.x;
  const unsigned lane = tid & 31, wid = tid >> 5;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile_id = tile_y * a.grid_x + tile_x;
  // warp footprint: 8x4 pixels
  const int fx0 = tile_x * TILE_X + (wid & 1) * WARP_FX;
  const int fy0 = tile_y * TILE_Y + (wid >> 1) * WARP_FY;
  const uint2 pix = {(unsigned)(fx0 + (lane & 7)), (unsigned)(fy0 + (lane >> 3))};
  const float2 pixf = {(float)pix.x, (float)pix.y};
  const bool inside = pix.x < (unsigned)a.W && pix.y < (unsigned)a.H;
  bool done = !inside;

  // tile / footprint boxes in pixel-centre coordinates (integer centres, base fork)
  const float tx0 = (float)(tile_x * TILE_X), ty0 = (float)(tile_y * TILE_Y);
  const float tx1 = tx0 + (float)(TILE_X - 1), ty1 = ty0 + (float)(TILE_Y - 1);
  const float wx0 = (float)fx0, wy0 = (float)fy0;
  const float wx1 = wx0 + (float)(WARP_FX - 1), wy1 = wy0 + (float)(WARP_FY - 1);

  const uint2 range = a.ranges[tile_id];
  const int total = (int)(range.y - range.x);

  float T = 1.0f;
  uint32_t last_contributor = 0;
  float C[3] = {0.f, 0.f, 0.f};
  float N[3] = {0.f, 0.f, 0.f};
  float D = 0.f, M1 = 0.f, M2 = 0.f, distortion = 0.f, median_depth = 0.f;
  float median_contributor = -1.f;

  for (int start = 0; start < total; start += RF_ROUND) {
    // CTA-wide early exit once every pixel is saturated
    if (__syncthreads_count(done) == TILE_PIX) break;

    // ---- stage: fetch candidate, cull against the tile, compact -------------
    const int cand = start + tid;
    bool keep = false;
    uint32_t id = 0;
    float4 box;
    if (cand < total) {
      id = a.point_list[range.x + cand];
      box = __ldg(&a.bbox[id]);
      keep = !(box.x > tx1 || box.z < tx0 || box.y > ty1 || box.w < ty0);
    }
    const unsigned kb = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp_cnt[wid] = __popc(kb);
    __syncthreads();
    int slot = __popc(kb & ((1u << lane) - 1));
    int nkept = 0;
#pragma unroll
    for (int w = 0; w < TILE_PIX / 32; w++) {
      int c = s_warp_cnt[w];
      if (w < (int)wid) slot += c;
      nkept += c;
    }
    if (keep) {
      const float4* src = a.rec + (size_t)id * REC_QUADS;
#pragma unroll
      for (int q = 0; q < REC_QUADS; q++) cp_async16(&s_rec[slot].q[q], src + q);
      s_box[slot] = box;
      s_pos[slot] = (uint32_t)(cand + 1);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // ---- blend: 32 survivors at a time, walk only those touching this warp --
    for (int c0 = 0; c0 < nkept; c0 += 32) {
      if (__all_sync(0xffffffffu, done)) break;
      bool hit = false;
      if (c0 + (int)lane < nkept) {
        const float4 b = s_box[c0 + lane];
        hit = !(b.x > wx1 || b.z < wx0 || b.y > wy1 || b.w < wy0);
      }
      unsigned m = __ballot_sync(0xffffffffu, hit);
      while (m) {
        const int j = c0 + __ffs(m) - 1;
        m &= m - 1;
        if (done) continue;

        const float4 q0 = s_rec[j].q[0];
        const float4 q1 = s_rec[j].q[1];
        const float4 q2 = s_rec[j].q[2];
        const float2 xy = {q0.w, q1.w};
        const float3 Tu = {q0.x, q0.y, q0.z};
        const float3 Tv = {q1.x, q1.y, q1.z};
        const float3 Tw = {q2.x, q2.y, q2.z};
        // two homogeneous planes through the pixel, intersected with the splat
        float3 k = {pixf.x * Tw.x - Tu.x, pixf.x * Tw.y - Tu.y, pixf.x * Tw.z - Tu.z};
        float3 l = {pixf.y * Tw.x - Tv.x, pixf.y * Tw.y - Tv.y, pixf.y * Tw.z - Tv.z};
        float3 p = {k.y * l.z - k.z * l.y, k.z * l.x - k.x * l.z, k.x * l.y - k.y * l.x};
        if (p.z == 0.0) continue;
        float2 s = {p.x / p.z, p.y / p.z};
        float rho3d = (s.x * s.x + s.y * s.y);
        float2 d = {xy.x - pixf.x, xy.y - pixf.y};
        float rho2d = PGS_FILTER_INV_SQUARE * (d.x * d.x + d.y * d.y);

        float rho = min(rho3d, rho2d);
        float depth = (rho3d <= rho2d) ? (s.x * Tw.x + s.y * Tw.y) + Tw.z : Tw.z;
        if (depth < PGS_NEAR_N) continue;
        const float opa = q2.w;

        float power = -0.5f * rho;
        if (power > 0.0f) continue;

        float alpha = min(0.99f, opa * exp(power));
        if (alpha < 1.0f / 255.0f) continue;
        float test_T = T * (1 - alpha);
        if (test_T < 0.0001f) {
          done = true;
          continue;
        }

        const float4 q3 = s_rec[j].q[3];
        const float4 q4 = s_rec[j].q[4];
        float w = alpha * T;
        float A = 1 - T;
        float mm = PGS_FAR_N / (PGS_FAR_N - PGS_NEAR_N) * (1 - PGS_NEAR_N / depth);
        distortion += (mm * mm * A + M2 - 2 * mm * M1) * w;
        D += depth * w;
        M1 += mm * w;
        M2 += mm * mm * w;

        const uint32_t contributor = s_pos[j];
        if (T > 0.5) {
          median_depth = depth;
          median_contributor = contributor;
        }
        N[0] += q3.x * w;
        N[1] += q3.y * w;
        N[2] += q3.z * w;
        C[0] += q4.x * w;
        C[1] += q4.y * w;
        C[2] += q4.z * w;
        T = test_T;
        last_contributor = contributor;
      }
    }
  }

  // per-pixel state for backward, tile-major so that a warp writes 128 contiguous bytes
  const size_t npt = (size_t)a.grid_x * a.grid_y * TILE_PIX;
  const size_t sidx = (size_t)tile_id * TILE_PIX + tid;
  a.final_T[sidx] = T;
  a.final_T[sidx + npt] = M1;
  a.final_T[sidx + 2 * npt] = M2;
  a.n_contrib[sidx] = last_contributor;
  // The reference converts its float -1 sentinel to uint32 (undefined in C++, garbage in
  // practice) for pixels nothing was blended into; those pixels have n_contrib == 0 and the
  // value is never consulted.  Store a defined 0 instead.
  a.n_contrib[sidx + npt] = median_contributor < 0.f ? 0u : (uint32_t)median_contributor;

  if (inside) {
    const size_t HW = (size_t)a.H * a.W;
    const size_t pix_id = (size_t)a.W * pix.y + pix.x;
    for (int ch = 0; ch < 3; ch++) a.out_color[ch * HW + pix_id] = C[ch] + T * a.bg_color[ch];
    a.out_others[pix_id + DEPTH_OFFSET * HW] = D;
    a.out_others[pix_id + ALPHA_OFFSET * HW] = 1 - T;
    for (int ch = 0; ch < 3; ch++) a.out_others[pix_id + (NORMAL_OFFSET + ch) * HW] = N[ch];
    a.out_others[pix_id + MIDDEPTH_OFFSET * HW] = median_depth;
    a.out_others[pix_id + DISTORTION_OFFSET * HW] = distortion;
  }
}

void launch_render_fwd(const RenderFwdArgs& a, cudaStream_t s) {
  dim3 grid(a.grid_x, a.grid_y, 1);
  render_fwd_kernel<<<grid, TILE_PIX, 0, s>>>(a);
  count_launch();
}

}  // namespace pgs
