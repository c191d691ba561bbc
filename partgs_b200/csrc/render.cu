// Per-tile alpha blending of 2D-Gaussian surfels, forward and backward, for both forks
// (sm_100a).
//
// Replaces reference renderCUDA forward (cuda_rasterizer/forward.cu:256-441; `_part` fork
// DSRP/cuda_rasterizer/forward.cu:265-474) and backward (backward.cu:143-440; `_part`
// :143-471).  The reference runs one 16x16 CTA per tile in lock-step rounds of 256 surfels
// (three __syncthreads per round, every pixel visits every surfel binned to the tile) and
// issues 16 global float atomics per pixel x surfel fragment in the backward pass.
//
// B200 design:
//   * a tile is still one 256-thread CTA, but its 8 warps are independent streams: each warp
//     owns an 8x4 pixel footprint and walks the tile's depth-sorted list on its own, 32
//     candidates per step — one candidate per lane: fetch id (coalesced) and the surfel's
//     16-byte cull box, test the box (then the exact conic) against the warp's footprint,
//     ballot-compact the survivors and stage only their 80-byte records into the warp's
//     private shared-memory ring with 128-bit cp.async (LDGSTS).  The ring is double
//     buffered: the survivors of step k+1 are in flight while those of step k are blended;
//     ids are prefetched two steps ahead, cull boxes one.  No __syncthreads anywhere in the
//     loop, so a warp never waits for a sibling with more work, and a surfel costs ALU time
//     only in the warps whose footprint it can touch;
//   * tiles are issued longest-list-first (tile_order, built by tile_order_kernel) so the
//     heavy tiles do not form the tail of the launch (giving the deepest tiles SMs of their own, on a forked
//     high-priority stream, was measured too: no gain — their warps are bound by their own dependency chains);
//   * warp-level early termination (all 32 pixels saturated) in forward;
//   * the forward pass records, per warp and list position, the 32-bit mask of pixels that
//     actually blended the surfel (`frag_mask`, 4 B per warp x instance, written coalesced).
//     The backward pass walks the same list back to front but never culls or re-decides
//     anything: a candidate is staged iff its mask is non-zero, and a lane takes part iff its
//     bit is set.  That removes the replay of the reference's decision chain (which would have
//     to be bit-identical to forward, i.e. IEEE divisions and precise expf) from the backward
//     pass: its per-fragment values are recomputed with MUFU.RCP / MUFU.EX2 (gradients are
//     gated at 1e-4 relative, measured ~1e-6);
//   * backward: the 18 (+S) per-fragment gradient components are summed over the 32 pixels
//     through the warp's shared memory (column-wise stores, four rotated 128-bit loads and one
//     shuffle per lane pair; two colour sums take a register butterfly); 18 lanes then hold one
//     finished component each and issue one red.global.add.f32 into the surfel's 80-byte record.
// The forward per-fragment arithmetic is pinned in frag_math.cuh; accumulations below use the
// reference's rounding sequence (explicit fma/mul), so images are bit-identical.
#include "common.cuh"
#include "frag_math.cuh"
#include "kernels.h"
#include <cstdlib>

namespace pgs {

constexpr unsigned RFULL = 0xffffffffu;
constexpr int NWARP = TILE_PIX / 32;
constexpr int CHUNK = 32;  // candidates per warp step (one per lane)
// Fragments evaluated ahead of the in-order blend.  Measured on C3: 2 -> 1.11 ms, 4 -> 1.14-1.23 ms, 8 -> 1.37-1.43 ms
// (a deeper look-ahead evaluates up to ILP-1 slots past the end of every step and costs registers).
constexpr int FWD_ILP = 2;

struct __align__(16) WarpStage {
  float4 rec[CHUNK][REC_QUADS];  // 32 x 80 B
  uint32_t pos[CHUNK];           // list position of the staged surfel
  uint32_t id[CHUNK];            // surfel index (backward only)
  uint32_t mask[CHUNK];          // pixels that blended it in forward (backward only)
};

__device__ __forceinline__ bool box_hits(const float4 b, float x0, float y0, float x1, float y1) {
  return !(b.x > x1 || b.z < x0 || b.y > y1 || b.w < y0);
}

// Second-level cull (after the box): can any pixel centre of the footprint [x0,x1]x[y0,y1] pass
// the reference's alpha >= 1/255 test?  Only if the footprint meets {f <= 0} (f = the surfel's
// conic, see common.cuh) or the low-pass disc around (cx, cy).  The minimum of the quadratic f
// over the rectangle is attained at the interior critical point (convex case) or on an edge.
__device__ __forceinline__ bool conic_hits(const float4 c1, const float4 c2, float x0, float y0, float x1, float y1) {
  const float a = c1.x, b = c1.y, c = c1.z, d = c1.w, e = c2.x, g = c2.y;
  const float X0 = x0 - c2.z, X1 = x1 - c2.z, Y0 = y0 - c2.w, Y1 = y1 - c2.w;
  // low-pass disc: distance from the centre to the rectangle
  const float ddx = fmaxf(fmaxf(X0, -X1), 0.f), ddy = fmaxf(fmaxf(Y0, -Y1), 0.f);
  if (ddx * ddx + ddy * ddy <= PGS_LOWPASS_RADIUS * PGS_LOWPASS_RADIUS) return true;
  auto f = [&](float X, float Y) { return (a * X + 2.f * (b * Y + d)) * X + (c * Y + 2.f * e) * Y + g; };
  float fmin_ = fminf(fminf(f(X0, Y0), f(X1, Y0)), fminf(f(X0, Y1), f(X1, Y1)));
  // edges Y = const: a X^2 + 2 (bY + d) X + ...,  vertex at X = -(bY + d)/a when a > 0
  if (a > 0.f) {
    const float ia = rcp_approx(a);
    const float xa = fminf(fmaxf(-(b * Y0 + d) * ia, X0), X1), xb = fminf(fmaxf(-(b * Y1 + d) * ia, X0), X1);
    fmin_ = fminf(fmin_, fminf(f(xa, Y0), f(xb, Y1)));
  }
  if (c > 0.f) {
    const float ic = rcp_approx(c);
    const float ya = fminf(fmaxf(-(b * X0 + e) * ic, Y0), Y1), yb = fminf(fmaxf(-(b * X1 + e) * ic, Y0), Y1);
    fmin_ = fminf(fmin_, fminf(f(X0, ya), f(X1, yb)));
  }
  // interior critical point (global minimum when the quadratic part is positive definite)
  const float det = a * c - b * b;
  if (a > 0.f && det > 0.f) {
    const float idet = rcp_approx(det);
    const float xc = (b * e - c * d) * idet, yc = (b * d - a * e) * idet;
    if (xc >= X0 && xc <= X1 && yc >= Y0 && yc <= Y1) fmin_ = fminf(fmin_, f(xc, yc));
  }
  // slack for the fp32 evaluation: 1e-5 x (sum of |terms| at the farthest corner); coefficients are
  // normalised to max(|a|,|b|,|c|) = 1.  NaN compares false -> keep the surfel.
  const float mx = fmaxf(fmaxf(fabsf(X0), fabsf(X1)), fmaxf(fabsf(Y0), fabsf(Y1)));
  const float slack = 0.05f + 1e-5f * (4.f * mx * mx + 2.f * (fabsf(d) + fabsf(e)) * mx + fabsf(g));
  return !(fmin_ > slack);
}

// =============================================================================
// tile order: longest list first (approximate LPT by counting sort on len/16)
// =============================================================================
constexpr int TO_BUCKETS = 1024;
__global__ void __launch_bounds__(1024) tile_order_kernel(const uint2* __restrict__ ranges, int ntiles,
                                                          uint32_t* __restrict__ order) {
  __shared__ uint32_t s_cnt[TO_BUCKETS];
  __shared__ uint32_t s_warp[32];
  const int tid = threadIdx.x;
  s_cnt[tid] = 0;
  __syncthreads();
  for (int t = tid; t < ntiles; t += blockDim.x) {
    const uint2 r = ranges[t];
    const uint32_t b = TO_BUCKETS - 1 - min((r.y - r.x) >> 4, (uint32_t)(TO_BUCKETS - 1));  // long lists -> small bucket
    atomicAdd(&s_cnt[b], 1u);
  }
  __syncthreads();
  // exclusive scan of the 1024 bucket counts
  const uint32_t c = s_cnt[tid];
  uint32_t w = c;
  const unsigned lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(RFULL, w, o);
    if (lane >= (unsigned)o) w += t;
  }
  if (lane == 31) s_warp[wid] = w;
  __syncthreads();
  uint32_t off = 0;
  for (int i = 0; i < 32; i++)
    if (i < (int)wid) off += s_warp[i];
  __syncthreads();
  s_cnt[tid] = off + w - c;
  __syncthreads();
  for (int t = tid; t < ntiles; t += blockDim.x) {
    const uint2 r = ranges[t];
    const uint32_t b = TO_BUCKETS - 1 - min((r.y - r.x) >> 4, (uint32_t)(TO_BUCKETS - 1));
    order[atomicAdd(&s_cnt[b], 1u)] = (uint32_t)t;
  }
}

void launch_tile_order(const uint2* ranges, int ntiles, uint32_t* order, cudaStream_t s) {
  if (ntiles <= 0) return;
  tile_order_kernel<<<1, 1024, 0, s>>>(ranges, ntiles, order);
  count_launch();
}

// =============================================================================
// forward
// =============================================================================
template <bool PART>
__global__ void __launch_bounds__(TILE_PIX, PART ? 2 : 4) render_fwd_kernel(RenderFwdArgs a) {
  // dynamic shared memory: WarpStage [2][nw], then (PART) the semantic staging [2][nw][CHUNK][MAX_SEMANTIC];
  // nw = warps per CTA (8 = whole tile; 4/2/1 when the image has too few tiles to fill the GPU)
  extern __shared__ __align__(16) unsigned char fwd_smem[];
  const int nw = blockDim.x >> 5;
  WarpStage* s_stage = reinterpret_cast<WarpStage*>(fwd_smem);
  float* s_sem_dyn = reinterpret_cast<float*>(fwd_smem + (size_t)2 * nw * sizeof(WarpStage));

  const int S = PART ? a.S : 0;
  const unsigned lane = threadIdx.x & 31, lw = threadIdx.x >> 5;  // lw: warp within the CTA
  const int groups = NWARP / nw;                                  // CTAs per tile
  const unsigned wid = (blockIdx.x % groups) * nw + lw;           // footprint (0..7) within the tile
  const int tid = wid * 32 + lane;                                // pixel slot within the tile
  const unsigned lt_mask = (1u << lane) - 1u;
  const int tile_slot = blockIdx.x / groups;
  const int tile_id = a.tile_order ? (int)a.tile_order[tile_slot] : tile_slot;
  const int tile_x = tile_id % a.grid_x, tile_y = tile_id / a.grid_x;
  const int fx0 = tile_x * TILE_X + (wid & 1) * WARP_FX;
  const int fy0 = tile_y * TILE_Y + (wid >> 1) * WARP_FY;
  const uint2 pix = {(unsigned)(fx0 + (lane & 7)), (unsigned)(fy0 + (lane >> 3))};
  const float poff = PART ? 0.5f : 0.0f;  // pixel centres: integer (base) / +0.5 (`_part`)
  const float2 pixf = {(float)pix.x + poff, (float)pix.y + poff};
  const bool inside = pix.x < (unsigned)a.W && pix.y < (unsigned)a.H;
  bool done = !inside;
  // Cull region: bounding box of the footprint's pixels that are still blending.  It starts as the
  // whole 8x4 footprint and shrinks as pixels saturate, so the long tail of a deep tile (a few
  // unsaturated pixels walking thousands of candidates) only evaluates surfels that can reach them.
  float wx0 = (float)fx0 + poff, wy0 = (float)fy0 + poff;
  float wx1 = wx0 + (float)(WARP_FX - 1), wy1 = wy0 + (float)(WARP_FY - 1);

  const uint2 range = a.ranges[tile_id];
  const int total = (int)(range.y - range.x);
  const uint32_t* __restrict__ list = a.point_list + range.x;
  uint32_t* __restrict__ fmask = a.frag_mask + (size_t)wid * a.mask_stride + range.x;

  float T = 1.0f;
  uint32_t last_contributor = 0;
  float C[3] = {0.f, 0.f, 0.f};
  float N[3] = {0.f, 0.f, 0.f};
  float D = 0.f, M1 = 0.f, M2 = 0.f, distortion = 0.f, median_depth = 0.f, median_weight = 0.f;
  float median_contributor = -1.f;
  float Sem[MAX_SEMANTIC];
  if (PART) {
#pragma unroll
    for (int i = 0; i < MAX_SEMANTIC; i++) Sem[i] = 0.f;
  }

  const float4 nobox = make_float4(3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f);
  auto load_id = [&](int base) -> uint32_t { return (base + (int)lane < total) ? list[base + lane] : 0u; };
  auto load_box = [&](int base, uint32_t id) -> float4 {
    return (base + (int)lane < total) ? __ldg(&a.bbox[(size_t)id * CULL_QUADS]) : nobox;  // lanes past the end never hit
  };
  // cull the 32 candidates of one step and start the copy of the survivors' records into ring buffer `buf`
  auto stage = [&](int base, int buf, uint32_t id, float4 box, bool& hit, int& slot) -> int {
    hit = box_hits(box, wx0, wy0, wx1, wy1);
    if (hit) {
      const float4* cr = a.bbox + (size_t)id * CULL_QUADS;
      hit = conic_hits(__ldg(cr + 1), __ldg(cr + 2), wx0, wy0, wx1, wy1);
    }
    const unsigned m = __ballot_sync(RFULL, hit);
    slot = __popc(m & lt_mask);
    if (hit) {
      WarpStage& st = s_stage[buf * nw + lw];
      const float4* src = a.rec + (size_t)id * REC_QUADS;
#pragma unroll
      for (int q = 0; q < REC_QUADS; q++) cp_async16(&st.rec[slot][q], src + q);
      st.pos[slot] = (uint32_t)(base + lane + 1);
      if (PART) {
        float* sem_stage = s_sem_dyn + ((size_t)(buf * nw + lw) * CHUNK + slot) * MAX_SEMANTIC;
        const float* sem = a.semantics + (size_t)id * S;
        for (int ch = 0; ch < S; ch++) sem_stage[ch] = __ldg(sem + ch);
      }
    }
    cp_async_commit();
    return __popc(m);
  };

  // software pipeline: step k is blended while the records of step k+1 are in flight, the cull box of
  // step k+2 and the ids of step k+3 are being fetched
  uint32_t id_n = load_id(0);
  float4 box_n = load_box(0, id_n);
  uint32_t id_n2 = load_id(CHUNK);
  bool hit_c = false;
  int slot_c = 0;
  int n_c = stage(0, 0, id_n, box_n, hit_c, slot_c);
  id_n = id_n2;
  box_n = load_box(CHUNK, id_n);
  id_n2 = load_id(2 * CHUNK);

  for (int base = 0, buf = 0; base < total; base += CHUNK, buf ^= 1) {
    const unsigned act = __ballot_sync(RFULL, !done);
    if (act == 0u) break;
    {
      const unsigned cols = (act | (act >> 8) | (act >> 16) | (act >> 24)) & 0xffu;
      wx0 = (float)(fx0 + (__ffs(cols) - 1)) + poff;
      wx1 = (float)(fx0 + (31 - __clz(cols))) + poff;
      wy0 = (float)(fy0 + ((__ffs(act) - 1) >> 3)) + poff;
      wy1 = (float)(fy0 + ((31 - __clz(act)) >> 3)) + poff;
    }
    bool hit_n = false;
    int slot_n = 0, n_n = 0;
    if (base + CHUNK < total) n_n = stage(base + CHUNK, buf ^ 1, id_n, box_n, hit_n, slot_n);
    else cp_async_commit();
    id_n = id_n2;
    box_n = load_box(base + 2 * CHUNK, id_n);
    id_n2 = load_id(base + 3 * CHUNK);
    cp_async_wait<1>();
    __syncwarp();

    uint32_t my_mask = 0;  // lane j: pixels that blended the survivor in slot j
    if (n_c > 0) {
      const WarpStage& st = s_stage[buf * nw + lw];
      const float* sem_stage = PART ? (s_sem_dyn + (size_t)(buf * nw + lw) * CHUNK * MAX_SEMANTIC) : nullptr;
      auto blend = [&](const int j, const float alpha, const float depth) -> bool {
        const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
        if (test_T < 0.0001f) {
          done = true;
          return false;
        }
        const float4 q3 = st.rec[j][3];
        const float4 q4 = st.rec[j][4];
        const uint32_t contributor = st.pos[j];
        const float A = __fsub_rn(1.0f, T);
        if (!PART) {
          const float w = __fmul_rn(alpha, T);
          const float mm = __fmul_rn(__fadd_rn(__fdiv_rn(-PGS_NEAR_N, depth), 1.0f),
                                     PGS_FAR_N / (PGS_FAR_N - PGS_NEAR_N));
          const float mm2 = __fmul_rn(mm, mm);
          const float err = __fmaf_rn(-M1, __fadd_rn(mm, mm), __fmaf_rn(A, mm2, M2));
          distortion = __fmaf_rn(err, w, distortion);
          D = __fmaf_rn(depth, w, D);
          M1 = __fmaf_rn(mm, w, M1);
          M2 = __fmaf_rn(mm2, w, M2);
          if (T > 0.5f) {
            median_depth = depth;
            median_contributor = (float)contributor;
          }
          N[0] = __fmaf_rn(q3.x, w, N[0]);
          N[1] = __fmaf_rn(q3.y, w, N[1]);
          N[2] = __fmaf_rn(q3.z, w, N[2]);
          C[0] = __fmaf_rn(q4.x, w, C[0]);
          C[1] = __fmaf_rn(q4.y, w, C[1]);
          C[2] = __fmaf_rn(q4.z, w, C[2]);
        } else {
          // `_part`: mapped depth in double, every accumulation is fma(T, x*alpha, acc)
          const double dd = (double)depth;
          const float md = (float)__ddiv_rn(__fma_rn(dd, 100.0, -20.0), __dmul_rn(dd, 100.0 - 0.2));
          const float md2 = __fmul_rn(md, md);
          const float err = __fmaf_rn(-M1, __fadd_rn(md, md), __fmaf_rn(A, md2, M2));
          distortion = __fmaf_rn(T, __fmul_rn(err, alpha), distortion);
          if (T > 0.5f) {
            median_depth = depth;
            median_weight = __fmul_rn(alpha, T);
            median_contributor = (float)contributor;
          }
          N[0] = __fmaf_rn(T, __fmul_rn(q3.x, alpha), N[0]);
          N[1] = __fmaf_rn(T, __fmul_rn(q3.y, alpha), N[1]);
          N[2] = __fmaf_rn(T, __fmul_rn(q3.z, alpha), N[2]);
          D = __fmaf_rn(T, __fmul_rn(depth, alpha), D);
          M1 = __fmaf_rn(T, __fmul_rn(md, alpha), M1);
          M2 = __fmaf_rn(T, __fmul_rn(md2, alpha), M2);
          C[0] = __fmaf_rn(T, __fmul_rn(q4.x, alpha), C[0]);
          C[1] = __fmaf_rn(T, __fmul_rn(q4.y, alpha), C[1]);
          C[2] = __fmaf_rn(T, __fmul_rn(q4.z, alpha), C[2]);
#pragma unroll
          for (int ch = 0; ch < MAX_SEMANTIC; ch++)
            if (ch < S) Sem[ch] = __fmaf_rn(T, __fmul_rn(sem_stage[j * MAX_SEMANTIC + ch], alpha), Sem[ch]);
        }
        T = test_T;
        last_contributor = contributor;
        return true;
      };
      // FWD_ILP fragments are evaluated (geometry, alpha: independent of the pixel's running state)
      // before they are blended in order.  The deepest tiles (thousands of fragments on the same
      // pixels, one warp alone on its scheduler at the end of the launch) are bound by the latency
      // of this chain, not by issue slots.
      for (int j = 0; j < n_c; j += FWD_ILP) {
        float alpha[FWD_ILP], depth[FWD_ILP];
        bool ok[FWD_ILP];
#pragma unroll
        for (int u = 0; u < FWD_ILP; u++) {
          const int ju = min(j + u, CHUNK - 1);
          ok[u] = frag_eval<PART>(pixf, st.rec[ju][0], st.rec[ju][1], st.rec[ju][2], alpha[u], depth[u]) && (j + u < n_c);
        }
#pragma unroll
        for (int u = 0; u < FWD_ILP; u++) {
          bool b = false;
          if (ok[u] && !done) b = blend(j + u, alpha[u], depth[u]);
          const unsigned m = __ballot_sync(RFULL, b);
          if ((int)lane == j + u) my_mask = m;
        }
      }
    }
    // what backward needs to know about this step: per candidate, the pixels that blended it
    const uint32_t mv = __shfl_sync(RFULL, my_mask, slot_c & 31);
    if (base + (int)lane < total) fmask[base + lane] = hit_c ? mv : 0u;
    __syncwarp();  // ring buffer `buf` is refilled by the next iteration's stage()
    n_c = n_n;
    hit_c = hit_n;
    slot_c = slot_n;
  }
  cp_async_wait<0>();

  // per-pixel state for backward, tile-major so that a warp writes 128 contiguous bytes
  const size_t npt = (size_t)a.grid_x * a.grid_y * TILE_PIX;
  const size_t sidx = (size_t)tile_id * TILE_PIX + tid;
  a.final_T[sidx] = T;
  a.final_T[sidx + npt] = M1;
  a.final_T[sidx + 2 * npt] = M2;
  a.n_contrib[sidx] = last_contributor;
  // The reference converts its float -1 sentinel to uint32 (undefined in C++, garbage in
  // practice) for pixels nothing was blended into; those have n_contrib == 0 and the value is
  // never consulted.  Store a defined 0.
  a.n_contrib[sidx + npt] = median_contributor < 0.f ? 0u : (uint32_t)median_contributor;

  if (inside) {
    const size_t HW = (size_t)a.H * a.W;
    const size_t pix_id = (size_t)a.W * pix.y + pix.x;
    for (int ch = 0; ch < 3; ch++) a.out_color[ch * HW + pix_id] = __fmaf_rn(T, a.bg_color[ch], C[ch]);
    a.out_others[pix_id + DEPTH_OFFSET * HW] = D;
    a.out_others[pix_id + ALPHA_OFFSET * HW] = 1 - T;
    for (int ch = 0; ch < 3; ch++) a.out_others[pix_id + (NORMAL_OFFSET + ch) * HW] = N[ch];
    a.out_others[pix_id + MIDDEPTH_OFFSET * HW] = median_depth;
    a.out_others[pix_id + DISTORTION_OFFSET * HW] = distortion;
    if (PART) {
      a.out_others[pix_id + MEDIAN_WEIGHT_OFFSET * HW] = median_weight;
#pragma unroll
      for (int ch = 0; ch < MAX_SEMANTIC; ch++)
        if (ch < S) a.out_semantic[ch * HW + pix_id] = Sem[ch];
    }
  }
}

// =============================================================================
// backward
// =============================================================================
// Sum v[0..15] over the warp; every lane returns component (lane >> 1).
__device__ __forceinline__ float warp_reduce16(float (&v)[16], unsigned lane) {
#pragma unroll
  for (int step = 0; step < 4; step++) {
    const int half = 8 >> step;
    const unsigned bit = 16u >> step;
    const bool hi = lane & bit;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (i < half) {
        const float keep = hi ? v[i + half] : v[i];
        const float send = hi ? v[i] : v[i + half];
        v[i] = keep + __shfl_xor_sync(RFULL, send, bit);
      }
    }
  }
  v[0] += __shfl_xor_sync(RFULL, v[0], 1);
  return v[0];
}
// Shared-memory plan of the backward kernel (dynamic): the double-buffered record ring of every warp, then
// one [RED_COLS][32] float transposition buffer per warp.
constexpr int RED_COLS = 16;
constexpr size_t bwd_smem_bytes(int nw) {
  return (size_t)2 * nw * sizeof(WarpStage) + (size_t)nw * RED_COLS * 32 * sizeof(float);
}

template <bool PART>
__global__ void __launch_bounds__(TILE_PIX, PART ? 2 : 3) render_bwd_kernel(RenderBwdArgs a) {
  extern __shared__ __align__(16) unsigned char bwd_smem[];
  const int nw = blockDim.x >> 5;  // warps per CTA (8 = whole tile)
  WarpStage* s_stage = reinterpret_cast<WarpStage*>(bwd_smem);  // [2][nw]
  // per-warp transposition buffer of the gradient reduction: [RED_COLS][32 lanes]
  float* red = reinterpret_cast<float*>(bwd_smem + (size_t)2 * nw * sizeof(WarpStage)) + (threadIdx.x >> 5) * (RED_COLS * 32);

  const int S = PART ? a.S : 0;
  const unsigned lane = threadIdx.x & 31, lw = threadIdx.x >> 5;
  const int groups = NWARP / nw;
  const unsigned wid = (blockIdx.x % groups) * nw + lw;
  const int tid = wid * 32 + lane;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int tile_slot = blockIdx.x / groups;
  const int tile_id = a.tile_order ? (int)a.tile_order[tile_slot] : tile_slot;
  const int tile_x = tile_id % a.grid_x, tile_y = tile_id / a.grid_x;
  const int fx0 = tile_x * TILE_X + (wid & 1) * WARP_FX;
  const int fy0 = tile_y * TILE_Y + (wid >> 1) * WARP_FY;
  const uint2 pix = {(unsigned)(fx0 + (lane & 7)), (unsigned)(fy0 + (lane >> 3))};
  const float poff = PART ? 0.5f : 0.0f;
  const float2 pixf = {(float)pix.x + poff, (float)pix.y + poff};
  const bool inside = pix.x < (unsigned)a.W && pix.y < (unsigned)a.H;

  const uint2 range = a.ranges[tile_id];
  const uint32_t* __restrict__ list = a.point_list + range.x;
  const uint32_t* __restrict__ fmask = a.frag_mask + (size_t)wid * a.mask_stride + range.x;

  const size_t npt = (size_t)a.grid_x * a.grid_y * TILE_PIX;
  const size_t sidx = (size_t)tile_id * TILE_PIX + tid;
  const size_t HW = (size_t)a.H * a.W;
  const size_t pix_id = (size_t)a.W * pix.y + pix.x;

  const float T_final = inside ? a.final_T[sidx] : 0;
  float T = T_final;
  const uint32_t last_contributor = inside ? a.n_contrib[sidx] : 0;
  const uint32_t median_contributor = inside ? a.n_contrib[sidx + npt] : 0;
  const float final_D = inside ? a.final_T[sidx + npt] : 0;
  const float final_D2 = inside ? a.final_T[sidx + 2 * npt] : 0;
  const float final_A = 1 - T_final;

  float dL_dpixel[3] = {0.f, 0.f, 0.f};
  float dL_dreg = 0.f, dL_ddepth = 0.f, dL_daccum = 0.f, dL_dmedian_depth = 0.f, dL_dmax_dweight = 0.f;
  float dL_dnormal2D[3] = {0.f, 0.f, 0.f};
  float dL_dsem[MAX_SEMANTIC];
  if (PART) {
#pragma unroll
    for (int i = 0; i < MAX_SEMANTIC; i++) dL_dsem[i] = 0.f;
  }
  if (inside) {
    for (int i = 0; i < 3; i++) dL_dpixel[i] = a.dL_dpixels[i * HW + pix_id];
    dL_ddepth = a.dL_dothers[DEPTH_OFFSET * HW + pix_id];
    dL_daccum = a.dL_dothers[ALPHA_OFFSET * HW + pix_id];
    dL_dreg = a.dL_dothers[DISTORTION_OFFSET * HW + pix_id];
    for (int i = 0; i < 3; i++) dL_dnormal2D[i] = a.dL_dothers[(NORMAL_OFFSET + i) * HW + pix_id];
    dL_dmedian_depth = a.dL_dothers[MIDDEPTH_OFFSET * HW + pix_id];
    if (PART) {
      dL_dmax_dweight = a.dL_dothers[MEDIAN_WEIGHT_OFFSET * HW + pix_id];
#pragma unroll
      for (int i = 0; i < MAX_SEMANTIC; i++)
        if (i < S) dL_dsem[i] = a.dL_dsemantic[i * HW + pix_id];
    }
  }
  float bg_dot_dpixel = 0;
  for (int i = 0; i < 3; i++) bg_dot_dpixel += a.bg_color[i] * dL_dpixel[i];
  const float Tf_bg = T_final * bg_dot_dpixel;

  // The reference keeps one "blend of everything behind me" recurrence per output channel
  // (accum_rec[3], accum_depth_rec, accum_alpha_rec, accum_normal_rec[3], last_dL_dT; backward.cu:316-372).
  // They all have the form A <- last_alpha * last_x + (1 - last_alpha) * A and enter dL_dalpha only through
  // sum_ch g_ch * (x_ch - A_ch) with per-pixel constant upstream gradients g_ch, so one scalar recurrence on
  // v = sum_ch g_ch * x_ch carries the same information.
  float last_alpha = 0.f, last_v = 0.f, accum_v = 0.f;

  // deepest contributing fragment over the warp's 32 pixels: positions [0, top) matter
  uint32_t top = last_contributor;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) top = max(top, __shfl_xor_sync(RFULL, top, o));
  const int total = (int)top;

  // candidate c of step `base` sits at list position total-1-(base+lane)  (back to front)
  auto load_cand = [&](int base, uint32_t& id, uint32_t& mk) {
    id = 0;
    mk = 0;
    if (base + (int)lane < total) {
      const int p = total - 1 - (base + (int)lane);
      id = list[p];
      mk = fmask[p];
    }
  };
  auto stage = [&](int base, int buf, uint32_t id, uint32_t mk) -> int {
    const bool hit = mk != 0u;
    const unsigned m = __ballot_sync(RFULL, hit);
    if (hit) {
      const int slot = __popc(m & lt_mask);
      WarpStage& st = s_stage[buf * nw + lw];
      const float4* src = a.rec + (size_t)id * REC_QUADS;
#pragma unroll
      for (int q = 0; q < REC_QUADS; q++) cp_async16(&st.rec[slot][q], src + q);
      st.pos[slot] = (uint32_t)(total - 1 - (base + (int)lane));
      st.id[slot] = id;
      st.mask[slot] = mk;
    }
    cp_async_commit();
    return __popc(m);
  };

  uint32_t id_n, mk_n, id_n2, mk_n2;
  load_cand(0, id_n, mk_n);
  int n_c = stage(0, 0, id_n, mk_n);
  load_cand(CHUNK, id_n, mk_n);
  load_cand(2 * CHUNK, id_n2, mk_n2);

  const float fis = PART ? (float)(1 / (0.7071067811865476 * 0.7071067811865476)) : PGS_FILTER_INV_SQUARE;
  // which float of the surfel's gradient record this lane adds to (see the reduction below)
  const int red_off = (lane & 1) ? 17 + (int)(lane >> 4) : (int)(lane >> 1) + (lane == 30 ? 1 : 0);
  const bool red_on = !(lane & 1) || lane == 1 || lane == 17;

  for (int base = 0, buf = 0; base < total; base += CHUNK, buf ^= 1) {
    int n_n = 0;
    if (base + CHUNK < total) n_n = stage(base + CHUNK, buf ^ 1, id_n, mk_n);
    else cp_async_commit();
    id_n = id_n2;
    mk_n = mk_n2;
    load_cand(base + 3 * CHUNK, id_n2, mk_n2);
    cp_async_wait<1>();
    __syncwarp();

    const WarpStage& st = s_stage[buf * nw + lw];
    for (int j = 0; j < n_c; j++) {
      const uint32_t contributor = st.pos[j];
      const bool valid = (st.mask[j] >> lane) & 1u;  // this pixel blended the surfel in forward
      const float4 q0 = st.rec[j][0];
      const float4 q1 = st.rec[j][1];
      const float4 q2 = st.rec[j][2];
      const float3 Tu = {q0.x, q0.y, q0.z};
      const float3 Tv = {q1.x, q1.y, q1.z};
      const float3 Tw = {q2.x, q2.y, q2.z};
      const float opa = q2.w;

      // Fragment values (forward.cu:344-387) recomputed with approximate reciprocal / exp2: no decision
      // depends on them any more (forward recorded which pixels blended), gradients are gated at 1e-4.
      const float3 k = {pixf.x * Tw.x - Tu.x, pixf.x * Tw.y - Tu.y, pixf.x * Tw.z - Tu.z};
      const float3 l = {pixf.y * Tw.x - Tv.x, pixf.y * Tw.y - Tv.y, pixf.y * Tw.z - Tv.z};
      const float3 p = {k.y * l.z - k.z * l.y, k.z * l.x - k.x * l.z, k.x * l.y - k.y * l.x};
      // lanes that did not blend this surfel run the arithmetic below with w = dL_dalpha = dL_dz = 0;
      // inv_pz = 0 keeps every intermediate finite for them
      const float inv_pz = valid ? rcp_approx(p.z) : 0.f;
      const float2 s = {p.x * inv_pz, p.y * inv_pz};
      const float rho3d = s.x * s.x + s.y * s.y;
      const float2 d = {q0.w - pixf.x, q1.w - pixf.y};
      const float rho2d = fis * (d.x * d.x + d.y * d.y);
      const bool use3d = rho3d <= rho2d;
      const float c_d = use3d ? (s.x * Tw.x + s.y * Tw.y) + Tw.z : Tw.z;
      const float G = ex2_approx(-0.5f * 1.4426950408889634f * fminf(rho3d, rho2d));
      const float alpha = fminf(0.99f, opa * G);

      float w = 0.f, dL_dalpha = 0.f, dL_dz = 0.f;
      const float4 q3 = st.rec[j][3];
      const float4 q4 = st.rec[j][4];
      if (valid) {
        const float inv_1ma = rcp_approx(1.f - alpha);
        T = T * inv_1ma;
        w = alpha * T;
        const float inv_cd = rcp_approx(c_d);
        float m_d, dmd_dd;
        if (PART) {
          m_d = (float)(100.0 / (100.0 - 0.2)) * (1.f - 0.2f * inv_cd);
          dmd_dd = (float)(100.0 * 0.2 / (100.0 - 0.2)) * inv_cd * inv_cd;
        } else {
          m_d = (PGS_FAR_N / (PGS_FAR_N - PGS_NEAR_N)) * (1.f - PGS_NEAR_N * inv_cd);
          dmd_dd = ((PGS_FAR_N * PGS_NEAR_N) / (PGS_FAR_N - PGS_NEAR_N)) * inv_cd * inv_cd;
        }
        // v = sum over channels of (upstream gradient x this fragment's attribute); the distortion weight
        // (and, in `_part`, the median-weight gradient) is the attribute of a channel with unit gradient
        float v = (final_D2 + m_d * m_d * final_A - 2 * m_d * final_D) * dL_dreg;
        if (contributor == median_contributor - 1) {
          dL_dz += dL_dmedian_depth;
          if (PART) v += dL_dmax_dweight;
        }
        v += q4.x * dL_dpixel[0] + q4.y * dL_dpixel[1] + q4.z * dL_dpixel[2];
        v += c_d * dL_ddepth + dL_daccum;
        v += q3.x * dL_dnormal2D[0] + q3.y * dL_dnormal2D[1] + q3.z * dL_dnormal2D[2];
        accum_v = last_alpha * last_v + (1.f - last_alpha) * accum_v;
        last_v = v;
        last_alpha = alpha;
        dL_dalpha = (v - accum_v) * T - Tf_bg * inv_1ma;
        dL_dz += (2.0f * w * (m_d * final_A - final_D) * dL_dreg) * dmd_dd + w * dL_ddepth;
      }

      float g[16];
      float gc[3];
#pragma unroll
      for (int ch = 0; ch < 3; ch++) {
        gc[ch] = w * dL_dpixel[ch];
        g[12 + ch] = w * dL_dnormal2D[ch];
      }
      g[11] = G * dL_dalpha;
      const float dL_dG = opa * dL_dalpha;
      // ray-splat branch: gradient w.r.t. the 3x3 transform through s = p.xy / p.z
      const float mGG = -G * dL_dG;
      const float dsx_pz = (mGG * s.x + dL_dz * Tw.x) * inv_pz;
      const float dsy_pz = (mGG * s.y + dL_dz * Tw.y) * inv_pz;
      const float3 dL_dp = {dsx_pz, dsy_pz, -(dsx_pz * s.x + dsy_pz * s.y)};
      const float3 dL_dk = {l.y * dL_dp.z - l.z * dL_dp.y, l.z * dL_dp.x - l.x * dL_dp.z,
                            l.x * dL_dp.y - l.y * dL_dp.x};
      const float3 dL_dl = {dL_dp.y * k.z - dL_dp.z * k.y, dL_dp.z * k.x - dL_dp.x * k.z,
                            dL_dp.x * k.y - dL_dp.y * k.x};
      // low-pass branch: gradient w.r.t. the screen-space centre
      const float mGGf = mGG * fis;
      g[0] = use3d ? -dL_dk.x : 0.f;
      g[1] = use3d ? -dL_dk.y : 0.f;
      g[2] = use3d ? -dL_dk.z : 0.f;
      g[3] = use3d ? -dL_dl.x : 0.f;
      g[4] = use3d ? -dL_dl.y : 0.f;
      g[5] = use3d ? -dL_dl.z : 0.f;
      g[6] = use3d ? pixf.x * dL_dk.x + pixf.y * dL_dl.x + dL_dz * s.x : 0.f;
      g[7] = use3d ? pixf.x * dL_dk.y + pixf.y * dL_dl.y + dL_dz * s.y : 0.f;
      g[8] = use3d ? pixf.x * dL_dk.z + pixf.y * dL_dl.z + dL_dz : dL_dz;
      g[9] = use3d ? 0.f : mGGf * d.x;
      g[10] = use3d ? 0.f : mGGf * d.y;

      // Sum the 18 gradient components over the warp's 32 pixels.  16 of them (g[0..14], gc[0]) go through
      // shared memory: every lane stores its values column-wise (conflict-free), then lanes (2c, 2c+1) each
      // add one half of column c with four rotated 128-bit loads and exchange halves with one shuffle —
      // ~40 instructions instead of the ~90 of a select-and-shuffle butterfly.  The two remaining colour
      // components take a 2-value butterfly in registers.
      g[15] = gc[0];
#pragma unroll
      for (int c = 0; c < RED_COLS; c++) red[c * 32 + lane] = g[c];
      __syncwarp();
      float r16;
      {
        const int col = lane >> 1;
        const float4* colp = reinterpret_cast<const float4*>(red + col * 32 + (lane & 1) * 16);
        const float4 t0 = colp[(0 + col) & 3], t1 = colp[(1 + col) & 3], t2 = colp[(2 + col) & 3],
                     t3 = colp[(3 + col) & 3];
        r16 = ((t0.x + t0.y) + (t0.z + t0.w)) + ((t1.x + t1.y) + (t1.z + t1.w)) +
              (((t2.x + t2.y) + (t2.z + t2.w)) + ((t3.x + t3.y) + (t3.z + t3.w)));
        r16 += __shfl_xor_sync(RFULL, r16, 1);
      }
      float r2;
      {
        const bool hi = lane & 16;
        const float keep = hi ? gc[2] : gc[1];
        const float send = hi ? gc[1] : gc[2];
        r2 = keep + __shfl_xor_sync(RFULL, send, 16);
        r2 += __shfl_xor_sync(RFULL, r2, 8);
        r2 += __shfl_xor_sync(RFULL, r2, 4);
        r2 += __shfl_xor_sync(RFULL, r2, 2);
        r2 += __shfl_xor_sync(RFULL, r2, 1);
      }
      __syncwarp();  // `red` is rewritten by the next fragment
      // One reduction instruction per fragment: even lanes own dL/dT[9], dL/dmean2D[2], dL/dopacity, dL/dnormal[3]
      // (lane 30: colour 0); lanes 1 and 17 carry the other two colour sums (every lane of a half holds its r2).
      const uint32_t gid = st.id[j];
      float* dst = a.grad + (size_t)gid * GRAD_FLOATS + red_off;
      if (red_on) red_add_f32(dst, (lane & 1) ? r2 : r16);
      if (PART && S > 0) {
        // dL/dsem[ch] = sum_pixels alpha*T * dL/dpixel_sem[ch]  (no alpha gradient in the reference fork)
        float gs[16];
#pragma unroll
        for (int i = 0; i < 16; i++) gs[i] = w * dL_dsem[i];
        const float rs = warp_reduce16(gs, lane);
        const int ch = lane >> 1;
        if ((lane & 1) == 0 && ch < S) atomicAdd(a.grad_semantics + (size_t)gid * S + ch, rs);
      }
    }
    __syncwarp();  // ring buffer `buf` is refilled by the next iteration's stage()
    n_c = n_n;
  }
  cp_async_wait<0>();
}

// =============================================================================
// launchers
// =============================================================================
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// Warps per CTA.  A whole tile (8 warps) per CTA is the default; images with few tiles (e.g. the
// 400x300 DTU training resolution: 475 tiles for 592 CTA slots) are launched in finer units so that
// the block scheduler can balance the warps of heavy tiles over all SMs.
static int warps_per_cta(int ntiles) {
  static const int forced = env_int("PGS_WARPS_PER_CTA", 0);
  if (forced == 1 || forced == 2 || forced == 4 || forced == 8) return forced;
  if (ntiles >= 3000) return 8;
  if (ntiles >= 1500) return 4;
  if (ntiles >= 750) return 2;
  return 1;
}

template <bool PART> static void launch_fwd(const RenderFwdArgs& a, cudaStream_t s) {
  const int ntiles = a.grid_x * a.grid_y;
  const int nw = warps_per_cta(ntiles);
  const int groups = NWARP / nw;
  const size_t smem = (size_t)2 * nw * sizeof(WarpStage) +
                      (PART ? (size_t)2 * nw * CHUNK * MAX_SEMANTIC * sizeof(float) : 0);
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    const size_t mx = (size_t)2 * NWARP * sizeof(WarpStage) +
                      (PART ? (size_t)2 * NWARP * CHUNK * MAX_SEMANTIC * sizeof(float) : 0);
    cudaFuncSetAttribute(render_fwd_kernel<PART>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx);
  }
  render_fwd_kernel<PART><<<ntiles * groups, 32 * nw, smem, s>>>(a);
  count_launch();
}
void launch_render_fwd(const RenderFwdArgs& a, cudaStream_t s) { launch_fwd<false>(a, s); }
void launch_render_fwd_part(const RenderFwdArgs& a, cudaStream_t s) { launch_fwd<true>(a, s); }

template <bool PART> static void launch_bwd(const RenderBwdArgs& a, cudaStream_t s) {
  const int ntiles = a.grid_x * a.grid_y;
  const int nw = warps_per_cta(ntiles);
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(render_bwd_kernel<PART>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)bwd_smem_bytes(NWARP));
  }
  render_bwd_kernel<PART><<<ntiles * (NWARP / nw), 32 * nw, bwd_smem_bytes(nw), s>>>(a);
  count_launch();
}
void launch_render_bwd(const RenderBwdArgs& a, cudaStream_t s) { launch_bwd<false>(a, s); }
void launch_render_bwd_part(const RenderBwdArgs& a, cudaStream_t s) { launch_bwd<true>(a, s); }

}  // namespace pgs
