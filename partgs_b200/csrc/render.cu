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
#include "f32x2.cuh"
#include "frag_math.cuh"
#include "kernels.h"
#include "render_common.cuh"

namespace pgs {

// Fragments evaluated ahead of the in-order blend.  Measured on C3: 2 -> 1.11 ms, 4 -> 1.14-1.23 ms, 8 -> 1.37-1.43 ms
// (a deeper look-ahead evaluates up to ILP-1 slots past the end of every step and costs registers).
#ifndef PGS_FWD_ILP
#define PGS_FWD_ILP 2
#endif
constexpr int FWD_ILP = PGS_FWD_ILP;

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
__global__ void __launch_bounds__(1024) tile_order_kernel(uint2* __restrict__ ranges, int ntiles,
                                                          uint32_t* __restrict__ order) {
  __shared__ uint32_t s_cnt[TO_BUCKETS];
  __shared__ uint32_t s_warp[32];
  const int tid = threadIdx.x;
  s_cnt[tid] = 0;
  __syncthreads();
  // Tiles without instances still carry the {0xffffffff, 0} the fused range detection starts from (binning.cu):
  // they become the reference's {0, 0} here.  Eight tiles per thread and round, so that the loads overlap.
  constexpr int U = 8;
  auto bucket_of = [](uint2 r) -> uint32_t {  // long lists -> small bucket
    return TO_BUCKETS - 1 - min((r.y - r.x) >> 4, (uint32_t)(TO_BUCKETS - 1));
  };
  for (int t0 = 0; t0 < ntiles; t0 += U * 1024) {
    uint2 r[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int t = t0 + u * 1024 + tid;
      r[u] = t < ntiles ? ranges[t] : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int t = t0 + u * 1024 + tid;
      if (t < ntiles) {
        if (r[u].x > r[u].y) {
          r[u] = make_uint2(0u, 0u);
          ranges[t] = r[u];
        }
        atomicAdd(&s_cnt[bucket_of(r[u])], 1u);
      }
    }
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
  for (int t0 = 0; t0 < ntiles; t0 += U * 1024) {
    uint2 r[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int t = t0 + u * 1024 + tid;
      r[u] = t < ntiles ? ranges[t] : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int t = t0 + u * 1024 + tid;
      if (t < ntiles) order[atomicAdd(&s_cnt[bucket_of(r[u])], 1u)] = (uint32_t)t;
    }
  }
}

void launch_tile_order(uint2* ranges, int ntiles, uint32_t* order, cudaStream_t s) {
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
  // `_part`: the S <= 16 part channels are blended as 8 packed pairs (FMUL2 + FFMA2 per pair: each half is rounded
  // exactly like the scalar fma(T, sem * alpha, acc) of the reference build)
  P2 Sem2[MAX_SEMANTIC / 2];
  if (PART) {
#pragma unroll
    for (int i = 0; i < MAX_SEMANTIC / 2; i++) Sem2[i] = bc(0.f);
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
        // the survivor's part row rides the same cp.async group as its record: 16-byte copies when the rows are
        // 16-byte aligned (S = 4, 8, 12, 16), scalar loads otherwise; channels >= S of the slot are never consumed
        float* sem_stage = s_sem_dyn + ((size_t)(buf * nw + lw) * CHUNK + slot) * MAX_SEMANTIC;
        const float* sem = a.semantics + (size_t)id * S;
        if ((S & 3) == 0 && (reinterpret_cast<size_t>(a.semantics) & 15) == 0) {
          for (int q = 0; q < (S >> 2); q++) cp_async16(sem_stage + 4 * q, sem + 4 * q);
        } else {
          for (int ch = 0; ch < S; ch++) sem_stage[ch] = __ldg(sem + ch);
        }
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
          const float4* srow = reinterpret_cast<const float4*>(sem_stage + j * MAX_SEMANTIC);
          const P2 a2 = bc(alpha), T2 = bc(T);
#pragma unroll
          for (int q = 0; q < MAX_SEMANTIC / 4; q++)
            if (4 * q < S) {   // channels in [S, 4 ceil(S/4)) of a quad accumulate stale values that are never stored
              const float4 v = srow[q];
              Sem2[2 * q] = fma2(T2, mul2(pk(v.x, v.y), a2), Sem2[2 * q]);
              Sem2[2 * q + 1] = fma2(T2, mul2(pk(v.z, v.w), a2), Sem2[2 * q + 1]);
            }
        }
        T = test_T;
        last_contributor = contributor;
        return true;
      };
      // FWD_ILP fragments are evaluated (geometry, alpha: independent of the pixel's running state)
      // before they are blended in order.  The deepest tiles (thousands of fragments on the same
      // pixels, one warp alone on its scheduler at the end of the launch) are bound by the latency
      // of this chain, not by issue slots.
      const int n_even = n_c - n_c % FWD_ILP;
      for (int j = 0; j < n_even; j += FWD_ILP) {
        float alpha[FWD_ILP], depth[FWD_ILP];
        bool ok[FWD_ILP];
#pragma unroll
        for (int u = 0; u < FWD_ILP; u++)
          ok[u] = frag_eval<PART>(pixf, st.rec[j + u][0], st.rec[j + u][1], st.rec[j + u][2], alpha[u], depth[u]);
#pragma unroll
        for (int u = 0; u < FWD_ILP; u++) {
          bool b = false;
          if (ok[u] && !done) b = blend(j + u, alpha[u], depth[u]);
          const unsigned m = __ballot_sync(RFULL, b);
          if ((int)lane == j + u) my_mask = m;
        }
      }
      // an odd survivor count: the last one alone (half of all steps: evaluating a stale second slot there used to
      // cost 4 % of the kernel's instructions)
      for (int j = n_even; j < n_c; j++) {
        float alpha, depth;
        const bool ok = frag_eval<PART>(pixf, st.rec[j][0], st.rec[j][1], st.rec[j][2], alpha, depth);
        bool b = false;
        if (ok && !done) b = blend(j, alpha, depth);
        const unsigned m = __ballot_sync(RFULL, b);
        if ((int)lane == j) my_mask = m;
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
        if (ch < S) a.out_semantic[ch * HW + pix_id] = (ch & 1) ? hi(Sem2[ch >> 1]) : lo(Sem2[ch >> 1]);
    }
  }
}

// =============================================================================
// launchers
// =============================================================================
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

}  // namespace pgs
