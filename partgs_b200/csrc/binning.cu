// Binning for the surfel rasteriser (sm_100a): prefix sum of tile counts, key
// duplication, a hand-written onesweep LSD radix sort and tile-range detection.
//
// Replaces, with bit-identical results:
//   cub::DeviceScan::InclusiveSum            rasterizer_impl.cu:278
//   duplicateWithKeys                        rasterizer_impl.cu:70-111
//   cub::DeviceRadixSort::SortPairs          rasterizer_impl.cu:304-309   (stable, bits [0, 32+tile_bits))
//   identifyTileRanges                       rasterizer_impl.cu:116-138
#include "common.cuh"
#include "kernels.h"

namespace pgs {

// Element count of a launch: the host value `n_cap`, or — when the host launched speculatively for a
// capacity, before the count was known to it — the device value *n_dev.  Returns -1 when the device
// count exceeds the capacity the buffers were sized for (the launch then does nothing; the host
// re-launches with larger buffers once it has read the count).
__device__ __forceinline__ int resolve_count(int n_cap, const uint32_t* __restrict__ n_dev) {
  if (n_dev == nullptr) return n_cap;
  const uint32_t v = __ldg(n_dev);
  return v > (uint32_t)n_cap ? -1 : (int)v;
}

// =============================================================================
// Single-pass inclusive scan (decoupled look-back), u32.
// =============================================================================
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr unsigned long long SCAN_FLAG_AGG = 1ull << 32;
constexpr unsigned long long SCAN_FLAG_INC = 2ull << 32;

size_t scan_temp_bytes(int n) {
  int tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  return 256 + (size_t)tiles * sizeof(unsigned long long);
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(const uint32_t* __restrict__ in,
                                                             uint32_t* __restrict__ out, int n, uint32_t* counter,
                                                             unsigned long long* state) {
  __shared__ uint32_t s_vals[SCAN_TILE];
  __shared__ uint32_t s_warp[SCAN_THREADS / 32];
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_excl;
  const int tid = threadIdx.x;
  if (tid == 0) s_tile = atomicAdd(counter, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const int base = tile * SCAN_TILE;

#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    int e = i * SCAN_THREADS + tid;
    s_vals[e] = (base + e < n) ? in[base + e] : 0u;
  }
  __syncthreads();

  uint32_t v[SCAN_ITEMS];
  uint32_t tsum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    tsum += s_vals[tid * SCAN_ITEMS + i];
    v[i] = tsum;
  }
  // warp inclusive scan of thread sums
  uint32_t w = tsum;
  const unsigned lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
    if (lane >= o) w += t;
  }
  if (lane == 31) s_warp[wid] = w;
  __syncthreads();
  uint32_t warp_off = 0;
#pragma unroll
  for (int i = 0; i < SCAN_THREADS / 32; i++)
    if (i < wid) warp_off += s_warp[i];
  const uint32_t thread_excl = warp_off + w - tsum;

  if (tid == SCAN_THREADS - 1) {
    const uint32_t agg = thread_excl + tsum;  // tile aggregate
    uint32_t excl = 0;
    if (tile == 0) {
      atomicExch(&state[0], SCAN_FLAG_INC | agg);
    } else {
      atomicExch(&state[tile], SCAN_FLAG_AGG | agg);
      int t = (int)tile - 1;
      while (true) {
        unsigned long long sv = *((volatile unsigned long long*)&state[t]);
        if ((sv >> 32) == 0) continue;
        excl += (uint32_t)sv;
        if ((sv >> 32) == 2) break;
        t--;
      }
      atomicExch(&state[tile], SCAN_FLAG_INC | (unsigned long long)(uint32_t)(excl + agg));
    }
    s_excl = excl;
  }
  __syncthreads();
  const uint32_t off = s_excl + thread_excl;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) s_vals[tid * SCAN_ITEMS + i] = v[i] + off;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    int e = i * SCAN_THREADS + tid;
    if (base + e < n) out[base + e] = s_vals[e];
  }
}

void launch_inclusive_scan_u32(const uint32_t* in, uint32_t* out, int n, void* temp, cudaStream_t s) {
  if (n <= 0) return;
  int tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  cudaMemsetAsync(temp, 0, scan_temp_bytes(n), s);
  uint32_t* counter = (uint32_t*)temp;
  unsigned long long* state = (unsigned long long*)((char*)temp + 256);
  scan_kernel<<<tiles, SCAN_THREADS, 0, s>>>(in, out, n, counter, state);
  count_launch();
}

// =============================================================================
// duplicateWithKeys: key = tile_id << 32 | bits(depth), value = surfel index,
// emitted row-major over the surfel's tile rectangle.
// =============================================================================
__device__ __forceinline__ void tile_rect_dup(const float2 p, int max_radius, uint2& rect_min, uint2& rect_max,
                                              unsigned gx, unsigned gy) {
  rect_min = {min(gx, max((int)0, (int)((p.x - max_radius) / TILE_X))),
              min(gy, max((int)0, (int)((p.y - max_radius) / TILE_Y)))};
  rect_max = {min(gx, max((int)0, (int)((p.x + max_radius + TILE_X - 1) / TILE_X))),
              min(gy, max((int)0, (int)((p.y + max_radius + TILE_Y - 1) / TILE_Y)))};
}

__global__ void __launch_bounds__(256) duplicate_with_keys_kernel(int P, const float4* __restrict__ rec,
                                                                  const uint32_t* __restrict__ offsets,
                                                                  uint64_t* __restrict__ keys,
                                                                  uint32_t* __restrict__ values,
                                                                  const int* __restrict__ radii, unsigned gx,
                                                                  unsigned gy, uint32_t capacity) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  if (__ldg(&offsets[P - 1]) > capacity) return;  // speculative launch whose buffers are too small
  const int radius = radii[idx];
  if (radius > 0) {
    uint32_t off = (idx == 0) ? 0 : offsets[idx - 1];
    const float4* r = rec + (size_t)idx * REC_QUADS;
    const float2 xy = {__ldg(&r[0]).w, __ldg(&r[1]).w};
    const uint32_t depth_bits = __float_as_uint(__ldg(&r[3]).w);
    uint2 rect_min, rect_max;
    tile_rect_dup(xy, radius, rect_min, rect_max, gx, gy);
    for (unsigned y = rect_min.y; y < rect_max.y; y++) {
      for (unsigned x = rect_min.x; x < rect_max.x; x++) {
        uint64_t key = y * gx + x;
        key <<= 32;
        key |= depth_bits;
        keys[off] = key;
        values[off] = idx;
        off++;
      }
    }
  }
}

void launch_duplicate_with_keys(int P, const float4* rec, const uint32_t* offsets, uint64_t* keys, uint32_t* values,
                                const int* radii, int grid_x, int grid_y, cudaStream_t s, uint32_t capacity) {
  if (P <= 0) return;
  duplicate_with_keys_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, rec, offsets, keys, values, radii, grid_x, grid_y,
                                                              capacity);
  count_launch();
}

// =============================================================================
// Onesweep LSD radix sort (Adinets & Merrill 2022): one global histogram pass,
// then one kernel per 8-bit digit that ranks a tile in shared memory, resolves its
// global offsets by decoupled look-back over per-digit tile counts, and scatters.
// Stable.  KeyT = uint64_t (tile|depth keys) or uint32_t (kNN cell codes).
// =============================================================================
constexpr int RS_BITS = 8;
constexpr int RS_RADIX = 1 << RS_BITS;
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
#ifndef PGS_RS_ITEMS
#define PGS_RS_ITEMS 12
#endif
constexpr int RS_ITEMS = PGS_RS_ITEMS;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 3072 pairs / CTA
constexpr int RS_MAX_PASSES = PGS_RS_MAX_PASSES;
constexpr int RS_MAX_PASSES_TILE = 4;  // tile ids: at most 32 bits
constexpr int RS_LB = 8;  // look-back window
constexpr uint32_t RS_FLAG_AGG = 1u << 30;
constexpr uint32_t RS_FLAG_INC = 2u << 30;
constexpr uint32_t RS_VAL_MASK = (1u << 30) - 1;

static inline int rs_passes(int end_bit) { return (end_bit + RS_BITS - 1) / RS_BITS; }

// digit passes of an LSD sort on key bits [0, end_bit): 8-bit digits from the bottom (what CUB does) ...
static RsPlan rs_plan_lsd8(int end_bit) {
  RsPlan pl;
  pl.passes = rs_passes(end_bit);
  for (int p = 0; p < pl.passes; p++) {
    pl.shift[p] = p * RS_BITS;
    pl.bits[p] = min(RS_BITS, end_bit - pl.shift[p]);
  }
  return pl;
}
// ... or, for short keys (tile ids), as few passes as 8-bit digits allow with the bits split evenly (13 -> 7 + 6)
RsPlan rs_plan_even(int end_bit) {
  RsPlan pl;
  pl.passes = max(1, rs_passes(end_bit));
  int shift = 0;
  for (int p = 0; p < pl.passes; p++) {
    const int left = end_bit - shift, bits = (left + (pl.passes - p) - 1) / (pl.passes - p);
    pl.shift[p] = shift;
    pl.bits[p] = max(bits, 1);
    shift += pl.bits[p];
  }
  return pl;
}

static size_t rs_temp_bytes_passes(int n, int passes) {
  int tiles = (n + RS_TILE - 1) / RS_TILE;
  // [hist: passes*256 u32][counters: passes u32 (padded)][lookback: passes*tiles*256 u32]
  return (size_t)passes * RS_RADIX * 4 + 256 + (size_t)passes * tiles * RS_RADIX * 4 + 256;
}
template <typename KeyT> static size_t rs_temp_bytes(int n, int end_bit) { return rs_temp_bytes_passes(n, rs_passes(end_bit)); }
size_t radix_sort_temp_bytes(int n, int end_bit) { return rs_temp_bytes<uint64_t>(n, end_bit); }
size_t radix_sort32_temp_bytes(int n, int end_bit) { return rs_temp_bytes<uint32_t>(n, end_bit); }

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS) rs_histogram_kernel(const KeyT* __restrict__ keys, int n_cap, RsPlan pl,
                                                                   uint32_t* __restrict__ hist,
                                                                   const uint32_t* __restrict__ n_dev) {
  const int passes = pl.passes;
  __shared__ uint32_t s_hist[RS_MAX_PASSES * RS_RADIX];
  const int n = resolve_count(n_cap, n_dev);
  if (n < 0) return;
  for (int i = threadIdx.x; i < passes * RS_RADIX; i += RS_THREADS) s_hist[i] = 0;
  __syncthreads();
  const int stride = gridDim.x * RS_THREADS;
  for (int i = blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += stride) {
    KeyT k = keys[i];
#pragma unroll 1
    for (int p = 0; p < passes; p++) {
      uint32_t d = (uint32_t)(k >> pl.shift[p]) & ((1u << pl.bits[p]) - 1);
      atomicAdd(&s_hist[p * RS_RADIX + d], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * RS_RADIX; i += RS_THREADS) {
    uint32_t c = s_hist[i];
    if (c) atomicAdd(&hist[i], c);
  }
}

// exclusive scan of each pass's 256-bin histogram, in place
__global__ void __launch_bounds__(RS_RADIX) rs_scan_hist_kernel(uint32_t* hist) {
  __shared__ uint32_t s_warp[RS_RADIX / 32];
  uint32_t* h = hist + blockIdx.x * RS_RADIX;
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31, wid = tid >> 5;
  uint32_t c = h[tid];
  uint32_t w = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
    if (lane >= o) w += t;
  }
  if (lane == 31) s_warp[wid] = w;
  __syncthreads();
  uint32_t off = 0;
#pragma unroll
  for (int i = 0; i < RS_RADIX / 32; i++)
    if (i < wid) off += s_warp[i];
  h[tid] = off + w - c;
}

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS)
    rs_onesweep_kernel(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                       KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int n_cap, int shift, int bits,
                       const uint32_t* __restrict__ bin_base,  // [256] exclusive global digit offsets
                       uint32_t* lookback,                     // [tiles][256]
                       uint32_t* tile_counter, const uint32_t* __restrict__ n_dev,
                       uint2* __restrict__ ranges = nullptr) {  // last pass of the tile-id sort: see below
  extern __shared__ __align__(16) unsigned char rs_smem[];
  KeyT* s_keys = reinterpret_cast<KeyT*>(rs_smem);                                   // [RS_TILE]
  uint32_t* s_vals = reinterpret_cast<uint32_t*>(rs_smem + sizeof(KeyT) * RS_TILE);  // [RS_TILE]
  uint32_t* s_warp_hist = s_vals + RS_TILE;                                          // [RS_WARPS][256]
  uint32_t* s_digit_start = s_warp_hist + RS_WARPS * RS_RADIX;                       // [256] block-local start
  uint32_t* s_goff = s_digit_start + RS_RADIX;                                       // [256] global - local
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_scan_warp[RS_WARPS];

  const int tid = threadIdx.x;
  const unsigned lane = tid & 31, wid = tid >> 5;
  const int n = resolve_count(n_cap, n_dev);
  if (n < 0) return;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) s_warp_hist[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const int base = tile * RS_TILE;
  if (base >= n) return;  // grid was sized for the capacity; no tile ever looks back at this one
  const uint32_t mask = (1u << bits) - 1;

  // warp-striped load: element order e = wid*ITEMS*32 + i*32 + lane
  KeyT key[RS_ITEMS];
  uint32_t val[RS_ITEMS];
  uint32_t rank[RS_ITEMS];
  const int wbase = base + wid * (RS_ITEMS * 32);
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    int g = wbase + i * 32 + lane;
    if (g < n) {
      key[i] = keys_in[g];
      val[i] = vals_in ? vals_in[g] : (uint32_t)g;  // first pass of an index sort: the values are the positions
    } else {
      key[i] = ~(KeyT)0;  // sentinel: last digit, ranked after every valid element
      val[i] = 0;
    }
  }
  // rank within warp, in element order (ballot-per-bit matching was measured: no faster than MATCH.ANY here)
  uint32_t* my_hist = s_warp_hist + wid * RS_RADIX;
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    uint32_t d = (uint32_t)(key[i] >> shift) & mask;
    unsigned peers = __match_any_sync(0xffffffffu, d);
    unsigned leader = __ffs(peers) - 1;
    uint32_t pre = 0;
    if (lane == leader) {
      pre = my_hist[d];
      my_hist[d] = pre + __popc(peers);
    }
    pre = __shfl_sync(0xffffffffu, pre, leader);
    rank[i] = pre + __popc(peers & ((1u << lane) - 1));
    __syncwarp();
  }
  __syncthreads();

  // thread d owns digit d: exclusive scan across warps -> block count
  uint32_t count = 0;
  {
    const int d = tid;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) {
      uint32_t t = s_warp_hist[w * RS_RADIX + d];
      s_warp_hist[w * RS_RADIX + d] = count;
      count += t;
    }
  }
  // publish + decoupled look-back per digit
  uint32_t excl = 0;
  {
    uint32_t* lb = lookback + (size_t)tile * RS_RADIX + tid;
    if (tile == 0) {
      atomicExch(lb, RS_FLAG_INC | count);
    } else {
      atomicExch(lb, RS_FLAG_AGG | count);
      // Walk back RS_LB tiles per round trip (the loads of a window are independent; they are consumed in order and
      // the walk resumes at the first tile that has published nothing yet).  The pass is bound by this chain: the
      // front of finished prefixes advances one tile per L2 round trip, the walk now RS_LB tiles.
      int t = (int)tile - 1;
      bool done = false;
      while (!done) {
        uint32_t v[RS_LB];
#pragma unroll
        for (int u = 0; u < RS_LB; u++)
          v[u] = (t - u >= 0) ? *((volatile uint32_t*)(lookback + (size_t)(t - u) * RS_RADIX + tid)) : (uint32_t)(2u << 30);
        int used = RS_LB;
#pragma unroll
        for (int u = 0; u < RS_LB; u++) {
          if (!done && used == RS_LB) {
            const uint32_t f = v[u] >> 30;
            if (f == 0) used = u;  // not published yet: poll again from this tile
            else {
              excl += v[u] & RS_VAL_MASK;
              if (f == 2) done = true;
            }
          }
        }
        t -= used;
      }
      atomicExch(lb, RS_FLAG_INC | ((excl + count) & RS_VAL_MASK));
    }
  }
  // block-local exclusive scan over digits
  {
    uint32_t w = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    if (lane == 31) s_scan_warp[wid] = w;
    __syncthreads();
    uint32_t off = 0;
#pragma unroll
    for (int i = 0; i < RS_WARPS; i++)
      if (i < (int)wid) off += s_scan_warp[i];
    const uint32_t start = off + w - count;
    s_digit_start[tid] = start;
    s_goff[tid] = bin_base[tid] + excl - start;
  }
  __syncthreads();

  // scatter into shared memory in sorted order
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    uint32_t d = (uint32_t)(key[i] >> shift) & mask;
    uint32_t pos = s_digit_start[d] + my_hist[d] + rank[i];
    s_keys[pos] = key[i];
    s_vals[pos] = val[i];
  }
  __syncthreads();

  // coalesced write-out; sentinels occupy the tail [nvalid, RS_TILE)
  const int nvalid = min(RS_TILE, n - base);
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    int j = i * RS_THREADS + tid;
    if (j < nvalid) {
      KeyT k = s_keys[j];
      uint32_t d = (uint32_t)(k >> shift) & mask;
      uint32_t g = s_goff[d] + j;
      keys_out[g] = k;
      vals_out[g] = s_vals[j];
      if (sizeof(KeyT) == 4 && ranges != nullptr) {
        // identifyTileRanges (rasterizer_impl.cu:116-138) fused into the final scatter of the tile-id sort: the
        // keys are tile ids and g is the element's final position.  Inside this CTA's run of a digit the
        // neighbours are at hand; a tile may continue in another CTA's run, so starts / ends are combined with
        // min / max (ranges are initialised to {0xffffffff, 0} by the emission kernel; tile_order_kernel turns
        // untouched entries into the reference's {0, 0}).
        const uint32_t t = (uint32_t)k;
        bool first = (j == 0), last = (j == nvalid - 1);
        if (!first) {
          const uint32_t kp = (uint32_t)s_keys[j - 1];
          first = kp != t;  // a different digit implies a different tile id
        }
        if (!last) {
          const uint32_t kn = (uint32_t)s_keys[j + 1];
          last = kn != t;
        }
        if (first) atomicMin(&ranges[t].x, g);
        if (last) atomicMax(&ranges[t].y, g + 1);
      }
    }
  }
}

// `hist_ready`: the digit histograms of `pl` were already accumulated into `temp` (which the caller zeroed with
// rs_prepare before) by the kernel that produced the keys — no histogram pass here.
// `iota_vals`: vals_a holds nothing yet; the first pass takes each element's position as its value.
template <typename KeyT>
static int rs_sort(KeyT* keys_a, uint32_t* vals_a, KeyT* keys_b, uint32_t* vals_b, int n, const RsPlan& pl, void* temp,
                   cudaStream_t s, const uint32_t* n_dev, bool hist_ready = false, bool iota_vals = false,
                   uint2* ranges = nullptr) {
  if (n <= 0) return 0;
  const int passes = pl.passes;
  const int tiles = (n + RS_TILE - 1) / RS_TILE;
  uint32_t* hist = (uint32_t*)temp;
  uint32_t* counters = (uint32_t*)((char*)temp + (size_t)passes * RS_RADIX * 4);
  uint32_t* lookback = (uint32_t*)((char*)counters + 256);

  if (!hist_ready) {
    cudaMemsetAsync(temp, 0, rs_temp_bytes_passes(n, passes), s);
    int hist_blocks = min(tiles, 148 * 8);
    rs_histogram_kernel<KeyT><<<hist_blocks, RS_THREADS, 0, s>>>(keys_a, n, pl, hist, n_dev);
    count_launch();
  }
  rs_scan_hist_kernel<<<passes, RS_RADIX, 0, s>>>(hist);
  count_launch();

  const size_t smem = sizeof(KeyT) * RS_TILE + 4 * RS_TILE + 4 * (RS_WARPS * RS_RADIX + 2 * RS_RADIX);
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(rs_onesweep_kernel<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  KeyT* kin = keys_a;
  uint32_t* vin = vals_a;
  KeyT* kout = keys_b;
  uint32_t* vout = vals_b;
  for (int p = 0; p < passes; p++) {
    rs_onesweep_kernel<KeyT><<<tiles, RS_THREADS, smem, s>>>(kin, (p == 0 && iota_vals) ? nullptr : vin, kout, vout, n,
                                                              pl.shift[p], pl.bits[p], hist + p * RS_RADIX,
                                                              lookback + (size_t)p * tiles * RS_RADIX, counters + p,
                                                              n_dev, p == passes - 1 ? ranges : nullptr);
    count_launch();
    KeyT* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
  }
  return (passes & 1) ? 1 : 0;
}

int launch_radix_sort_pairs(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, int n, int end_bit,
                            void* temp, cudaStream_t s, const uint32_t* n_dev) {
  return rs_sort<uint64_t>(keys_a, vals_a, keys_b, vals_b, n, rs_plan_lsd8(end_bit), temp, s, n_dev);
}
int launch_radix_sort_pairs32(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, int n,
                              int end_bit, void* temp, cudaStream_t s) {
  return rs_sort<uint32_t>(keys_a, vals_a, keys_b, vals_b, n, rs_plan_lsd8(end_bit), temp, s, nullptr);
}
int launch_radix_sort_index32(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, int n,
                              int end_bit, void* temp, cudaStream_t s, bool hist_ready) {
  return rs_sort<uint32_t>(keys_a, vals_a, keys_b, vals_b, n, rs_plan_lsd8(end_bit), temp, s, nullptr, hist_ready, true);
}
void radix_sort32_prepare(int n, int end_bit, void* temp, cudaStream_t s) {
  if (n > 0) cudaMemsetAsync(temp, 0, rs_temp_bytes<uint32_t>(n, end_bit), s);
}
size_t radix_sort_plan_temp_bytes(int n, const RsPlan& pl) { return rs_temp_bytes_passes(n, pl.passes); }
void radix_sort_plan_prepare(int n, const RsPlan& pl, void* temp, cudaStream_t s) {
  if (n > 0) cudaMemsetAsync(temp, 0, rs_temp_bytes_passes(n, pl.passes), s);
}
int launch_radix_sort_plan32(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, int n,
                             const RsPlan& pl, void* temp, cudaStream_t s, const uint32_t* n_dev, uint2* ranges) {
  return rs_sort<uint32_t>(keys_a, vals_a, keys_b, vals_b, n, pl, temp, s, n_dev, /*hist_ready=*/true, false, ranges);
}

// =============================================================================
// Instance emission in depth order (production binning path).
//
// The reference sorts R = sum(tiles_touched) (tile | depth) 64-bit keys in one go (6 digit passes over R pairs at
// 1600x1200).  Only the tile part of the key needs an R-sized sort: the surfels are first stable-sorted by their
// depth bits (P pairs, 4 passes), the (tile, surfel) instances are emitted in that order, and a stable sort on the
// tile id alone (<= 16 bits: 2 passes over R pairs of 32-bit keys) then yields exactly the reference's order
// (tile, depth bits, surfel index): LSD radix sorting is stable, so sorting by the minor key first and the major
// key second is the same permutation as sorting by the concatenated key.
//
// This kernel fuses what the reference does in three steps (InclusiveSum over tiles_touched, duplicateWithKeys,
// the sort's histogram pass): a single-pass decoupled-look-back scan of the per-surfel tile counts in depth order,
// emission of the instances through shared memory (coalesced stores), and the digit histograms of the tile sort.
// =============================================================================
constexpr int EM_THREADS = 256;
constexpr int EM_ITEMS = 4;
constexpr int EM_TILE = EM_THREADS * EM_ITEMS;  // surfels per CTA
constexpr int EM_STAGE = 4096;                  // instances staged per round

size_t emit_state_bytes(int P) {
  const int tiles = (P + EM_TILE - 1) / EM_TILE;
  return 256 + (size_t)tiles * sizeof(unsigned long long);
}

__global__ void __launch_bounds__(EM_THREADS) emit_instances_kernel(EmitArgs a) {
  __shared__ uint32_t s_keys[EM_STAGE];
  __shared__ uint32_t s_vals[EM_STAGE];
  __shared__ uint32_t s_hist[RS_MAX_PASSES_TILE * RS_RADIX];
  __shared__ uint32_t s_warp[EM_THREADS / 32];
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_excl;
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(a.counter, 1u);
  for (int i = tid; i < a.plan.passes * RS_RADIX; i += EM_THREADS) s_hist[i] = 0;
  // tile ranges start as {max, 0}: the last pass of the tile-id sort combines starts / ends with min / max
  for (int i = blockIdx.x * EM_THREADS + tid; i < a.ntiles; i += gridDim.x * EM_THREADS)
    a.ranges[i] = make_uint2(0xffffffffu, 0u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const int base = tile * EM_TILE + tid * EM_ITEMS;

  uint32_t id[EM_ITEMS], x0[EM_ITEMS], y0[EM_ITEMS], w[EM_ITEMS], cnt[EM_ITEMS];
  uint32_t tsum = 0;
#pragma unroll
  for (int i = 0; i < EM_ITEMS; i++) {
    id[i] = 0; x0[i] = 0; y0[i] = 0; w[i] = 1; cnt[i] = 0;
    if (base + i < a.P) {
      id[i] = a.sorted_ids[base + i];
      const uint2 rc = __ldg(&a.rect[id[i]]);
      x0[i] = rc.x & 0xffffu;
      y0[i] = rc.y & 0xffffu;
      w[i] = (rc.x >> 16) - x0[i];
      cnt[i] = w[i] * ((rc.y >> 16) - y0[i]);
      if (w[i] == 0) w[i] = 1;
    }
    tsum += cnt[i];
  }
  // block-wide exclusive scan of the thread sums
  uint32_t ws = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, ws, o);
    if (lane >= (unsigned)o) ws += t;
  }
  if (lane == 31) s_warp[wid] = ws;
  __syncthreads();
  uint32_t warp_off = 0, agg = 0;
#pragma unroll
  for (int i = 0; i < EM_THREADS / 32; i++) {
    if (i < (int)wid) warp_off += s_warp[i];
    agg += s_warp[i];
  }
  const uint32_t thread_excl = warp_off + ws - tsum;

  // decoupled look-back over the CTA aggregates, one warp wide: lane l polls tile (t - l)
  if (wid == 0) {
    uint32_t excl = 0;
    if (tile == 0) {
      if (lane == 0) atomicExch(&a.state[0], SCAN_FLAG_INC | agg);
    } else {
      if (lane == 0) atomicExch(&a.state[tile], SCAN_FLAG_AGG | agg);
      int t = (int)tile - 1;
      while (true) {
        const int mine = t - (int)lane;
        const unsigned long long sv = mine >= 0 ? *((volatile unsigned long long*)&a.state[mine]) : (unsigned long long)(2ull << 32);
        const unsigned f = (unsigned)(sv >> 32);
        const unsigned pending = __ballot_sync(0xffffffffu, f == 0), inc = __ballot_sync(0xffffffffu, f == 2);
        // usable prefix of the window: up to the first inclusive value, and before the first unpublished tile
        const int n_inc = inc ? __ffs(inc) : 33, n_pend = pending ? __ffs(pending) - 1 : 32;
        const int take = min(n_inc, n_pend);  // lanes [0, take) contribute
        uint32_t c = ((int)lane < take) ? (uint32_t)sv : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        excl += c;
        if (n_inc <= n_pend) break;  // reached an inclusive prefix
        t -= take;
      }
      if (lane == 0) atomicExch(&a.state[tile], SCAN_FLAG_INC | (unsigned long long)(uint32_t)(excl + agg));
    }
    if (lane == 0) {
      s_excl = excl;
      if ((int)tile == (a.P + EM_TILE - 1) / EM_TILE - 1) *a.total = excl + agg;  // number of instances of the frame
    }
  }
  __syncthreads();
  const uint32_t cta_excl = s_excl;
  // A speculative launch whose arena is too small: nothing may be written past the capacity.  (The frame is then
  // re-launched with a larger arena once the host has read the total; sums beyond 2^32 are caught there as well.)
  if ((unsigned long long)cta_excl + agg > (unsigned long long)a.capacity) return;

  const int passes = a.plan.passes;
  for (uint32_t win = 0; win < agg; win += EM_STAGE) {
    uint32_t o = thread_excl;
#pragma unroll
    for (int i = 0; i < EM_ITEMS; i++) {
      const uint32_t lo = max(win, o), hi = min(win + (uint32_t)EM_STAGE, o + cnt[i]);
      if (lo < hi) {
        const uint32_t k0 = lo - o;
        uint32_t ty = y0[i] + k0 / w[i], tx = k0 % w[i];
        for (uint32_t k = lo; k < hi; k++) {
          s_keys[k - win] = ty * a.gx + x0[i] + tx;
          s_vals[k - win] = id[i];
          if (++tx == w[i]) { tx = 0; ty++; }
        }
      }
      o += cnt[i];
    }
    __syncthreads();
    const uint32_t n_here = min((uint32_t)EM_STAGE, agg - win);
    for (uint32_t j0 = 0; j0 < n_here; j0 += EM_THREADS) {
      const uint32_t j = j0 + tid;
      const bool ok = j < n_here;
      uint32_t dh = 0xffffffffu;  // lanes past the end: a digit nobody else has
      if (ok) {
        const uint32_t key = s_keys[j];
        a.keys[cta_excl + win + j] = key;
        a.vals[cta_excl + win + j] = s_vals[j];
        // digit histograms of the tile sort: the low digits of a surfel's row of tiles are distinct ...
        for (int p = 0; p + 1 < passes; p++)
          atomicAdd(&s_hist[p * RS_RADIX + ((key >> a.plan.shift[p]) & ((1u << a.plan.bits[p]) - 1u))], 1u);
        dh = (key >> a.plan.shift[passes - 1]) & ((1u << a.plan.bits[passes - 1]) - 1u);
      }
      // ... but neighbours share the high digit: one shared-memory atomic per distinct digit in the warp
      const unsigned peers = __match_any_sync(0xffffffffu, dh);
      if (ok && (int)lane == __ffs(peers) - 1) atomicAdd(&s_hist[(passes - 1) * RS_RADIX + dh], (uint32_t)__popc(peers));
    }
    __syncthreads();
  }
  for (int i = tid; i < passes * RS_RADIX; i += EM_THREADS) {
    const uint32_t c = s_hist[i];
    if (c) atomicAdd(&a.hist[i], c);
  }
}

void launch_emit_instances(const EmitArgs& a, void* state_mem, cudaStream_t s) {
  if (a.P <= 0) return;
  EmitArgs b = a;
  cudaMemsetAsync(state_mem, 0, emit_state_bytes(a.P), s);
  b.counter = (uint32_t*)state_mem;
  b.state = (unsigned long long*)((char*)state_mem + 256);
  emit_instances_kernel<<<(a.P + EM_TILE - 1) / EM_TILE, EM_THREADS, 0, s>>>(b);
  count_launch();
}

// (tile << 32 | depth bits) of the sorted instance list, rebuilt from the sorted tile ids and the point list: the
// reference's BinningState::point_list_keys (rasterizer_impl.cu:187-194) for stage-wise parity checks.  The
// production path never materialises these keys.
__global__ void __launch_bounds__(256) rebuild_sorted_keys_kernel(int L, const uint32_t* __restrict__ tile_keys,
                                                                  const uint32_t* __restrict__ point_list,
                                                                  const float4* __restrict__ rec,
                                                                  uint64_t* __restrict__ keys) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L) return;
  const uint32_t depth_bits = __float_as_uint(__ldg(&rec[(size_t)point_list[idx] * REC_QUADS + 3]).w);
  keys[idx] = ((uint64_t)tile_keys[idx] << 32) | depth_bits;
}
void launch_rebuild_sorted_keys(int L, const uint32_t* tile_keys, const uint32_t* point_list, const float4* rec,
                                uint64_t* keys, cudaStream_t s) {
  if (L <= 0) return;
  rebuild_sorted_keys_kernel<<<(L + 255) / 256, 256, 0, s>>>(L, tile_keys, point_list, rec, keys);
  count_launch();
}

// =============================================================================
// identifyTileRanges
// =============================================================================
template <typename KeyT>
__global__ void __launch_bounds__(256) identify_tile_ranges_kernel(int L_cap, const KeyT* __restrict__ keys,
                                                                   uint2* ranges, const uint32_t* __restrict__ n_dev) {
  constexpr int TSHIFT = sizeof(KeyT) == 8 ? 32 : 0;  // u64: tile | depth keys; u32: plain tile ids
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int L = resolve_count(L_cap, n_dev);
  if (idx >= L) return;
  uint32_t currtile = (uint32_t)(keys[idx] >> TSHIFT);
  if (idx == 0)
    ranges[currtile].x = 0;
  else {
    uint32_t prevtile = (uint32_t)(keys[idx - 1] >> TSHIFT);
    if (currtile != prevtile) {
      ranges[prevtile].y = idx;
      ranges[currtile].x = idx;
    }
  }
  if (idx == L - 1) ranges[currtile].y = L;
}

void launch_identify_tile_ranges(int L, const uint64_t* keys, uint2* ranges, cudaStream_t s, const uint32_t* n_dev) {
  if (L <= 0) return;
  identify_tile_ranges_kernel<uint64_t><<<(L + 255) / 256, 256, 0, s>>>(L, keys, ranges, n_dev);
  count_launch();
}
void launch_identify_tile_ranges32(int L, const uint32_t* tile_keys, uint2* ranges, cudaStream_t s,
                                   const uint32_t* n_dev) {
  if (L <= 0) return;
  identify_tile_ranges_kernel<uint32_t><<<(L + 255) / 256, 256, 0, s>>>(L, tile_keys, ranges, n_dev);
  count_launch();
}

}  // namespace pgs
