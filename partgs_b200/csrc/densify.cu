// Densification of the point-level surfel model as one planned compaction (sm_100a).
//
// Replaces TwoGaussianModel.densify_and_prune (games/block_mesh_splatting/scene/two_gaussian_model.py:341-423;
// densify_and_prune / _prune_optimizer / cat_tensors_to_optimizer: scene/gaussian_model.py:384-436,495-509).  The
// reference runs clone -> split -> prune as ~150 ATen kernels: five boolean-mask gathers and `torch.cat`s of every
// parameter and both Adam moments (each a full copy of the 232 B/surfel model + 464 B/surfel optimiser state) and a
// dozen host synchronisations for the mask sizes.  Here every surfel is classified once, the row order of the
// reference's final tensors
//       [surviving originals | surviving clones | surviving split children, replica-major]
// is produced by one scan, and every tensor is moved exactly once by a multi-tensor gather:
//   densify_classify_kernel   20 B read / surfel: code byte + per-CTA counts
//   densify_scan_kernel       exclusive scan of the per-CTA counts (1 CTA), totals for the host (its only read-back)
//   densify_map_kernel        output row -> source row map (4 B / output row), normal-draw row of every child
//   densify_gather_kernel     all parameters, moments and the semantic table in ONE launch, coalesced
//   densify_children_kernel   position / scale of the split children
// HBM-bound streaming; algorithmic bytes = 20 P + (4 + 2 * 4 * sum of row widths) * P_out.
//
// Decisions replay ATen's float arithmetic (IEEE division, precise expf/logf, `x / python_scalar` evaluated as
// x * (1.f / scalar)), so the selected / pruned sets equal the reference's.
#include "common.cuh"
#include "kernels.h"

namespace pgs {

constexpr unsigned DN_KEEP = 1u;       // original row survives (not split, not pruned)
constexpr unsigned DN_CLONE = 2u;      // its clone survives
constexpr unsigned DN_SPLIT = 4u;      // selected for splitting (consumes N rows of normal draws)
constexpr unsigned DN_CHILD = 8u;      // its split children survive
constexpr unsigned DN_CLONE_SEL = 16u; // selected for cloning (statistic)
constexpr int DN_THREADS = 256;

// torch.max propagates NaN, fmaxf drops it
__device__ __forceinline__ float nan_max(float a, float b) { return (a != a || b != b) ? (a + b) : fmaxf(a, b); }

__global__ void __launch_bounds__(DN_THREADS) densify_classify_kernel(
    int P, const float* __restrict__ accum, const float* __restrict__ denom, const float* __restrict__ scaling,
    const float* __restrict__ opacity, float max_grad, float dense_thr, float min_opacity, int use_ws, float ws_thr,
    float inv_divisor, unsigned char* __restrict__ code, uint32_t* __restrict__ block_tot /* [4][nblk] */,
    uint32_t* __restrict__ counts) {
  __shared__ uint32_t s_cnt[DN_THREADS / 32][5];
  const int i = blockIdx.x * DN_THREADS + threadIdx.x;
  unsigned c = 0;
  if (i < P) {
    float g = __fdiv_rn(accum[i], denom[i]);  // grads = xyz_gradient_accum / denom; grads[isnan] = 0
    if (g != g) g = 0.f;
    const float2 sc = reinterpret_cast<const float2*>(scaling)[i];
    const float s0 = expf(sc.x), s1 = expf(sc.y);
    const float smax = nan_max(s0, s1);
    const float gnorm = sqrtf(__fmul_rn(g, g));  // torch.norm(grads, dim=-1) over one element
    const bool clone_sel = (gnorm >= max_grad) && (smax <= dense_thr);
    const bool split_sel = (g >= max_grad) && (smax > dense_thr);
    const float sig = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-opacity[i])));  // torch.sigmoid
    const bool transparent = sig < min_opacity;
    const bool pruned = transparent || (use_ws && smax > ws_thr);
    // children: new_scaling = log(get_scaling / (0.8 N)); the final prune looks at exp(new_scaling)
    const float c0 = expf(logf(__fmul_rn(s0, inv_divisor))), c1 = expf(logf(__fmul_rn(s1, inv_divisor)));
    const bool child_pruned = transparent || (use_ws && nan_max(c0, c1) > ws_thr);
    if (!split_sel && !pruned) c |= DN_KEEP;
    if (clone_sel) c |= DN_CLONE_SEL;
    if (clone_sel && !pruned) c |= DN_CLONE;
    if (split_sel) c |= DN_SPLIT;
    if (split_sel && !child_pruned) c |= DN_CHILD;
    code[i] = (unsigned char)c;
  }
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned b0 = __ballot_sync(0xffffffffu, c & DN_KEEP), b1 = __ballot_sync(0xffffffffu, c & DN_CLONE),
                 b2 = __ballot_sync(0xffffffffu, c & DN_SPLIT), b3 = __ballot_sync(0xffffffffu, c & DN_CHILD),
                 b4 = __ballot_sync(0xffffffffu, c & DN_CLONE_SEL);
  if (lane == 0) {
    s_cnt[w][0] = __popc(b0);
    s_cnt[w][1] = __popc(b1);
    s_cnt[w][2] = __popc(b2);
    s_cnt[w][3] = __popc(b3);
    s_cnt[w][4] = __popc(b4);
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    uint32_t t = 0;
    for (int k = 0; k < DN_THREADS / 32; k++) t += s_cnt[k][threadIdx.x];
    if (threadIdx.x < 4) block_tot[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = t;
    else if (t) atomicAdd(&counts[4], t);
  }
}

// In-place exclusive scan of the four per-CTA count rows; totals -> counts[0..3].  One CTA of 1024 threads.
__global__ void __launch_bounds__(1024) densify_scan_kernel(int nblk, uint32_t* __restrict__ block_tot,
                                                            uint32_t* __restrict__ counts) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int row = 0; row < 4; row++) {
    uint32_t* v = block_tot + (size_t)row * nblk;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += 1024) {
      const int idx = base + (int)threadIdx.x;
      const uint32_t x = idx < nblk ? v[idx] : 0u;
      uint32_t incl = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += t;
      }
      if (lane == 31) s_warp[w] = incl;
      __syncthreads();
      uint32_t woff = 0;
      for (unsigned k = 0; k < w; k++) woff += s_warp[k];
      const uint32_t carry = s_carry;
      if (idx < nblk) v[idx] = carry + woff + incl - x;
      __syncthreads();
      if (threadIdx.x == 1023) s_carry = carry + woff + incl;
      __syncthreads();
    }
    if (threadIdx.x == 0) counts[row] = s_carry;
    __syncthreads();
  }
}

// Output row -> source row.  Within a CTA the ranks come from ballots; the CTA's offsets from the scan above.
__global__ void __launch_bounds__(DN_THREADS) densify_map_kernel(int P, const unsigned char* __restrict__ code,
                                                                 const uint32_t* __restrict__ block_off,
                                                                 const uint32_t* __restrict__ counts, int n_split,
                                                                 int* __restrict__ src_row,
                                                                 int* __restrict__ sample_row) {
  __shared__ uint32_t s_cnt[DN_THREADS / 32][4];
  const int i = blockIdx.x * DN_THREADS + threadIdx.x;
  const unsigned c = i < P ? code[i] : 0u;
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const unsigned flags[4] = {c & DN_KEEP, c & DN_CLONE, c & DN_SPLIT, c & DN_CHILD};
  uint32_t rank[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const unsigned b = __ballot_sync(0xffffffffu, flags[k]);
    rank[k] = __popc(b & lt);
    if (lane == 0) s_cnt[w][k] = __popc(b);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; k++) {
    for (unsigned q = 0; q < w; q++) rank[k] += s_cnt[q][k];
    rank[k] += block_off[(size_t)k * gridDim.x + blockIdx.x];
  }
  const uint32_t n_keep = counts[0], n_clone = counts[1], n_sel = counts[2], n_child = counts[3];
  if (flags[0]) src_row[rank[0]] = i;
  if (flags[1]) src_row[n_keep + rank[1]] = i;
  if (flags[3]) {
    for (int r = 0; r < n_split; r++) {
      const uint32_t child = (uint32_t)r * n_child + rank[3];  // replica-major, like .repeat(N, 1)
      src_row[n_keep + n_clone + child] = i;
      sample_row[child] = (int)((uint32_t)r * n_sel + rank[2]);  // row of the [N*Ns,3] normal draw
    }
  }
}

__global__ void __launch_bounds__(256) densify_gather_kernel(GatherTable t, int n_keep,
                                                             const int* __restrict__ src_row) {
  int k = 0;
#pragma unroll
  for (int i = 1; i < PGS_GATHER_MAX_TENSORS; i++)
    if (i < t.n && (int)blockIdx.x >= t.block_start[i]) k = i;
  const size_t base = ((size_t)blockIdx.x - t.block_start[k]) * (256 * 4);
  const float* __restrict__ src = t.src[k];
  float* __restrict__ dst = t.dst[k];
  const unsigned w = (unsigned)t.width[k];
  const bool zero_new = t.zero_new[k] != 0;
  const size_t n = t.numel[k];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const size_t o = base + (size_t)j * 256 + threadIdx.x;
    if (o < n) {
      const size_t row = o / w;
      const unsigned e = (unsigned)(o - row * w);
      float v = 0.f;
      if (!(zero_new && row >= (size_t)n_keep)) v = src[(size_t)src_row[row] * w + e];
      dst[o] = v;
    }
  }
}

// new_xyz = R(rotation) @ (z * [exp(scaling), 0]) + xyz;  new_scaling = log(exp(scaling) / (0.8 N))
// (two_gaussian_model.py:385-391; build_rotation: utils/general_utils.py:149-170)
__global__ void __launch_bounds__(256) densify_children_kernel(
    int n_children, const uint32_t* __restrict__ counts, const int* __restrict__ src_row,
    const int* __restrict__ sample_row, const float* __restrict__ z, const float* __restrict__ xyz_in,
    const float* __restrict__ scaling_in, const float* __restrict__ rotation_in, float inv_divisor,
    float* __restrict__ xyz_out, float* __restrict__ scaling_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_children) return;
  const size_t row = (size_t)counts[0] + counts[1] + (size_t)c;
  const size_t p = (size_t)src_row[row];
  const size_t zr = (size_t)sample_row[c];
  const float2 sc = reinterpret_cast<const float2*>(scaling_in)[p];
  const float s0 = expf(sc.x), s1 = expf(sc.y);
  const float a0 = __fmul_rn(z[3 * zr], s0), a1 = __fmul_rn(z[3 * zr + 1], s1), a2 = __fmul_rn(z[3 * zr + 2], 0.f);
  const float4 rq = reinterpret_cast<const float4*>(rotation_in)[p];
  const float norm = sqrtf(rq.x * rq.x + rq.y * rq.y + rq.z * rq.z + rq.w * rq.w);
  const float r = __fdiv_rn(rq.x, norm), x = __fdiv_rn(rq.y, norm), y = __fdiv_rn(rq.z, norm),
              q = __fdiv_rn(rq.w, norm);
  const float R00 = 1.f - 2.f * (y * y + q * q), R01 = 2.f * (x * y - r * q), R02 = 2.f * (x * q + r * y);
  const float R10 = 2.f * (x * y + r * q), R11 = 1.f - 2.f * (x * x + q * q), R12 = 2.f * (y * q - r * x);
  const float R20 = 2.f * (x * q - r * y), R21 = 2.f * (y * q + r * x), R22 = 1.f - 2.f * (x * x + y * y);
  xyz_out[3 * row] = (R00 * a0 + R01 * a1 + R02 * a2) + xyz_in[3 * p];
  xyz_out[3 * row + 1] = (R10 * a0 + R11 * a1 + R12 * a2) + xyz_in[3 * p + 1];
  xyz_out[3 * row + 2] = (R20 * a0 + R21 * a1 + R22 * a2) + xyz_in[3 * p + 2];
  scaling_out[2 * row] = logf(__fmul_rn(s0, inv_divisor));
  scaling_out[2 * row + 1] = logf(__fmul_rn(s1, inv_divisor));
}

int densify_blocks(int P) { return P > 0 ? (P + DN_THREADS - 1) / DN_THREADS : 0; }

void launch_densify_plan(int P, const float* accum, const float* denom, const float* scaling, const float* opacity,
                         float max_grad, float dense_thr, float min_opacity, int use_ws, float ws_thr,
                         float inv_divisor, unsigned char* code, uint32_t* block_off, uint32_t* counts,
                         cudaStream_t s) {
  cudaMemsetAsync(counts, 0, 8 * sizeof(uint32_t), s);
  const int nblk = densify_blocks(P);
  if (nblk == 0) return;
  densify_classify_kernel<<<nblk, DN_THREADS, 0, s>>>(P, accum, denom, scaling, opacity, max_grad, dense_thr,
                                                      min_opacity, use_ws, ws_thr, inv_divisor, code, block_off,
                                                      counts);
  densify_scan_kernel<<<1, 1024, 0, s>>>(nblk, block_off, counts);
  count_launch(2);
}

void launch_densify_map(int P, const unsigned char* code, const uint32_t* block_off, const uint32_t* counts,
                        int n_split, int* src_row, int* sample_row, cudaStream_t s) {
  const int nblk = densify_blocks(P);
  if (nblk == 0) return;
  densify_map_kernel<<<nblk, DN_THREADS, 0, s>>>(P, code, block_off, counts, n_split, src_row, sample_row);
  count_launch();
}

void launch_densify_gather(GatherTable& t, int n_keep, const int* src_row, cudaStream_t s) {
  int blocks = 0;
  for (int i = 0; i < t.n; i++) {
    t.block_start[i] = blocks;
    blocks += (int)((t.numel[i] + 1023) / 1024);
  }
  if (blocks == 0) return;
  densify_gather_kernel<<<blocks, 256, 0, s>>>(t, n_keep, src_row);
  count_launch();
}

void launch_densify_children(int n_children, const uint32_t* counts, const int* src_row, const int* sample_row,
                             const float* z, const float* xyz_in, const float* scaling_in, const float* rotation_in,
                             float inv_divisor, float* xyz_out, float* scaling_out, cudaStream_t s) {
  if (n_children <= 0) return;
  densify_children_kernel<<<(n_children + 255) / 256, 256, 0, s>>>(n_children, counts, src_row, sample_row, z, xyz_in,
                                                                   scaling_in, rotation_in, inv_divisor, xyz_out,
                                                                   scaling_out);
  count_launch();
}

}  // namespace pgs
