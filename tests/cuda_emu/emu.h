// TEST INFRASTRUCTURE: a tiny lock-step CUDA emulator for functional tests of SIMPLE kernels on the CPU.
//
// The kernel source (.cu) is compiled unchanged by g++ (tests/cuda_emu/build.py only rewrites the
// `kernel<<<grid, block, smem, stream>>>(args)` launch statements into EMU_LAUNCH).  Every CUDA thread of a block is
// an OS thread; blocks run one after the other; __syncthreads is a block barrier, warp collectives (__ballot_sync,
// __shfl_*_sync) rendezvous the 32 lanes of a warp.  `__shared__` becomes `static` (one instance per kernel, which
// is what a block sees because blocks are sequential).  Supported: 1-D grids / blocks that are multiples of 32,
// static shared memory, atomicAdd, the arithmetic intrinsics below.  NOT a performance model and not bit-exact for
// transcendental functions (glibc expf/logf vs the GPU's) — it checks indexing, scans, barriers and data movement.
#pragma once
#include <cuda_runtime.h>  // vector types, dim3, cudaStream_t (host-compilable header of the toolkit)

#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)

namespace emu {
struct Barrier {
  std::mutex m;
  std::condition_variable cv;
  int n = 0, waiting = 0;
  unsigned long gen = 0;
  void reset(int count) { n = count; waiting = 0; }
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const unsigned long g = gen;
    if (++waiting == n) { waiting = 0; gen++; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};
struct WarpState {
  Barrier bar;
  unsigned vals[32];
};
inline Barrier g_block_bar;
inline std::vector<WarpState>* g_warps = nullptr;
inline thread_local WarpState* t_warp = nullptr;
inline thread_local unsigned t_lane = 0;
}  // namespace emu

inline thread_local uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

inline void __syncthreads() { emu::g_block_bar.wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::t_warp->bar.wait(); }

inline unsigned emu_exchange(unsigned v, unsigned (&out)[32]) {
  emu::WarpState& w = *emu::t_warp;
  w.vals[emu::t_lane] = v;
  w.bar.wait();
  for (int i = 0; i < 32; i++) out[i] = w.vals[i];
  w.bar.wait();  // nobody overwrites vals before everyone has read them
  return v;
}
inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned all[32];
  emu_exchange(pred ? 1u : 0u, all);
  unsigned b = 0;
  for (int i = 0; i < 32; i++) b |= (all[i] & 1u) << i;
  return b;
}
template <typename T> inline T __shfl_sync(unsigned, T v, int src) {
  static_assert(sizeof(T) == 4, "4-byte shuffles only");
  unsigned all[32], u;
  memcpy(&u, &v, 4);
  emu_exchange(u, all);
  T r;
  memcpy(&r, &all[src & 31], 4);
  return r;
}
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
  unsigned all[32], u;
  memcpy(&u, &v, 4);
  emu_exchange(u, all);
  const int src = (int)emu::t_lane - (int)delta;
  T r;
  memcpy(&r, &all[src < 0 ? emu::t_lane : src], 4);
  return r;
}
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int mask) {
  unsigned all[32], u;
  memcpy(&u, &v, 4);
  emu_exchange(u, all);
  T r;
  memcpy(&r, &all[(emu::t_lane ^ mask) & 31], 4);
  return r;
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
template <typename T> inline T __ldg(const T* p) { return *p; }
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }  // only so that unused helpers parse
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline float atomicAdd(float* p, float v) {
  float old = *p, want;
  do { want = old + v; } while (!__atomic_compare_exchange(p, &old, &want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return old;
}
using std::max;
using std::min;

#define cudaMemsetAsync(p, v, n, s) (memset((p), (v), (n)), cudaSuccess)

// Run `body` for every thread of a 1-D grid of 1-D blocks.
inline void emu_run(unsigned grid, unsigned block, const std::function<void()>& body) {
  if (block % 32 != 0 || block == 0) throw std::runtime_error("emu: block size must be a multiple of 32");
  blockDim = dim3(block, 1, 1);
  gridDim = dim3(grid, 1, 1);
  std::vector<emu::WarpState> warps(block / 32);
  for (auto& w : warps) w.bar.reset(32);
  for (unsigned b = 0; b < grid; b++) {
    emu::g_block_bar.reset((int)block);
    std::vector<std::thread> ts;
    ts.reserve(block);
    for (unsigned t = 0; t < block; t++) {
      ts.emplace_back([&, t, b] {
        threadIdx = {t, 0, 0};
        blockIdx = {b, 0, 0};
        emu::t_warp = &warps[t / 32];
        emu::t_lane = t % 32;
        body();
      });
    }
    for (auto& th : ts) th.join();
  }
}
#define EMU_LAUNCH(kernel, grid, block, ...) emu_run((unsigned)(grid), (unsigned)(block), [&] { kernel(__VA_ARGS__); })
