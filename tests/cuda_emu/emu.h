// TEST INFRASTRUCTURE: a tiny lock-step CUDA emulator for functional tests of SIMPLE kernels on the CPU.
//
// The kernel source (.cu) is compiled unchanged by g++ (tests/cuda_emu/build.py only rewrites the
// `kernel<<<grid, block, smem, stream>>>(args)` launch statements into EMU_LAUNCH).  Every CUDA thread of a block is
// an OS thread; blocks run one after the other; __syncthreads is a block barrier, warp collectives (__ballot_sync,
// __shfl_*_sync) rendezvous the 32 lanes of a warp.  `__shared__` becomes `static` (one instance per kernel, which
// is what a block sees because blocks are sequential).  Supported: 1-D to 3-D grids / blocks (warps are 32 consecutive
// linear thread ids), static and dynamic shared memory, the atomics / intrinsics below, cp.async as an immediate copy
// (the product headers select their non-PTX helper bodies under PGS_EMU), and — for whole-library builds — a stub CUDA
// runtime (emu_runtime.cpp: memcpy/memset, no-op events and streams).  NOT a performance model and not bit-exact for
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
// Sense-reversing barrier: spin briefly, then yield (hundreds of OS threads share a few cores, so a waiting thread
// must give its core away quickly; a mutex + condition variable cost ~10x more per rendezvous here).
struct Barrier {
  std::atomic<int> waiting{0};
  std::atomic<unsigned> gen{0};
  int n = 0;
  void reset(int count) { n = count; waiting.store(0); }
  void wait() {
    const unsigned g = gen.load(std::memory_order_acquire);
    if (waiting.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
      waiting.store(0, std::memory_order_relaxed);
      gen.fetch_add(1, std::memory_order_release);
      return;
    }
    for (int spin = 0; gen.load(std::memory_order_acquire) == g; spin++)
      if (spin > 64) std::this_thread::yield();
  }
};
struct WarpState {
  Barrier bar;
  unsigned vals[32];
};
inline Barrier g_block_bar;
inline std::vector<WarpState>* g_warps = nullptr;
inline thread_local WarpState* t_warp = nullptr;
inline unsigned char* g_dyn_smem = nullptr;  // dynamic shared memory of the running block
inline thread_local unsigned t_lane = 0;
}  // namespace emu

inline thread_local uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

inline void __syncthreads() { emu::g_block_bar.wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::t_warp->bar.wait(); }
namespace emu { inline std::atomic<int> g_block_count{0}; }
inline int __syncthreads_count(int pred) {
  if (pred) emu::g_block_count.fetch_add(1);
  emu::g_block_bar.wait();
  const int n = emu::g_block_count.load();
  emu::g_block_bar.wait();
  if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) emu::g_block_count.store(0);
  emu::g_block_bar.wait();
  return n;
}

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
template <typename T> inline unsigned __match_any_sync(unsigned, T v) {
  static_assert(sizeof(T) == 4, "4-byte values only");
  unsigned all[32], u;
  memcpy(&u, &v, 4);
  emu_exchange(u, all);
  unsigned m = 0;
  for (int i = 0; i < 32; i++) m |= (unsigned)(all[i] == u) << i;
  return m;
}
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
  unsigned all[32], u;
  memcpy(&u, &v, 4);
  emu_exchange(u, all);
  const unsigned src = emu::t_lane + delta;
  T r;
  memcpy(&r, &all[src > 31 ? emu::t_lane : src], 4);
  return r;
}
inline void __trap() { abort(); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
inline float __double2float_rd(double d) {  // round towards -inf
  float f = (float)d;
  return ((double)f > d) ? nextafterf(f, -INFINITY) : f;
}
inline float __double2float_ru(double d) {  // round towards +inf
  float f = (float)d;
  return ((double)f < d) ? nextafterf(f, INFINITY) : f;
}
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float saturate(float x) { return fminf(fmaxf(x, 0.f), 1.f); }
using std::isfinite;
using std::isnan;
template <typename T> inline T __ldg(const T* p) { return *p; }
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }  // only so that unused helpers parse
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicExch(unsigned long long* p, unsigned long long v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline int atomicMin(int* p, int v) { int o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
inline int atomicMax(int* p, int v) { int o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
inline unsigned atomicMin(unsigned* p, unsigned v) { unsigned o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
inline double atomicAdd(double* p, double v) {
  double old, want;
  __atomic_load(p, &old, __ATOMIC_RELAXED);
  do { want = old + v; } while (!__atomic_compare_exchange(p, &old, &want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return old;
}
inline float atomicAdd(float* p, float v) {
  float old, want;
  __atomic_load(p, &old, __ATOMIC_RELAXED);
  do { want = old + v; } while (!__atomic_compare_exchange(p, &old, &want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return old;
}
using std::max;
using std::min;
// CUDA's mixed-type overloads (the int operand is converted like the hardware min/max do)
inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
inline unsigned max(unsigned a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
inline unsigned max(int a, unsigned b) { return (unsigned)a > b ? (unsigned)a : b; }
inline float max(float a, float b) { return fmaxf(a, b); }
inline float min(float a, float b) { return fminf(a, b); }
inline double max(float a, double b) { return fmax((double)a, b); }
inline double max(double a, float b) { return fmax(a, (double)b); }
inline double min(float a, double b) { return fmin((double)a, b); }
inline double min(double a, float b) { return fmin(a, (double)b); }

#define cudaMemsetAsync(p, v, n, s) (memset((p), (v), (n)), cudaSuccess)
template <class T> inline cudaError_t cudaFuncSetAttribute(T*, cudaFuncAttribute, int) { return cudaSuccess; }

// Run `body` for every thread of the grid; blocks one after the other in x-fastest order.  (A kernel thread that
// returns early simply waits at the end-of-block barrier; kernels that call __syncthreads after some threads have
// returned are not supported.)
inline void emu_run(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const unsigned nthreads = block.x * block.y * block.z;
  if (nthreads % 32 != 0 || nthreads == 0) throw std::runtime_error("emu: block size must be a multiple of 32");
  blockDim = block;
  gridDim = grid;
  std::vector<emu::WarpState> warps(nthreads / 32);
  for (auto& w : warps) w.bar.reset(32);
  std::vector<unsigned char> dyn(smem + 64);
  emu::g_dyn_smem = (unsigned char*)(((size_t)dyn.data() + 63) & ~(size_t)63);
  // one OS thread per CUDA thread of a block, reused for every block of the launch; a block barrier separates
  // consecutive blocks (their `__shared__` statics are the same storage)
  emu::g_block_bar.reset((int)nthreads);
  std::vector<std::thread> ts;
  ts.reserve(nthreads);
  for (unsigned t = 0; t < nthreads; t++) {
    ts.emplace_back([&, t] {
      threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
      emu::t_warp = &warps[t / 32];
      emu::t_lane = t % 32;
      for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
          for (unsigned bx = 0; bx < grid.x; bx++) {
            blockIdx = {bx, by, bz};
            body();
            emu::g_block_bar.wait();
          }
    });
  }
  for (auto& th : ts) th.join();
  emu::g_dyn_smem = nullptr;
}
#define EMU_LAUNCH(kernel, grid, block, smem, ...) \
  emu_run(dim3(grid), dim3(block), (size_t)(smem), [&] { kernel(__VA_ARGS__); })
