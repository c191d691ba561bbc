// TEST INFRASTRUCTURE: the sliver of cooperative_groups the reference rasteriser uses (this_grid().thread_rank(),
// this_thread_block().{sync, thread_rank, group_index, thread_index}), on top of the CPU emulator (emu.h).
#pragma once
namespace cooperative_groups {
struct grid_group {
  unsigned long long thread_rank() const {
    const unsigned long long b = blockIdx.x + (unsigned long long)gridDim.x * (blockIdx.y + (unsigned long long)gridDim.y * blockIdx.z);
    const unsigned t = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    return b * (blockDim.x * blockDim.y * blockDim.z) + t;
  }
};
struct thread_block {
  void sync() const { __syncthreads(); }
  unsigned thread_rank() const { return threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z); }
  dim3 group_index() const { return dim3(blockIdx.x, blockIdx.y, blockIdx.z); }
  dim3 thread_index() const { return dim3(threadIdx.x, threadIdx.y, threadIdx.z); }
};
inline grid_group this_grid() { return grid_group(); }
inline thread_block this_thread_block() { return thread_block(); }
}  // namespace cooperative_groups
