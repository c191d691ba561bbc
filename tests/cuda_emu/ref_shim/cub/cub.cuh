// TEST INFRASTRUCTURE: host stand-ins for the two CUB device algorithms the reference rasteriser calls
// (rasterizer_impl.cu:166,188,280,306), for running the reference's own CUDA source on the CPU emulator.
// Both are exact integer algorithms, so any correct implementation gives CUB's result: an inclusive prefix sum and a
// STABLE sort of (key, value) pairs on key bits [begin_bit, end_bit).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <numeric>
#include <vector>

namespace cub {
struct DeviceScan {
  template <typename In, typename Out>
  static cudaError_t InclusiveSum(void* temp, size_t& temp_bytes, In in, Out out, int n) {
    if (temp == nullptr) { temp_bytes = 16; return cudaSuccess; }
    std::partial_sum(in, in + n, out);
    return cudaSuccess;
  }
};
struct DeviceReduce {
  template <typename In, typename Out, typename Op, typename T>
  static cudaError_t Reduce(void* temp, size_t& temp_bytes, In in, Out out, int n, Op op, T init) {
    if (temp == nullptr) { temp_bytes = 16; return cudaSuccess; }
    T acc = init;
    for (int i = 0; i < n; i++) acc = op(acc, in[i]);
    *out = acc;
    return cudaSuccess;
  }
};
struct DeviceRadixSort {
  template <typename K, typename V>
  static cudaError_t SortPairs(void* temp, size_t& temp_bytes, const K* keys_in, K* keys_out, const V* vals_in,
                               V* vals_out, int n, int begin_bit = 0, int end_bit = sizeof(K) * 8) {
    if (temp == nullptr) { temp_bytes = 16; return cudaSuccess; }
    const K mask = (end_bit - begin_bit >= (int)sizeof(K) * 8) ? ~K(0) : (((K(1) << (end_bit - begin_bit)) - 1) << begin_bit);
    std::vector<int> idx(n);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return (keys_in[a] & mask) < (keys_in[b] & mask); });
    for (int i = 0; i < n; i++) { keys_out[i] = keys_in[idx[i]]; vals_out[i] = vals_in[idx[i]]; }
    return cudaSuccess;
  }
};
}  // namespace cub
