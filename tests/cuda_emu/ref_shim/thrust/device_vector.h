// TEST INFRASTRUCTURE: the sliver of thrust::device_vector that simple_knn.cu uses, over host memory (CPU emulator).
#pragma once
#include <vector>
namespace thrust {
template <typename T> struct device_ptr_shim {
  T* p;
  T* get() const { return p; }
};
template <typename T> class device_vector {
  std::vector<T> v_;
 public:
  device_vector() {}
  explicit device_vector(size_t n) : v_(n) {}
  device_ptr_shim<T> data() { return {v_.data()}; }
  typename std::vector<T>::iterator begin() { return v_.begin(); }
  typename std::vector<T>::iterator end() { return v_.end(); }
  size_t size() const { return v_.size(); }
  void resize(size_t n) { v_.resize(n); }
};
}  // namespace thrust
