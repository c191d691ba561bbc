// TEST INFRASTRUCTURE: stub CUDA runtime for whole-library emulator builds (tests/cuda_emu).  "Device" memory is host
// memory, streams are synchronous, events are no-ops.  Only the calls partgs_b200/csrc makes are provided.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>

extern "C" {
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t) {
  memmove(dst, src, n);
  return cudaSuccess;
}
cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind) {
  memmove(dst, src, n);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
  memset(p, v, n);
  return cudaSuccess;
}
cudaError_t cudaMemset(void* p, int v, size_t n) {
  memset(p, v, n);
  return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaMalloc(void** p, size_t n) {
  *p = malloc(n ? n : 1);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void* p) {
  free(p);
  return cudaSuccess;
}
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
cudaError_t cudaGetDevice(int* d) {
  *d = 0;
  return cudaSuccess;
}
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "emulator"; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) {
  *e = (cudaEvent_t)malloc(1);
  return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) {
  *ms = 0.f;
  return cudaSuccess;
}
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
}
