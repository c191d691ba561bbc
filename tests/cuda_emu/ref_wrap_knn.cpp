// TEST INFRASTRUCTURE: extern "C" door into the UNMODIFIED reference simple-knn (SimpleKNN::knn, simple_knn.cu:185-221)
// compiled for the CPU emulator (tests/cuda_emu/build.py::build_reference("knn")).
#include "simple_knn.h"

extern "C" void ref_knn(int P, float* points /* [P,3] */, float* mean_dists /* [P] */) {
  SimpleKNN::knn(P, (float3*)points, mean_dists);
}
