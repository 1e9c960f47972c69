// taps_golden_driver.cpp -- calls the REFERENCE's own CudaOperationConvolution3D::ComputeGaussianKernel
// (src/cuda_operations/entire_data/cuda_operation_convolution.cpp:85-108; private, reached by re-declaring the
// access specifier for this translation unit only; compiled from /root/reference by
// scripts/make_taps_golden.sh, linked against the CUDA driver STUB because nothing here touches a device)
// for a list of sigmas, as Execute() calls it (precision 3, pixel size 1.0, :159), and prints per sigma
//   sigma_bits radius tap_bits...
#include <cstdint>
#include <cstdio>
#include <cstring>

#define private public
#include "src/cuda_operations/entire_data/cuda_operation_convolution.h"
#undef private

int main() {
  const float sigmas[] = {0.34f, 0.5f, 0.75f, 1.0f, 1.3f, 1.5f, 2.0f, 2.5f, 3.0f, 3.14159f, 4.0f, 5.0f, 5.33f, 7.9f, 10.5f};
  for (float s : sigmas) {
    CudaOperationConvolution3D op;
    op.ComputeGaussianKernel(s, 3, 1.0);
    uint32_t b;
    std::memcpy(&b, &s, 4);
    std::printf("%u %zu", b, op.kernel_radius_);
    for (size_t i = 0; i < op.kernel_length_; ++i) {
      std::memcpy(&b, &op.kernel_[i], 4);
      std::printf(" %u", b);
    }
    std::printf("\n");
  }
  std::fflush(nullptr);
  std::_Exit(0);
}
