// levels_golden_driver.cpp -- calls the REFERENCE's own OpticalFlowBase::GetMaxWarpLevel
// (src/optical_flow/optical_flow_base.cpp:31-56, compiled from /root/reference by
// scripts/make_levels_golden.sh) on a grid of volume sizes and scale factors and prints one line per case:
//   W H D scale_bits max_level
// The level count decides the whole pyramid, so flow3d_max_warp_level must agree with it exactly
// (tests/test_cabi_cpu.py::test_max_warp_level_equals_the_reference_function).
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "src/optical_flow/optical_flow_base.h"

class Probe : public OpticalFlowBase {
 public:
  Probe() : OpticalFlowBase("probe") {}
  bool Initialize(const DataSize4&) override { return true; }
  size_t MaxLevel(size_t w, size_t h, size_t d, float s) const { return GetMaxWarpLevel(w, h, d, s); }
};

int main() {
  Probe p;
  const size_t dims[] = {1, 2, 3, 4, 5, 7, 8, 16, 17, 31, 64, 100, 128, 257, 388, 512, 584, 1000, 1024, 2048};
  const float scales[] = {0.95f, 0.9f, 0.8f, 0.75f, 0.5f, 0.99f, 0.3f, 1.0f, 1.5f, 0.949999f, 0.6180339f};
  const size_t nd = sizeof(dims) / sizeof(dims[0]);
  for (float s : scales) {
    uint32_t bits;
    std::memcpy(&bits, &s, 4);
    for (size_t i = 0; i < nd; ++i)
      for (size_t j = 0; j < nd; j += (i % 3) + 1)
        for (size_t k = 0; k < nd; k += (j % 4) + 1)
          std::printf("%zu %zu %zu %u %zu\n", dims[i], dims[j], dims[k], bits, p.MaxLevel(dims[i], dims[j], dims[k], s));
  }
  return 0;
}
