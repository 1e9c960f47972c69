// oracle_sanitized_driver.cpp -- runs full pyramid solves of the CPU oracle (oracle/flow3d_oracle.cpp) compiled
// with AddressSanitizer + UndefinedBehaviorSanitizer on shapes that stress its index arithmetic (thin slab,
// 4-voxel levels, steep pyramid, every median radius) and dumps the flows; tests/test_oracle_sanitized_cpu.py
// checks that the sanitizers stay silent and that the flows equal the production oracle's bit for bit.
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <string>
#include <vector>

struct OracleParams {
  size_t warp_levels_count;
  float warp_scale_factor;
  size_t outer_iterations_count;
  size_t inner_iterations_count;
  float equation_alpha, equation_smoothness, equation_data;
  size_t median_radius;
  float gaussian_sigma;
};
typedef void (*o_level_cb)(int, const size_t*, const float*, const float*, const float*, void*);
extern "C" int o_compute_flow(const float*, const float*, size_t, size_t, size_t, const OracleParams*, float*, float*,
                              float*, o_level_cb, void*);

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string dir = argv[1];
  const size_t shapes[][3] = {{28, 24, 20}, {70, 40, 5}, {9, 7, 33}, {4, 4, 4}, {17, 5, 6}};  // W, H, D
  const size_t meds[] = {5, 3, 7, 1, 5};
  for (int c = 0; c < 5; ++c) {
    const size_t W = shapes[c][0], H = shapes[c][1], D = shapes[c][2], N = W * H * D;
    std::vector<float> f0(N), f1(N), u(N), v(N), w(N);
    for (size_t i = 0; i < N; ++i) {
      f0[i] = 120.f + 60.f * std::sin(0.37f * (float)i);
      f1[i] = 120.f + 60.f * std::sin(0.37f * (float)(i + 1));
    }
    const OracleParams p = {40, c == 2 ? 0.5f : 0.95f, 3, 5, 7.5f, 0.001f, 0.001f, meds[c], c == 3 ? 0.f : 2.f};
    if (o_compute_flow(f0.data(), f1.data(), W, H, D, &p, u.data(), v.data(), w.data(), nullptr, nullptr) != 0) return 3;
    std::FILE* f = std::fopen((dir + "/case" + std::to_string(c) + ".raw").c_str(), "wb");
    if (!f) return 4;
    std::fwrite(f0.data(), 4, N, f);
    std::fwrite(f1.data(), 4, N, f);
    std::fwrite(u.data(), 4, N, f);
    std::fwrite(v.data(), 4, N, f);
    std::fwrite(w.data(), 4, N, f);
    std::fclose(f);
  }
  return 0;
}
