// flow3d_synth -- standalone generator of the analytic synthetic volume pair.
//
// usage: flow3d_synth W H D SEED frame0.raw frame1.raw
// Writes the two frames as tight float32 RAW (x fastest), the format Data3D::ReadRAWFromFileF32 reads
// (reference: src/data_types/data3d.cpp:136-165).
//
// The kernel source (csrc/kernels_synth.cu) is compiled INTO this binary, which does not link
// libflow3d_b200.so: bench.py's reference arm runs it as a separate process, so the baseline process
// tree never maps the library under test while both arms still solve bit-identical inputs.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../cuda_flow3d_b200/csrc/kernels_synth.cu"

namespace f3d {
static std::string g_err;
void note_cuda_error(cudaError_t e, const char* what) { g_err = std::string(what) + ": " + cudaGetErrorString(e); }
void count_launch(unsigned) {}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e, what); return FLOW3D_ERR_CUDA; }
  return FLOW3D_OK;
}
}  // namespace f3d

int main(int argc, char** argv) {
  if (argc < 7) {
    std::fprintf(stderr, "usage: %s W H D SEED frame0.raw frame1.raw\n", argv[0]);
    return 2;
  }
  const size_t W = std::strtoull(argv[1], nullptr, 10), H = std::strtoull(argv[2], nullptr, 10),
               D = std::strtoull(argv[3], nullptr, 10);
  const uint64_t seed = std::strtoull(argv[4], nullptr, 10);
  if (W == 0 || H == 0 || D == 0) return 2;
  const size_t plane = W * H;
  // a slab of planes at a time: bounded device and host memory whatever the volume size
  const size_t slab = std::max<size_t>(1, std::min<size_t>(D, (size_t(256) << 20) / (plane * sizeof(float))));
  float *d0 = nullptr, *d1 = nullptr;
  if (cudaMalloc(&d0, slab * plane * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&d1, slab * plane * sizeof(float)) != cudaSuccess) {
    std::fprintf(stderr, "cudaMalloc failed\n");
    return 1;
  }
  std::vector<float> h0(slab * plane), h1(slab * plane);
  FILE* f0 = std::fopen(argv[5], "wb");
  FILE* f1 = std::fopen(argv[6], "wb");
  if (!f0 || !f1) { std::fprintf(stderr, "cannot open outputs\n"); return 1; }
  for (size_t z0 = 0; z0 < D; z0 += slab) {
    const size_t nz = std::min(slab, D - z0);
    if (f3d::launch_synth(W, H, D, z0, nz, W, seed, d0, d1, nullptr, nullptr, nullptr, nullptr) != FLOW3D_OK ||
        cudaMemcpy(h0.data(), d0, nz * plane * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(h1.data(), d1, nz * plane * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) {
      std::fprintf(stderr, "synthesis failed: %s\n", f3d::g_err.c_str());
      return 1;
    }
    if (std::fwrite(h0.data(), sizeof(float), nz * plane, f0) != nz * plane ||
        std::fwrite(h1.data(), sizeof(float), nz * plane, f1) != nz * plane) {
      std::fprintf(stderr, "short write\n");
      return 1;
    }
  }
  if (std::fclose(f0) != 0 || std::fclose(f1) != 0) return 1;
  cudaFree(d0);
  cudaFree(d1);
  return 0;
}
