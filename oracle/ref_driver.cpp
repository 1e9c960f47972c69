// ref_driver.cpp -- TEST/BENCH INFRASTRUCTURE ONLY.
//
// Minimal driver around the UNMODIFIED reference sources (compiled by build_ref.sh straight from
// /root/reference into oracle/_ref/).  The reference's own main.cpp is a lab script with hard-coded
// dims and paths (src/main.cpp:72-74,106-138), so this file reproduces only its call sequence
// (src/main.cpp:118,150-185,236) with dims/paths/params taken from argv.
//
// usage: flow3d_ref <frame0.raw> <frame1.raw> <W> <H> <D> <u8|f32> <out_prefix|-> <reps> [key=value ...]
// Prints one line per repetition:  REF_SOLVE rep=<i> seconds=<wall time of ComputeFlow>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include <cuda.h>

#include "src/data_types/data3d.h"
#include "src/data_types/data_structs.h"
#include "src/data_types/operation_parameters.h"
#include "src/optical_flow/optical_flow_e.h"
#include "src/utils/cuda_utils.h"

int main(int argc, char** argv) {
  if (argc < 9) {
    std::fprintf(stderr, "usage: %s f0 f1 W H D u8|f32 out_prefix|- reps [key=value ...]\n", argv[0]);
    return 2;
  }
  const char* path0 = argv[1];
  const char* path1 = argv[2];
  size_t W = std::strtoull(argv[3], nullptr, 10);
  size_t H = std::strtoull(argv[4], nullptr, 10);
  size_t D = std::strtoull(argv[5], nullptr, 10);
  bool is_u8 = std::strcmp(argv[6], "u8") == 0;
  std::string out_prefix = argv[7];
  int reps = std::atoi(argv[8]);

  /* defaults: src/main.cpp:77-85 */
  size_t warp_levels_count = 40;
  float warp_scale_factor = 0.95f;
  size_t outer_iterations_count = 40;
  size_t inner_iterations_count = 5;
  float equation_alpha = 7.5f;
  float equation_smoothness = 0.001f;
  float equation_data = 0.001f;
  size_t median_radius = 5;
  float gaussian_sigma = 2.0f;
  for (int i = 9; i < argc; ++i) {
    std::string kv = argv[i];
    size_t eq = kv.find('=');
    if (eq == std::string::npos) continue;
    std::string k = kv.substr(0, eq), v = kv.substr(eq + 1);
    if (k == "warp_levels_count") warp_levels_count = std::strtoull(v.c_str(), nullptr, 10);
    else if (k == "warp_scale_factor") warp_scale_factor = std::strtof(v.c_str(), nullptr);
    else if (k == "outer_iterations_count") outer_iterations_count = std::strtoull(v.c_str(), nullptr, 10);
    else if (k == "inner_iterations_count") inner_iterations_count = std::strtoull(v.c_str(), nullptr, 10);
    else if (k == "equation_alpha") equation_alpha = std::strtof(v.c_str(), nullptr);
    else if (k == "equation_smoothness") equation_smoothness = std::strtof(v.c_str(), nullptr);
    else if (k == "equation_data") equation_data = std::strtof(v.c_str(), nullptr);
    else if (k == "median_radius") median_radius = std::strtoull(v.c_str(), nullptr, 10);
    else if (k == "gaussian_sigma") gaussian_sigma = std::strtof(v.c_str(), nullptr);
    else { std::fprintf(stderr, "unknown parameter %s\n", k.c_str()); return 2; }
  }

  CUcontext cu_context;
  if (!InitCudaContextWithFirstAvailableDevice(&cu_context)) return 1;

  Data3D frame_0, frame_1;
  bool ok = is_u8 ? (frame_0.ReadRAWFromFileU8(path0, W, H, D) && frame_1.ReadRAWFromFileU8(path1, W, H, D))
                  : (frame_0.ReadRAWFromFileF32(path0, W, H, D) && frame_1.ReadRAWFromFileF32(path1, W, H, D));
  if (!ok) { std::fprintf(stderr, "cannot read input frames\n"); return 1; }

  DataSize4 data_size = { W, H, D, 0 };
  OpticalFlowE optical_flow_e;
  optical_flow_e.silent = true;
  if (!optical_flow_e.Initialize(data_size)) { std::fprintf(stderr, "Initialize failed\n"); return 1; }

  Data3D flow_u(W, H, D), flow_v(W, H, D), flow_w(W, H, D);

  OperationParameters params;
  params.PushValuePtr("warp_levels_count", &warp_levels_count);
  params.PushValuePtr("warp_scale_factor", &warp_scale_factor);
  params.PushValuePtr("outer_iterations_count", &outer_iterations_count);
  params.PushValuePtr("inner_iterations_count", &inner_iterations_count);
  params.PushValuePtr("equation_alpha", &equation_alpha);
  params.PushValuePtr("equation_smoothness", &equation_smoothness);
  params.PushValuePtr("equation_data", &equation_data);
  params.PushValuePtr("median_radius", &median_radius);
  params.PushValuePtr("gaussian_sigma", &gaussian_sigma);

  for (int r = 0; r < reps; ++r) {
    auto t0 = std::chrono::steady_clock::now();
    optical_flow_e.ComputeFlow(frame_0, frame_1, flow_u, flow_v, flow_w, params);
    auto t1 = std::chrono::steady_clock::now();
    std::printf("REF_SOLVE rep=%d seconds=%.6f\n", r, std::chrono::duration<double>(t1 - t0).count());
    std::fflush(stdout);
  }

  if (out_prefix != "-") {
    flow_u.WriteRAWToFileF32((out_prefix + "_u.raw").c_str());
    flow_v.WriteRAWToFileF32((out_prefix + "_v.raw").c_str());
    flow_w.WriteRAWToFileF32((out_prefix + "_w.raw").c_str());
  }
  optical_flow_e.Destroy();
  cuCtxDestroy(cu_context);
  return 0;
}
