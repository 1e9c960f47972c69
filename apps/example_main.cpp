// example_main.cpp -- the reference's driver sequence (src/main.cpp:118,150-185,236) written against
// this repo's headers: the same class names, calls and parameter bag compile and run unchanged.
//
// usage: example_main <frame0_u8.raw> <frame1_u8.raw> <W> <H> <D> <output_dir>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "flow3d/cuda_utils.h"
#include "flow3d/data3d.h"
#include "flow3d/data_structs.h"
#include "flow3d/operation_parameters.h"
#include "flow3d/optical_flow_e.h"

int main(int argc, char** argv) {
  if (argc < 7) {
    std::printf("usage: %s frame0_u8.raw frame1_u8.raw W H D output_dir\n", argv[0]);
    return 2;
  }
  const size_t width = std::strtoull(argv[3], nullptr, 10);
  const size_t height = std::strtoull(argv[4], nullptr, 10);
  const size_t depth = std::strtoull(argv[5], nullptr, 10);
  const std::string output_path = std::string(argv[6]) + "/";

  /* Optical flow variables (src/main.cpp:77-85) */
  size_t warp_levels_count = 40;
  float warp_scale_factor = 0.95f;
  size_t outer_iterations_count = 40;
  size_t inner_iterations_count = 5;
  float equation_alpha = 7.5f;
  float equation_smoothness = 0.001f;
  float equation_data = 0.001f;
  size_t median_radius = 5;
  float gaussian_sigma = 2.0f;

  CUcontext cu_context;
  if (!InitCudaContextWithFirstAvailableDevice(&cu_context)) return 1;

  Data3D frame_0, frame_1;
  if (!frame_0.ReadRAWFromFileU8(argv[1], width, height, depth) ||
      !frame_1.ReadRAWFromFileU8(argv[2], width, height, depth)) {
    return 2;
  }

  DataSize4 data_size = {width, height, depth, 0};
  OpticalFlowE optical_flow_e;
  if (optical_flow_e.Initialize(data_size)) {
    Data3D flow_u(width, height, depth);
    Data3D flow_v(width, height, depth);
    Data3D flow_w(width, height, depth);

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

    optical_flow_e.silent = true;
    optical_flow_e.ComputeFlow(frame_0, frame_1, flow_u, flow_v, flow_w, params);

    const std::string tail = "-" + std::to_string(width) + "-" + std::to_string(height) + "-" +
                             std::to_string(depth) + ".raw";
    flow_u.WriteRAWToFileF32((output_path + "flow-u" + tail).c_str());
    flow_v.WriteRAWToFileF32((output_path + "flow-v" + tail).c_str());
    flow_w.WriteRAWToFileF32((output_path + "flow-w" + tail).c_str());

    optical_flow_e.Destroy();
  }
  return 0;
}
