// warp_golden_driver.cpp -- runs the REFERENCE's own CPU backward registration: the loop inside
// CudaOperationRegistrationP::Execute (src/cuda_operations/partial_data/cuda_operation_register_p.cpp:96-139;
// everything after it in that function is commented out, so with `initialized_` forced it runs without a
// device).  Compiled from /root/reference by scripts/make_warp_golden.sh, linked against the CUDA driver stub.
// Reads f0, f1, u, v, w (float32 RAW, W x H x D) from argv[1], writes the warped volume next to them.
#include <cstdio>
#include <cstdlib>
#include <string>

#define private public
#define protected public
#include "src/cuda_operations/partial_data/cuda_operation_register_p.h"
#undef private
#undef protected
#include "src/data_types/data3d.h"
#include "src/data_types/data_structs.h"
#include "src/data_types/operation_parameters.h"

int main(int argc, char** argv) {
  if (argc < 8) return 2;
  const std::string dir = argv[1];
  size_t W = std::strtoull(argv[2], nullptr, 10), H = std::strtoull(argv[3], nullptr, 10), D = std::strtoull(argv[4], nullptr, 10);
  float hx = std::strtof(argv[5], nullptr), hy = std::strtof(argv[6], nullptr), hz = std::strtof(argv[7], nullptr);
  Data3D *f0 = new Data3D, *f1 = new Data3D, *u = new Data3D, *v = new Data3D, *w = new Data3D, *tmp = new Data3D(W, H, D);
  if (!f0->ReadRAWFromFileF32((dir + "/f0.raw").c_str(), W, H, D) || !f1->ReadRAWFromFileF32((dir + "/f1.raw").c_str(), W, H, D) ||
      !u->ReadRAWFromFileF32((dir + "/u.raw").c_str(), W, H, D) || !v->ReadRAWFromFileF32((dir + "/v.raw").c_str(), W, H, D) ||
      !w->ReadRAWFromFileF32((dir + "/w.raw").c_str(), W, H, D))
    std::_Exit(3);
  DataSize4 size = {W, H, D, 0};
  size_t max_mag = 0;
  OperationParameters p;
  p.PushValuePtr("frame_0", f0); p.PushValuePtr("frame_1", f1);
  p.PushValuePtr("flow_u", u); p.PushValuePtr("flow_v", v); p.PushValuePtr("flow_w", w);
  p.PushValuePtr("temp", tmp);
  p.PushValuePtr("hx", &hx); p.PushValuePtr("hy", &hy); p.PushValuePtr("hz", &hz);
  p.PushValuePtr("data_size", &size); p.PushValuePtr("max_mag", &max_mag);
  CudaOperationRegistrationP* op = new CudaOperationRegistrationP;
  op->initialized_ = true;
  op->Execute(p);
  const bool ok = f1->WriteRAWToFileF32((dir + "/warped_ref_cpu.raw").c_str());  // Execute swapped the result into frame_1
  std::fflush(nullptr);
  std::_Exit(ok ? 0 : 4);
}
