// optical_flow.cpp -- OpticalFlowBase / OpticalFlowE with the reference's interface and messages
// (src/optical_flow/optical_flow_base.cpp, optical_flow_e.cpp), implemented on the C ABI.
#include <cstdio>

#include "flow3d/cuda_utils.h"
#include "flow3d/optical_flow_e.h"
#include "flow3d_c.h"

OpticalFlowBase::OpticalFlowBase(const char* name) : name_(name) {}
OpticalFlowBase::~OpticalFlowBase() {}
const char* OpticalFlowBase::GetName() const { return name_; }

size_t OpticalFlowBase::GetMaxWarpLevel(size_t width, size_t height, size_t depth, float scale_factor) const {
  return flow3d_max_warp_level(width, height, depth, scale_factor);
}

bool OpticalFlowBase::IsInitialized() const {
  if (!initialized_) std::printf("Error: '%s' was not initialized.\n", name_);
  return initialized_;
}

void OpticalFlowBase::ComputeFlow(Data3D&, Data3D&, Data3D&, Data3D&, Data3D&, OperationParameters&) {
  std::printf("Warning: '%s' ComputeFlow() was not defined.\n", name_);
}

void OpticalFlowBase::Destroy() { initialized_ = false; }

OpticalFlowE::OpticalFlowE() : OpticalFlowBase("Optical Flow Single GPU") {}

OpticalFlowE::~OpticalFlowE() { Destroy(); }

bool OpticalFlowE::Initialize(const DataSize4& data_size) {
  Destroy();
  size_ = data_size;
  std::printf("Allocating memory on the device...\n");
  const double mb = flow3d_solver_workspace_bytes(data_size.width, data_size.height, data_size.depth) / (1024.0 * 1024.0);
  std::printf("Needed\t\t:\t%.0fMB\n", mb);
  last_status_ = flow3d_solver_create(data_size.width, data_size.height, data_size.depth, device_, &solver_);
  if (last_status_ != FLOW3D_OK) {
    std::printf("Initialization failed: %s. %s\n", flow3d_status_string(last_status_), flow3d_last_cuda_error());
    solver_ = nullptr;
    initialized_ = false;
    return false;
  }
  size_.pitch = flow3d_aligned_ld(data_size.width) * sizeof(float);
  std::printf("Allocated\t:\t%.0fMB\n", mb);
  initialized_ = true;
  return true;
}

namespace {
template <typename T>
bool get_param(const OperationParameters& p, const char* solver, const char* key, T* out) {
  void* v = p.GetValuePtr(key);
  if (!v) {
    std::printf("Operation: '%s'. Missing parameter '%s'.\n", solver, key);
    return false;
  }
  *out = *static_cast<T*>(v);
  return true;
}
}  // namespace

void OpticalFlowE::ComputeFlow(Data3D& frame_0, Data3D& frame_1, Data3D& flow_u, Data3D& flow_v,
                               Data3D& flow_w, OperationParameters& params) {
  if (!IsInitialized()) {
    last_status_ = FLOW3D_ERR_NOT_INITIALIZED;
    return;
  }
  flow3d_params p;
  const char* n = GetName();
  if (!get_param(params, n, "warp_levels_count", &p.warp_levels_count) ||
      !get_param(params, n, "warp_scale_factor", &p.warp_scale_factor) ||
      !get_param(params, n, "outer_iterations_count", &p.outer_iterations_count) ||
      !get_param(params, n, "inner_iterations_count", &p.inner_iterations_count) ||
      !get_param(params, n, "equation_alpha", &p.equation_alpha) ||
      !get_param(params, n, "equation_smoothness", &p.equation_smoothness) ||
      !get_param(params, n, "equation_data", &p.equation_data) ||
      !get_param(params, n, "median_radius", &p.median_radius) ||
      !get_param(params, n, "gaussian_sigma", &p.gaussian_sigma)) {
    last_status_ = FLOW3D_ERR_INVALID_ARG;
    return;
  }
  Data3D* vols[5] = {&frame_0, &frame_1, &flow_u, &flow_v, &flow_w};
  for (Data3D* v : vols) {
    if (v->Width() != size_.width || v->Height() != size_.height || v->Depth() != size_.depth || !v->DataPtr()) {
      std::printf("Error: '%s' volume dimensions do not match the initialized size.\n", n);
      last_status_ = FLOW3D_ERR_INVALID_ARG;
      return;
    }
  }
  std::printf("\nStarting optical flow computation...\n");
  last_status_ = flow3d_solver_compute_host(solver_, frame_0.DataPtr(), frame_1.DataPtr(), &p, flow_u.DataPtr(),
                                            flow_v.DataPtr(), flow_w.DataPtr());
  if (last_status_ != FLOW3D_OK) {
    std::printf("Error: '%s' failed: %s. %s\n", n, flow3d_status_string(last_status_), flow3d_last_cuda_error());
    return;
  }
  flow3d_solver_last_timing(solver_, last_ms_);
  std::printf("Total GPU computation time: % 4.4fs\n", last_ms_[0] / 1000.);
}

void OpticalFlowE::Destroy() {
  if (solver_) {
    flow3d_solver_destroy(solver_);
    solver_ = nullptr;
  }
  OpticalFlowBase::Destroy();
}

bool InitCudaContextWithFirstAvailableDevice(CUcontext* cu_context) {
  const int n = flow3d_device_count();
  if (n <= 0) {
    std::printf("There are no cuda capable devices.");
    return false;
  }
  if (flow3d_set_device(0) != FLOW3D_OK) return false;
  char name[128] = "?";
  flow3d_device_name(0, name, sizeof(name));
  std::printf("CUDA Device: %s. Launch timeout: %s\n", name, "No");
  // touch the device so the primary context exists; hand back an opaque non-null token
  void* p = nullptr;
  if (flow3d_malloc(&p, 256) != FLOW3D_OK) return false;
  flow3d_free(p);
  if (cu_context) *cu_context = reinterpret_cast<CUcontext>(static_cast<size_t>(1));
  return true;
}
