// optical_flow.cpp -- OpticalFlowBase / OpticalFlowE with the reference's interface and messages
// (src/optical_flow/optical_flow_base.cpp, optical_flow_e.cpp), implemented on the C ABI.
#include <dlfcn.h>

#include <cstdio>
#include <string>
#include <vector>

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

namespace {
// flow3d_mgpu_compute_host of libflow3d_b200_mgpu.so, which sits next to this library; loaded on first
// use so that single-GPU programs never need NCCL
typedef int (*MgpuHostFn)(size_t, size_t, size_t, int, const int*, const float*, const float*, const flow3d_params*, float*,
                          float*, float*, float*, int);
MgpuHostFn mgpu_host_fn() {
  static MgpuHostFn fn = [] {
    Dl_info info;
    std::string dir;
    if (dladdr(reinterpret_cast<void*>(&flow3d_version), &info) && info.dli_fname) {
      dir = info.dli_fname;
      const size_t slash = dir.rfind('/');
      dir = slash == std::string::npos ? std::string() : dir.substr(0, slash + 1);
    }
    void* h = dlopen((dir + "libflow3d_b200_mgpu.so").c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libflow3d_b200_mgpu.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      std::printf("Error: cannot load libflow3d_b200_mgpu.so (%s)\n", dlerror());
      return static_cast<MgpuHostFn>(nullptr);
    }
    return reinterpret_cast<MgpuHostFn>(dlsym(h, "flow3d_mgpu_compute_host"));
  }();
  return fn;
}
}  // namespace

OpticalFlowE::OpticalFlowE() : OpticalFlowBase("Optical Flow Single GPU") {}

OpticalFlowE::~OpticalFlowE() { Destroy(); }

bool OpticalFlowE::Initialize(const DataSize4& data_size) {
  Destroy();
  size_ = data_size;
  if (devices_.size() > 1) {  // sharded solve: the ranks allocate their slabs on first use
    if (!mgpu_host_fn()) { last_status_ = FLOW3D_ERR_UNSUPPORTED; return false; }
    const int n = flow3d_device_count();
    for (int d : devices_)
      if (d < 0 || d >= n) { std::printf("Error: no CUDA device %d.\n", d); last_status_ = FLOW3D_ERR_INVALID_ARG; return false; }
    std::printf("Sharding along z over %zu devices.\n", devices_.size());
    size_.pitch = flow3d_aligned_ld(data_size.width) * sizeof(float);
    last_status_ = FLOW3D_OK;
    initialized_ = true;
    return true;
  }
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
  if (devices_.size() > 1) {
    float ms = 0.f;
    last_status_ = mgpu_host_fn()(size_.width, size_.height, size_.depth, (int)devices_.size(), devices_.data(),
                                  frame_0.DataPtr(), frame_1.DataPtr(), &p, flow_u.DataPtr(), flow_v.DataPtr(),
                                  flow_w.DataPtr(), &ms, 1);
    if (last_status_ != FLOW3D_OK) {
      std::printf("Error: '%s' failed: %s. %s\n", n, flow3d_status_string(last_status_), flow3d_last_cuda_error());
      return;
    }
    last_ms_[0] = last_ms_[1] = ms;
    std::printf("Total GPU computation time: % 4.4fs\n", ms / 1000.);
    return;
  }
  flow3d_solver_set_verbose(solver_, silent ? 0 : 1);
  last_status_ = flow3d_solver_compute_host(solver_, frame_0.DataPtr(), frame_1.DataPtr(), &p, flow_u.DataPtr(),
                                            flow_v.DataPtr(), flow_w.DataPtr());
  if (last_status_ != FLOW3D_OK) {
    std::printf("Error: '%s' failed: %s. %s\n", n, flow3d_status_string(last_status_), flow3d_last_cuda_error());
    return;
  }
  flow3d_solver_last_timing(solver_, last_ms_);
  std::printf("Total GPU computation time: % 4.4fs\n", last_ms_[0] / 1000.);
}

bool OpticalFlowE::SetDiagnostics(bool enable, float update_tolerance) {
  if (!solver_) return false;
  last_status_ = flow3d_solver_set_diagnostics(solver_, enable ? 1 : 0, update_tolerance);
  return last_status_ == FLOW3D_OK;
}

void OpticalFlowE::PrintDiagnostics() const {
  if (!solver_) return;
  size_t nl = 0, nr = 0;
  if (flow3d_solver_diagnostics(solver_, &nl, nullptr, nullptr, nullptr, 0, &nr) != FLOW3D_OK || nl == 0) return;
  const size_t cap = nl > nr ? nl : nr;
  std::vector<size_t> per(cap);
  std::vector<double> rms(cap), mx(cap);
  if (flow3d_solver_diagnostics(solver_, &nl, per.data(), rms.data(), mx.data(), cap, &nr) != FLOW3D_OK) return;
  size_t k = 0;
  for (size_t l = 0; l < nl; ++l) {
    if (per[l] > 0)
      std::printf("level %2zu: outer %3zu  rms update %.3e -> %.3e  max %.3e\n", nl - 1 - l, per[l], rms[k],
                  rms[k + per[l] - 1], mx[k + per[l] - 1]);
    k += per[l];
  }
}

bool OpticalFlowE::WarpFrame(Data3D& frame_0, Data3D& frame_1, Data3D& flow_u, Data3D& flow_v, Data3D& flow_w,
                             Data3D& warped, Data3D* abs_error) {
  Data3D* vols[6] = {&frame_0, &frame_1, &flow_u, &flow_v, &flow_w, &warped};
  const size_t W = frame_0.Width(), H = frame_0.Height(), D = frame_0.Depth();
  for (Data3D* v : vols)
    if (v->Width() != W || v->Height() != H || v->Depth() != D || !v->DataPtr()) return false;
  if (abs_error && (abs_error->Width() != W || abs_error->Height() != H || abs_error->Depth() != D)) return false;
  const size_t dims[3] = {W, H, D};
  const size_t ld = flow3d_aligned_ld(W);
  const size_t bytes = ld * H * D * sizeof(float);
  void* dev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int rc = flow3d_set_device(device_);
  for (int i = 0; i < 6 && rc == FLOW3D_OK; ++i) rc = flow3d_malloc(&dev[i], bytes);
  for (int i = 0; i < 5 && rc == FLOW3D_OK; ++i)
    rc = flow3d_upload(vols[i]->DataPtr(), static_cast<float*>(dev[i]), dims, ld, nullptr);
  const float h[3] = {1.f, 1.f, 1.f};  // full resolution
  if (rc == FLOW3D_OK)
    rc = flow3d_warp(static_cast<float*>(dev[0]), static_cast<float*>(dev[1]), static_cast<float*>(dev[2]),
                     static_cast<float*>(dev[3]), static_cast<float*>(dev[4]), dims, ld, h,
                     static_cast<float*>(dev[5]), nullptr);
  if (rc == FLOW3D_OK) rc = flow3d_download(static_cast<float*>(dev[5]), warped.DataPtr(), dims, ld, nullptr);
  if (rc == FLOW3D_OK) rc = flow3d_stream_synchronize(nullptr);
  for (void* p : dev)
    if (p) flow3d_free(p);
  last_status_ = rc;
  if (rc != FLOW3D_OK) return false;
  if (abs_error) {
    const float* a = warped.DataPtr();
    const float* b = frame_0.DataPtr();
    float* e = abs_error->DataPtr();
    for (size_t i = 0; i < W * H * D; ++i) e[i] = a[i] > b[i] ? a[i] - b[i] : b[i] - a[i];
  }
  return true;
}

void OpticalFlowE::Destroy() {
  if (devices_.size() > 1 && initialized_ && mgpu_host_fn())  // release the persistent ranks
    mgpu_host_fn()(0, 0, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0);
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
