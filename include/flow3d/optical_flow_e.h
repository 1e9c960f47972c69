// optical_flow_e.h -- the all-on-GPU solver with the reference's interface
// (src/optical_flow/optical_flow_e.h:38-66): Initialize(DataSize4), ComputeFlow(...), Destroy(),
// public `silent`.  Everything below the class is the B200-native library (include/flow3d_c.h).
#ifndef FLOW3D_OPTICAL_FLOW_E_H_
#define FLOW3D_OPTICAL_FLOW_E_H_

#include <vector>

#include "flow3d/optical_flow_base.h"

struct flow3d_solver;

class OpticalFlowE : public OpticalFlowBase {
 public:
  OpticalFlowE();
  bool Initialize(const DataSize4& data_size) override;
  // Reads the nine named parameters (warp_levels_count, warp_scale_factor, outer_iterations_count,
  // inner_iterations_count, equation_alpha, equation_smoothness, equation_data, median_radius,
  // gaussian_sigma); a missing one prints "Missing parameter" and returns without computing, as the
  // reference does (optical_flow_e.cpp:150-158).  Blocks until flow_u/v/w hold the result.
  void ComputeFlow(Data3D& frame_0, Data3D& frame_1, Data3D& flow_u, Data3D& flow_v, Data3D& flow_w,
                   OperationParameters& params) override;
  void Destroy() override;
  ~OpticalFlowE() override;

  bool silent = false;

  // additions (not in the reference)
  int last_status() const { return last_status_; }      // FLOW3D_OK or the C-ABI error code
  float last_total_ms() const { return last_ms_[0]; }   // H2D + levels + D2H (the reference's bracket)
  float last_device_ms() const { return last_ms_[1]; }  // levels only
  void SetDevice(int device) { device_ = device; }      // before Initialize(); default 0
  // Several GPUs of this box (before Initialize()): ComputeFlow then z-shards the volume over them through
  // libflow3d_b200_mgpu.so (one host thread + one NCCL rank per device, loaded on first use; the result is
  // bit-identical to the single-GPU one).  Takes the place of the reference's second solver class for
  // volumes beyond one device, OpticalFlowP (src/optical_flow/optical_flow_p.cpp:58-323).
  void SetDevices(const std::vector<int>& devices) { devices_ = devices; if (!devices.empty()) device_ = devices[0]; }
  // Convergence diagnostics (flow3d_solver_set_diagnostics): record the Jacobi update norm per level
  // and outer iteration; update_tolerance > 0 stops a level early (results then differ from the
  // reference's fixed iteration count).  Call after Initialize().
  bool SetDiagnostics(bool enable, float update_tolerance = 0.f);
  // prints "level <k>: outer <n> rms <first> -> <last> max <last>" for the last ComputeFlow
  void PrintDiagnostics() const;
  // Registered volume: frame_1 warped back by the flow (what the reference's disabled debug block
  // optical_flow_e.cpp:535-571 wrote out), and optionally |warped - frame_0| as an error map.
  bool WarpFrame(Data3D& frame_0, Data3D& frame_1, Data3D& flow_u, Data3D& flow_v, Data3D& flow_w,
                 Data3D& warped, Data3D* abs_error = nullptr);

 private:
  flow3d_solver* solver_ = nullptr;
  DataSize4 size_{0, 0, 0, 0};
  int device_ = 0;
  std::vector<int> devices_;  // > 1 entries: sharded solve
  int last_status_ = 0;
  float last_ms_[2] = {0.f, 0.f};
};

#endif  // FLOW3D_OPTICAL_FLOW_E_H_
