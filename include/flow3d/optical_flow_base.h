// optical_flow_base.h -- abstract solver base (reference: src/optical_flow/optical_flow_base.h:24-45).
#ifndef FLOW3D_OPTICAL_FLOW_BASE_H_
#define FLOW3D_OPTICAL_FLOW_BASE_H_

#include "flow3d/data3d.h"
#include "flow3d/data_structs.h"
#include "flow3d/operation_parameters.h"

class OpticalFlowBase {
 public:
  const char* GetName() const;
  virtual bool Initialize(const DataSize4& data_size) = 0;
  virtual void ComputeFlow(Data3D& frame_0, Data3D& frame_1, Data3D& flow_u, Data3D& flow_v,
                           Data3D& flow_w, OperationParameters& params);
  virtual void Destroy();
  virtual ~OpticalFlowBase();

 protected:
  explicit OpticalFlowBase(const char* name);
  // pyramid depth rule of the reference (optical_flow_base.cpp:31-56)
  size_t GetMaxWarpLevel(size_t width, size_t height, size_t depth, float scale_factor) const;
  bool IsInitialized() const;
  bool initialized_ = false;

 private:
  const char* name_ = nullptr;
};

#endif  // FLOW3D_OPTICAL_FLOW_BASE_H_
