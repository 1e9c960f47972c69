// operation_parameters.h -- the reference's named-parameter bag: name -> pointer to a caller-owned
// variable (src/data_types/operation_parameters.h:24-33).  The pointed-to variables must outlive
// the call that reads them.  Solver keys and types: see include/flow3d_c.h (flow3d_params).
#ifndef FLOW3D_OPERATION_PARAMETERS_H_
#define FLOW3D_OPERATION_PARAMETERS_H_

#include <string>
#include <unordered_map>

class OperationParameters {
 public:
  OperationParameters() = default;
  bool PushValuePtr(std::string key, void* value_ptr);  // false if the key already exists
  void* GetValuePtr(std::string key) const;             // nullptr if absent
  void Clear();

 private:
  std::unordered_map<std::string, void*> map_;
};

#endif  // FLOW3D_OPERATION_PARAMETERS_H_
