// data_structs.h -- plain aggregates of the reference's host API
// (reference: src/data_types/data_structs.h:20-33; same names, same field order and types).
#ifndef FLOW3D_DATA_STRUCTS_H_
#define FLOW3D_DATA_STRUCTS_H_

#include <cstddef>

struct DataSize4 {
  size_t width;
  size_t height;
  size_t depth;
  size_t pitch;  // bytes per device row; filled in by the solver, callers pass 0
};

struct Stat3 {
  float min;
  float max;
  float avg;
};

#endif  // FLOW3D_DATA_STRUCTS_H_
