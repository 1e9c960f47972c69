// data3d_check.cpp -- replays tests/data3d_ops.inc on THIS repo's Data3D (include/flow3d/data3d.h);
// tests/test_data3d_cpp_cpu.py compares what it writes with the files the reference's own Data3D wrote
// (tests/golden/data3d/, made by scripts/make_data3d_golden.sh).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <string>

#include "flow3d/data3d.h"

#define DATA3D_DELETE_AT_END 1
#include "data3d_ops.inc"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  return run_ops(std::string(argv[1]));  // every volume is deleted at the end: a failed read must leave nothing to double-free
}
