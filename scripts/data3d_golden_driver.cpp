// data3d_golden_driver.cpp -- runs the REFERENCE's own Data3D (src/data_types/data3d.cpp, compiled from
// /root/reference by scripts/make_data3d_golden.sh) on a small deterministic volume and writes what it
// produces: the u8 / f32 RAW files, the VTK flow file and the outcome of reads of files of the right and
// of the wrong size.  tests/test_data3d_cpp_cpu.py replays the same operations on this repo's Data3D
// (include/flow3d/data3d.h) and compares byte for byte.  The operation list is shared: tests/data3d_ops.inc.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <string>

#include "src/data_types/data3d.h"

#define DATA3D_WIDTH(d) (d).Width()
#include "../tests/data3d_ops.inc"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const int rc = run_ops(std::string(argv[1]));
  std::fflush(nullptr);
  std::_Exit(rc);  // no destructors: the reference double-frees a volume whose read failed
}
