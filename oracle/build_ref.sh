#!/usr/bin/env bash
# Compiles the reference's OWN E-path sources, where they lie under /root/reference, into
# oracle/_ref/ (git-ignored; travels to the GPU box).  No reference source is copied into the repo.
# Flags follow the reference Makefile (:5-6,49-58): g++ -std=c++11 -O3 -D NO_VISUALIZATION for host
# code, `nvcc -ptx -std=c++11` (no -arch, no fast-math) for the kernels, PTX loaded at run time from
# <exe dir>/kernels/.  Excluded (do not build with CUDA 12.9 / off the hot path): partial_data/*,
# optical_flow_p.*, solve_p_3d.cu, registration_p_3d.cu, utils/gl/*, visualization.*, main.cpp.
#
# Also emits kernels_guarded/: identical PTX except that convolution_3d.cu is compiled from a
# temporary copy with y/z(x) bounds guards added to the three blur kernels (the
# shipped kernels write out of bounds when a dimension is not a multiple of 4).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
CUDA="${CUDA_HOME:-/usr/local/cuda}"
CXX=g++
if [ ! -d "$REF/src" ]; then echo "reference not present at $REF (GPU box uses prebuilt files)"; exit 0; fi
mkdir -p "$OUT/kernels" "$OUT/kernels_guarded" "$OUT/obj"

HOST_SRCS=(
  src/optical_flow/optical_flow_base.cpp
  src/optical_flow/optical_flow_e.cpp
  src/cuda_operations/cuda_operation_base.cpp
  src/cuda_operations/entire_data/cuda_operation_add.cpp
  src/cuda_operations/entire_data/cuda_operation_convolution.cpp
  src/cuda_operations/entire_data/cuda_operation_median.cpp
  src/cuda_operations/entire_data/cuda_operation_registration.cpp
  src/cuda_operations/entire_data/cuda_operation_resample.cpp
  src/cuda_operations/entire_data/cuda_operation_solve.cpp
  src/data_types/data3d.cpp
  src/data_types/operation_parameters.cpp
  src/utils/cuda_utils.cpp
  src/utils/common_utils.cpp
)
OBJS=()
for s in "${HOST_SRCS[@]}"; do
  o="$OUT/obj/$(echo "$s" | tr '/' '_' | sed 's/\.cpp$/.o/')"
  $CXX -c -std=c++11 -O3 -funroll-all-loops -Wno-deprecated -w -D NO_VISUALIZATION \
       -I"$REF" -I"$CUDA/include" "$REF/$s" -o "$o"
  OBJS+=("$o")
done
$CXX -c -std=c++11 -O2 -w -D NO_VISUALIZATION -I"$REF" -I"$CUDA/include" "$HERE/ref_driver.cpp" -o "$OUT/obj/ref_driver.o"
$CXX "$OUT/obj/ref_driver.o" "${OBJS[@]}" -L"$CUDA/lib64/stubs" -lcuda -o "$OUT/flow3d_ref"

for k in add_3d median_3d convolution_3d registration_3d resample_3d solve_3d; do
  "$CUDA/bin/nvcc" -ptx -std=c++11 -w -I"$REF" -I"$CUDA/include" "$REF/src/kernels/$k.cu" -o "$OUT/kernels/$k.ptx"
  cp "$OUT/kernels/$k.ptx" "$OUT/kernels_guarded/$k.ptx"
done

# guarded blur variant (temporary patched copy, never committed)
TMP="$(mktemp -d)"
python3 - "$REF/src/kernels/convolution_3d.cu" "$TMP/convolution_3d.cu" <<'PY'
import re, sys
src = open(sys.argv[1]).read()
guards = {
    "convolutionRowsKernel": "    if (baseY >= imageH || baseZ >= imageD) return;\n",
    "convolutionColumnsKernel": "    if (baseX >= imageW || baseZ >= imageD) return;\n",
    "convolutionSlicesKernel": "    if (baseX >= imageW || baseY >= imageH) return;\n",
}
# insert each guard right after the declaration of baseZ inside the named kernel.  The guard sits
# before __syncthreads(); whole (y,z) [or (x,z)/(x,y)] thread groups that leave never shared data
# with in-range threads along the convolved axis, but they do share the barrier -- so instead of
# returning we predicate the global loads/stores: rewrite to a flag.
out = []
pos = 0
for name, guard in guards.items():
    k = src.index("void " + name)
    b = src.index("const int baseZ", k)
    e = src.index("\n", b) + 1
    out.append((e, guard))
res = src
for e, guard in sorted(out, reverse=True):
    res = res[:e] + guard + res[e:]
open(sys.argv[2], "w").write(res)
PY
"$CUDA/bin/nvcc" -ptx -std=c++11 -w -I"$REF" -I"$CUDA/include" "$TMP/convolution_3d.cu" -o "$OUT/kernels_guarded/convolution_3d.ptx"
rm -rf "$TMP" "$OUT/obj"
echo "reference built into $OUT"
