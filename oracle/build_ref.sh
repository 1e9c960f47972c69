#!/usr/bin/env bash
# Compiles the reference's OWN E-path sources, where they lie under /root/reference, into
# oracle/_ref/ (git-ignored; travels to the GPU box).  No reference source is copied into the repo.
# Flags follow the reference Makefile (:5-6,49-58): g++ -std=c++11 -O3 -D NO_VISUALIZATION for host
# code, `nvcc -ptx -std=c++11` (no -arch, no fast-math) for the kernels, PTX loaded at run time from
# <exe dir>/kernels/.  Excluded (do not build with CUDA 12.9 / off the hot path): partial_data/*,
# optical_flow_p.*, solve_p_3d.cu, registration_p_3d.cu, utils/gl/*, visualization.*, main.cpp.
#
# Also emits kernels_guarded/: the same kernels compiled from temporary copies with bounds guards
# (the shipped blur kernels write out of bounds when a dimension is not a multiple of 4,
# and the shipped median reads a wild address on thin volumes; details below).
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

# guarded variants (temporary patched copies, never committed):
#  * convolution_3d.cu: the three blur kernels get bounds guards on the two non-convolved axes;
#  * median_3d.cu, solve_3d.cu: the IND() macro clamps its coordinates to the level extent.  The
#    shipped kernels mirror halo cells as `2*dim - i - 2`, which goes negative (-> a wild size_t
#    index) for cells beyond 2*dim-2; those cells are never used by an in-range voxel, but the LOAD
#    faults on a B200 (depth 5, blockDim.z 4 => illegal address in median_3d on the 584x388x5 pair).
#    Clamping changes no value that any output depends on.
TMP="$(mktemp -d)"
python3 - "$REF/src/kernels" "$TMP" <<'PY'
import sys
src_dir, tmp = sys.argv[1], sys.argv[2]
src = open(src_dir + "/convolution_3d.cu").read()
guards = {
    "convolutionRowsKernel": "    if (baseY >= imageH || baseZ >= imageD) return;\n",
    "convolutionColumnsKernel": "    if (baseX >= imageW || baseZ >= imageD) return;\n",
    "convolutionSlicesKernel": "    if (baseX >= imageW || baseY >= imageH) return;\n",
}
# whole thread groups that leave share no shared-memory cell with in-range threads (s_Data is
# indexed by the two guarded thread coordinates) and exited threads do not block the barrier.
ins = []
for name, guard in guards.items():
    k = src.index("void " + name)
    last = max(src.index("const int base" + ax, k) for ax in "XYZ")
    ins.append((src.index("\n", last) + 1, guard))
for e, guard in sorted(ins, reverse=True):
    src = src[:e] + guard + src[e:]
open(tmp + "/convolution_3d.cu", "w").write(src)

helper = """
__device__ __forceinline__ size_t f3d_clamp(long long v, size_t n) {
  return v < 0 ? 0 : (v >= (long long)n ? n - 1 : (size_t)v);
}
#undef IND
#define IND(X, Y, Z) ((f3d_clamp((long long)(Z), depth) * container_size.height + f3d_clamp((long long)(Y), height)) * (container_size.pitch / sizeof(float)) + f3d_clamp((long long)(X), width))
"""
for name in ("median_3d.cu", "solve_3d.cu"):
    s = open(src_dir + "/" + name).read()
    k = s.index("__constant__ DataSize4 container_size;")
    e = s.index("\n", k) + 1
    open(tmp + "/" + name, "w").write(s[:e] + helper + s[e:])
PY
for k in convolution_3d median_3d solve_3d; do
  "$CUDA/bin/nvcc" -ptx -std=c++11 -w -I"$REF" -I"$CUDA/include" "$TMP/$k.cu" -o "$OUT/kernels_guarded/$k.ptx"
done
rm -rf "$TMP" "$OUT/obj"
echo "reference built into $OUT"
