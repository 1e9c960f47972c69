#!/usr/bin/env bash
# Golden Gaussian taps printed by the reference's own host function, compiled from /root/reference (never
# copied): tests/golden/gauss_taps.txt
set -euo pipefail
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${REF:-/root/reference}
TMP=$(mktemp -d)
ln -s /usr/local/cuda/lib64/stubs/libcuda.so "$TMP/libcuda.so.1"
g++ -std=c++11 -O1 -w -DNO_VISUALIZATION -I"$REF" -I/usr/local/cuda/include "$ROOT/scripts/taps_golden_driver.cpp" \
    "$REF/src/cuda_operations/entire_data/cuda_operation_convolution.cpp" "$REF/src/cuda_operations/cuda_operation_base.cpp" \
    "$REF/src/utils/cuda_utils.cpp" "$REF/src/utils/common_utils.cpp" "$REF/src/data_types/operation_parameters.cpp" \
    "$REF/src/data_types/data3d.cpp" -L/usr/local/cuda/lib64/stubs -lcuda -o "$TMP/gen"
LD_LIBRARY_PATH="$TMP" "$TMP/gen" > "$ROOT/tests/golden/gauss_taps.txt"
wc -l "$ROOT/tests/golden/gauss_taps.txt"
rm -rf "$TMP"
