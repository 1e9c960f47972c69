#!/usr/bin/env bash
# Golden table of the reference's own level-count function, compiled from /root/reference (never copied):
# tests/golden/max_warp_level.txt.xz
set -euo pipefail
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${REF:-/root/reference}
TMP=$(mktemp -d)
g++ -std=c++11 -O1 -w -DNO_VISUALIZATION -I"$REF" -I/usr/local/cuda/include "$ROOT/scripts/levels_golden_driver.cpp" \
    "$REF/src/optical_flow/optical_flow_base.cpp" "$REF/src/data_types/data3d.cpp" \
    "$REF/src/data_types/operation_parameters.cpp" -o "$TMP/gen"
"$TMP/gen" | xz -9 > "$ROOT/tests/golden/max_warp_level.txt.xz"
xz -dc "$ROOT/tests/golden/max_warp_level.txt.xz" | wc -l
rm -rf "$TMP"
