#!/usr/bin/env bash
# Golden files for the host data formats (Data3D RAW u8 / f32, VTK flow): produced by the REFERENCE's own
# Data3D compiled from /root/reference (never copied into this repo), committed under tests/golden/data3d/.
set -euo pipefail
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${REF:-/root/reference}
OUT=$ROOT/tests/golden/data3d
TMP=$(mktemp -d)
g++ -std=c++11 -O1 -w -DNO_VISUALIZATION -I"$REF" -I/usr/local/cuda/include -I"$ROOT/scripts" \
    "$ROOT/scripts/data3d_golden_driver.cpp" "$REF/src/data_types/data3d.cpp" -o "$TMP/gen"
mkdir -p "$OUT" "$TMP/run"
"$TMP/gen" "$TMP/run" > "$TMP/stdout.txt"
for f in out_u8.raw out_f32.raw out_flow.vtk reread_u8_as_f32.raw reread_f32.raw swapped.raw results.json; do
  cp "$TMP/run/$f" "$OUT/$f"
done
cat "$OUT/results.json"
rm -rf "$TMP"
