#!/usr/bin/env bash
# Golden vectors for the backward warp from the REFERENCE's own CPU registration loop
# (cuda_operation_register_p.cpp:96-139), compiled from /root/reference (never copied):
# tests/golden/warp_cpu/{f0,f1,u,v,w,warped_ref_cpu}.raw + case.json
set -euo pipefail
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${REF:-/root/reference}
OUT=$ROOT/tests/golden/warp_cpu
TMP=$(mktemp -d)
mkdir -p "$OUT"
ln -s /usr/local/cuda/lib64/stubs/libcuda.so "$TMP/libcuda.so.1"
g++ -std=c++11 -O1 -w -DNO_VISUALIZATION -I"$REF" -I/usr/local/cuda/include "$ROOT/scripts/warp_golden_driver.cpp" \
    "$REF/src/cuda_operations/partial_data/cuda_operation_register_p.cpp" "$REF/src/cuda_operations/cuda_operation_base.cpp" \
    "$REF/src/utils/cuda_utils.cpp" "$REF/src/utils/common_utils.cpp" "$REF/src/data_types/operation_parameters.cpp" \
    "$REF/src/data_types/data3d.cpp" -L/usr/local/cuda/lib64/stubs -lcuda -o "$TMP/gen"
python3 - "$OUT" <<'PY'
import json, sys
import numpy as np
out = sys.argv[1]
W, H, D = 11, 9, 7
h = (1.0, 1.25, 2.0)
rng = np.random.default_rng(20240521)
f0 = rng.uniform(0, 255, (D, H, W)).astype(np.float32)
f1 = rng.uniform(0, 255, (D, H, W)).astype(np.float32)
u = rng.normal(0, 1.5, (D, H, W)).astype(np.float32)
v = rng.normal(0, 1.5, (D, H, W)).astype(np.float32)
w = rng.normal(0, 1.5, (D, H, W)).astype(np.float32)
# special voxels: integer shifts (zero fractions), exact borders, far outside, NaN, +-inf, -0
u[0, 0, :] = 1.0; v[0, 0, :] = 0.0; w[0, 0, :] = 0.0
u[1, 1, :] = 0.0; v[1, 1, :] = 1.25; w[1, 1, :] = 2.0
u[2, 2, 0] = -0.0; v[2, 2, 0] = -0.0; w[2, 2, 0] = -0.0
u[2, 2, 1] = np.nan
v[2, 2, 2] = np.inf
w[2, 2, 3] = -np.inf
u[2, 2, 4] = 1e9
u[3, 3, W - 1] = 0.0; v[3, 3, W - 1] = 0.0; w[3, 3, W - 1] = 0.0      # lands exactly on the last column
u[3, 4, 0] = (W - 1) * h[0]                                            # from x = 0 exactly to x = W-1
u[3, 5, 0] = np.nextafter(np.float32((W - 1) * h[0]), np.float32(np.inf))  # one ulp past it
for name, a in (("f0", f0), ("f1", f1), ("u", u), ("v", v), ("w", w)):
    a.tofile("%s/%s.raw" % (out, name))
json.dump({"W": W, "H": H, "D": D, "h": h}, open(out + "/case.json", "w"))
PY
LD_LIBRARY_PATH="$TMP" "$TMP/gen" "$OUT" 11 9 7 1.0 1.25 2.0
ls -la "$OUT"
rm -rf "$TMP"
