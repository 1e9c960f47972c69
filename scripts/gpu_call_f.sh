#!/bin/bash
O=gpurun_out; T=${1:-r1k}
timeout 300 python -m pytest tests/test_diagnostics_gpu.py tests/test_cpp_api.py -m gpu -q 2>&1 | tail -3
for rep in 1 2; do
timeout 100 python scripts/run_stage.py sweep --size 512 --reps 40
timeout 100 python scripts/run_stage.py sweep --size 512 --reps 40 --alias-f
timeout 100 python scripts/run_stage.py sweep --size 512 --reps 40 --alias-most
done
