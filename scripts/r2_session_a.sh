#!/usr/bin/env bash
# round-2 GPU session A: parity of the new kernel variants, then the bench with per-candidate tune log
set -u
mkdir -p gpurun_out
timeout -s KILL 420 python -m pytest tests/test_sweep_variants_gpu.py -q -x 2>&1 | tail -25 > gpurun_out/r2a_variants.log
echo "variants rc=$?" >> gpurun_out/r2a_variants.log
timeout -s KILL 600 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2a_pytest.log
FLOW3D_TUNE_LOG=1 timeout -s KILL 600 python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline \
   > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 600 gpurun_out/r2a_variants.log; tail -5 gpurun_out/r2a_pytest.log; head -c 1500 gpurun_out/r2a_bench.json
