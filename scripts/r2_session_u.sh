#!/usr/bin/env bash
set -u
O=gpurun_out
timeout -s KILL 600 python -m pytest tests/test_fast_div_gpu.py tests/test_stages_gpu.py tests/test_sweep_variants_gpu.py tests/test_golden.py tests/test_solve_gpu.py tests/test_slab_gpu.py -m gpu -q -x 2>&1 | tail -4
rm -f $O/r2u_stage.txt
run() { echo "$*" >> $O/r2u_stage.txt; env "$@" >> $O/r2u_stage.txt 2>&1; }
run FOO=1 python scripts/run_stage.py sweep --variant 0 --nchunks 10 --reps 10
run FLOW3D_SWEEP_ROT=1 python scripts/run_stage.py sweep --variant 0 --nchunks 10 --reps 10
run FLOW3D_SWEEP_SPEC=1 python scripts/run_stage.py sweep --variant 0 --nchunks 10 --reps 10
run FLOW3D_SWEEP_ROT=1 FLOW3D_SWEEP_SPEC=1 python scripts/run_stage.py sweep --variant 0 --nchunks 10 --reps 10
run FOO=1 python scripts/run_stage.py sweep --variant 0 --nchunks 10 --reps 10 --ksi
run FOO=1 python scripts/run_stage.py sweep --variant 0 --vec 2 --nchunks 6 --reps 10 --ksi
run FOO=1 python scripts/run_stage.py phi_ksi --reps 10
run FOO=1 python scripts/run_stage.py blur --reps 5
run FLOW3D_BLUR_SCALAR=1 python scripts/run_stage.py blur --reps 5
cat $O/r2u_stage.txt
timeout -s KILL 600 python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > $O/r2u_bench.json 2> $O/r2u_bench.err
python -c "
import json; d=json.load(open('$O/r2u_bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['stage_ms_per_step'])"
