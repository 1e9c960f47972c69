#!/usr/bin/env bash
set -u
O=gpurun_out
timeout -s KILL 420 python -m pytest tests/test_sweep_variants_gpu.py tests/test_fast_div_gpu.py tests/test_stages_gpu.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -8 > $O/r2c_tests.log
cat $O/r2c_tests.log
python scripts/run_stage.py sweep --variant 1 --nchunks 8 --reps 10 > $O/r2c_stage.txt 2>&1
python scripts/run_stage.py sweep --variant 1 --nchunks 4 --reps 10 >> $O/r2c_stage.txt 2>&1
python scripts/run_stage.py sweep --variant 2 --nchunks 8 --reps 10 >> $O/r2c_stage.txt 2>&1
python scripts/run_stage.py sweep --variant 0 --nchunks 10 --reps 10 >> $O/r2c_stage.txt 2>&1
python scripts/run_stage.py sweep --variant 0 --vec 2 --nchunks 9 --reps 10 >> $O/r2c_stage.txt 2>&1
python scripts/run_stage.py sweep --variant 1 --nchunks 8 --reps 10 --dims 487x487x487 >> $O/r2c_stage.txt 2>&1
python scripts/run_stage.py sweep --variant 0 --nchunks 10 --reps 10 --dims 487x487x487 >> $O/r2c_stage.txt 2>&1
cat $O/r2c_stage.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_tma -s 1 -c 1 -f -o $O/r2c_prof_sweep_tma \
   python scripts/run_stage.py sweep --variant 1 --nchunks 8 --reps 2 > $O/r2c_ncu_sweep_tma.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1 -c 1 -f -o $O/r2c_prof_sweep_reg \
   python scripts/run_stage.py sweep --variant 0 --nchunks 10 --reps 2 > $O/r2c_ncu_sweep_reg.log 2>&1
FLOW3D_TUNE_LOG=1 timeout -s KILL 600 python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline \
   > $O/r2c_bench.json 2> $O/r2c_bench.err
head -c 1200 $O/r2c_bench.json
