#!/usr/bin/env bash
set -u
O=gpurun_out
timeout -s KILL 300 python -m pytest tests/test_sweep_variants_gpu.py -q -x 2>&1 | tail -4 > $O/r2d_tests.log
cat $O/r2d_tests.log
for a in "--variant 1 --nchunks 8" "--variant 1 --nchunks 4" "--variant 1 --nchunks 16" "--variant 2 --nchunks 8" "--variant 0 --nchunks 10"; do
python scripts/run_stage.py sweep $a --reps 10 >> $O/r2d_stage.txt 2>&1
done
python scripts/run_stage.py sweep --variant 1 --nchunks 8 --reps 10 --dims 487x487x487 >> $O/r2d_stage.txt 2>&1
python scripts/run_stage.py sweep --variant 1 --nchunks 8 --reps 10 --ksi >> $O/r2d_stage.txt 2>&1
python scripts/run_stage.py sweep --variant 0 --nchunks 10 --reps 10 --ksi >> $O/r2d_stage.txt 2>&1
cat $O/r2d_stage.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_tma -s 1 -c 1 -f -o $O/r2d_prof_sweep_tma \
   python scripts/run_stage.py sweep --variant 1 --nchunks 8 --reps 2 > $O/r2d_ncu_sweep_tma.log 2>&1
