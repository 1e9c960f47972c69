#!/usr/bin/env bash
set -u
O=gpurun_out
timeout -s KILL 300 python -m pytest tests/test_sweep_variants_gpu.py -q -x 2>&1 | tail -4 > $O/r2e_tests.log
cat $O/r2e_tests.log
rm -f $O/r2e_stage.txt
for a in "--variant 0 --nchunks 10" "--variant 3 --nchunks 10" "--variant 3 --nchunks 16" "--variant 4 --nchunks 10" "--variant 4 --nchunks 20" \
         "--variant 0 --nchunks 10 --ksi" "--variant 3 --nchunks 10 --ksi" "--variant 4 --nchunks 10 --ksi" \
         "--variant 0 --nchunks 10 --dims 487x487x487" "--variant 3 --nchunks 10 --dims 487x487x487" "--variant 4 --nchunks 10 --dims 487x487x487"; do
echo "$a" >> $O/r2e_stage.txt
python scripts/run_stage.py sweep $a --reps 10 >> $O/r2e_stage.txt 2>&1
done
cat $O/r2e_stage.txt
FLOW3D_TUNE_LOG=1 timeout -s KILL 600 python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline \
   > $O/r2e_bench.json 2> $O/r2e_bench.err
head -c 900 $O/r2e_bench.json; echo; python -c "
import json; d=json.load(open('$O/r2e_bench.json')); print(d['stage_ms_per_step'], d['tune_seconds_untimed'], d['roofline']['frac'])"
