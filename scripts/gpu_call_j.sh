#!/bin/bash
O=gpurun_out; T=${1:-r1p}
( time timeout 600 python -m pytest tests -m gpu -q -x ) > $O/${T}_pytest.log 2>&1; tail -4 $O/${T}_pytest.log
for sz in 128; do
timeout 200 python bench.py --size $sz --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_bench_$sz.json 2> $O/${T}_bench_$sz.err
python -c "
import json;d=json.load(open('$O/${T}_bench_$sz.json'));s=d['stage_ms_per_step'];print($sz, d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], s)"
FLOW3D_AUTOTUNE=0 timeout 200 python bench.py --size $sz --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/${T}_bench_${sz}_notune.json 2> $O/${T}_bench_${sz}_notune.err
python -c "
import json;d=json.load(open('$O/${T}_bench_${sz}_notune.json'));s=d['stage_ms_per_step'];print('notune', $sz, d['ms_per_step'], d['value'], s)"
done
timeout 400 python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 2 > $O/${T}_bench.json 2> $O/${T}_bench.err
python -c "
import json;d=json.load(open('$O/${T}_bench.json'));s=d['stage_ms_per_step'];print(d['ms_per_step'], d['value'], s, d['clocks'])"
