#!/bin/bash
O=gpurun_out; T=${1:-r1i}
( time timeout 300 python -m pytest tests -m gpu -x -q ) > $O/${T}_pytest.log 2>&1; tail -4 $O/${T}_pytest.log
for f in 0 1 0 1; do
FLOW3D_FUSE_KSI=$f timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > $O/${T}_bench_fuse$f.json 2> $O/${T}_bench_fuse$f.err
python -c "
import json;d=json.load(open('$O/${T}_bench_fuse$f.json'));print('fuse=$f', d['ms_per_step'], d['stage_ms_per_step']['phi_ksi'], d['stage_ms_per_step']['sweep'], d['clocks']['sm_mhz'])"
done
