#!/bin/bash
O=gpurun_out; T=${1:-r1j}
( time timeout 400 python -m pytest tests -m gpu -q ) > $O/${T}_pytest.log 2>&1; tail -4 $O/${T}_pytest.log
for f in 1 0 1 0; do
FLOW3D_RESAMPLE_SCALAR=$f timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > $O/${T}_bench_scalar$f.json 2> $O/${T}_bench_scalar$f.err
python -c "
import json;d=json.load(open('$O/${T}_bench_scalar$f.json'));s=d['stage_ms_per_step'];print('scalar_resample=$f', d['ms_per_step'], s['resample'], s['phi_ksi'], s['sweep'], s['median'], d['clocks']['sm_mhz'])"
done
