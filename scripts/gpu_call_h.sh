#!/bin/bash
O=gpurun_out; T=${1:-r1m}
( time timeout 600 python -m pytest tests -m gpu -q ) > $O/${T}_pytest.log 2>&1; tail -4 $O/${T}_pytest.log
timeout 400 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
python -c "
import json;d=json.load(open('$O/${T}_bench.json'));s=d['stage_ms_per_step'];print(d['ms_per_step'], d['value'], d['e2e']['value'], s, d['clocks'], d['roofline']['frac'])"
export FLOW3D_AUTOTUNE=0
for st in sweep median; do
  timeout 120 python scripts/run_stage.py $st --size 512 --reps 3 > $O/${T}_plain_$st.log 2>&1 &&
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sweep_kernel|median" -s 1 -c 1 -f \
    -o $O/${T}_prof_$st python scripts/run_stage.py $st --size 512 --reps 2 > $O/${T}_ncu_$st.log 2>&1
  cat $O/${T}_plain_$st.log
done
