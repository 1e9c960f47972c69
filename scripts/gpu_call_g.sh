#!/bin/bash
O=gpurun_out; T=${1:-r1l}
( time timeout 600 python -m pytest tests -m gpu -q ) > $O/${T}_pytest.log 2>&1; tail -4 $O/${T}_pytest.log
timeout 400 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
python -c "
import json;d=json.load(open('$O/${T}_bench.json'));s=d['stage_ms_per_step'];print(d['ms_per_step'], d['value'], d['e2e']['value'], s, d['clocks'], d['roofline']['frac'])"
