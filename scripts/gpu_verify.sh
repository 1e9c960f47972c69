#!/bin/bash
# One gpurun call that re-verifies a build: GPU parity tests, smoke(), the default bench line.
O=gpurun_out; T=${1:-verify}
mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -q ) > $O/${T}_pytest.log 2>&1; tail -4 $O/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
python - <<PY
import json
d = json.load(open("$O/${T}_bench.json"))
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["stage_ms_per_step"], d["clocks"], d["roofline"]["frac"], d.get("endpoint_error", {}).get("interior_mean"))
PY
