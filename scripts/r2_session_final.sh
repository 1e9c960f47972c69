#!/usr/bin/env bash
# final single-GPU verification of the shipped tree: whole GPU suite, smoke(), both bench arms
set -u
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r2z_pytest_1gpu.log
tail -3 $O/r2z_pytest_1gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2z_ref.json 2> $O/r2z_ref.err; echo "ref rc=$?"
FLOW3D_TUNE_LOG=1 timeout -s KILL 900 python bench.py --steps 3 --warmup 3 > $O/r2z_bench.json 2> $O/r2z_bench.err; echo "ours rc=$?"
python - <<PY
import json
r=json.load(open("$O/r2z_ref.json")); o=json.load(open("$O/r2z_bench.json"))
print("ref", r["ms_per_step"], "ours", o["ms_per_step"], "e2e", o["e2e"]["ms_per_step"], "sha equal", r["flow_sha256"]==o["flow_sha256"], "frac", o["roofline"]["frac"], "clk", o["clocks"]["sm_mhz"])
print(o["stage_ms_per_step"])
print([ (e["workload"][:20], round(e["ms_per_solve"],1), e.get("matches_reference_build_sha256")) for e in o["extra_configs"]])
print([ (e["workload"][:20], round(e["ms_per_solve"],1), e.get("matches_golden_sha256")) for e in r["extra_configs"]])
PY
