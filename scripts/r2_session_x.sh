#!/usr/bin/env bash
# what is left of the GPU budget (75 s): one short bench.py run on the final tree -- the driver's own command
# with fewer steps -- to confirm the bench line, its flow sha256 (vs profiles/r02_bench_n1_512.json) and the
# effect of the PDL chain on the 512^3 solve
set -u
O=gpurun_out
timeout -s KILL 62 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $O/r2x_bench.json 2> $O/r2x_bench.err; echo "bench rc=$?"
python - <<PY
import json
o=json.load(open("$O/r2x_bench.json")); r=json.load(open("profiles/r02_bench_n1_512.json"))
print("ms", o["ms_per_step"], "e2e", o["e2e"]["ms_per_step"], "sha equal to r02 record", o["flow_sha256"]==r["flow_sha256"], "clk", o["clocks"]["sm_mhz"], "frac", o["roofline"]["frac"])
print(o["stage_ms_per_step"])
PY
