#!/usr/bin/env bash
# both bench arms back to back (as the driver does) + sha256 comparison of the 512^3 flows
set -u
O=gpurun_out
timeout -s KILL 900 python bench.py --impl reference --steps 1 --warmup 1 > $O/r2f_ref.json 2> $O/r2f_ref.err
echo "ref rc=$?"; head -c 1500 $O/r2f_ref.json; echo
timeout -s KILL 900 python bench.py --steps 2 --warmup 3 > $O/r2f_bench.json 2> $O/r2f_bench.err
echo "ours rc=$?"
python - <<'PY'
import json
r=json.load(open('gpurun_out/r2f_ref.json')); o=json.load(open('gpurun_out/r2f_bench.json'))
print("ref ms", r.get('ms_per_step'), "ours ms", o['ms_per_step'], "e2e", o['e2e']['ms_per_step'])
print("same workload:", r['config']['workload']==o['config']['workload'])
print("sha equal:", r.get('flow_sha256')==o.get('flow_sha256'), r.get('flow_sha256'), o.get('flow_sha256'))
print(json.dumps(o.get('extra_configs'),indent=1)[:1800])
print(json.dumps(r.get('extra_configs'),indent=1)[:1200])
print(o.get('cpu_baseline'))
PY
