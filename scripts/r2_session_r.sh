#!/usr/bin/env bash
# final single-GPU evidence: full GPU test suite, both bench arms, 1024^3 on one GPU, ncu captures
set -u
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/r2r_pytest_1gpu.log
tail -4 $O/r2r_pytest_1gpu.log
timeout -s KILL 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2r_ref.json 2> $O/r2r_ref.err; echo "ref rc=$?"
timeout -s KILL 900 python bench.py --steps 3 --warmup 3 > $O/r2r_bench.json 2> $O/r2r_bench.err; echo "ours rc=$?"
timeout -s KILL 600 python bench.py --size 1024 --steps 1 --warmup 1 --no-extra --no-cpu-baseline --no-e2e > $O/r2r_bench_1024_1gpu.json 2> $O/r2r_bench_1024_1gpu.err; echo "1024 rc=$?"
python - <<PY
import json
r=json.load(open("$O/r2r_ref.json")); o=json.load(open("$O/r2r_bench.json")); b=json.load(open("$O/r2r_bench_1024_1gpu.json"))
print("ref", r["ms_per_step"], "ours", o["ms_per_step"], "e2e", o["e2e"]["ms_per_step"], "sha equal", r["flow_sha256"]==o["flow_sha256"], "frac", o["roofline"]["frac"])
print(o["stage_ms_per_step"]); print("1024^3 1 GPU ms", b["ms_per_step"], b["stage_ms_per_step"])
print([ (e["workload"][:20], e["ms_per_solve"], e.get("matches_reference_build_sha256")) for e in o["extra_configs"]])
PY
# ncu: launch list of the two finest levels, then --set full captures of the main kernels
CMD="python bench.py --steps 1 --warmup 1 --levels 2 --no-e2e --no-cpu-baseline --no-extra"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2r_launches_levels2.csv $CMD > $O/r2r_ncu_launch.log 2>&1
wc -l $O/r2r_launches_levels2.csv
for k in "sweep_kernel" "phi_ksi_kernel" "warp_derivatives" "median5"; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 60 -c 2 -f -o $O/r2r_prof_$k $CMD > $O/r2r_ncu_$k.log 2>&1
  ls -la $O/r2r_prof_$k.ncu-rep
done
