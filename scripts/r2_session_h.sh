#!/usr/bin/env bash
set -u
O=gpurun_out
N=${1:-2}; SIZE=${2:-512}; STEPS=${3:-1}; EXTRA="${4:-}"
timeout -s KILL 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps $STEPS --warmup 1 --size $SIZE $EXTRA > $O/r2h_bench_n${N}_${SIZE}.json 2> $O/r2h_bench_n${N}_${SIZE}.err
echo "rc=$?"; tail -5 $O/r2h_bench_n${N}_${SIZE}.err
python - <<PY
import json
d=json.load(open("$O/r2h_bench_n${N}_${SIZE}.json"))
for k in ("value","ms_per_step","phase_ms_per_step_rank0","halo_exchange_fraction_of_step","parity_check","strong_scaling","endpoint_error","e2e","tune_seconds_untimed"):
    print(k, d.get(k))
print(d["roofline"]["frac"], d["config"]["sharded_levels_per_step"], d["config"]["replicated_levels_per_step"], d["config"]["device_bytes_high_water_max_over_ranks"]/1e9)
PY
