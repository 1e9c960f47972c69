#!/usr/bin/env bash
set -u
O=gpurun_out
N=8; SIZE=1024
timeout -s KILL 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
   bench.py --gpus $N --steps 2 --warmup 1 --size $SIZE --no-strong-ref --no-parity-check --no-e2e > $O/r2p_n${N}_$SIZE.json 2> $O/r2p_n${N}_$SIZE.err
python - <<PY
import json
d=json.load(open("$O/r2p_n${N}_$SIZE.json"))
print("N=$N size=$SIZE ms/step", d["ms_per_step"], "clocks", d["clocks"]); 
for r,p in enumerate(d["phase_ms_per_step_all_ranks"]): print(r,p)
PY
