#!/usr/bin/env bash
set -u
O=gpurun_out
SIZE=${1:-512}; N=${2:-2}
for ov in 1 0; do
FLOW3D_MGPU_OVERLAP=$ov timeout -s KILL 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
   bench.py --gpus $N --steps 2 --warmup 1 --size $SIZE --no-strong-ref --no-parity-check --no-e2e > $O/r2m_ov${ov}_n${N}_$SIZE.json 2> $O/r2m_ov${ov}_n${N}_$SIZE.err
python - <<PY
import json
d=json.load(open("$O/r2m_ov${ov}_n${N}_$SIZE.json"))
print("overlap=$ov N=$N size=$SIZE ms/step", d["ms_per_step"], "clocks", d["clocks"]); 
for r,p in enumerate(d["phase_ms_per_step_all_ranks"]): print(r,p)
PY
done
