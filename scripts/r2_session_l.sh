#!/usr/bin/env bash
set -u
O=gpurun_out
FLOW3D_MGPU_LOG=1 timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 \
   bench.py --gpus 2 --steps 1 --warmup 1 --size 512 --no-strong-ref --no-parity-check --no-e2e > $O/r2l_log.json 2> $O/r2l_log.err
grep "mgpu r0" $O/r2l_log.err | tail -40 | cut -c18-260
python -c "
import json; d=json.load(open('$O/r2l_log.json')); print(d['ms_per_step']); [print(r,p) for r,p in enumerate(d['phase_ms_per_step_all_ranks'])]"
