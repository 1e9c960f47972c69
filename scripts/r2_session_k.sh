#!/usr/bin/env bash
set -u
O=gpurun_out
timeout -s KILL 600 python -m pytest tests/test_slab_gpu.py tests/test_mgpu_gpu.py -q -x 2>&1 | tail -5 > $O/r2k_pytest.log
cat $O/r2k_pytest.log | tail -5
for ov in 1 0; do
FLOW3D_MGPU_OVERLAP=$ov timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
   bench.py --gpus 2 --steps 2 --warmup 1 --size 512 --no-strong-ref --no-parity-check --no-e2e > $O/r2k_ov$ov.json 2> $O/r2k_ov$ov.err
python - <<PY
import json
d=json.load(open("$O/r2k_ov$ov.json"))
print("overlap=$ov ms/step", d["ms_per_step"]); 
for r,p in enumerate(d["phase_ms_per_step_all_ranks"]): print(r,p)
PY
done
FLOW3D_MGPU_LOG=1 FLOW3D_MGPU_OVERLAP=1 timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 \
   bench.py --gpus 2 --steps 1 --warmup 1 --size 512 --no-strong-ref --no-parity-check --no-e2e > $O/r2k_log.json 2> $O/r2k_log.err
grep "mgpu r0" $O/r2k_log.err | tail -42 | cut -c1-150
