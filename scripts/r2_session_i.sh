#!/usr/bin/env bash
# NCCL send/recv bandwidth of the ghost-plane exchange under different channel settings (2 GPUs, 1024^3, 3 levels)
set -u
O=gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
     bench.py --gpus 2 --steps 2 --warmup 1 --size 1024 --warp-levels 3 --no-strong-ref --no-parity-check --no-e2e \
     > $O/r2i_$tag.json 2> $O/r2i_$tag.err
  python - <<PY
import json
d=json.load(open("$O/r2i_$tag.json"))
p=d["phase_ms_per_step_rank0"]; c=d["config"]
n=c["halo_exchanges_per_step"]; b=c["halo_bytes_sent_per_step_rank0"]
print("$tag", "ms/step %.0f"%d["ms_per_step"], "halo ms %.1f"%p["halo_exchange"], "exchanges", n, "GB/s sent %.0f"%(b/1e9/(p["halo_exchange"]/1e3)), "solver %.0f"%p["solver"])
PY
}
run default FOO=1
run chan32 NCCL_NCHANNELS_PER_PEER=32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=64
run chan8 NCCL_NCHANNELS_PER_PEER=8
run memcpy NCCL_P2P_USE_CUDA_MEMCPY=1
