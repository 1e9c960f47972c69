#!/usr/bin/env bash
# 8 GPUs: configs[3] (1024^3, full line incl. parity check + e2e) and configs[4] (2048^3, 60 levels)
set -u
O=gpurun_out
timeout -s KILL 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
   bench.py --gpus 8 --steps 2 --warmup 1 --size 1024 --no-strong-ref > $O/r2q_n8_1024.json 2> $O/r2q_n8_1024.err
echo "1024 rc=$?"
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 \
   bench.py --gpus 8 --steps 1 --warmup 1 --size 2048 --warp-levels 60 --no-strong-ref --no-parity-check --no-e2e \
   > $O/r2q_n8_2048.json 2> $O/r2q_n8_2048.err
echo "2048 rc=$?"; tail -3 $O/r2q_n8_2048.err | cut -c1-300
python - <<PY
import json
for f in ("r2q_n8_1024","r2q_n8_2048"):
    try:
        d=json.load(open("$O/%s.json"%f))
    except Exception as e:
        print(f, "no json", e); continue
    print(f, "ms/step", d["ms_per_step"], "value", d["value"], "clocks", d["clocks"]["sm_mhz"], "halo frac", d["halo_exchange_fraction_of_step"])
    print(" parity", d.get("parity_check"), "epe", d.get("endpoint_error"))
    c=d["config"]; print(" levels", c["pyramid_levels"], "sharded", c["sharded_levels_per_step"], "repl", c["replicated_levels_per_step"], "gathers", c["frame_gathers_per_step"], "mem GB", c["device_bytes_high_water_max_over_ranks"]/1e9, "tune s", d["tune_seconds_untimed"])
    for r,p in enumerate(d["phase_ms_per_step_all_ranks"]): print("  ",r,p)
PY
