#!/bin/bash
O=gpurun_out; T=${1:-r1o}
export FLOW3D_AUTOTUNE=0
for cfg in "0 0" "1 0" "0 1" "1 1" "0 0" "1 1"; do
set -- $cfg
echo "ROT=$1 SPEC=$2: $(FLOW3D_SWEEP_ROT=$1 FLOW3D_SWEEP_SPEC=$2 timeout 200 python scripts/level_table.py --reps 10 --every 2 | tail -1)"
FLOW3D_SWEEP_ROT=$1 FLOW3D_SWEEP_SPEC=$2 timeout 100 python scripts/run_stage.py sweep --size 512 --reps 40
done
