#!/bin/bash
O=gpurun_out; T=${1:-r1g}
( time timeout 300 python -m pytest tests -m gpu -x -q ) > $O/${T}_pytest.log 2>&1; tail -4 $O/${T}_pytest.log
FLOW3D_AUTOTUNE=1 timeout 200 python scripts/level_table.py --reps 10 > $O/${T}_levels_tuned.txt 2>&1
tail -1 $O/${T}_levels_tuned.txt
FLOW3D_AUTOTUNE=0 FLOW3D_SWEEP_PF=0 timeout 200 python scripts/level_table.py --reps 10 --every 3 > $O/${T}_levels_pf0.txt 2>&1
FLOW3D_AUTOTUNE=0 FLOW3D_SWEEP_PF=2 timeout 200 python scripts/level_table.py --reps 10 --every 3 > $O/${T}_levels_pf2.txt 2>&1
FLOW3D_AUTOTUNE=0 FLOW3D_SWEEP_PF=4 timeout 200 python scripts/level_table.py --reps 10 --every 3 > $O/${T}_levels_pf4.txt 2>&1
tail -1 $O/${T}_levels_pf0.txt $O/${T}_levels_pf2.txt $O/${T}_levels_pf4.txt
timeout 400 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; cut -c1-300 $O/${T}_bench.json
