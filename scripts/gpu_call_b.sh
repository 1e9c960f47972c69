#!/bin/bash
O=gpurun_out; T=${1:-r1f}
( time timeout 300 python -m pytest tests -m gpu -x -q ) > $O/${T}_pytest.log 2>&1; tail -4 $O/${T}_pytest.log
FLOW3D_GRID_MODEL=0 timeout 200 python scripts/level_table.py --reps 10 > $O/${T}_levels_model0.txt 2>&1
FLOW3D_GRID_MODEL=1 timeout 200 python scripts/level_table.py --reps 10 > $O/${T}_levels_model1.txt 2>&1
tail -1 $O/${T}_levels_model0.txt; tail -1 $O/${T}_levels_model1.txt
timeout 400 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; cut -c1-300 $O/${T}_bench.json
