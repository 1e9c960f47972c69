#!/usr/bin/env bash
# last GPU call of round 2 (2.5 GPU-minutes left): PDL chain A/B first (quick, its JSON is safe even if the
# suite is cut), then the whole single-GPU suite on the final tree, then smoke()
set -u
O=gpurun_out
T0=$(date +%s)
timeout -s KILL 45 python scripts/pdl_ab.py 4 256 > $O/r2w_pdl_ab.json 2> $O/r2w_pdl_ab.err; echo "ab rc=$? t=$(( $(date +%s) - T0 ))"
cat $O/r2w_pdl_ab.json
timeout -s KILL 85 python -u -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2w_pytest_1gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))"
tail -15 $O/r2w_pytest_1gpu.log
timeout -s KILL 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; echo "smoke t=$(( $(date +%s) - T0 ))"
