#!/usr/bin/env bash
set -u
O=gpurun_out
timeout -s KILL 600 python -m pytest tests/test_slab_gpu.py tests/test_mgpu_gpu.py -v -x 2>&1 | tail -25 > $O/r2j_pytest.log
cat $O/r2j_pytest.log | tail -22
bash scripts/r2_session_h.sh 2 512 2 "--no-strong-ref"
