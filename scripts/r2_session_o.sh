#!/usr/bin/env bash
set -u
O=gpurun_out
timeout -s KILL 600 python -m pytest tests/test_mgpu_gpu.py -q -x 2>&1 | tail -3
bash scripts/r2_session_m.sh 512 2
