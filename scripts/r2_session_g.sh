#!/usr/bin/env bash
set -u
O=gpurun_out
nvidia-smi -L > $O/r2g_gpus.txt
timeout -s KILL 900 python -m pytest tests/test_mgpu_gpu.py -v -x 2>&1 | tail -40 > $O/r2g_mgpu_pytest.log
cat $O/r2g_mgpu_pytest.log
