#!/usr/bin/env bash
# 4-GPU box: the sharded-solver tests with interior ranks (two neighbours), ranks as threads and as processes
set -u
O=gpurun_out
nvidia-smi -L > $O/r2v_gpus.txt
timeout -s KILL 600 python -m pytest tests/test_mgpu_gpu.py tests/test_cpp_api.py::test_cli_gpus_shards_over_two_devices -v 2>&1 | grep -E "PASSED|FAILED|SKIPPED|ERROR|passed|failed" > $O/r2v_pytest_4gpu.log
cat $O/r2v_pytest_4gpu.log
