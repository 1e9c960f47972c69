#!/usr/bin/env bash
# the whole GPU test suite on a 2-GPU box (the multi-GPU tests run instead of being skipped)
set -u
O=gpurun_out
nvidia-smi -L > $O/r2t_gpus.txt
timeout -s KILL 900 python -m pytest tests -m gpu -v 2>&1 | grep -E "mgpu|cpp_api|dist_gpu|passed|failed|skipped|FAILED|ERROR" | tail -40 > $O/r2t_pytest_2gpu.log
tail -30 $O/r2t_pytest_2gpu.log
python scripts/run_stage.py warp --reps 5
