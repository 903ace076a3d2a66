#!/bin/bash
# Sharded paths after a change to the distribution kernel: virtual shards + two processes on two GPUs + the 3.1 Gb four-shard
# bit-identity test, then the 2-GPU bench lines.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py "tests/test_gpu_configs.py::test_c3_full_size_sharded_is_bit_identical" -m gpu -q -x -o timeout=300 > gpurun_out/r02e_pytest_sharded.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02e_pytest_sharded.log
tail -4 gpurun_out/r02e_pytest_sharded.log
RUN_TESTS=0 bash scripts/gpu_multi.sh 2 "auto phases" 2>&1 | grep -v "^\*\|OMP_NUM" | cut -c1-200
