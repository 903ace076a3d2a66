#!/bin/bash
# Sharded path on the GPU box: parity tests (virtual shards + NCCL when 2 GPUs are visible) and the N-GPU bench line.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_shard.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_shard.log
tail -15 gpurun_out/pytest_gpu_shard.log
if [ "$N" -gt 1 ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu \
    > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err; echo "bench N=$N exit $?"
tail -c 2500 gpurun_out/bench_c3_n$N.json; tail -c 1500 gpurun_out/bench_c3_n$N.err
fi
