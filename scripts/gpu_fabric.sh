#!/bin/bash
# Sharded path with the peer-memory fabric on an N-GPU box: sharded parity tests (virtual shards on one GPU,
# NCCL + fabric between two processes, the 3.1 Gb bitwise test), then the N-GPU bench line with both exchange modes.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_sharded.py -k "sharded or two_gpus or fabric or virtual" -q --durations=10 > gpurun_out/pytest_gpu_shard.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_shard.log
tail -25 gpurun_out/pytest_gpu_shard.log
for X in fabric phases; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --exchange $X \
    > gpurun_out/bench_c3_n${N}_$X.json 2> gpurun_out/bench_c3_n${N}_$X.err; echo "bench N=$N $X exit $?"
tail -c 2500 gpurun_out/bench_c3_n${N}_$X.json; tail -c 800 gpurun_out/bench_c3_n${N}_$X.err
done
timeout 300 python bench.py --no-cpu > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3_n1.err; echo "bench N=1 exit $?"
tail -c 1800 gpurun_out/bench_c3_n1.json
