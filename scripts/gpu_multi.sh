#!/bin/bash
# N-GPU visit: the two-process sharded parity tests (NCCL + fabric), then the N-GPU bench line(s).
# usage: gpu_multi.sh N [exchange modes, default "auto"]
set -u
N=${1:-2}; MODES=${2:-auto}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
if [ "${RUN_TESTS:-1}" = "1" ]; then
  timeout 900 python -m pytest tests/test_gpu_sharded.py -k "two_gpus" -q > gpurun_out/pytest_gpu_n2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_n2.log
  tail -6 gpurun_out/pytest_gpu_n2.log
fi
for X in $MODES; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --no-cpu --exchange $X \
      > gpurun_out/bench_c3_n${N}_$X.json 2> gpurun_out/bench_c3_n${N}_$X.err; echo "bench N=$N $X exit $?"
  tail -c 1800 gpurun_out/bench_c3_n${N}_$X.json; tail -c 600 gpurun_out/bench_c3_n${N}_$X.err
done
