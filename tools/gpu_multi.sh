#!/bin/bash
# Multi-GPU round: 2-rank peer-memory tests, then the scene-parallel + view-sharded bench on N GPUs.
#   tools/gpurun_retry.sh --gpus N 900 'bash tools/gpu_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_p2p.py -m gpu -x -q > gpurun_out/pytest_p2p.log 2>&1; tail -15 gpurun_out/pytest_p2p.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 3000 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
