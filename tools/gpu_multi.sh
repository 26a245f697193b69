#!/bin/bash
# Multi-GPU sanity: scene-parallel bench (weak scaling) and the view-sharded test-time forward
# (V=80) on N GPUs of one box.  N from $1 (default 2).  Outputs under gpurun_out/.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_smi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
cat gpurun_out/bench_n$N.json | cut -c1-400; tail -2 gpurun_out/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tools/run_sharded.py --views 80 --p2p > gpurun_out/sharded_v80_n$N.json 2> gpurun_out/sharded_v80_n$N.err
cat gpurun_out/sharded_v80_n$N.json; tail -2 gpurun_out/sharded_v80_n$N.err
