#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
for n in 1 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/run_sharded.py --views 80 > gpurun_out/sharded_n$n.json 2> gpurun_out/sharded_n$n.err
cat gpurun_out/sharded_n$n.json; tail -3 gpurun_out/sharded_n$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
cat gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
