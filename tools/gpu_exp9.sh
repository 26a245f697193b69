#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
for n in 1 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/run_sharded.py --views 80 > gpurun_out/sharded_n$n.json 2> gpurun_out/sharded_n$n.err
cat gpurun_out/sharded_n$n.json; tail -2 gpurun_out/sharded_n$n.err
done
