#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_p2p.py -m gpu -x -q > gpurun_out/pytest_p2p.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_p2p.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 tools/run_sharded.py --views 80 --p2p > gpurun_out/sharded_p2p_n$N.json 2> gpurun_out/sharded_p2p_n$N.err
echo "rc=$?"; cat gpurun_out/sharded_p2p_n$N.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/sharded_p2p_n$N.err | tail -12
