#!/bin/bash
# GPU parity tests (incl. the dispatcher ops), smoke, both bench arms, eager-GPU comparison.
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? after ${SECONDS}s" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 300 python tools/eager_gpu.py > gpurun_out/eager_gpu.json 2> gpurun_out/eager_gpu.err; cat gpurun_out/eager_gpu.json; tail -2 gpurun_out/eager_gpu.err
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
echo "total ${SECONDS}s"
