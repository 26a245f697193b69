#!/bin/bash
# Experiment run 1 (round 1, session 3): packed-math sweep kernels + L2 prefetch, RED egress micro-benchmark.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 120 ./tools/microbench_red > gpurun_out/microbench_red.txt 2>&1; tail -40 gpurun_out/microbench_red.txt
for t in "" "6=1" "5=2" "5=1"; do
  echo "tuning=$t" >> gpurun_out/variants.txt
  timeout 300 python tools/run_steps.py --steps 12 --tuning "$t" >> gpurun_out/variants.txt 2>&1
done
cat gpurun_out/variants.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep -s 4 -c 2 -o gpurun_out/prof_sweep2 -f python tools/run_steps.py --steps 3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
