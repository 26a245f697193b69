#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
rm -f gpurun_out/variants.txt
for t in "" "4=2"; do
  echo "tuning=$t" >> gpurun_out/variants.txt
  timeout 120 python tools/run_steps.py --steps 12 --tuning "$t" >> gpurun_out/variants.txt 2>&1
done
cat gpurun_out/variants.txt
