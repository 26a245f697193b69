#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
rm -f gpurun_out/variants.txt
for t in "" "7=8" "7=16"; do
  echo "tuning=$t" >> gpurun_out/variants.txt
  timeout 120 python tools/run_steps.py --steps 12 --tuning "$t" >> gpurun_out/variants.txt 2>&1
done
cat gpurun_out/variants.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:v['ms'] for k,v in d['kernels'].items()})"; tail -3 gpurun_out/bench.err
