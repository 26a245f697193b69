#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/variants.txt
for t in "" "7=1" "7=2" "7=3"; do
  echo "tuning=$t" >> gpurun_out/variants.txt
  timeout 120 python tools/run_steps.py --steps 12 --tuning "$t" >> gpurun_out/variants.txt 2>&1
done
cat gpurun_out/variants.txt
