#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_bwd -s 2 -c 1 -o gpurun_out/prof_rows -f python tools/run_steps.py --steps 3 --tuning "${TUNING-5=5}" > gpurun_out/ncu_rows.log 2>&1
tail -3 gpurun_out/ncu_rows.log
