#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"topk|backproject|prob_norm|pack" -s 9 -c 9 -o gpurun_out/prof_small -f python tools/run_steps.py --steps 3 > gpurun_out/ncu_small.log 2>&1
tail -2 gpurun_out/ncu_small.log
