#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/sweep_chart.py --out gpurun_out/sweep_chart.json > gpurun_out/sweep_chart.log 2>&1
tail -40 gpurun_out/sweep_chart.log
