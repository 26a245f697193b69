#!/bin/bash
# Round-end measurement run: parity tests, smoke, both bench arms, ncu launch list of the bench
# command, full ncu capture of every kernel of the step.  The .ncu-rep is summarised ON the box
# (counter summary + per-kernel DRAM traffic) and then dropped without import-source so that
# gpurun_out/ stays under gpurun's 64 MiB copy-back limit.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
SECONDS=0
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? after ${SECONDS}s" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --sustained-steps 0 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"sweep|backproject|topk|pack|prob_norm" -s 18 -c 9 -o /tmp/prof_step -f python tools/run_steps.py --steps 4 > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/prof_step.ncu-rep > gpurun_out/ncu_step_summary.txt 2>&1
python tools/ncu_summary.py --traffic gpurun_out/traffic.json /tmp/prof_step.ncu-rep > gpurun_out/traffic.log 2>&1
ls -la /tmp/prof_step.ncu-rep; [ "$(stat -c %s /tmp/prof_step.ncu-rep)" -lt 40000000 ] && cp /tmp/prof_step.ncu-rep gpurun_out/
du -sh gpurun_out; echo "total ${SECONDS}s"
