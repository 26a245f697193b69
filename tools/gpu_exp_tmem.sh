#!/bin/bash
# Round-2 first measurement: the TMEM-pending 4-CTA/SM backward.
# On the CPU box first:   MVSD_EXTRA_NVCC_FLAGS=-DMVSD_EXP_TMEM_PENDING python -m mvsdet_b200.build
# then:                   gpurun --timeout 400 -- 'bash tools/gpu_exp_tmem.sh'
# (rebuild without the flag afterwards: the shipped library does not contain the experiment)
mkdir -p gpurun_out
MVSD_TEST_EXTRA_VARIANTS=17 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or ragged" 2>&1 | tail -3
timeout 150 python tools/compare_variants.py --variants 0,17,0,17 | tee gpurun_out/exp_tmem.txt
