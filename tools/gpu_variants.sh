#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/compare_variants.py --variants "${VARIANTS:-0,5,6}" > gpurun_out/variants.txt 2>&1
cat gpurun_out/variants.txt
