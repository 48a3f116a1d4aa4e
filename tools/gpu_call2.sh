#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_full.txt 2>&1; tail -8 gpurun_out/pytest_full.txt
grep -n "^E  " gpurun_out/pytest_full.txt | cut -c1-260 | head -40
for i in 0 4 1 5; do timeout 60 python tools/probe_bf16_kmajor.py $i 2>&1 | tail -6; done > gpurun_out/probe_kmajor.txt; cat gpurun_out/probe_kmajor.txt
