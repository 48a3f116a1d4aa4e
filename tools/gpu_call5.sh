#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_attn.py > gpurun_out/diag_attn.txt 2>&1; tail -20 gpurun_out/diag_attn.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_full.txt 2>&1; tail -12 gpurun_out/pytest_full.txt
grep -n "^E  .*assert\|^E  .*Error" gpurun_out/pytest_full.txt | cut -c1-200 | head -30
