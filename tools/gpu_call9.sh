#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_full.txt 2>&1; tail -6 gpurun_out/pytest_full.txt
grep -n "^E  .*assert\|^E  .*Error" gpurun_out/pytest_full.txt | cut -c1-220 | head -20
