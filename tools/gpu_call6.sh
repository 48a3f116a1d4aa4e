#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_attn2.py > gpurun_out/diag_attn2.txt 2>&1; tail -40 gpurun_out/diag_attn2.txt
