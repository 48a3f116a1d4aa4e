#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_writes.py > gpurun_out/diag_writes.txt 2>&1; tail -40 gpurun_out/diag_writes.txt
