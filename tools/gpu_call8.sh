#!/bin/bash
mkdir -p gpurun_out
bash tools/small_batch.sh 2>&1 | tee gpurun_out/small_batch.txt
