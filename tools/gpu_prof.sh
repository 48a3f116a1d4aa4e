#!/bin/bash
# launch list + full-set capture of the hot kernels of one benchmark step
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-small"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > /dev/null 2>&1
# the 4th step (after 3 warm-ups and before the kernel-timing / e2e steps): skip the launches of the warm-up steps
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_decoder|k_attn" -s 36 -c 12 -o gpurun_out/top $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -5
# the reference's own batch size: exact per-kernel durations of a 1000-ray step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_1000rays.csv python bench.py --rays 1000 --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-small > /dev/null 2>&1
ls -la gpurun_out | tail -4
