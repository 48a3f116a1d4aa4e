#!/bin/bash
# Round-2 end-of-round evidence on ONE GPU (final build): the GPU test suite twice in one process each (collection order,
# then shuffled), smoke, headline bench + reference arms, secondary workloads, the 1 048 576-ray variant, ncu launch lists
# and a full-set capture of the hot kernels.  Outputs under gpurun_out/final/.
O=gpurun_out/final; mkdir -p $O
timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee $O/pytest_gpu_run1.txt
DFP_PARITY_LOG=$O/parity_margins.json timeout 300 python -m pytest tests -q -m gpu -x --shuffle 20261017 2>&1 | tail -3 | tee $O/pytest_gpu_run2_shuffled.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.txt
timeout 300 python bench.py 2>$O/bench.err | tail -1 > $O/bench.json; head -c 250 $O/bench.json; echo
timeout 300 python bench.py --impl reference 2>/dev/null | tail -1 > $O/bench_reference_arm.json; head -c 200 $O/bench_reference_arm.json; echo
timeout 200 python bench.py --impl reference --device cuda:0 --ref-rays 8192 --steps 5 --warmup 2 2>/dev/null | tail -1 > $O/bench_reference_cuda.json
timeout 300 python bench_extra.py 2>/dev/null | tail -6 > $O/secondary_workloads.jsonl; cut -c1-220 $O/secondary_workloads.jsonl
timeout 300 python bench.py --rays 1048576 --steps 3 --warmup 3 --no-small --no-cpu-baseline --no-gpu-baseline 2>$O/bench_1M.err | tail -1 > $O/bench_1M_rays.json; head -c 250 $O/bench_1M_rays.json; echo
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-small"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_65536rays.csv $CMD > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_decoder|k_attn" -s 36 -c 12 -o $O/top $CMD > $O/ncu_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_1000rays.csv python bench.py --rays 1000 --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-small > /dev/null 2>&1
ls -la $O
