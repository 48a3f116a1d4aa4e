#!/bin/bash
# Round-end evidence on one GPU: full GPU test suite, smoke, headline bench (+ reference arm), secondary workloads,
# ncu launch list and a full-set capture of the top kernels.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
timeout 600 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench.json; head -c 600 gpurun_out/bench.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference.json; head -c 400 gpurun_out/bench_reference.json; echo
timeout 600 python bench_extra.py 2>/dev/null | tail -6 > gpurun_out/secondary.jsonl; cat gpurun_out/secondary.jsonl | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-small > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_decoder|k_attn" -c 14 -o gpurun_out/top python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-small > /dev/null 2>&1
ls -la gpurun_out
