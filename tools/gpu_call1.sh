#!/bin/bash
# GPU call: full GPU suite in one process (two file orders), layout probes, headline bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/pytest1.txt; cat gpurun_out/pytest1.txt | tail -5
(timeout 900 python -m pytest tests/test_gpu_extra.py tests/test_gpu_tc.py tests/test_gpu_mapper.py tests/test_gpu_parity.py tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -25) > gpurun_out/pytest2.txt; cat gpurun_out/pytest2.txt | tail -5
timeout 120 python tools/probe_bf16_kmajor.py > gpurun_out/probe_kmajor.txt 2>&1; tail -30 gpurun_out/probe_kmajor.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
