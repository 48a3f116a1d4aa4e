#!/bin/bash
# Run on the GPU box: tensor-core + SIMT parity suites, then a short headline bench with the per-kernel breakdown.
# usage: tools/gpu_check.sh [pytest-targets...]
targets=${@:-tests/test_gpu_parity.py tests/test_gpu_tc.py}
timeout 600 python -m pytest $targets -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read())
print("ms/step", round(d["ms_per_step"], 3), "rays/s", round(d["value"]), "e2e", round(d["e2e"]["value"]))
print(d["roofline"]["kernels_ms_per_step"])
print(d.get("mapping_1000_rays"))'
