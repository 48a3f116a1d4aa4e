#!/bin/bash
# per-kernel breakdown at the reference's own mapping batch sizes
for n in 1000 5000; do
timeout 200 python bench.py --rays $n --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-small 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read())
k = d["roofline"]["kernels_ms_per_step"]
print(d["config"]["rays_per_gpu_per_step"], "rays: ms/step", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "launches/step", d["gpu_launches"] / d["steps"], "kernel sum", round(sum(k.values()), 3))
print(k)'
done
