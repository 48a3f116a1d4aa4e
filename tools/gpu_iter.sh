#!/bin/bash
# One development iteration on the GPU box: parity suites (quick subset unless FULL=1), then the headline bench with the
# per-kernel breakdown and the small-batch lines.  Output also goes to gpurun_out/iter.txt.
mkdir -p gpurun_out
{
if [ "${FULL:-0}" = "1" ]; then targets="tests"; else targets="tests/test_gpu_parity.py tests/test_gpu_extra.py tests/test_gpu_tc.py"; fi
timeout 240 python -m pytest $targets -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ${BENCH_ARGS} 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read())
print("ms/step", round(d["ms_per_step"], 3), "rays/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"])
print(d["roofline"]["kernels_ms_per_step"])
if d.get("mapping_frozen_low_high"): print("frozen low+high", round(d["mapping_frozen_low_high"]["ms_per_step"], 3))
for k in ("mapping_1000_rays", "tracker_200_rays"):
    s = d.get(k)
    if s: print(k, round(s["ms_per_step"], 4), "graph", s.get("cuda_graph", {}).get("ms_per_step"), "graph e2e", s.get("cuda_graph", {}).get("e2e", {}).get("ms_per_step"))
PY
} 2>&1 | tee gpurun_out/iter.txt
