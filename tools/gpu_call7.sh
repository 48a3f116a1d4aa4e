#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_attn2.py > gpurun_out/diag_attn2.txt 2>&1; head -3 gpurun_out/diag_attn2.txt | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_full.txt 2>&1; tail -6 gpurun_out/pytest_full.txt
grep -n "^E  .*assert\|^E  .*Error" gpurun_out/pytest_full.txt | cut -c1-200 | head -20
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    print("ms/step", round(d["ms_per_step"], 3), "rays/s", round(d["value"]), "e2e", round(d["e2e"]["value"]))
    print(d["roofline"]["kernels_ms_per_step"])
    print(d.get("mapping_1000_rays")); print(d.get("tracker_200_rays"))
except Exception as e:
    print("bench parse failed", e)
PY
