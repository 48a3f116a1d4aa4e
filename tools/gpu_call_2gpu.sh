#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_dist.py -q > gpurun_out/pytest_dist.txt 2>&1; tail -5 gpurun_out/pytest_dist.txt
grep -n "^E  " gpurun_out/pytest_dist.txt | cut -c1-300 | head -10
for mode in "" "--no-overlap"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 10 --warmup 3 --no-small $mode > gpurun_out/bench_n2$mode.json 2> gpurun_out/bench_n2$mode.err
tail -2 gpurun_out/bench_n2$mode.err | cut -c1-300
python - "$mode" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/bench_n2%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("N=2", sys.argv[1], "ms/step", round(d["ms_per_step"], 3), "rays/s", round(d["value"]), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("bench parse failed", e)
PY
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-small --no-cpu-baseline --no-gpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=1 ms/step', round(d['ms_per_step'],3), 'rays/s', round(d['value']))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench_extra.py > gpurun_out/extra_n2.jsonl 2> gpurun_out/extra_n2.err; tail -3 gpurun_out/extra_n2.err | cut -c1-300; cut -c1-250 gpurun_out/extra_n2.jsonl
