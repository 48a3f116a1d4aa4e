#!/bin/bash
# One 8-GPU box, final build of the round: weak scaling N = 1, 2, 4, 8 (the driver's contract), strong scaling at 8 GPUs,
# the sharded secondary workloads (render_img, Mesher query) and the 2-GPU equivalence test.  Outputs under gpurun_out/scale/.
mkdir -p gpurun_out/scale
O=gpurun_out/scale
nvidia-smi -L | wc -l
run() {  # n tag extra-args
  n=$1; tag=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 150 python bench.py --gpus 1 --steps 20 --warmup 5 --no-small --no-cpu-baseline --no-gpu-baseline "$@" > $O/$tag.json 2> $O/$tag.err
  else
    timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + n)) bench.py --gpus $n --steps 20 --warmup 5 --no-small "$@" > $O/$tag.json 2> $O/$tag.err
  fi
  python - $O/$tag.json $tag <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "n_gpus", d["n_gpus"], d["scaling"], "rays/gpu", d["config"]["rays_per_gpu_per_step"], "ms/step", round(d["ms_per_step"], 3), "rays/s", round(d["value"]), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
for n in 1 2 4 8; do run $n weak_n$n; done
run 8 weak_n8_nooverlap --no-overlap
run 8 strong65536_n8 --scaling strong --rays 65536
run 8 strong8192_n8 --scaling strong --rays 8192
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29850 bench_extra.py > $O/extra_n8.jsonl 2> $O/extra_n8.err; cut -c1-260 $O/extra_n8.jsonl | tail -3
timeout 150 python -m pytest tests/test_gpu_dist.py -q 2>&1 | tail -2 | tee $O/pytest_dist.txt
