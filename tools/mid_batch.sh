for r in 2368 5000 16384; do
timeout 120 python bench.py --rays $r --steps 30 --warmup 8 --no-small --no-cpu-baseline --no-gpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d['roofline']['kernels_ms_per_step']
print('rays', $r, 'ms/step', round(d['ms_per_step'], 4), {a: round(b*1000) for a, b in k.items() if 'decoder' in a})"
done
