#!/bin/bash
# K1 timing experiments: skip one part at a time (results are wrong by construction; only k_decoder_dh_tc's time matters)
for d in ${DBG_LIST:-0 1 2 4 8 15}; do   # bits: 1 plane stores, 2 scatter, 4 MMAs, 8 sincos
DFP_DEBUG_SKIP=$d timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-small 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d['roofline']['kernels_ms_per_step']
print('skip=$d', 'K1', k['k_decoder_dh_tc'], 'step', round(d['ms_per_step'], 2))"
done
