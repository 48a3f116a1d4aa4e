#!/bin/bash
# Timing experiments (NOT the product build): rebuild the library with -DDFP_EXPERIMENTS on the GPU box, run the bench's
# kernel breakdown with parts of the kernels switched off (DFP_EXP bit mask, see DFP_EXP(...) in the kernels), then
# restore the product build.  Results are wrong by construction; only the durations are read.
# usage: tools/exp_run.sh "0 1 2 4 16 32 64"
mkdir -p gpurun_out
cp attentive_dfprior_b200/libdfprior_b200.so /tmp/lib_product.so
cd attentive_dfprior_b200/csrc
objs=""
for f in dfp_api dfp_dec_fwd dfp_dec_bwd_m0 dfp_dec_bwd_m1 dfp_dec_bwd_m2 dfp_attn dfp_attn_tc dfp_dec_fwd_tc dfp_dec_dh_tc dfp_dec_dw_tc dfp_mapper dfp_fusion; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DDFP_EXPERIMENTS -c $f.cu -o /tmp/exp_$f.o &
  objs="$objs /tmp/exp_$f.o"
done
wait
nvcc -shared -o ../libdfprior_b200.so $objs -lcudart
cd ../..
for m in $1; do
  DFP_EXP=$m timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-small 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d['roofline']['kernels_ms_per_step']
print('EXP $m: step', round(d['ms_per_step'], 3), 'fwd', k['k_decoder_fwd_tc'], 'K1', k['k_decoder_dh_tc'], 'K2', k['k_decoder_dw_tc'])"
done 2>&1 | tee gpurun_out/exp.txt
[ -n "$2" ] && python $2 2>&1 | tail -3 | tee -a gpurun_out/exp.txt
cp /tmp/lib_product.so attentive_dfprior_b200/libdfprior_b200.so
