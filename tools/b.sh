#!/bin/bash
# build here (cross-compile), then run one development iteration on the GPU box; refuses to go to the GPU if the build failed
cd /root/repo || exit 1
python -m attentive_dfprior_b200.build 2>&1 | grep -i " error\|error:" && { echo BUILD FAILED; exit 1; }
[ attentive_dfprior_b200/libdfprior_b200.so -nt attentive_dfprior_b200/csrc/dfp_api.cu ] || echo "warning: library older than dfp_api.cu"
gpurun --timeout 600 -- "${1:-bash tools/gpu_iter.sh}" 2>&1 | tail -${2:-8}
