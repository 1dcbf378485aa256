#!/bin/bash
# ncu --set full of the fused transformer kernels at the BERT-large shapes (one GPU, one capture).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 240 ncu --set full --clock-control none --import-source on -k regex:'add_ln_|bias_gelu_|qkv_pack|colsum_finalize' -c 12 -f -o gpurun_out/ncu_transformer python benchmarks/ncu_transformer.py ) > gpurun_out/r2j_ncu.log 2>&1
echo "ncu transformer rc=$?"; tail -3 gpurun_out/r2j_ncu.log
ls -la gpurun_out/ncu_transformer.ncu-rep
