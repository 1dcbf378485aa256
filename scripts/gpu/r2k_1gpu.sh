#!/bin/bash
# Occupancy tuning of the transformer kernels (packed parameters + launch bounds in the LayerNorm kernels, 4 CTAs/SM for GELU backward,
# 32-warp column-sum finalize): numerics first, then the BERT-large step.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
S=gpurun_out/r2k_summary.txt; : > $S
( timeout 150 python -m pytest tests/test_gpu_kernels.py -x -q -k "layernorm or gelu or qkv or bert" ) > gpurun_out/r2k_tests.log 2>&1
echo "kernel tests rc=$? $(tail -1 gpurun_out/r2k_tests.log)" | tee -a $S
( timeout 100 python benchmarks/bert_pairavg.py --impl local --steps 30 --out gpurun_out/bert_local_1gpu_tuned.json --breakdown gpurun_out/bert_local_1gpu_tuned_kernels.md ) > gpurun_out/r2k_bert.log 2>&1
echo "bert local N=1 rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2k_bert.log | head -1)" | tee -a $S
cat $S
