#!/bin/bash
# Round-2 BERT step work: new transformer kernels (linear bias folded into residual+LayerNorm, QKV gradient pack, rational-erf GELU,
# parallel column-sum finalize) -> numerics tests, BERT-large step time + kernel breakdown on one GPU.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
S=gpurun_out/r2g_summary.txt; : > $S
( timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "layernorm or gelu or qkv or bert or selftest" ) > gpurun_out/r2g_tests.log 2>&1
echo "kernel tests rc=$? $(tail -1 gpurun_out/r2g_tests.log)" | tee -a $S
( timeout 300 python benchmarks/bert_pairavg.py --impl local --steps 30 --out gpurun_out/bert_local_1gpu.json --breakdown gpurun_out/bert_local_1gpu_kernels.md ) > gpurun_out/r2g_bert.log 2>&1
echo "bert local N=1 rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2g_bert.log | head -1)" | tee -a $S
( timeout 300 python benchmarks/bert_pairavg.py --impl local --steps 30 --eager-ops --out gpurun_out/bert_local_1gpu_eagerops.json ) > gpurun_out/r2g_bert_eager.log 2>&1
echo "bert local N=1 torch ops rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2g_bert_eager.log | head -1)" | tee -a $S
cat $S
