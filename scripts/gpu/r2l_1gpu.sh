#!/bin/bash
# Last GPU call of round 2 (one GPU): the whole kernel test file on the final kernels (LayerNorm backward restored, 4 loads in flight in the
# hand-over copy), the BERT-large step, the flagship bench line.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
S=gpurun_out/r2l_summary.txt; : > $S
( timeout 110 python -m pytest tests/test_gpu_kernels.py -x -q ) > gpurun_out/r2l_tests.log 2>&1
echo "kernel tests rc=$? $(tail -1 gpurun_out/r2l_tests.log)" | tee -a $S
( timeout 60 python benchmarks/bert_pairavg.py --impl local --steps 30 --out gpurun_out/bert_local_1gpu_final.json --breakdown gpurun_out/bert_local_1gpu_final_kernels.md ) > gpurun_out/r2l_bert.log 2>&1
echo "bert local N=1 rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2l_bert.log | head -1)" | tee -a $S
( timeout 70 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2l_bench1.log 2>&1
echo "bench1 rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/r2l_bench1.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2l_bench1.log | head -1) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2l_bench1.log)" | tee -a $S
cat $S
