#!/bin/bash
# quick 2-GPU check: GEMM epilogue rework, hand-over copy with sum of squares, NVLink hardware counters
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export KUNGFU_DEVICE_TIMEOUT_MS=8000
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
S=gpurun_out/r2f_summary.txt; : > $S
( timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "tcgen05 or linear_act or noise or monitor or ssgd" ) > gpurun_out/r2f_kernels.log 2>&1
echo "kernel tests rc=$? $(tail -1 gpurun_out/r2f_kernels.log)" | tee -a $S
( timeout 300 python benchmarks/gemm_bias_gelu.py --out gpurun_out/gemm_bias_gelu.json ) > gpurun_out/r2f_gemm_bench.log 2>&1
echo "gemm bench rc=$?" | tee -a $S; grep '^{' gpurun_out/r2f_gemm_bench.log | cut -c1-330 | tee -a $S
( timeout 400 $TR --nproc-per-node 2 --master-port 29611 tests/mp_gpu_worker.py ssgd ) > gpurun_out/r2f_mp2.log 2>&1
echo "mp2 ssgd rc=$? $(grep -o '"fails": [0-9]*' gpurun_out/r2f_mp2.log | tail -1)" | tee -a $S
( timeout 300 $TR --nproc-per-node 2 --master-port 29641 bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e ) > gpurun_out/r2f_ours2.log 2>&1
echo "ours2 rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2f_ours2.log | head -1) $(grep -o '"buckets": [0-9]*' gpurun_out/r2f_ours2.log)" | tee -a $S
( timeout 300 $TR --nproc-per-node 2 --master-port 29702 benchmarks/fused_paths.py --models resnet50 --out gpurun_out/fused_paths_2gpu.json ) > gpurun_out/r2f_fused.log 2>&1
echo "fused paths rc=$?" | tee -a $S; grep -o '"path": "[^"]*", "us": [0-9.]*\|counted": [0-9]*' gpurun_out/r2f_fused.log | head -30 | tr '\n' ' ' | tee -a $S
nvidia-smi nvlink -gt d -i 0 2>&1 | head -8 | tee -a $S
( timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -c 1 -f -o gpurun_out/ncu_gemm_tc5 python benchmarks/ncu_gemm.py ) > gpurun_out/r2f_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?" | tee -a $S
( timeout 600 bash scripts/sanitize.sh cuda gpurun_out/compute_sanitizer ) > gpurun_out/r2f_sanitize.log 2>&1
echo "compute-sanitizer rc=$?" | tee -a $S; cat gpurun_out/r2f_sanitize.log | tail -5 | tee -a $S
cat $S
