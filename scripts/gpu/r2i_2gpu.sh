#!/bin/bash
# Round-2 closing check on the final tree (2 GPUs, bounded): pair / sma with the gradient hand-over against their simulations, the whole
# single-GPU kernel test file, smoke, the 1-GPU bench line.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export KUNGFU_DEVICE_TIMEOUT_MS=10000
S=gpurun_out/r2i_summary.txt; : > $S
( timeout 110 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29611 tests/mp_gpu_worker.py extra ) > gpurun_out/r2i_mp2.log 2>&1
echo "mp2 extra rc=$? $(grep -o '{"mp_gpu_worker.*}' gpurun_out/r2i_mp2.log | tail -1) $(grep -c 'done, fails = 0' gpurun_out/r2i_mp2.log) checks clean" | tee -a $S
( CUDA_VISIBLE_DEVICES=0 timeout 120 python -m pytest tests/test_gpu_kernels.py -x -q ) > gpurun_out/r2i_kernels.log 2>&1
echo "kernel tests rc=$? $(tail -1 gpurun_out/r2i_kernels.log)" | tee -a $S
( CUDA_VISIBLE_DEVICES=0 timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/r2i_smoke.log 2>&1
echo "smoke rc=$? $(tail -1 gpurun_out/r2i_smoke.log | cut -c1-200)" | tee -a $S
( CUDA_VISIBLE_DEVICES=0 timeout 100 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2i_bench1.log 2>&1
echo "bench1 rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/r2i_bench1.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2i_bench1.log | head -1) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2i_bench1.log)" | tee -a $S
cat $S
