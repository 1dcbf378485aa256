#!/bin/bash
# Round-2 second GPU pass (2 GPUs): kernel tests, N=1 trace + box-coupling check, comm-only trainer, fused paths, BERT gossip arms.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export KUNGFU_DEVICE_TIMEOUT_MS=8000
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
S=gpurun_out/r2b_summary.txt; : > $S
( timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q ) > gpurun_out/r2b_kernels.log 2>&1
echo "kernel tests rc=$? $(tail -1 gpurun_out/r2b_kernels.log)" | tee -a $S
( timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-e2e --trace gpurun_out/trace_ours1 ) > gpurun_out/r2b_ours1.log 2>&1
echo "ours1 rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2b_ours1.log | head -1)" | tee -a $S
# two INDEPENDENT single-GPU jobs at the same time: does merely running both GPUs cost per-GPU speed?
( CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --gpus 1 --steps 60 --warmup 5 --no-e2e > gpurun_out/r2b_indep0.log 2>&1 ) &
( CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --gpus 1 --steps 60 --warmup 5 --no-e2e > gpurun_out/r2b_indep1.log 2>&1 ) &
wait
echo "indep0 $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2b_indep0.log | head -1) indep1 $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2b_indep1.log | head -1)" | tee -a $S
( CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-e2e > gpurun_out/r2b_gpu1alone.log 2>&1 )
echo "gpu1 alone $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2b_gpu1alone.log | head -1)" | tee -a $S
( timeout 400 $TR --nproc-per-node 2 --master-port 29701 benchmarks/fake_trainer.py --model resnet50 --out gpurun_out/fake_trainer_resnet50_2gpu.json ) > gpurun_out/r2b_fake.log 2>&1
echo "fake trainer rc=$? $(grep -c RESULT gpurun_out/r2b_fake.log) results" | tee -a $S
( timeout 400 $TR --nproc-per-node 2 --master-port 29702 benchmarks/fused_paths.py --out gpurun_out/fused_paths_2gpu.json ) > gpurun_out/r2b_fused.log 2>&1
echo "fused paths rc=$?" | tee -a $S
for impl in ours nccl local; do
  ( timeout 500 $TR --nproc-per-node 2 --master-port $((29710 + RANDOM % 200)) benchmarks/bert_pairavg.py --impl $impl --steps 20 --out gpurun_out/bert_${impl}_2gpu.json ) > gpurun_out/r2b_bert_$impl.log 2>&1
  echo "bert $impl rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2b_bert_$impl.log | head -1) $(grep -o '"cuda_graph": [a-z]*' gpurun_out/r2b_bert_$impl.log | head -1) $(grep -o '"torn_reads_max_over_ranks": [0-9]*' gpurun_out/r2b_bert_$impl.log | head -1)" | tee -a $S
done
( timeout 500 $TR --nproc-per-node 2 --master-port 29731 benchmarks/bert_pairavg.py --impl ours --eager-ops --no-graph --steps 10 ) > gpurun_out/r2b_bert_eager.log 2>&1
echo "bert ours eager-ops no-graph rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2b_bert_eager.log | head -1)" | tee -a $S
cat $S
