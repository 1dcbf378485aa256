#!/bin/bash
# Round-2 third GPU pass (2 GPUs): tcgen05 GEMM first light, native engine, LL one-shot, ILP'd NVLS kernels, diagnostics, ncu.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export KUNGFU_DEVICE_TIMEOUT_MS=8000
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
S=gpurun_out/r2c_summary.txt; : > $S
( timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -k "tcgen05 or linear_act" ) > gpurun_out/r2c_gemm_test.log 2>&1
echo "gemm test rc=$? $(tail -1 gpurun_out/r2c_gemm_test.log)" | tee -a $S
( timeout 300 python benchmarks/gemm_bias_gelu.py --out gpurun_out/gemm_bias_gelu.json ) > gpurun_out/r2c_gemm_bench.log 2>&1
echo "gemm bench rc=$?" | tee -a $S
( timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "not tcgen05 and not linear_act" ) > gpurun_out/r2c_kernels.log 2>&1
echo "kernel tests rc=$? $(tail -1 gpurun_out/r2c_kernels.log)" | tee -a $S
( timeout 700 $TR --nproc-per-node 2 --master-port 29611 tests/mp_gpu_worker.py allreduce bcast ssgd extra ) > gpurun_out/r2c_mp2.log 2>&1
echo "mp2 rc=$? $(grep -o '"fails": [0-9]*' gpurun_out/r2c_mp2.log | tail -1)" | tee -a $S
run() { local name=$1 n=$2; shift 2; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  if [ "$n" = 1 ]; then ( env "${envs[@]}" timeout 400 python bench.py --gpus 1 "$@" ) > gpurun_out/r2c_$name.log 2>&1
  else ( env "${envs[@]}" timeout 400 $TR --nproc-per-node $n --master-port $((29620 + RANDOM % 300)) bench.py --gpus $n "$@" ) > gpurun_out/r2c_$name.log 2>&1; fi
  echo "$name rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/r2c_$name.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2c_$name.log | head -1)" | tee -a $S; }
run ours1 1 X=1 -- --steps 40 --warmup 5 --no-e2e
run ours2 2 X=1 -- --steps 40 --warmup 5 --no-e2e
run nocomm2 2 KUNGFU_BENCH_NO_COMM=1 KUNGFU_NATIVE_ENGINE=0 -- --steps 40 --warmup 5 --no-e2e
run ours2_b16 2 KUNGFU_COMM_BLOCKS=16 -- --steps 40 --warmup 5 --no-e2e
run nccl2 2 X=1 -- --impl nccl --steps 40 --warmup 5 --no-e2e
run eager2_py 2 KUNGFU_NATIVE_ENGINE=0 -- --steps 20 --warmup 5 --no-e2e --no-graph
run eager2_native 2 KUNGFU_NATIVE_ENGINE=1 -- --steps 20 --warmup 5 --no-e2e --no-graph
( timeout 400 $TR --nproc-per-node 2 --master-port 29701 benchmarks/fake_trainer.py --model resnet50 --out gpurun_out/fake_trainer_resnet50_2gpu.json ) > gpurun_out/r2c_fake.log 2>&1
echo "fake trainer rc=$?" | tee -a $S; grep RESULT gpurun_out/r2c_fake.log | tee -a $S
( timeout 400 $TR --nproc-per-node 2 --master-port 29702 benchmarks/fused_paths.py --out gpurun_out/fused_paths_2gpu.json ) > gpurun_out/r2c_fused.log 2>&1
echo "fused paths rc=$?" | tee -a $S
( timeout 400 $TR --nproc-per-node 2 --master-port 29703 benchmarks/allreduce_bw.py --nccl --max-mb 256 --out gpurun_out/allreduce_bw_2gpu.json ) > gpurun_out/r2c_arbw.log 2>&1
echo "allreduce bw rc=$?" | tee -a $S
for pb in 8 24; do
  ( KUNGFU_PULL_BLOCKS=$pb timeout 400 $TR --nproc-per-node 2 --master-port $((29710 + RANDOM % 200)) benchmarks/bert_pairavg.py --impl ours --steps 20 --out gpurun_out/bert_ours_pb${pb}_2gpu.json ) > gpurun_out/r2c_bert_ours_pb$pb.log 2>&1
  echo "bert ours pull_blocks=$pb rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2c_bert_ours_pb$pb.log | head -1)" | tee -a $S
done
( timeout 240 $TR --nproc-per-node 2 --master-port 29733 benchmarks/bert_pairavg.py --impl nccl --no-graph --steps 10 --out gpurun_out/bert_nccl_2gpu.json ) > gpurun_out/r2c_bert_nccl.log 2>&1
echo "bert nccl eager rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2c_bert_nccl.log | head -1)" | tee -a $S
bash scripts/gpu/ncu_collectives.sh 2 2>&1 | tee -a $S
cat $S
