#!/bin/bash
# Round-2 fourth GPU pass (2 GPUs): re-validation after the native-engine / polling fixes, ncu captures, elastic with standby workers.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export KUNGFU_DEVICE_TIMEOUT_MS=8000
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
S=gpurun_out/r2d_summary.txt; : > $S
( timeout 600 python -m pytest tests/test_gpu_kernels.py -q ) > gpurun_out/r2d_kernels.log 2>&1
echo "kernel tests rc=$? $(tail -1 gpurun_out/r2d_kernels.log)" | tee -a $S
( timeout 700 $TR --nproc-per-node 2 --master-port 29611 tests/mp_gpu_worker.py allreduce bcast ssgd extra ) > gpurun_out/r2d_mp2.log 2>&1
echo "mp2 rc=$? $(grep -o '"fails": [0-9]*' gpurun_out/r2d_mp2.log | tail -1)" | tee -a $S
run() { local name=$1 n=$2; shift 2; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  if [ "$n" = 1 ]; then ( env "${envs[@]}" timeout 400 python bench.py --gpus 1 "$@" ) > gpurun_out/r2d_$name.log 2>&1
  else ( env "${envs[@]}" timeout 400 $TR --nproc-per-node $n --master-port $((29620 + RANDOM % 300)) bench.py --gpus $n "$@" ) > gpurun_out/r2d_$name.log 2>&1; fi
  echo "$name rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/r2d_$name.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2d_$name.log | head -1)" | tee -a $S; }
run ours1 1 X=1 -- --steps 40 --warmup 5
run ours2 2 X=1 -- --steps 40 --warmup 5 --no-e2e
run ours2_b16 2 KUNGFU_COMM_BLOCKS=16 -- --steps 40 --warmup 5 --no-e2e
run nccl2 2 X=1 -- --impl nccl --steps 40 --warmup 5 --no-e2e
run sma2 2 X=1 -- --optimizer sma --steps 40 --warmup 5 --no-e2e
run eager2_py 2 KUNGFU_NATIVE_ENGINE=0 -- --steps 20 --warmup 5 --no-e2e --no-graph
run eager2_native 2 KUNGFU_NATIVE_ENGINE=1 -- --steps 20 --warmup 5 --no-e2e --no-graph
( timeout 400 $TR --nproc-per-node 2 --master-port 29701 benchmarks/fake_trainer.py --model resnet50 --skip-simple-test --out gpurun_out/fake_trainer_resnet50_2gpu.json ) > gpurun_out/r2d_fake.log 2>&1
echo "fake trainer rc=$?" | tee -a $S; grep RESULT gpurun_out/r2d_fake.log | tee -a $S
( timeout 400 $TR --nproc-per-node 2 --master-port 29703 benchmarks/allreduce_bw.py --nccl --max-mb 16 --out gpurun_out/allreduce_bw_2gpu.json ) > gpurun_out/r2d_arbw.log 2>&1
echo "allreduce bw rc=$?" | tee -a $S; grep '"bytes"' gpurun_out/r2d_arbw.log | cut -c1-400 | tee -a $S
( timeout 400 $TR --nproc-per-node 2 --master-port 29702 benchmarks/fused_paths.py --models resnet50 --out gpurun_out/fused_paths_2gpu.json ) > gpurun_out/r2d_fused.log 2>&1
echo "fused paths rc=$?" | tee -a $S
# elastic 1 -> 2 -> 1 with a warm standby worker (resize latency)
( timeout 300 kungfu_b200/bin/kungfu-run -q -w -np 1 -H 127.0.0.1:2 -standby 1 -builtin-config-port 9100 python benchmarks/adaptation/elastic_resize.py --schedule 1:10,2:10,1:10 --model resnet50 ) > gpurun_out/r2d_elastic_standby.log 2>&1
echo "elastic standby rc=$? $(grep -o '"latency_s": [0-9.]*' gpurun_out/r2d_elastic_standby.log | tr '\n' ' ')" | tee -a $S
( timeout 300 kungfu_b200/bin/kungfu-run -q -w -np 1 -H 127.0.0.1:2 -builtin-config-port 9101 python benchmarks/adaptation/elastic_resize.py --schedule 1:10,2:10,1:10 --model resnet50 ) > gpurun_out/r2d_elastic_cold.log 2>&1
echo "elastic cold rc=$? $(grep -o '"latency_s": [0-9.]*' gpurun_out/r2d_elastic_cold.log | tr '\n' ' ')" | tee -a $S
# ncu: hand-written GEMM (1 GPU), collectives on a one-rank mesh (full set), and a light metric set on rank 0 of a live 2-rank job
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -c 2 -f -o gpurun_out/ncu_gemm_tc5 python benchmarks/ncu_gemm.py ) > gpurun_out/r2d_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?" | tee -a $S
( WORLD_SIZE=1 RANK=0 LOCAL_RANK=0 KF_NCU_ROUNDS=1 timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:ssgd_fused|allreduce_twoshot|pull_tma|pair_avg|sma_blend|allreduce_ll' -c 8 -f -o gpurun_out/ncu_collectives_1gpu python benchmarks/ncu_collectives.py ) > gpurun_out/r2d_ncu_coll1.log 2>&1
echo "ncu collectives 1 rank rc=$?" | tee -a $S
export MASTER_ADDR=127.0.0.1 MASTER_PORT=29877 WORLD_SIZE=2 KUNGFU_DEVICE_TIMEOUT_MS=60000 KF_NCU_ROUNDS=2
( RANK=1 LOCAL_RANK=1 timeout 300 python benchmarks/ncu_collectives.py > gpurun_out/r2d_ncu_coll2.rank1.log 2>&1 ) &
( RANK=0 LOCAL_RANK=0 timeout 300 ncu --metrics gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --cache-control none -k 'regex:ssgd_fused|allreduce_nvls|pull_tma|pair_avg' -c 10 --csv --log-file gpurun_out/ncu_collectives_2gpu_light.csv python benchmarks/ncu_collectives.py ) > gpurun_out/r2d_ncu_coll2.rank0.log 2>&1
echo "ncu collectives 2 ranks (light) rc=$?" | tee -a $S
wait
cat $S
