#!/bin/bash
# Round-2 8-GPU pass: N=8 correctness of every collective / fused path, both bench arms for S-SGD and SMA, comm-only trainer,
# fused-path rooflines, all-reduce sweep, BERT-large gossip, elastic 4 -> 8 -> 4 with standby workers, step trace, NVLink counters.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export KUNGFU_DEVICE_TIMEOUT_MS=10000
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=8
S=gpurun_out/r2e_summary.txt; : > $S
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2e_gpus.csv 2>&1
( timeout 500 $TR --nproc-per-node $N --master-port 29611 tests/mp_gpu_worker.py allreduce bcast ssgd extra ) > gpurun_out/mp8_round2.log 2>&1
echo "mp8 rc=$? $(grep -o '{"mp_gpu_worker.*}' gpurun_out/mp8_round2.log | tail -1)" | tee -a $S
run() { local name=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  ( env "${envs[@]}" timeout 300 $TR --nproc-per-node $N --master-port $((29620 + RANDOM % 300)) bench.py --gpus $N "$@" ) > gpurun_out/r2e_$name.log 2>&1
  echo "$name rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/r2e_$name.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2e_$name.log | head -1)" | tee -a $S; }
run ours8 X=1 -- --steps 40 --warmup 5 --trace gpurun_out/trace_ours8
run nccl8 X=1 -- --impl nccl --steps 40 --warmup 5 --no-e2e
run sma8 X=1 -- --optimizer sma --steps 40 --warmup 5 --no-e2e
run sma8_nccl X=1 -- --optimizer sma --impl nccl --steps 40 --warmup 5 --no-e2e
( timeout 300 $TR --nproc-per-node $N --master-port 29701 benchmarks/fake_trainer.py --model resnet50 --out gpurun_out/fake_trainer_resnet50_8gpu.json ) > gpurun_out/r2e_fake.log 2>&1
echo "fake trainer rc=$?" | tee -a $S; grep RESULT gpurun_out/r2e_fake.log | tee -a $S
( timeout 300 $TR --nproc-per-node $N --master-port 29702 benchmarks/fused_paths.py --out gpurun_out/fused_paths_8gpu.json ) > gpurun_out/r2e_fused.log 2>&1
echo "fused paths rc=$?" | tee -a $S
( timeout 400 $TR --nproc-per-node $N --master-port 29703 benchmarks/allreduce_bw.py --nccl --max-mb 256 --blocks 32,64 --out gpurun_out/allreduce_bw_8gpu.json ) > gpurun_out/r2e_arbw.log 2>&1
echo "allreduce bw rc=$?" | tee -a $S
for impl in ours local; do
  ( timeout 300 $TR --nproc-per-node $N --master-port $((29710 + RANDOM % 200)) benchmarks/bert_pairavg.py --impl $impl --steps 30 --out gpurun_out/bert_${impl}_8gpu.json --breakdown gpurun_out/bert_${impl}_8gpu_kernels.md ) > gpurun_out/r2e_bert_$impl.log 2>&1
  echo "bert $impl rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2e_bert_$impl.log | head -1) $(grep -o '"value": [0-9.]*' gpurun_out/r2e_bert_$impl.log | head -1) $(grep -o '"torn_reads_max_over_ranks": [0-9]*' gpurun_out/r2e_bert_$impl.log | head -1)" | tee -a $S
done
# BASELINE.json config #5: elastic 4 -> 8 -> 4 with the gradient-noise-scale monitor, warm standby workers
( timeout 300 kungfu_b200/bin/kungfu-run -q -w -np 4 -H 127.0.0.1:8 -standby 4 -builtin-config-port 9100 python benchmarks/adaptation/elastic_resize.py --schedule 4:15,8:15,4:15 --model resnet50 --batch-size 256 ) > gpurun_out/r2e_elastic.log 2>&1
echo "elastic 4-8-4 rc=$? $(grep -o '{"resize_to[^}]*}' gpurun_out/r2e_elastic.log | tr '\n' ' ') $(grep -o '{"phase_np[^}]*}' gpurun_out/r2e_elastic.log | tr '\n' ' ')" | tee -a $S
# NVLink / DRAM counters of the body kernels on rank 0 of the live 8-rank job (light metric set)
export MASTER_ADDR=127.0.0.1 MASTER_PORT=29877 WORLD_SIZE=$N KUNGFU_DEVICE_TIMEOUT_MS=60000 KF_NCU_ROUNDS=2
for r in $(seq 1 $((N - 1))); do ( RANK=$r LOCAL_RANK=$r timeout 200 python benchmarks/ncu_collectives.py > gpurun_out/r2e_ncu.rank$r.log 2>&1 ) & done
( RANK=0 LOCAL_RANK=0 timeout 200 ncu --metrics gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --cache-control none -k 'regex:ssgd_fused|allreduce_nvls|pull_tma|pair_avg' -c 10 --csv --log-file gpurun_out/ncu_collectives_8gpu_light.csv python benchmarks/ncu_collectives.py ) > gpurun_out/r2e_ncu.rank0.log 2>&1
echo "ncu light N=8 rc=$?" | tee -a $S
wait
cat $S
