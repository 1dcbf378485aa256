#!/bin/bash
# Round-2 last 2-GPU pass: gradient hand-over for the model-averaging engines (pair / sma), shrink-and-continue on the GPU plane, eager
# standby workers, P2P sweep with clocks, BERT-large gossip with the reworked transformer kernels.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export KUNGFU_DEVICE_TIMEOUT_MS=15000
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=2
S=gpurun_out/r2h_summary.txt; : > $S
( timeout 400 $TR --nproc-per-node $N --master-port 29611 tests/mp_gpu_worker.py allreduce ssgd extra ) > gpurun_out/r2h_mp2.log 2>&1
echo "mp2 rc=$? $(grep -o '{"mp_gpu_worker.*}' gpurun_out/r2h_mp2.log | tail -1)" | tee -a $S
( timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k "shrink_and_continue" ) > gpurun_out/r2h_shrink.log 2>&1
echo "gpu shrink rc=$? $(tail -1 gpurun_out/r2h_shrink.log) $(grep -o '"recover_s": [^]]*]' gpurun_out/r2h_shrink.log | head -1)" | tee -a $S
for impl in ours local; do
  ( timeout 300 $TR --nproc-per-node $N --master-port $((29710 + RANDOM % 200)) benchmarks/bert_pairavg.py --impl $impl --steps 30 --out gpurun_out/bert_${impl}_2gpu.json --breakdown gpurun_out/bert_${impl}_2gpu_kernels.md ) > gpurun_out/r2h_bert_$impl.log 2>&1
  echo "bert $impl rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2h_bert_$impl.log | head -1) $(grep -o '"value": [0-9.]*' gpurun_out/r2h_bert_$impl.log | head -1) $(grep -o '"torn_reads_max_over_ranks": [0-9]*' gpurun_out/r2h_bert_$impl.log | head -1)" | tee -a $S
done
run() { local name=$1; shift
  ( timeout 300 $TR --nproc-per-node $N --master-port $((29620 + RANDOM % 300)) bench.py --gpus $N "$@" ) > gpurun_out/r2h_$name.log 2>&1
  echo "$name rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/r2h_$name.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2h_$name.log | head -1)" | tee -a $S; }
run sma2 --optimizer sma --steps 30 --warmup 5 --no-e2e
run ours2 --steps 30 --warmup 5 --no-e2e
( timeout 300 kungfu_b200/bin/kungfu-run -q -w -np 1 -H 127.0.0.1:2 -standby 1 -builtin-config-port 9100 python benchmarks/adaptation/elastic_resize.py --schedule 1:10,2:10,1:10 --model resnet50 ) > gpurun_out/r2h_elastic_standby.log 2>&1
echo "elastic standby (eager) rc=$? $(grep -o '"latency_s": [0-9.]*' gpurun_out/r2h_elastic_standby.log | tr '\n' ' ')" | tee -a $S
( timeout 300 $TR --nproc-per-node $N --master-port 29705 benchmarks/p2p_bw.py --out gpurun_out/p2p_bw_2gpu.json ) > gpurun_out/r2h_p2p.log 2>&1
echo "p2p bw rc=$? $(grep -c clocks gpurun_out/p2p_bw_2gpu.json)" | tee -a $S
cat $S
