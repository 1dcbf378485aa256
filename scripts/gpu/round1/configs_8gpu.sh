# BASELINE.json configs #2-#5 on 8 x B200 (one gpurun call)
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "=== #2 S-SGD N=8"; timeout 240 $T --nproc-per-node 8 --master-port 29801 bench.py --gpus 8 --steps 30 --warmup 5 2>&1 | grep "^{" | tail -1
echo "=== #2 S-SGD N=4"; timeout 240 $T --nproc-per-node 4 --master-port 29802 bench.py --gpus 4 --steps 30 --warmup 5 2>&1 | grep "^{" | tail -1
echo "=== #3 SMA N=8"; timeout 300 $T --nproc-per-node 8 --master-port 29803 benchmarks/system/benchmark_kungfu_torch.py --kf-optimizer sma --num-iters 3 --num-batches-per-iter 10 --num-warmup-batches 5 2>&1 | grep -E "RESULT|Error|Traceback" | tail -3
echo "=== GNS N=8"; timeout 300 $T --nproc-per-node 8 --master-port 29804 benchmarks/system/benchmark_kungfu_torch.py --kf-optimizer noise-scale --num-iters 3 --num-batches-per-iter 10 --num-warmup-batches 5 2>&1 | grep -E "RESULT|Error|Traceback" | tail -3
echo "=== #4 BERT-large PairAveraging N=8"; timeout 400 $T --nproc-per-node 8 --master-port 29805 benchmarks/bert_pairavg.py --model large --batch 32 --seq 128 --steps 10 --warmup 3 2>&1 | grep -E "RESULT|Error|Traceback" | tail -3
echo "=== #5 elastic 4->8->4 + GNS"; timeout 500 kungfu_b200/bin/kungfu-run -q -w -np 4 -H 127.0.0.1:8 -port 31777 -port-range 21777-21830 -builtin-config-port 9777 python benchmarks/adaptation/elastic_resize.py --schedule 4:10,8:10,4:10 --batch-size 64 2>&1 | grep -E "\{|Error|Traceback" | sed 's/\x1b\[[0-9;]*m//g' | tail -12
