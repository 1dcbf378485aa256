# 4-GPU check: collectives (bit-identical one-shot across ranks), fused optimizers, SMA/pair, headline bench at N=4
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "=== multi-GPU tests (N=4)"; timeout 700 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "collectives or fused or sma_pair" 2>&1 | tail -4
echo "=== bench N=4"; timeout 240 $T --nproc-per-node 4 --master-port 29832 bench.py --gpus 4 --steps 30 --warmup 5 2>&1 | grep "^{" | tail -1
