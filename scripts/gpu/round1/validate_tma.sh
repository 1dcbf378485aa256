# 2-GPU check of the TMA pull kernel + p2p bandwidth table + N=2 headline bench (one gpurun call)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "=== multi-GPU tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
echo "=== p2p bw"; timeout 200 $T --nproc-per-node 2 --master-port 29821 benchmarks/p2p_bw.py --out gpurun_out/p2p_bw_2gpu.json 2>&1 | grep -E "^\{|RESULT|Error|error|Traceback" | cut -c1-400
echo "=== bench N=2"; timeout 240 $T --nproc-per-node 2 --master-port 29822 bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | grep "^{" | tail -1
