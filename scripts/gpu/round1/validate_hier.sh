mkdir -p gpurun_out
echo "=== host-spanning paths (forced groups) + sma/pair"; timeout 420 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "host_spanning or sma_pair" 2>&1 | tail -15
