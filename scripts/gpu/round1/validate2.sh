python benchmarks/profile_step.py --out gpurun_out/resnet50_step_cupti.md 2>&1 | grep -v Warn | head -26 | cut -c1-160
echo "=== tests"; python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "maxpool or bn or resnet or trainer" 2>&1 | tail -3
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "sma_pair" 2>&1 | tail -3
echo "=== BERT pair"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29703 benchmarks/bert_pairavg.py --model base --steps 5 --warmup 2 2>&1 | grep -E "RESULT|Error|error|Traceback" | tail -5
echo "=== elastic"; timeout 400 kungfu_b200/bin/kungfu-run -q -w -np 1 -H 127.0.0.1:2 -port 31555 -port-range 21555-21600 -builtin-config-port 9555 python benchmarks/adaptation/elastic_resize.py --schedule 1:6,2:6,1:6 --batch-size 64 2>&1 | grep -E "\{|Error|error|Traceback" | tail -12
echo "=== bench1"; python bench.py --steps 30 --warmup 5 2>&1 | grep "^{" | tail -1
