# final 1-GPU validation of the round: full GPU suite, smoke, headline bench (both arms), tcgen05 GEMM table, ncu of the BN kernels
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== bench ours"; timeout 400 python bench.py 2>&1 | grep "^{" | tail -1
echo "=== bench reference"; timeout 100 python bench.py --impl reference 2>&1 | grep "^{" | tail -1
echo "=== gemm"; timeout 300 python benchmarks/gemm_bias_gelu.py --out gpurun_out/gemm_bias_gelu.json 2>&1 | grep -E "^\{|Error|error" | cut -c1-500
echo "=== ncu"; timeout 500 ncu --set full --clock-control none --import-source on -k regex:'bn_|maxpool' --launch-skip 14 -c 14 -f -o gpurun_out/prof_bn2 python benchmarks/ncu_bn.py 2>&1 | tail -3
