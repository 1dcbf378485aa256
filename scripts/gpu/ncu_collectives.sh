#!/bin/bash
# Nsight Compute capture of the collective / fused BODY kernels on rank 0 of a live N-rank job (the other ranks run unprofiled).
# usage: scripts/gpu/ncu_collectives.sh N [out-prefix]
set -u
N=${1:-2}; OUT=${2:-gpurun_out/ncu_collectives_${N}gpu}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export MASTER_ADDR=127.0.0.1 MASTER_PORT=$((29800 + RANDOM % 100)) WORLD_SIZE=$N KUNGFU_DEVICE_TIMEOUT_MS=180000 KF_NCU_ROUNDS=2
pids=()
for r in $(seq 1 $((N - 1))); do
  RANK=$r LOCAL_RANK=$r timeout 900 python benchmarks/ncu_collectives.py > ${OUT}.rank$r.log 2>&1 &
  pids+=($!)
done
RANK=0 LOCAL_RANK=0 timeout 900 ncu --set full --clock-control none --import-source on \
  -k 'regex:ssgd_fused|allreduce_nvls|allreduce_twoshot|pull_tma|pair_avg|sma_blend|allreduce_ll' -c 14 -f -o ${OUT} \
  python benchmarks/ncu_collectives.py > ${OUT}.rank0.log 2>&1
rc=$?
for p in "${pids[@]}"; do wait $p; done
echo "ncu collectives N=$N rc=$rc $(ls -la ${OUT}.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
