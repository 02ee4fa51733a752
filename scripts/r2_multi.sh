#!/usr/bin/env bash
# Multi-GPU validation in ONE call:  gpurun --gpus N --timeout 1200 -- 'bash scripts/r2_multi.sh N'
# 1. tests/test_gpu_multi.py (NCCL all-to-all, fused het batch, peer exchange over symmetric memory, DP replicas)
# 2. bench.py --gpus N (config 4 table-sharded: nccl + peer exchange, eager + CUDA graph, parity check inside)
set -u
cd "$(dirname "${BASH_SOURCE[0]}")/.."
N=${1:-2}
OUT=gpurun_out/r2_multi_$N
mkdir -p "$OUT"
step() {
  local name=$1 t=$2; shift 2
  echo "== $name" | tee -a "$OUT/summary.txt"
  local t0=$SECONDS
  timeout -k 10 "$t" "$@" >"$OUT/$name.log" 2>&1
  echo "   rc=$? $((SECONDS - t0))s" | tee -a "$OUT/summary.txt"
}
nvidia-smi topo -m >"$OUT/topo.txt" 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  step tests_multi 600 python -m pytest tests/test_gpu_multi.py -q -m gpu
fi
step bench_ref 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
  --master-port 29583 bench.py --impl reference --gpus "$N" --steps 2 --warmup 1 --cpu-budget-s 20
step bench 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
  --master-port 29582 bench.py --gpus "$N" --steps ${STEPS:-30} --warmup 5 --no-refcuda
tail -c 3000 "$OUT/bench.log"
cat "$OUT/summary.txt"
