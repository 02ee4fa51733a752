#!/usr/bin/env bash
# The multi-GPU paths that have never run on a multi-GPU box, in ONE call:
#   gpurun --gpus 2 --timeout 1200 -- 'bash scripts/round2_multi_gpu.sh 2'
# (then the same with 8 for the config-4 scaling numbers).  Logs under gpurun_out/r2_multi_N/.
set -u
cd "$(dirname "${BASH_SOURCE[0]}")/.."
N=${1:-2}
OUT=gpurun_out/r2_multi_$N
mkdir -p "$OUT"
step() {
  local name=$1 t=$2; shift 2
  echo "== $name" | tee -a "$OUT/summary.txt"
  local t0=$SECONDS
  timeout "$t" "$@" >"$OUT/$name.log" 2>&1
  echo "   rc=$? $((SECONDS - t0))s" | tee -a "$OUT/summary.txt"
}
nvidia-smi topo -m >"$OUT/topo.txt" 2>&1
# parity first: table-sharded (per-table / fused / fused + peer exchange) and data-parallel replicas, 2 ranks
step tests_multi 900 python -m pytest tests/test_gpu_multi.py -q -m gpu
# config 4, table-sharded over N GPUs: NCCL all-to-all with per-table modules, table groups, fused batch; then
# the fused batch with the exchange folded into the kernels (symmetric memory, 2 rank barriers per step)
run4() {  # run4 <tag> <env...>
  local tag=$1; shift
  step "cfg4_$tag" 400 env "$@" STEPS=30 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" \
    --master-addr 127.0.0.1 --master-port 29581 scripts/bench_config4.py
}
run4 per_table CFG4_GROUPED=0
run4 grouped CFG4_GROUPED=1 CFG4_LANES=2
run4 fused_nccl CFG4_FUSED=1
run4 fused_peer CFG4_FUSED=1 CFG4_EXCHANGE=peer
# headline bench, N replicas
step bench_n 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
  --master-port 29582 bench.py --gpus "$N"
cat "$OUT/summary.txt"
