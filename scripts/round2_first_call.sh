#!/usr/bin/env bash
# Everything that was written after round 1's GPU budget ran out, in ONE gpurun call (1 GPU, ~15 min):
#   gpurun --timeout 1500 -- 'bash scripts/round2_first_call.sh'
# Each step has its own timeout and log under gpurun_out/r2_first/; a failing step does not stop the rest.
# Read the logs afterwards, copy what is worth keeping into profiles/ (r2_*).
set -u
cd "$(dirname "${BASH_SOURCE[0]}")/.."
OUT=gpurun_out/r2_first
mkdir -p "$OUT"
step() {  # step <name> <timeout_s> <command...>
  local name=$1 t=$2; shift 2
  echo "== $name" | tee -a "$OUT/summary.txt"
  local t0=$SECONDS
  timeout "$t" "$@" >"$OUT/$name.log" 2>&1
  echo "   rc=$? $((SECONDS - t0))s" | tee -a "$OUT/summary.txt"
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv >"$OUT/gpu.csv" 2>&1

# 1. parity: the measured suites first, then the never-run ones one file at a time (a failure in one must not hide the others)
step tests_measured 900 python -m pytest tests -q -m gpu --ignore-glob="tests/test_zz*"
for f in tests/test_zz1_gpu_group.py tests/test_zz2_gpu_async_cache.py tests/test_zz4_gpu_fused.py tests/test_zz9_gpu_fuzz.py; do
  step "$(basename "$f" .py)" 600 python -m pytest "$f" -q -m gpu
done
step smoke 300 python -c "import __graft_entry__ as g; g.smoke()"

# 2. tcgen05 feature probe for the backward revision (MN-major tf32 with SWIZZLE_128B_BASE32B, A operand from TMEM)
step probe2_build 300 nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Ifbtt_embedding_b200/csrc -Iinclude \
  tests/cuda/mma_probe2.cu -o "$OUT/mma_probe2"
step probe2_run 60 "$OUT/mma_probe2"
# ... and for bf16 operands (kind::f16): K-major / MN-major under the standard swizzle, hi/lo split precision vs tf32
step probe3_build 300 nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Ifbtt_embedding_b200/csrc -Iinclude \
  tests/cuda/mma_probe3.cu -o "$OUT/mma_probe3"
step probe3_run 60 "$OUT/mma_probe3"

# 3. headline bench, then the opt-in kernel variants A/B (TTB_BWD_VEC_FLUSH, TTB_PDL)
step bench_n1 600 python bench.py
step ab_variants 900 python scripts/ab_variants.py --steps 200

# 4. config 4 on one GPU: per-table modules vs table group (1 and 4 lanes) vs fused heterogeneous batch
for mode in "CFG4_GROUPED=0 CFG4_FUSED=0" "CFG4_GROUPED=1 CFG4_LANES=1" "CFG4_GROUPED=1 CFG4_LANES=4" "CFG4_FUSED=1"; do
  tag=$(echo "$mode" | tr ' =' '__')
  step "cfg4_$tag" 300 env $mode STEPS=20 python scripts/bench_config4.py
done

# 5. config 3: reference flow vs async cache front-end
step cfg3 600 python scripts/bench_configs.py cfg3

# 6. ncu: launch list of the bench command + one full capture of the fused config-4 step
step ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
  --log-file "$OUT/launches_bench.csv" python bench.py --steps 4 --warmup 3 --no-cpu --no-refcuda
step ncu_cfg4_fused 900 env CFG4_FUSED=1 STEPS=2 ncu --set full --clock-control none --import-source on -c 12 \
  -o "$OUT/cfg4_fused" -f python scripts/bench_config4.py
cat "$OUT/summary.txt"
