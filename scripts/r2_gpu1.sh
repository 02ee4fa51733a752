#!/usr/bin/env bash
# One-GPU validation in ONE call:  gpurun --timeout 1800 -- 'bash scripts/r2_gpu1.sh [steps...]'
# steps: xtests (new kernel family first: fail fast), tests (whole -m gpu suite), smoke, bench, ncu
set -u
cd "$(dirname "${BASH_SOURCE[0]}")/.."
OUT=gpurun_out/${R2_TAG:-r2_gpu1}
mkdir -p "$OUT"
STEPS=${*:-xtests tests smoke bench}
step() {
  local name=$1 t=$2; shift 2
  echo "== $name" | tee -a "$OUT/summary.txt"
  local t0=$SECONDS
  timeout -k 10 "$t" "$@" >"$OUT/$name.log" 2>&1
  echo "   rc=$? $((SECONDS - t0))s" | tee -a "$OUT/summary.txt"
}
for s in $STEPS; do
  case $s in
    xtests) step xtests 600 python -m pytest tests/test_gpu_x_kernels.py -q -m gpu -x; tail -30 "$OUT/xtests.log";;
    tests) step tests 1500 python -m pytest tests -q -m gpu; tail -40 "$OUT/tests.log";;
    tests_noseam) step tests_noseam 1200 python -m pytest tests -q -m gpu --deselect tests/test_gpu_reference_seam.py; tail -40 "$OUT/tests_noseam.log";;
    smoke) step smoke 300 python -c "import __graft_entry__ as g; g.smoke()"; tail -3 "$OUT/smoke.log";;
    bench) step bench 900 python bench.py; tail -c 6000 "$OUT/bench.log";;
    benchq) step benchq 600 python bench.py --steps 50 --no-cpu --no-refcuda --no-config4; tail -c 4000 "$OUT/benchq.log";;
    benchsw) step benchsw 600 env TTB_TAIL_SWEEP_FLOATS=0 python bench.py --steps 50 --no-cpu --no-refcuda --no-config4; tail -c 1500 "$OUT/benchsw.log";;
    legacy) step legacy 600 env TTB_LEGACY_TC=1 python bench.py --steps 50 --no-cpu --no-refcuda --no-config4; tail -c 3000 "$OUT/legacy.log";;
    cfgs) step cfgs 900 python scripts/bench_configs.py all; tail -c 4000 "$OUT/cfgs.log";;
    ncu_list) step ncu_list 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
        --log-file "$OUT/launches_bench.csv" python bench.py --steps 4 --warmup 3 --no-cpu --no-refcuda --no-config4;;
    ncu_cache) step ncu_cache 900 ncu --set full --clock-control none -k regex:"cache|partition|update_cache|frontend|rowidx|mark_popular" --launch-skip 40 -c 24 \
        -o "$OUT/cfg3_cache" -f python scripts/bench_configs.py cfg3;;
    ncu_big) step ncu_big 900 ncu --set full --clock-control none --import-source on -k regex:"x_bwd|x_fwd|plan_" --launch-skip 10 -c 5 \
        -o "$OUT/s1_65536" -f python scripts/profile_big.py;;
    ncu_ranks) step ncu_ranks 900 ncu --set full --clock-control none --import-source on \
        -k regex:"tt_fwd_generic|tt_bwd_generic|tt_fwd_bk|tt_bwd_bk|x_fwd|x_bwd|optimizer_sweep" -c ${NCU_COUNT:-16} \
        -o "$OUT/cfg5_ranks" -f python scripts/profile_rank_sweep.py ${RANKS:-8 16 128};;
    seam) step seam 900 python -m pytest tests/test_gpu_reference_seam.py -q -x -k benchmark -s;;
    ncu_full) step ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:"x_bwd|x_fwd|plan_onepass" -c 9 \
        -o "$OUT/bench_full" -f python bench.py --steps 2 --warmup 3 --no-cpu --no-refcuda --no-config4 --no-graph;;
  esac
done
# .ncu-rep files are too large to travel back (gpurun merges <= 64 MiB): reduce them to CSV on the box
for rep in "$OUT"/*.ncu-rep; do
  [ -f "$rep" ] || continue
  base="${rep%.ncu-rep}"
  ncu -i "$rep" --page raw --csv > "$base.raw.csv" 2>/dev/null
  ncu -i "$rep" --page source --csv > "$base.source.csv" 2>/dev/null
  python scripts/ncu_reduce.py "$base.raw.csv" "$base.source.csv" > "$base.summary.txt" 2>&1
  gzip -f "$base.source.csv"
  rm -f "$rep"
done
cat "$OUT/summary.txt"
