#!/usr/bin/env bash
# tools/gpu_variants.sh TAG -- A/Bs on one B200 (bench.py --headline-only, so each point is a few seconds):
#   config 2 (N = 10,000, 1,000 steps): chained steps vs grid-wide waits, batches vs single calls, every small shape
#   config 3 (N = 262,144): every launch shape forced, and the A/B instantiations of csrc/force_shapes.inc
# Appends a table to gpurun_out/<tag>_shape_variants.txt.  (The sweeps that chose the current table are recorded in
# profiles/r02_shape_variants.txt; their extra instantiations are no longer compiled in.)
set -u
TAG="${1:-rXX}"; OUT=gpurun_out; mkdir -p $OUT; cd "$(dirname "$0")/.."
B2="python bench.py --bodies 10000 --steps 1000 --warmup 50 --no-cpu-baseline --headline-only"
B3="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --headline-only"
run() { local name=$1; shift; env "$@" > $OUT/${TAG}_v_$name.json 2>> $OUT/${TAG}_v.err; }
for chain in 1 0; do
  run n10000_batched_chain$chain MAPC_CHAIN=$chain timeout 300 $B2 --batch 50
  run n10000_single_chain$chain MAPC_CHAIN=$chain timeout 300 $B2 --no-l2-flush
done
for shape in "1 128" "1 64" "1 32" "2 64"; do set -- $shape
  run n10000_batched_p$1_t$2 MAPC_PLAN_PAIRS=$1 MAPC_PLAN_THREADS=$2 timeout 300 $B2 --batch 50
done
for shape in "2 128" "2 64" "4 256" "4 128" "1 128"; do set -- $shape
  run n262144_p$1_t$2 MAPC_PLAN_PAIRS=$1 MAPC_PLAN_THREADS=$2 timeout 300 $B3
done
run n262144_p4_t256_pairmajor MAPC_PLAN_PAIRS=4 MAPC_PLAN_THREADS=256 MAPC_SHAPE_VARIANT=1 timeout 300 $B3
run n262144_p4_t256_u16 MAPC_PLAN_PAIRS=4 MAPC_PLAN_THREADS=256 MAPC_SHAPE_VARIANT=2 timeout 300 $B3
run n262144_p2_t128_u4 MAPC_PLAN_PAIRS=2 MAPC_PLAN_THREADS=128 MAPC_SHAPE_VARIANT=1 timeout 300 $B3
run n262144_p2_t128_tj128 MAPC_PLAN_PAIRS=2 MAPC_PLAN_THREADS=128 MAPC_SHAPE_VARIANT=2 timeout 300 $B3
run n262144_p2_t128_u8_pairmajor MAPC_PLAN_PAIRS=2 MAPC_PLAN_THREADS=128 MAPC_SHAPE_VARIANT=3 timeout 300 $B3
run n262144_p2_t64_tj64_u4 MAPC_PLAN_PAIRS=2 MAPC_PLAN_THREADS=64 MAPC_SHAPE_VARIANT=1 timeout 300 $B3
python - "$TAG" <<'PY' | tee $OUT/${TAG}_shape_variants.txt >&2
import json,glob,sys
tag=sys.argv[1]
for f in sorted(glob.glob(f'gpurun_out/{tag}_v_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split(tag+'_v_')[1].replace('.json',''), 'ms_per_step', round(d['ms_per_step'],5), 'frac_fp32_peak', round(d['roofline']['frac'],4), 'kernel_ms_in_kernel_stamps', round(d['roofline']['kernel_ms_in_kernel_stamps'],5), 'shape', (d['config']['plan']['pairs_per_thread'], d['config']['plan']['threads_per_block']))
    except Exception as e: print(f, 'ERR', e)
PY
tail -n 3 $OUT/${TAG}_v.err >&2
