#!/usr/bin/env bash
# GPU pass 4: the gpu suite with chained steps, config 2 with and without the chain, shape variants at config 3
set -u
OUT=gpurun_out; mkdir -p $OUT; cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu -x -s > $OUT/r02d_pytest_gpu.log 2>&1; tail -4 $OUT/r02d_pytest_gpu.log >&2
B2="python bench.py --bodies 10000 --steps 1000 --warmup 50 --no-cpu-baseline --headline-only"
for chain in 1 0; do
  MAPC_CHAIN=$chain timeout 300 $B2 --batch 50 > $OUT/r02d_n10000_batched_chain$chain.json 2>> $OUT/r02d.err
  MAPC_CHAIN=$chain timeout 300 $B2 --no-l2-flush > $OUT/r02d_n10000_single_chain$chain.json 2>> $OUT/r02d.err
done
for shape in "1 32" "2 64" "2 128"; do set -- $shape
  MAPC_PLAN_PAIRS=$1 MAPC_PLAN_THREADS=$2 timeout 300 $B2 --batch 50 > $OUT/r02d_n10000_batched_p$1_t$2.json 2>> $OUT/r02d.err
done
B3="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --headline-only"
for v in 0 1 2 3 4; do
  MAPC_SHAPE_VARIANT=$v timeout 300 $B3 > $OUT/r02d_n262144_variant$v.json 2>> $OUT/r02d.err
done
for v in 0 1 2 3; do
  MAPC_PLAN_PAIRS=4 MAPC_PLAN_THREADS=128 MAPC_SHAPE_VARIANT=$v timeout 300 $B3 > $OUT/r02d_n262144_t128_variant$v.json 2>> $OUT/r02d.err
done
python - <<'PY' >&2
import json,glob
for f in sorted(glob.glob('gpurun_out/r02d_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('r02d_')[1], 'ms', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],4), 'kernel', round(d['roofline']['kernel_ms_in_kernel_stamps'],5), d['config']['plan']['pairs_per_thread'], d['config']['plan']['threads_per_block'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 $OUT/r02d.err >&2
