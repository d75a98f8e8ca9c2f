#!/usr/bin/env bash
# GPU pass 7: config 2 after the batched combine loads, the (1,128) shape, config 3 sanity
set -u
OUT=gpurun_out; mkdir -p $OUT; cd "$(dirname "$0")/.."
B2="python bench.py --bodies 10000 --steps 1000 --warmup 50 --no-cpu-baseline --headline-only"
timeout 300 $B2 --batch 50 > $OUT/r02g_n10000_batched.json 2>> $OUT/r02g.err
timeout 300 $B2 --no-l2-flush > $OUT/r02g_n10000_single.json 2>> $OUT/r02g.err
for shape in "1 128" "1 32" "2 64"; do set -- $shape
  MAPC_PLAN_PAIRS=$1 MAPC_PLAN_THREADS=$2 timeout 300 $B2 --batch 50 > $OUT/r02g_n10000_batched_p$1_t$2.json 2>> $OUT/r02g.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --headline-only > $OUT/r02g_n262144.json 2>> $OUT/r02g.err
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "geometry or chained or batched or timer" > $OUT/r02g_pytest.log 2>&1; tail -n 3 $OUT/r02g_pytest.log >&2
python - <<'PY' >&2
import json,glob
for f in sorted(glob.glob('gpurun_out/r02g_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('r02g_')[1], 'ms', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],4), 'kernel', round(d['roofline']['kernel_ms_in_kernel_stamps'],5), d['config']['plan']['pairs_per_thread'], d['config']['plan']['threads_per_block'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -n 3 $OUT/r02g.err >&2
