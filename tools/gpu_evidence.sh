#!/usr/bin/env bash
# tools/gpu_evidence.sh -- one pass over everything profiles/ holds for the single-GPU path, on a B200 box.
#
#   gpurun --timeout 1700 -- 'bash tools/gpu_evidence.sh r02'
#
# Writes into gpurun_out/<tag>_* (scratch; copy what should be judged into profiles/):
#   <tag>_pytest_gpu.log            python -m pytest tests -q -m gpu -s
#   <tag>_bench_default.json        bench.py default line (config 3 + strong / determinism / well legs + cpu_baseline)
#   <tag>_bench_reference.json      bench.py --impl reference
#   <tag>_bench_n10000*.json        config 2, single calls and batches of 50
#   <tag>_bench_well.json           bench.py --mode well (the reference's shipped kernel, HBM roofline)
#   <tag>_soak.txt                  tools/soak.py: 20,000 chained steps / 100 ring steps, bitwise against the protocols off
#   <tag>_compute_sanitizer.txt     tools/gpu_sanitizer.sh (memcheck / racecheck / synccheck), when SANITIZER=1
#   <tag>_launches.csv              ncu launch list of the bench command (gpu__time_duration.sum per launch)
#   <tag>_force_full.ncu-rep/.csv   ncu --set full of one force kernel launch at N = 262,144 + its raw page as CSV
#   <tag>_force_noring_dram.csv     dram bytes of the same launch with MAPC_RING=0 (one scratch slot per target block)
#   <tag>_force_4m_dram.csv         dram bytes of one force launch at N = 4,194,304
#   <tag>_well_full.ncu-rep/.csv    ncu --set full of one well_step_kernel launch at N = 4,194,304
#   <tag>_ncu_traffic.json          dram bytes per launch of those captures, in the format bench.py reads
# Numbers printed under ncu are attribution only; bench values come from the un-profiled runs above.
# Every step is bounded by `timeout`; a failing step does not stop the following ones.
set -u
TAG="${1:-rXX}"
OUT=gpurun_out
mkdir -p "$OUT"
cd "$(dirname "$0")/.."
ERR="$OUT/${TAG}_bench.err"; : > "$ERR"
step() { echo "=== $*" >&2; }

if [ "${SKIP_TESTS:-0}" != 1 ]; then
step "pytest -m gpu"
timeout 1500 python -m pytest tests -q -m gpu -x -s > "$OUT/${TAG}_pytest_gpu.log" 2>&1
tail -3 "$OUT/${TAG}_pytest_gpu.log" >&2
fi

step "bench, default line"
timeout 900 python bench.py --steps 10 --warmup 3 > "$OUT/${TAG}_bench_default.json" 2>> "$ERR"
tail -c 600 "$OUT/${TAG}_bench_default.json" >&2
step "bench, reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > "$OUT/${TAG}_bench_reference.json" 2>> "$ERR"
step "bench, config 2 (N = 10,000)"
timeout 300 python bench.py --bodies 10000 --steps 1000 --warmup 50 --no-l2-flush --no-cpu-baseline --headline-only \
    > "$OUT/${TAG}_bench_n10000.json" 2>> "$ERR"
timeout 300 python bench.py --bodies 10000 --steps 1000 --warmup 50 --batch 50 --no-cpu-baseline --headline-only \
    > "$OUT/${TAG}_bench_n10000_batched.json" 2>> "$ERR"
step "soak: chained steps and scratch ring, bit-identity over many steps"
timeout 600 python tools/soak.py > "$OUT/${TAG}_soak.txt" 2>> "$ERR"; cat "$OUT/${TAG}_soak.txt" >&2
step "bench, well mode"
timeout 300 python bench.py --mode well --steps 20 --warmup 5 > "$OUT/${TAG}_bench_well.json" 2>> "$ERR"
tail -c 400 "$OUT/${TAG}_bench_well.json" >&2

if [ "${SKIP_NCU:-0}" != 1 ]; then
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --headline-only"
step "ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file "$OUT/${TAG}_launches.csv" $B > "$OUT/${TAG}_launches_bench.log" 2>&1
step "ncu --set full, one force kernel launch (N = 262,144)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:force_cells -s 2 -c 1 -f \
    -o "$OUT/${TAG}_force_full" $B > "$OUT/${TAG}_force_full_bench.log" 2>&1
[ -f "$OUT/${TAG}_force_full.ncu-rep" ] && ncu -i "$OUT/${TAG}_force_full.ncu-rep" --page raw --csv > "$OUT/${TAG}_force_full.csv" 2>/dev/null
DRAM="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum"
step "dram bytes: MAPC_RING=0 at N = 262,144, ring at N = 4,194,304"
MAPC_RING=0 timeout 600 ncu --metrics $DRAM --clock-control none -k regex:force_cells -s 2 -c 1 --csv \
    --log-file "$OUT/${TAG}_force_noring_dram.csv" $B > /dev/null 2>&1
timeout 900 ncu --metrics $DRAM --clock-control none -k regex:force_cells -s 1 -c 1 --csv \
    --log-file "$OUT/${TAG}_force_4m_dram.csv" python bench.py --bodies 4194304 --steps 1 --warmup 3 --no-cpu-baseline --headline-only > /dev/null 2>&1
step "ncu --set full, one well_step_kernel launch (N = 4,194,304)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:well_step -s 3 -c 1 -f \
    -o "$OUT/${TAG}_well_full" python bench.py --mode well --steps 10 --warmup 3 > "$OUT/${TAG}_well_full_bench.log" 2>&1
[ -f "$OUT/${TAG}_well_full.ncu-rep" ] && ncu -i "$OUT/${TAG}_well_full.ncu-rep" --page raw --csv > "$OUT/${TAG}_well_full.csv" 2>/dev/null
python tools/ncu_traffic.py "$OUT" "$TAG" > "$OUT/${TAG}_ncu_traffic.json" 2>> "$ERR"
cat "$OUT/${TAG}_ncu_traffic.json" >&2
fi
if [ "${SANITIZER:-0}" = 1 ]; then
step "compute-sanitizer"
bash tools/gpu_sanitizer.sh "$TAG" 2>/dev/null
tail -n 40 "$OUT/${TAG}_compute_sanitizer.txt" >&2
fi
step "done"
tail -5 "$ERR" >&2
ls -la "$OUT" | grep "${TAG}_" >&2
