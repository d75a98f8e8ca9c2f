#!/usr/bin/env bash
# tools/gpu_evidence.sh -- one pass over everything profiles/ holds for the single-GPU path, on a B200 box.
#
#   gpurun --timeout 1500 -- 'bash tools/gpu_evidence.sh r02'
#
# Writes into gpurun_out/<tag>_* (scratch; copy what should be judged into profiles/):
#   <tag>_pytest_gpu.log            python -m pytest tests -q -m gpu
#   <tag>_bench_n262144.json        bench.py default line (config 3)
#   <tag>_bench_reference.json      bench.py --impl reference
#   <tag>_bench_n10000*.json        config 2, single calls and batches of 50
#   <tag>_bench_*_chunk.json        the same lines with MAPC_CHUNK=1 (bounded chains)
#   <tag>_ubench_s32.txt            FP32-pipe microbenchmarks and the library's launch shapes at S = 32
#   <tag>_launches.csv              ncu launch list of the bench command (gpu__time_duration.sum per launch)
#   <tag>_force_full.ncu-rep/.csv   ncu --set full of one force kernel launch + its raw page as CSV
#   <tag>_ncu_traffic.json          dram bytes per launch of that capture, in the format bench.py reads
# Numbers printed under ncu are attribution only; bench values come from the un-profiled runs above.
# Every step is bounded by `timeout`; a failing step does not stop the following ones.
set -u
TAG="${1:-rXX}"
OUT=gpurun_out
mkdir -p "$OUT"
cd "$(dirname "$0")/.."

step() { echo "=== $*" >&2; }

step "pytest -m gpu"
timeout 1200 python -m pytest tests -q -m gpu -x > "$OUT/${TAG}_pytest_gpu.log" 2>&1
tail -5 "$OUT/${TAG}_pytest_gpu.log" >&2

step "bench, config 3"
timeout 600 python bench.py --steps 10 --warmup 3 > "$OUT/${TAG}_bench_n262144.json" 2> "$OUT/${TAG}_bench_n262144.err"
step "bench, reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > "$OUT/${TAG}_bench_reference.json" 2>> "$OUT/${TAG}_bench_n262144.err"
step "bench, config 2 (N = 10,000)"
timeout 300 python bench.py --bodies 10000 --steps 1000 --warmup 50 --no-l2-flush --no-cpu-baseline \
    > "$OUT/${TAG}_bench_n10000.json" 2>> "$OUT/${TAG}_bench_n262144.err"
timeout 300 python bench.py --bodies 10000 --steps 1000 --warmup 50 --batch 50 --no-cpu-baseline \
    > "$OUT/${TAG}_bench_n10000_batched.json" 2>> "$OUT/${TAG}_bench_n262144.err"

step "A/B: bounded-chain order (MAPC_CHUNK=1, experimental) at config 3 and at N = 4,194,304"
MAPC_CHUNK=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline \
    > "$OUT/${TAG}_bench_n262144_chunk.json" 2>> "$OUT/${TAG}_bench_n262144.err"
timeout 600 python bench.py --bodies 4194304 --steps 2 --warmup 3 --no-cpu-baseline \
    > "$OUT/${TAG}_bench_n4194304.json" 2>> "$OUT/${TAG}_bench_n262144.err"
MAPC_CHUNK=1 timeout 600 python bench.py --bodies 4194304 --steps 2 --warmup 3 --no-cpu-baseline \
    > "$OUT/${TAG}_bench_n4194304_chunk.json" 2>> "$OUT/${TAG}_bench_n262144.err"

step "microbenchmarks + library shapes with the current S (tools/ubench N targets S)"
if [ -x tools/ubench ]; then
    timeout 600 tools/ubench 262144 0 32 > "$OUT/${TAG}_ubench_s32.txt" 2>&1
fi

step "ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file "$OUT/${TAG}_launches.csv" python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
    > "$OUT/${TAG}_launches_bench.log" 2>&1

step "ncu --set full, one force kernel launch"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:force_cells -s 2 -c 1 -f \
    -o "$OUT/${TAG}_force_full" python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
    > "$OUT/${TAG}_force_full_bench.log" 2>&1
if [ -f "$OUT/${TAG}_force_full.ncu-rep" ]; then
    ncu -i "$OUT/${TAG}_force_full.ncu-rep" --page raw --csv > "$OUT/${TAG}_force_full.csv" 2>/dev/null
    python - "$OUT/${TAG}_force_full.csv" "$OUT/${TAG}_ncu_traffic.json" "$TAG" <<'EOF'
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
# raw page: header row (metric names), unit row, then one row per profiled launch
hdr = next(r for r in rows if "Kernel Name" in r)
units = rows[rows.index(hdr) + 1]
data = rows[rows.index(hdr) + 2]
col = {name: k for k, name in enumerate(hdr)}
def metric(name):
    v = float(data[col[name]].replace(",", ""))
    u = units[col[name]].lower()
    scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
    return v * scale
total = metric("dram__bytes_read.sum") + metric("dram__bytes_write.sum")
grid = data[col["Grid Size"]] if "Grid Size" in col else ""
out = {"kernel": data[col["Kernel Name"]], "n": 262144, "dram_bytes_per_launch": total,
       "grid": grid, "source": f"profiles/{sys.argv[3]}_force_full.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
# the bench workload's canonical segment count: grid.y of the launch
try:
    out["segments"] = int(grid.strip("() ").split(",")[1])
except Exception:
    pass
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(out)
EOF
fi
step "done"
ls -la "$OUT" | grep "${TAG}_" >&2
