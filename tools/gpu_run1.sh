#!/usr/bin/env bash
# first GPU pass of round 2: the -m gpu suite on the frozen canonical order, bench config 3, N = 4 M, the
# operand-fetch probes of tools/ubench (first 60 lines: the FMA-pipe patterns), ring on/off A/B
set -u
OUT=gpurun_out; mkdir -p $OUT; cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/r02a_smi.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -x -s > $OUT/r02a_pytest_gpu.log 2>&1; tail -5 $OUT/r02a_pytest_gpu.log >&2
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/r02a_bench_n262144.json 2> $OUT/r02a_bench.err; tail -c 1500 $OUT/r02a_bench_n262144.json >&2
MAPC_RING=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r02a_bench_n262144_noring.json 2>> $OUT/r02a_bench.err
timeout 600 python bench.py --bodies 4194304 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/r02a_bench_n4194304.json 2>> $OUT/r02a_bench.err
timeout 300 python bench.py --bodies 10000 --steps 1000 --warmup 50 --batch 50 --no-cpu-baseline > $OUT/r02a_bench_n10000_batched.json 2>> $OUT/r02a_bench.err
timeout 300 tools/ubench 262144 0 32 2>&1 | head -80 > $OUT/r02a_ubench_probes.txt
tail -3 $OUT/r02a_bench.err >&2
ls -la $OUT >&2
