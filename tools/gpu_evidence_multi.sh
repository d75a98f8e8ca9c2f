#!/usr/bin/env bash
# tools/gpu_evidence_multi.sh -- the multi-GPU evidence of profiles/ in one call on an N-GPU box:
#
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_evidence_multi.sh r02 2'
#   gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_evidence_multi.sh r02 8 quick'
#
# Writes gpurun_out/<tag>_*: the multi-GPU tests (bit-identity sharded vs unsharded with every exchange, sharded
# InitializeParticles, a consumer per rank; cross-device consumer + migration), then bench.py at the box's GPU
# count: the default line (weak-scaled headline + strong / determinism / config-4 legs) and the headline with the
# peer exchanges.  "quick" skips the tests that have a 2-GPU log already and the peer A/Bs.
# Every step is bounded by `timeout`; a failing step does not stop the following ones.
set -u
TAG="${1:-rXX}"
GPUS="${2:-8}"
QUICK="${3:-}"
OUT=gpurun_out
mkdir -p "$OUT"
cd "$(dirname "$0")/.."
PORT=29540

run() {   # run <name> <bench args...>
    local name="$1"; shift
    PORT=$((PORT + 1))
    echo "=== $name" >&2
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$GPUS" --master-addr 127.0.0.1 \
        --master-port "$PORT" bench.py --gpus "$GPUS" "$@" > "$OUT/${TAG}_${name}.json" 2> "$OUT/${TAG}_${name}.err"
    tail -c 700 "$OUT/${TAG}_${name}.json" >&2
    grep -v "^W\|^\[W\|warn" "$OUT/${TAG}_${name}.err" | tail -3 >&2
}

nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > "$OUT/${TAG}_${GPUS}gpu_smi.txt" 2>&1
if [ "$QUICK" != quick ]; then
echo "=== multi-GPU tests" >&2
timeout 1500 python -m pytest tests -q -m gpu -s > "$OUT/${TAG}_pytest_multi_gpu_${GPUS}.log" 2>&1
tail -5 "$OUT/${TAG}_pytest_multi_gpu_${GPUS}.log" >&2
fi
echo "=== bit-identity across exchanges, sharded ICs, per-rank consumers" >&2
PORT=$((PORT + 1))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$GPUS" --master-addr 127.0.0.1 \
    --master-port "$PORT" tests/mgpu_worker.py --exchange all > "$OUT/${TAG}_mgpu_worker_${GPUS}gpu.log" 2>&1
grep "bit-identical\|skipped" "$OUT/${TAG}_mgpu_worker_${GPUS}gpu.log" | tail -20 >&2

run "bench_${GPUS}gpu_default" --steps 10 --warmup 3
if [ "$QUICK" != quick ]; then
run "bench_${GPUS}gpu_weak_peer" --steps 10 --warmup 3 --exchange peer --headline-only
run "bench_${GPUS}gpu_weak_peer_single" --steps 10 --warmup 3 --exchange peer-single --headline-only
fi
ls -la "$OUT" | grep "${TAG}_" >&2
