#!/usr/bin/env bash
# tools/gpu_evidence_multi.sh -- the multi-GPU evidence of profiles/ in one call on an N-GPU box:
#
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_evidence_multi.sh r02 8'
#
# Writes gpurun_out/<tag>_*: the multi-GPU tests (bit-identity sharded vs unsharded with both exchanges,
# cross-device consumer + migration), then bench lines for BASELINE config 4 (N = 1,048,576, NCCL and peer
# exchange), the weak-scaled size and config 5 (N = 4,194,304, strong) at the box's GPU count.
# Every step is bounded by `timeout`; a failing step does not stop the following ones.
set -u
TAG="${1:-rXX}"
GPUS="${2:-8}"
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
    tail -c 600 "$OUT/${TAG}_${name}.json" >&2
}

echo "=== multi-GPU tests" >&2
timeout 1200 python -m pytest tests/test_multi_gpu.py -q -m gpu > "$OUT/${TAG}_pytest_multi_gpu.log" 2>&1
tail -5 "$OUT/${TAG}_pytest_multi_gpu.log" >&2
echo "=== bit-identity incl. the experimental one-grid peer exchange" >&2
PORT=$((PORT + 1))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$GPUS" --master-addr 127.0.0.1 \
    --master-port "$PORT" tests/mgpu_worker.py --exchange all > "$OUT/${TAG}_mgpu_worker_all.log" 2>&1
tail -15 "$OUT/${TAG}_mgpu_worker_all.log" >&2

run "bench_${GPUS}gpu_config4_nccl" --bodies 1048576 --steps 10 --warmup 3 --exchange nccl
run "bench_${GPUS}gpu_config4_peer" --bodies 1048576 --steps 10 --warmup 3 --exchange peer
run "bench_${GPUS}gpu_weak_nccl" --steps 10 --warmup 3 --exchange nccl
run "bench_${GPUS}gpu_weak_peer" --steps 10 --warmup 3 --exchange peer
run "bench_${GPUS}gpu_weak_peer_single" --steps 10 --warmup 3 --exchange peer-single
run "bench_${GPUS}gpu_strong_nccl" --scaling strong --steps 3 --warmup 3 --exchange nccl
ls -la "$OUT" | grep "${TAG}_" >&2
