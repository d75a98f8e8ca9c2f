#!/usr/bin/env bash
# tools/gpu_sanitizer.sh TAG -- compute-sanitizer memcheck / racecheck / synccheck over the paths of the library:
# the reference's frame loop with the copying and the async consumer (small N), chained steps without a consumer,
# the scratch ring with its ticket (N = 98,304: 48 target blocks share 32 slots), the well kernel and device-side ICs.
set -u
TAG="${1:-rXX}"; OUT=gpurun_out; mkdir -p $OUT; cd "$(dirname "$0")/.."
LOG=$OUT/${TAG}_compute_sanitizer.txt; : > $LOG
run() { echo "== $*" >> $LOG; timeout 600 compute-sanitizer "$@" 2>&1 | grep -v "^=========     \|^=========$" | tail -n 6 >> $LOG; echo "rc=${PIPESTATUS[0]}" >> $LOG; }
for tool in memcheck racecheck synccheck; do
  run --tool $tool tools/mapc_run --numparticles 3000 --steps 4 --radius 1300
  run --tool $tool tools/mapc_run --numparticles 3000 --steps 4 --radius 1300 --async
  run --tool $tool tools/mapc_run --numparticles 3000 --steps 6 --radius 1300 --no-consumer
done
run --tool memcheck tools/mapc_run --numparticles 98304 --steps 2 --radius 5800 --no-consumer
run --tool synccheck tools/mapc_run --numparticles 98304 --steps 2 --radius 5800 --no-consumer
run --tool racecheck tools/mapc_run --numparticles 98304 --steps 1 --radius 5800 --no-consumer
run --tool memcheck tools/mapc_run --numparticles 70000 --steps 2 --mode well --ic shells
run --tool racecheck tools/mapc_run --numparticles 70000 --steps 2 --mode well --ic shells
cat $LOG >&2
