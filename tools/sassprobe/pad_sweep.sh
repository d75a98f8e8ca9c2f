#!/usr/bin/env bash
# tools/sassprobe/pad_sweep.sh -- step time of libraries built with -DMAPC_LOOP_PAD=k (k instructions of padding in front
# of the hot loop: the same 31 instructions at every 16-byte alignment mod 256), one line per k; k = 0 is the in-tree build.
cd "$(dirname "$0")/../.."
ARGS="${@:-262144 8}"
echo "pad0 $(python tools/sassprobe/time_lib.py $ARGS)"
for k in $(seq 1 15); do
  [ -f tools/sassprobe/pad_libs/pad$k/libmapc.so ] || continue
  echo "pad$k $(MAPC_LIB_PATH=$PWD/tools/sassprobe/pad_libs/pad$k/libmapc.so python tools/sassprobe/time_lib.py $ARGS)"
done
