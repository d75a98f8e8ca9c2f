#!/usr/bin/env bash
# variant_pad_sweep.sh "P T V" ... -- step time (N = 262,144) of forced launch shapes / variants under every 16-byte code
# alignment: pad0 = the in-tree build, pad_libs/padK = the same source with -DMAPC_LOOP_PAD=K.
cd "$(dirname "$0")/../.."
for spec in "$@"; do set -- $spec
  for k in 0 1 2 3 4 5 6 7; do
    lib=$PWD/multi-adapter-particles_b200/lib/libmapc.so; [ $k -gt 0 ] && lib=$PWD/tools/sassprobe/pad_libs/pad$k/libmapc.so
    [ -f $lib ] || continue
    echo "P=$1 T=$2 variant=$3 pad$k $(MAPC_LIB_PATH=$lib MAPC_PLAN_PAIRS=$1 MAPC_PLAN_THREADS=$2 MAPC_SHAPE_VARIANT=$3 python tools/sassprobe/time_lib.py 262144 5)"
  done
done
