#!/bin/bash
# build everything in-tree (the .so files travel with the snapshot), then run a script on the B200 box
set -e
cd "$(dirname "$0")/.."
make -C gaussian-splatting-toolkit_b200/csrc -j8 -s 2>&1 | grep -v "warning\|deprecated" || true
make -C oracle -s
T=${GRUN_TIMEOUT:-1500}
/usr/local/graft/bin/gpurun --timeout $T ${GRUN_GPUS:+--gpus $GRUN_GPUS} -- "$@"
