#!/bin/bash
# build the CUDA library + reference tools locally, then run a command on the GPU box:  tools/grun.sh [--gpus N] <timeout> '<cmd>'
set -e
cd "$(dirname "$0")/.."
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
make -s -j8 -C dspfun_b200/csrc 2>&1 | grep -E "error" && exit 1
make -s -C oracle reftools > /dev/null 2>&1 || true
/usr/local/graft/bin/gpurun $GP --timeout "$1" -- "$2"
