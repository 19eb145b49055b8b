#!/bin/bash
# build the in-tree library (if stale), then run a command on a B200 box: tools/gpu.sh [--gpus N] <timeout-seconds> '<command>'
set -e
cd "$(dirname "$0")/.."
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
T=$1; shift
python -c "from circkit_b200 import build; build.build()" > /tmp/ck_build.log 2>&1 || { tail -30 /tmp/ck_build.log; exit 1; }
exec tools/gpurun_retry.sh $G --timeout "$T" -- "$@"
