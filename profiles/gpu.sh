#!/bin/bash
# Build in-tree (the .so travels with the snapshot), then run the given command on the GPU box.
# usage: profiles/gpu.sh <timeout-seconds> [--gpus N] -- '<command>'
set -e
cd "$(dirname "$0")/.."
python rabitq-ann-search_b200/build.py > /dev/null
make -C oracle port > /dev/null
T=$1; shift
exec /usr/local/graft/bin/gpurun --timeout "$T" "$@"
