#!/bin/bash
# Build everything locally (the GPU box only uses prebuilt in-tree .so files), then run a command on the B200 box.
# usage: tools/gpu.sh [--timeout S] [--gpus N] -- '<command>'
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" | tail -1
exec /usr/local/graft/bin/gpurun "$@"
