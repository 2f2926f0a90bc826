#!/bin/bash
# Build, and only if the build succeeds run the given command on a B200 through gpurun.
#   tools/gpu.sh [gpurun timeout seconds] '<command>'
cd "$(dirname "$0")/.."
T=900
if [[ "$1" =~ ^[0-9]+$ ]]; then T=$1; shift; fi
./tools/build.sh > /tmp/build.log 2>&1 || { grep -E "error|Error" /tmp/build.log | head -20; echo "BUILD FAILED"; exit 1; }
exec gpurun --timeout "$T" -- "$1"
