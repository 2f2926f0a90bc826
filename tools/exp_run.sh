#!/bin/bash
# Development helper: (re)build the product and experiment libraries from the repo root, then hand a command to gpurun.
#   tools/exp_run.sh [gpurun timeout seconds] '<command run on the GPU box>'
set -e
cd "$(dirname "$0")/.."
T=1200
if [[ "$1" =~ ^[0-9]+$ ]]; then T=$1; shift; fi
python -c "import b200ctc; b200ctc._lib.load()"
B200CTC_EXPERIMENT=1 python -c "import b200ctc; b200ctc._lib.load()"
exec gpurun --timeout "$T" -- "$1"
