#!/bin/bash
# Build libb200ctc.so in-tree (same as __graft_entry__.build()), from any working directory.
set -e
cd "$(dirname "$0")/.."
python -c "
import importlib
pkg = importlib.import_module('chainer-speech-recognition_b200'); print(pkg.build())"
