#!/bin/bash
# Freeze a copy of the working tree under gpurun_stage/<name>/ so that a queued gpurun call runs exactly this state
# even if the tree is edited while the call waits for a GPU slot.  usage: tools/stage.sh <name>
set -e
cd "$(dirname "$0")/.."
rm -rf gpurun_stage/$1; mkdir -p gpurun_stage/$1
tar --exclude=.git --exclude=gpurun_out --exclude=gpurun_stage --exclude=build_variants/r1src --exclude=__pycache__ --exclude=.pytest_cache -cf - . | tar -xf - -C gpurun_stage/$1
echo "staged gpurun_stage/$1 ($(du -sh gpurun_stage/$1 | cut -f1))"
