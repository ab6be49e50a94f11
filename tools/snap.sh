#!/bin/bash
# Freeze a consistent copy of the working tree under _snap/ (git-ignored, travels with gpurun) so that a queued GPU
# call runs exactly this state while editing continues:  tools/snap.sh && gpurun -- 'cd _snap && ...  (outputs: ../gpurun_out)'
set -e
cd "$(dirname "$0")/.."
rm -rf _snap && mkdir _snap
tar --exclude=./.git --exclude=./gpurun_out --exclude=./_snap --exclude='*.o' --exclude=__pycache__ --exclude=.pytest_cache -cf - . | tar -xf - -C _snap
ln -s ../gpurun_out _snap/gpurun_out
echo "snapshot: $(du -sh _snap | cut -f1)"
