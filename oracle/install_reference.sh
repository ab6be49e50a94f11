#!/bin/bash
# Install the UNMODIFIED reference into baseline/_ref/ (git-ignored; it travels to the GPU box with gpurun).
# Run in the build container, where /root/reference exists.  No network: --no-index, --no-deps.
# The reference's setup.py lists `setup_requires=['git-python']`, which cannot be fetched offline; an empty
# dist-info directory on PYTHONPATH satisfies setuptools without touching the reference's files.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d)"
cp -r "$SRC" "$TMP/refcopy"                      # the build writes nasbench_asr/_dist_info.py into its source tree
mkdir -p "$TMP/fake/git_python-3.0.0.dist-info"
printf 'Metadata-Version: 2.1\nName: git-python\nVersion: 3.0.0\n' > "$TMP/fake/git_python-3.0.0.dist-info/METADATA"
touch "$TMP/fake/git_python-3.0.0.dist-info/RECORD"
rm -rf "$ROOT/baseline/_ref"
PYTHONPATH="$TMP/fake" python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
  --target "$ROOT/baseline/_ref" "$TMP/refcopy"
rm -rf "$TMP"
echo "installed: $ROOT/baseline/_ref"
