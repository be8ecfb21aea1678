#!/bin/bash
# Stage the UNMODIFIED reference sources under baseline/_ref/ (git-ignored, never committed; it travels to the GPU
# box with the gpurun snapshot) so that tests/test_reference_dropin.py can run the reference's own modules on a
# GPU.  Nothing in the product imports from there.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
mkdir -p "$ROOT/baseline/_ref"
rm -rf "$ROOT/baseline/_ref/scripts" "$ROOT/baseline/_ref/config"
cp -r "$SRC/scripts" "$ROOT/baseline/_ref/scripts"
cp -r "$SRC/config" "$ROOT/baseline/_ref/config"
find "$ROOT/baseline/_ref" -name '__pycache__' -prune -exec rm -rf {} +
echo "staged $(find "$ROOT/baseline/_ref" -name '*.py' | wc -l) reference modules under baseline/_ref (git-ignored)"
