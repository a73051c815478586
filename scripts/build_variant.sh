#!/bin/bash
# Build an A/B variant of the library: scripts/build_variant.sh <name> "<nvcc -D flags>"  -> gpurun_in/<name>.so
# ("default" rebuilds the in-tree library)
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_in
if [ "$1" == "default" ]; then
  python -c "import sys; sys.path.insert(0,'.'); from boxer_b200 import _native; _native.build(force=True)"
else
  BOXER_B200_LIB=$PWD/gpurun_in/$1.so BOXER_B200_NVCC_EXTRA="$2" python -c "import sys; sys.path.insert(0,'.'); from boxer_b200 import _native; _native.build(force=True)"
fi
