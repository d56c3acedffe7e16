#!/bin/bash
# A/B timing of library variants built with build.build(lib=..., defines=[...]):
#   scripts/ab_libs.sh <case list> scratch_libs/libX.so scratch_libs/libY.so ...
# (B200DET_LIB makes the package load a variant instead of the in-tree libb200det.so)
cases=$1; shift
for lib in pytorch_object_detection_b200/libb200det.so "$@"; do
  echo "== $lib"
  B200DET_LIB=$lib python scripts/profile_kernels.py --time --only "$cases" 2>&1 | grep median
done
