#!/bin/bash
# A/B timing of library variants: scripts/ab_libs.sh <case list> scratch_libs/libX.so ...   (restores the first at the end)
cases=$1; shift
keep=/tmp/lib_keep.so
cp pytorch_object_detection_b200/libb200det.so $keep
for lib in "$@"; do
  cp "$lib" pytorch_object_detection_b200/libb200det.so
  echo "== $lib"
  python scripts/profile_kernels.py --time --only "$cases" 2>&1 | grep median
done
cp $keep pytorch_object_detection_b200/libb200det.so
