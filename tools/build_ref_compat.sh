#!/usr/bin/env bash
# Drop-in check: compile the reference's OWN callers -- its example driver and its two GoogleTest files -- UNCHANGED,
# from where they lie under /root/reference, against THIS repository's headers (include/) and library
# (cp-cals_b200/libcals.so), with tests/cpp/gtest_shim standing in for GoogleTest (not installed in this image).
# Outputs (git-ignored, shipped to the GPU box by gpurun like every built file):
#   cp-cals_b200/bin/ref_driver  cp-cals_b200/bin/ref_test_cals  cp-cals_b200/bin/ref_test_als
# No reference source is copied into the repository.  Where /root/reference is absent (GPU box) the prebuilt files stay.
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
REF="${CALS_REFERENCE_DIR:-/root/reference}"
OUT="$ROOT/cp-cals_b200/bin"
if [ ! -d "$REF/src" ]; then
  echo "build_ref_compat: $REF not present -- keeping prebuilt binaries" >&2
  exit 0
fi
mkdir -p "$OUT"
FLAGS=(-O2 -std=c++17 -w -DWITH_TIME=1 -I"$ROOT/include" -I"$ROOT/include/utils" -pthread)
LINK=(-L"$ROOT/cp-cals_b200" -lcals -lcals_b200 -Wl,-rpath,'$ORIGIN/..')
/usr/bin/g++ "${FLAGS[@]}" -o "$OUT/ref_driver" "$REF/src/examples/driver.cpp" "${LINK[@]}"
/usr/bin/g++ "${FLAGS[@]}" -I"$ROOT/tests/cpp/gtest_shim" -o "$OUT/ref_test_cals" "$REF/tests/cals/test_cals.cpp" "${LINK[@]}"
/usr/bin/g++ "${FLAGS[@]}" -I"$ROOT/tests/cpp/gtest_shim" -o "$OUT/ref_test_als" "$REF/tests/als/test_als.cpp" "${LINK[@]}"
echo "build_ref_compat: built ref_driver ref_test_cals ref_test_als in $OUT"
