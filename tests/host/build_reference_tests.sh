#!/usr/bin/env bash
# Builds the reference's OWN test suite (tests/00-fft.cpp, tests/01-real.cpp + its harness) against
# THIS repository's include/signalsmith-fft.h and libssfft.so -- the drop-in source-compatibility check
# (SURVEY.md section 8f row 2).  The reference sources are compiled from where they lie: they are only
# copied into a throw-away temp dir (their `#include "../signalsmith-fft.h"` must resolve to our header),
# never into the repository.  Output: tests/host/_build/reference_tests (git-ignored, travels to the GPU
# box, where tests/test_gpu_reference_suite.py runs it).
set -euo pipefail
REF=${REFERENCE_DIR:-/root/reference}
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
OUT="$ROOT/tests/host/_build"
[ -f "$REF/tests/00-fft.cpp" ] || { echo "reference not present at $REF: keeping prebuilt binary if any"; exit 0; }
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/tests" "$OUT"
cp "$REF"/tests/*.cpp "$REF"/tests/*.h "$TMP/tests/"
cp -r "$REF/common" "$TMP/common"
# the tests include "../signalsmith-fft.h": give them OUR front end under that name
cp "$ROOT/include/signalsmith-fft.h" "$ROOT/include/ssfft.h" "$TMP/"
g++ -std=c++11 -Wall -Wextra -O1 "$TMP/common/test/main.cpp" -I "$TMP/common" -I "$TMP/tests" \
    "$TMP"/tests/*.cpp -o "$OUT/reference_tests" \
    -L"$ROOT/fft_b200" -lssfft -Wl,-rpath,'$ORIGIN/../../../fft_b200'
echo "built $OUT/reference_tests"
