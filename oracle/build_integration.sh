#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles tests/integration/stepper_check.cpp (the reference's VectorizedEnvironment driven through
# its CustomForwardDynamicsStepper plugin point by libtds_b200.so) against the reference headers in /root/reference.
# Output: tests/integration/stepper_check.bin (git-ignored, travels to the GPU box with gpurun).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$HERE/.."
REF="${TDS_REFERENCE_ROOT:-/root/reference}"
OUT="$ROOT/tests/integration/stepper_check.bin"
if [ ! -d "$REF/src" ]; then
  echo "build_integration.sh: $REF not present (GPU box?) - keeping prebuilt $OUT" >&2
  [ -f "$OUT" ] && exit 0 || exit 3
fi
CXX="${TDS_CXX:-/usr/bin/g++}"
LIBDIR="$ROOT/tiny-differentiable-simulator_b200"
$CXX -std=c++17 -O2 -DNDEBUG -fopenmp -w -I"$REF/src" -I"$REF/third_party/tinyxml2/include" -I"$REF/examples" -I"$REF" \
  -I"$ROOT/include" "$ROOT/tests/integration/stepper_check.cpp" "$REF/third_party/tinyxml2/tinyxml2.cpp" \
  -L"$LIBDIR" -ltds_b200 -Wl,-rpath,'$ORIGIN/../../tiny-differentiable-simulator_b200' -o "$OUT"
echo "built $OUT"
