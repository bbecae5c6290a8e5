#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles the UNMODIFIED reference hot path, in place from
# /root/reference (read-only), into oracle/_ref/libtds_ref.so.  Nothing is copied from the
# reference; only the thin shims oracle/ref/*.cpp are ours.  The reference's own build system
# (cmake) is not used: the path compiles from its headers + vendored tinyxml2 alone.
# The resulting .so is git-ignored but travels to the GPU box with gpurun.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${TDS_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: $REF not present (GPU box?) - keeping prebuilt $OUT/libtds_ref.so" >&2
  [ -f "$OUT/libtds_ref.so" ] && exit 0 || exit 3
fi
CXX="${TDS_CXX:-/usr/bin/g++}"
FLAGS="-std=c++17 -O3 -march=x86-64-v3 -DNDEBUG -fPIC -fopenmp -w"
INC="-I$REF/src -I$REF/third_party/tinyxml2/include -I$REF/examples -I$HERE/../include"
$CXX $FLAGS $INC -c "$HERE/ref/ref_core.cpp" -o "$OUT/ref_core.o" &
$CXX $FLAGS $INC -c "$HERE/ref/ref_laikago.cpp" -o "$OUT/ref_laikago.o" &
$CXX $FLAGS $INC -c "$HERE/ref/ref_ant.cpp" -o "$OUT/ref_ant.o" &
$CXX $FLAGS $INC -c "$HERE/ref/ref_world.cpp" -o "$OUT/ref_world.o" &
$CXX $FLAGS $INC -c "$HERE/ref/ref_rigid.cpp" -o "$OUT/ref_rigid.o" &
$CXX $FLAGS -I"$REF/third_party/tinyxml2/include" -c "$REF/third_party/tinyxml2/tinyxml2.cpp" -o "$OUT/tinyxml2.o" &
wait
$CXX -shared -fopenmp -o "$OUT/libtds_ref.so" "$OUT/ref_core.o" "$OUT/ref_laikago.o" "$OUT/ref_ant.o" "$OUT/ref_world.o" "$OUT/ref_rigid.o" "$OUT/tinyxml2.o"
rm -f "$OUT"/*.o
echo "built $OUT/libtds_ref.so"
