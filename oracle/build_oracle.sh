#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Builds the plain-C restatement into oracle/libtds_oracle.so.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
CC="${TDS_CC:-/usr/bin/gcc}"
$CC -std=c11 -O2 -fPIC -shared -Wall -Wno-unused-function -I"$HERE/../include" -I"$HERE" \
    "$HERE/tds_oracle.c" -o "$HERE/libtds_oracle.so" -lm
echo "built $HERE/libtds_oracle.so"
