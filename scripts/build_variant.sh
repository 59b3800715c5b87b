#!/bin/bash
# Build libqmcnn_b200 with extra nvcc flags into a side library (A/B measurements; bench.py --lib <path>).
# Usage: scripts/build_variant.sh <out.so> "<extra nvcc flags>"
set -e
OUT=$(realpath -m "$1"); EXTRA="$2"
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
cp "$ROOT"/qmcnn_b200/csrc/*.cu "$ROOT"/qmcnn_b200/csrc/*.cuh "$ROOT"/qmcnn_b200/csrc/*.h "$ROOT"/qmcnn_b200/csrc/Makefile "$TMP"/
mkdir -p "$TMP/inc" && cp "$ROOT"/include/*.h "$TMP/inc/"
sed -i "s#-I../../include#-I$TMP/inc $EXTRA#; s#../../include/qmcnn_b200.h#$TMP/inc/qmcnn_b200.h#; s#^OUT := .*#OUT := $OUT#" "$TMP/Makefile"
make -C "$TMP" -j8 > "$TMP/build.log" 2>&1 || { tail -20 "$TMP/build.log"; exit 1; }
rm -rf "$TMP"
echo "built $OUT"
