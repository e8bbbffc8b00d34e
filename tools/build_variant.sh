#!/bin/bash
# Build an alternative libprometheus_gs.so with extra nvcc flags for A/B timing (tools/stage_bench.py).
#   tools/build_variant.sh NAME "-DPM_COMPOSITE_ILP2=0"   ->  gpurun_variants/NAME.so
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NAME="$1"; shift
TMP="$(mktemp -d)"
mkdir -p "$ROOT/gpurun_variants" "$TMP/prometheus_b200" "$TMP/include"
cp -r "$ROOT/prometheus_b200/csrc" "$TMP/prometheus_b200/csrc"
cp "$ROOT/include/"*.h "$TMP/include/"
rm -f "$TMP/prometheus_b200/csrc/"*.o
PATH=/usr/local/cuda/bin:$PATH make -C "$TMP/prometheus_b200/csrc" -j8 $MAKEVARS EXTRA="$*" OUT="$ROOT/gpurun_variants/$NAME.so" >/dev/null
rm -rf "$TMP"
echo "built gpurun_variants/$NAME.so"
