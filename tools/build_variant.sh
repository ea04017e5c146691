#!/bin/bash
# usage: tools/build_variant.sh <name> <nvcc extra flags...>   -> rs_pathtracing_b200/variants/<name>.so
# (the default librt_b200.so is rebuilt afterwards by the next plain build; objects of a variant are not reused)
set -e
N=$1; shift
cd "$(dirname "$0")/.."
mkdir -p rs_pathtracing_b200/variants
cp rs_pathtracing_b200/librt_b200.so /tmp/librt_b200.default.so 2>/dev/null || true
RT_B200_NVCC_EXTRA="$*" python -c "from rs_pathtracing_b200 import build as b; b.build_core()"
cp rs_pathtracing_b200/librt_b200.so rs_pathtracing_b200/variants/$N.so
echo "built rs_pathtracing_b200/variants/$N.so"
