#!/usr/bin/env bash
# tools/build_variant.sh NAME "EXTRA_NVCC_FLAGS" -- an alternative build of libbfg_b200.so with compile-time switches, for A/B runs:
#     bash tools/build_variant.sh unroll2_c6 "-DBFG_SHELL_UNROLL2 -DBFG_SHELL_MIN_CTAS=6"
#     gpurun -- 'for v in unroll2_c7 unroll2_c6 unroll2_c5; do BFG_LIB=$PWD/baryonforge_b200/variants/libbfg_$v.so \
#                python bench.py --no-cpu-baseline --no-particles --no-e2e --steps 3; done'
# The variants are git-ignored (*.so) but travel to the GPU box with the snapshot.  BFG_LIB selects the library (_lib.py).
set -eu
cd "$(dirname "$0")/.."
NAME=$1; shift
FLAGS=${1:-}
mkdir -p baryonforge_b200/variants
OUT=baryonforge_b200/variants/libbfg_${NAME}.so
# shellcheck disable=SC2086
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared $FLAGS \
    -o "$OUT" baryonforge_b200/csrc/*.cu
echo "$OUT"
