#!/usr/bin/env bash
# tools/build_variant.sh NAME "EXTRA_NVCC_FLAGS" -- an alternative build of libbfg_b200.so with compile-time switches, for A/B runs:
#     bash tools/build_variant.sh c6 "-DBFG_SHELL_MIN_CTAS=6";  bash tools/build_variant.sh nored "-DBFG_SHELL_NO_RED"
#     gpurun -- 'bash tools/ab_check.sh'      (one bench line per variant + the shell parity tests on the default library)
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
