#!/usr/bin/env bash
# tools/staged_check.sh -- first GPU call of the next round: everything DESIGN.md section 8 lists as written-but-unmeasured, in
# ONE gpurun call (1 GPU).  Build the variants HERE first (the box has nvcc too, but box minutes are the scarce resource):
#     for c in 7 6 5; do bash tools/build_variant.sh unroll2_c$c "-DBFG_SHELL_UNROLL2 -DBFG_SHELL_MIN_CTAS=$c"; done
#     gpurun --timeout 900 -- 'bash tools/staged_check.sh > gpurun_out/staged_check.log 2>&1; tail -40 gpurun_out/staged_check.log'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
echo "=== gated parity tests (two-pass cell list, P(k) from the cell list, C_l kernels)"
BFG_TEST_EXPERIMENTAL=1 python -m pytest tests -m gpu -q -k "two_pass or cell_ordered or harmonics or kept_cell" 2>&1 | tail -15
echo "=== cell-list build: default vs BFG_CELL_SORT=2"
python tools/bench_configs.py --which c4 2>/dev/null | tee $OUT/staged_c4_default.json | cut -c1-400
BFG_CELL_SORT=2 python tools/bench_configs.py --which c4 2>/dev/null | tee $OUT/staged_c4_twopass.json | cut -c1-400
echo "=== headline kernel: default vs two-chain variants"
python bench.py --no-cpu-baseline --no-particles --no-e2e --steps 3 2>/dev/null | tee $OUT/staged_bench_default.json | cut -c1-200
for v in baryonforge_b200/variants/libbfg_unroll2_c*.so; do
    [ -e "$v" ] || continue
    echo "--- $v"
    BFG_LIB=$PWD/$v python bench.py --no-cpu-baseline --no-particles --no-e2e --steps 3 2>/dev/null \
        | tee $OUT/staged_bench_$(basename $v .so).json | cut -c1-200
    BFG_LIB=$PWD/$v python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
        -k "map_runners or config1 or shell_invariants or pipelined or full_size or query_disc" 2>&1 | tail -2
done
echo "=== C_l step timing"
timeout 300 python tools/bench_anafast.py 2>$OUT/staged_anafast.err | tee $OUT/staged_anafast.json | cut -c1-600
echo "=== compute-sanitizer"
SANITIZE_TIMEOUT=240 timeout 800 bash tools/sanitize.sh 2>&1 | tail -12
