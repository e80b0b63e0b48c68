#!/usr/bin/env bash
# tools/sanitize.sh -- compute-sanitizer passes over the small GPU parity cases (the reference is
# single-threaded, the GPU scatter-adds are not, so memcheck + racecheck belong to the parity suite).  Run on a GPU box:
#     gpurun --timeout 900 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1; tail -5 gpurun_out/sanitize.log'
# Each pass runs a handful of fixture-sized tests (the sanitizer slows kernels down 10-100x); the log ends with one
# "ERROR SUMMARY" line per pass.
set -u
cd "$(dirname "$0")/.."
CS=${COMPUTE_SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
SMALL='map_runners_match_reference_fixture or snapshot_matches_reference_fixture or process_to_map or pk_matches_notebook_fixture or folded_deposit_edge_cases or raw_record or small_angle or warp_per_halo'
for tool in memcheck racecheck initcheck; do
    echo "=== compute-sanitizer --tool $tool"
    timeout ${SANITIZE_TIMEOUT:-600} "$CS" --tool "$tool" --target-processes all --error-exitcode 9 \
        python -m pytest tests/test_gpu_parity.py tests/test_gpu_spectrum.py -m gpu -q -x -k "$SMALL" 2>&1 \
        | grep -E "ERROR SUMMARY|passed|failed|error|Error|=====" | tail -20
    echo "exit code of the $tool pass: ${PIPESTATUS[0]}"
done
