#!/bin/bash
# tools/sweep_r2.sh -- one GPU call: (1) ring threshold of the warp-per-halo kernel on both catalogues, (2) latitude chunks of the
# single-GPU end-to-end pipeline.  Prints one line per run.
cd "$(dirname "$0")/.."
common="--no-particles --no-extra-configs --no-cpu-baseline --steps 3"
show='import json,sys; d=json.loads(sys.stdin.read()); e=d.get("e2e") or {}; print("%s value %.4e step %.2f ms kernel %.2f ms e2e %s" % (sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], e.get("ms_per_step")))'
for cat in "" "--mass-function"; do
  for r in off 12 20 32 48 64; do
    if [ "$r" = off ]; then w=0; rr=64; else w=1; rr=$r; fi
    BFG_SHELL_WARP_KERNEL=$w BFG_SHELL_WARP_MAX_RINGS=$rr python bench.py $common --no-e2e $cat 2>/dev/null | python -c "$show" "warp=$r${cat:+ mass-function}"
  done
done
for k in 12 16 24; do
  BFG_PIPELINE_CHUNKS=$k python bench.py $common 2>/dev/null | python -c "$show" "chunks=$k"
done
