#!/usr/bin/env bash
# tools/ab_check.sh -- A/B of the headline kernel: the default library and every baryonforge_b200/variants/libbfg_*.so, one bench line
# each (device-resident step only), then the shell parity tests on the default library.  One gpurun call, 1 GPU.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-ab}
B="python bench.py --no-cpu-baseline --no-particles --no-e2e --steps 5"
echo "--- default"; $B 2>$OUT/${TAG}_default.err | tee $OUT/${TAG}_default.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d.get('roofline'))"
for v in baryonforge_b200/variants/libbfg_*.so; do
    [ -e "$v" ] || continue
    n=$(basename $v .so)
    echo "--- $n"; BFG_LIB=$PWD/$v $B 2>$OUT/${TAG}_$n.err | tee $OUT/${TAG}_$n.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d.get('roofline'))"
done
echo "--- mass-function catalogue, default"; $B --mass-function 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -15
