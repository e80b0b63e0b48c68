#!/usr/bin/env bash
# tools/round_check.sh [tag] -- everything the round-end driver looks at, in ONE gpurun call (1 GPU, ~4 minutes of box time):
#     gpurun --timeout 600 -- 'bash tools/round_check.sh r2a'
# smoke(), the -m gpu suite, the default bench line, the reference arm (short), the ncu launch list of the bench command and one
# full capture of the dominant kernel.  Results land in gpurun_out/<tag>_*; copy what should be judged into profiles/.
set -u
cd "$(dirname "$0")/.."
TAG=${1:-check}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py 2>$OUT/${TAG}_bench_n1.err >$OUT/${TAG}_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 2>$OUT/${TAG}_bench_ref.err >$OUT/${TAG}_bench_ref.json
python - <<PY
import json
for f in ("$OUT/${TAG}_bench_n1.json", "$OUT/${TAG}_bench_ref.json"):
    try:
        d = json.load(open(f))
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"),
              (d.get("particles") or {}).get("value"), ((d.get("particles") or {}).get("to_map") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-particles --no-extra-configs >$OUT/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shell_halos -s 1 -c 1 -o $OUT/${TAG}_shell_halos \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-particles --no-extra-configs >$OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | grep "${TAG}_"
