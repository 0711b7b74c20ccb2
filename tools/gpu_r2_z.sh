#!/bin/bash
# round 2, call Z: block tail with the hidden chunk read by one x64 TMEM load; suite; bench
mkdir -p gpurun_out
L=gpurun_out/r2z.log
: > $L
VITOCM_MLP_TL_ITEM=3 timeout 200 python tools/tail_timeline.py 175 2 >> $L 2>&1
VITOCM_MLP_TL_ITEM=3 timeout 200 python tools/tail_timeline.py 175 0 2>&1 | head -1 >> $L
VITOCM_MLP_TL_ITEM=20 timeout 200 python tools/tail_timeline.py 1225 2 2>&1 | head -1 >> $L
echo "=== suite" >> $L
timeout 1200 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:" | head -40 >> $L
echo "=== bench" >> $L
for ft in 1 0; do
VITOCM_FUSE_TAIL=$ft timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2z_bench_$ft.json 2> gpurun_out/r2z_bench_$ft.err
tail -3 gpurun_out/r2z_bench_$ft.err >> $L
python - $ft >> $L <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2z_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("bench tail=" + sys.argv[1], d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"])
except Exception as e:
    print("bench ERR", repr(e))
PY
done
echo "=== done" >> $L
cat $L
