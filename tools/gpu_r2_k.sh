#!/bin/bash
# round 2, call K: patch-embedding epilogue through TMA boxes: whole GPU suite, bench A/B
mkdir -p gpurun_out
L=gpurun_out/r2k.log
: > $L
timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:" | head -40 >> $L
for pt in 1 0; do
VITOCM_PATCH_TMA=$pt timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2k_bench_$pt.json 2> gpurun_out/r2k_bench_$pt.err
tail -3 gpurun_out/r2k_bench_$pt.err >> $L
python - $pt >> $L <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2k_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("bench patch_tma=" + sys.argv[1], d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"])
except Exception as e:
    print("bench ERR", repr(e))
PY
done
echo "=== done" >> $L
cat $L
