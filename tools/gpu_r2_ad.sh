#!/bin/bash
# round 2, call AD: full GPU suite + default bench line (with extras and the CPU baseline) in the final state
mkdir -p gpurun_out
L=gpurun_out/r2ad.log
: > $L
echo "=== suite" >> $L
timeout 1500 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:" | head -40 >> $L
echo "=== smoke" >> $L
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 >> $L
echo "=== bench (driver's default command)" >> $L
( time timeout 1200 python bench.py > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err ) 2>&1 | tail -3 >> $L
tail -3 gpurun_out/r2ad_bench.err >> $L
python - >> $L <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2ad_bench.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("bench", d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"])
    print("roofline", d["roofline"]); print("cpu_baseline", d.get("cpu_baseline")); print("mask_agreement", d.get("mask_agreement"))
    print("precision_modes", json.dumps(d.get("precision_modes"))[:1500]); print("mim_train", json.dumps(d.get("mim_train"))[:400])
except Exception as e:
    print("bench ERR", repr(e))
PY
echo "=== reference arm" >> $L
( time timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2ad_ref.json 2> gpurun_out/r2ad_ref.err ) 2>&1 | tail -3 >> $L
tail -c 700 gpurun_out/r2ad_ref.json >> $L
echo "=== done" >> $L
cat $L
