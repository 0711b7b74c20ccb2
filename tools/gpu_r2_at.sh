#!/bin/bash
# round 2, call AT: whole GPU suite, smoke and the driver's default bench command on the folded block tail
mkdir -p gpurun_out
L=gpurun_out/r2at.log
: > $L
echo "=== suite" >> $L
timeout 1200 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|vitocm:" | head -30 >> $L
echo "=== smoke" >> $L
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 >> $L
echo "=== bench (driver's default command)" >> $L
( time timeout 900 python bench.py > gpurun_out/r2at_bench.json 2> gpurun_out/r2at_bench.err ) 2>> $L
python - >> $L <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2at_bench.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("bench", d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"])
    print("roofline", d["roofline"])
    print("step_tensor", d["step_tensor"])
    print("mask_agreement", d.get("mask_agreement"))
    print("cpu_baseline", d.get("cpu_baseline"))
except Exception as e:
    print("bench ERR", repr(e))
PY
echo "=== done" >> $L
cat $L
