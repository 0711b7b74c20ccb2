#!/bin/bash
# round 2, call AL: final block-tail kernel (plain remote arrives): suite, ncu capture of the kernel, default bench
mkdir -p gpurun_out
L=gpurun_out/r2al.log
: > $L
echo "=== suite" >> $L
timeout 1500 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:" | head -40 >> $L
echo "=== smoke" >> $L
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> $L
echo "=== ncu (block tail, 175-tile launch)" >> $L
P="python tools/profile_step.py 175 vit_small fp16"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:block_tail" -s 3 -c 1 -o /tmp/prof_tail -f $P > gpurun_out/ncu4_tail.log 2>&1
python tools/ncu_summary.py /tmp/prof_tail.ncu-rep >> $L 2>&1
ncu -i /tmp/prof_tail.ncu-rep --page details --csv > gpurun_out/prof4_tail_details.csv 2>/dev/null
echo "=== bench (driver's default command)" >> $L
( time timeout 1200 python bench.py > gpurun_out/r2al_bench.json 2> gpurun_out/r2al_bench.err ) 2>&1 | tail -3 >> $L
tail -3 gpurun_out/r2al_bench.err >> $L
python - >> $L <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2al_bench.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("bench", d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"])
    print("roofline", d["roofline"]); print("step_tensor", d["step_tensor"]); print("mask_agreement", d.get("mask_agreement"))
except Exception as e:
    print("bench ERR", repr(e))
PY
echo "=== done" >> $L
cat $L
