#!/bin/bash
# round 2, call AG: L2 prefetches moved close to their use: DRAM traffic of the block-tail kernel, timing, tests
mkdir -p gpurun_out
L=gpurun_out/r2ag.log
: > $L
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -x -k "block_tail" 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert|timeout" | head -30 >> $L
VITOCM_MLP_TL_ITEM=3 timeout 200 python tools/tail_timeline.py 175 2 1 2>&1 | head -1 >> $L
VITOCM_MLP_TL_ITEM=20 timeout 200 python tools/tail_timeline.py 1225 2 1 2>&1 | head -1 >> $L
P="python tools/profile_step.py 175 vit_small fp16"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:block_tail" -s 3 -c 1 -o /tmp/prof_tail -f $P > gpurun_out/ncu3_tail.log 2>&1
python tools/ncu_summary.py /tmp/prof_tail.ncu-rep >> $L 2>&1
ncu -i /tmp/prof_tail.ncu-rep --page details --csv > gpurun_out/prof3_tail_details.csv 2>/dev/null
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2ag_bench.json 2> gpurun_out/r2ag_bench.err
python - >> $L <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2ag_bench.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("bench", d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"], d["roofline"])
except Exception as e:
    print("bench ERR", repr(e))
PY
cat $L
