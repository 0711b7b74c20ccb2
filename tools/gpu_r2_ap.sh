#!/bin/bash
# round 2, call AP: five-coefficient GELU without the clamp: kernel tests, parity, timing, bench
mkdir -p gpurun_out
L=gpurun_out/r2ap.log
: > $L
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fp16.py tests/test_gpu_parity.py -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert " | head -30 >> $L
VITOCM_MLP_TL_ITEM=4 timeout 200 python tools/tail_timeline.py 175 2 1 2>&1 | head -1 >> $L
VITOCM_MLP_TL_ITEM=20 timeout 200 python tools/tail_timeline.py 1225 2 1 2>&1 | head -1 >> $L
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2ap_bench.json 2> gpurun_out/r2ap_bench.err
python - >> $L <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2ap_bench.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("bench", d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"])
    print("mask_agreement", d.get("mask_agreement"))
except Exception as e:
    print("bench ERR", repr(e))
PY
cat $L
