#!/bin/bash
# round 2, call AH: training forward attention through the four-pipeline kernel (LSE rows): training tests, attention tests, MIM bench A/B
mkdir -p gpurun_out
L=gpurun_out/r2ah.log
: > $L
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_kernels.py tests/test_gpu_kernels.py -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert " | head -30 >> $L
for qd in 1 0; do
VITOCM_ATTN_QUAD=$qd timeout 600 python bench.py --workload mim_train --steps 8 --warmup 3 --batch-per-gpu 32 --no-cpu-baseline > gpurun_out/r2ah_mim_$qd.json 2> gpurun_out/r2ah_mim_$qd.err
tail -2 gpurun_out/r2ah_mim_$qd.err >> $L
python - $qd >> $L <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2ah_mim_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 3) for k, v in d["kernel_classes"].items()}
    print("mim quad=" + sys.argv[1], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 3), "ms", kc)
except Exception as e:
    print("bench ERR", repr(e))
PY
done
cat $L
