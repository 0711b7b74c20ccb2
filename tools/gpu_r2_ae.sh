#!/bin/bash
# round 2, call AE: block tail + next block's QKV projection in one kernel: tests, timelines, suite, bench A/B
mkdir -p gpurun_out
L=gpurun_out/r2ae.log
: > $L
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -x -k "block_tail" 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert|timeout" | head -30 >> $L
echo "=== timeline (with QKV / without)" >> $L
VITOCM_MLP_TL_ITEM=3 timeout 200 python tools/tail_timeline.py 175 2 1 >> $L 2>&1
VITOCM_MLP_TL_ITEM=3 timeout 200 python tools/tail_timeline.py 175 2 0 2>&1 | head -1 >> $L
VITOCM_MLP_TL_ITEM=20 timeout 200 python tools/tail_timeline.py 1225 2 1 2>&1 | head -1 >> $L
if [ "$1" != "quick" ]; then
echo "=== suite" >> $L
timeout 1500 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:" | head -40 >> $L
echo "=== bench" >> $L
for fq in 1 0; do
VITOCM_FUSE_QKV=$fq timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2ae_bench_$fq.json 2> gpurun_out/r2ae_bench_$fq.err
tail -3 gpurun_out/r2ae_bench_$fq.err >> $L
python - $fq >> $L <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2ae_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("bench fuse_qkv=" + sys.argv[1], d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"], d["gpu_launches"])
except Exception as e:
    print("bench ERR", repr(e))
PY
done
fi
echo "=== done" >> $L
cat $L
