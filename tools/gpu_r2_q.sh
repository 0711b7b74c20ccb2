#!/bin/bash
# round 2, call Q: block-tail kernel (proj + LN2 + MLP + next LN in one kernel): kernel tests, timeline, suite, bench A/B
mkdir -p gpurun_out
L=gpurun_out/r2q.log
: > $L
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -x -k "block_tail" 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert|timeout" | head -30 >> $L
echo "=== timeline" >> $L
timeout 120 python tools/tail_timeline.py 175 2 >> $L 2>&1
if [ "$1" != "quick" ]; then
echo "=== suite" >> $L
timeout 1200 python -m pytest tests -m gpu -q --no-header -x 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:" | head -40 >> $L
echo "=== bench" >> $L
for ft in 1 0; do
VITOCM_FUSE_TAIL=$ft timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2q_bench_$ft.json 2> gpurun_out/r2q_bench_$ft.err
tail -3 gpurun_out/r2q_bench_$ft.err >> $L
python - $ft >> $L <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2q_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("bench tail=" + sys.argv[1], d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"])
except Exception as e:
    print("bench ERR", repr(e))
PY
done
fi
echo "=== done" >> $L
cat $L
