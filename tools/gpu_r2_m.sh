#!/bin/bash
# round 2, call M: four-pipeline forward attention (attention_quad_sm100.cuh): attention tests, micro-benchmark A/B, suite, bench A/B
mkdir -p gpurun_out
L=gpurun_out/r2m.log
: > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fp16.py -m gpu -q --no-header -x -k "attention" 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert" | head -30 >> $L
echo "=== attention bench (175 tiles)" >> $L
for qd in 1 0; do for pr in 2 0; do
  VITOCM_ATTN_QUAD=$qd TILES=175 PRECISION=$pr timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad=$qd /" >> $L
done; done
VITOCM_ATTN_QUAD=1 TILES=32 TOKENS=3137 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad=1 /" >> $L
VITOCM_ATTN_QUAD=0 TILES=32 TOKENS=3137 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad=0 /" >> $L
if [ "$1" != "quick" ]; then
echo "=== suite" >> $L
timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:" | head -40 >> $L
echo "=== bench" >> $L
for qd in 1 0; do
VITOCM_ATTN_QUAD=$qd timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2m_bench_$qd.json 2> gpurun_out/r2m_bench_$qd.err
tail -3 gpurun_out/r2m_bench_$qd.err >> $L
python - $qd >> $L <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2m_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("bench quad=" + sys.argv[1], d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"])
except Exception as e:
    print("bench ERR", repr(e))
PY
done
fi
echo "=== done" >> $L
cat $L
