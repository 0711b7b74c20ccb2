#!/bin/bash
# round 2, call P: quad attention with S_{j+1} issued right behind PV_j (no wait for PV_j's retirement in between): tests, A/B, timeline, bench
mkdir -p gpurun_out
L=gpurun_out/r2p.log
: > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fp16.py -m gpu -q --no-header -x -k "attention" 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert" | head -30 >> $L
echo "=== attention bench (175 tiles)" >> $L
for qd in 1 0; do
  VITOCM_ATTN_QUAD=$qd TILES=175 TOKENS=785 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad=$qd /" >> $L
done
VITOCM_ATTN_QUAD=1 TILES=175 TOKENS=768 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad=1 /" >> $L
VITOCM_ATTN_QUAD=1 TILES=32 TOKENS=3137 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad=1 /" >> $L
VITOCM_ATTN_TL_ITEM=3 timeout 120 python tools/attn_quad_timeline.py 175 6 785 2>&1 | head -16 >> $L
if [ "$1" != "quick" ]; then
echo "=== suite" >> $L
timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:" | head -40 >> $L
echo "=== bench" >> $L
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
tail -3 gpurun_out/r2p_bench.err >> $L
python - >> $L <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r2p_bench.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("bench", d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"])
except Exception as e:
    print("bench ERR", repr(e))
PY
fi
echo "=== done" >> $L
cat $L
