#!/bin/bash
# round 2, call X: quad attention tail items grouped with their pairs, compile-time lane masks
mkdir -p gpurun_out
L=gpurun_out/r2x.log
: > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fp16.py -m gpu -q --no-header -x -k "attention" 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert" | head -30 >> $L
for pk in 2 4; do
  VITOCM_ATTN_QUAD_PACK=$pk TILES=1225 TOKENS=785 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad_pack=$pk /" >> $L
  VITOCM_ATTN_QUAD_PACK=$pk TILES=175 TOKENS=785 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad_pack=$pk /" >> $L
done
VITOCM_ATTN_QUAD_TAILS=0 TILES=1225 TOKENS=785 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad_tails=0 /" >> $L
TILES=175 TOKENS=820 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 >> $L
echo "== timeline, tail item (item 12 of pipeline 0: group 592/13...), pack 2" >> $L
VITOCM_ATTN_TL_ITEM=0 timeout 120 python tools/attn_quad_timeline.py 175 6 785 2>&1 | sed -n 1,4p >> $L
echo "=== suite" >> $L
timeout 1200 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:" | head -40 >> $L
echo "=== bench" >> $L
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
tail -3 gpurun_out/r2x_bench.err >> $L
python - >> $L <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r2x_bench.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("bench", d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"])
except Exception as e:
    print("bench ERR", repr(e))
PY
echo "=== done" >> $L
cat $L
