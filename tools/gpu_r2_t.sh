#!/bin/bash
# round 2, call T: packed tail items inside the quad attention kernel + block tail with compile-time formats: tests, A/B, suite, bench
mkdir -p gpurun_out
L=gpurun_out/r2t.log
: > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fp16.py -m gpu -q --no-header -x -k "attention or block_tail" 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert" | head -30 >> $L
echo "=== attention bench (175 tiles)" >> $L
for qt in 1 0; do
  VITOCM_ATTN_QUAD_TAILS=$qt TILES=175 TOKENS=785 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad_tails=$qt /" >> $L
  VITOCM_ATTN_QUAD_TAILS=$qt TILES=1225 TOKENS=785 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad_tails=$qt /" >> $L
done
VITOCM_ATTN_QUAD_TAILS=1 TILES=175 TOKENS=800 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad_tails=1 /" >> $L
VITOCM_ATTN_QUAD_TAILS=1 TILES=175 TOKENS=820 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad_tails=1 /" >> $L
echo "=== suite" >> $L
timeout 1200 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:" | head -40 >> $L
echo "=== bench" >> $L
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
tail -3 gpurun_out/r2t_bench.err >> $L
python - >> $L <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r2t_bench.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("bench", d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"])
except Exception as e:
    print("bench ERR", repr(e))
PY
echo "=== done" >> $L
cat $L
