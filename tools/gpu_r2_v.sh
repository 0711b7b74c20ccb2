#!/bin/bash
# round 2, call V: quad attention with two-pair tail items; restructured parity test
mkdir -p gpurun_out
L=gpurun_out/r2v.log
: > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fp16.py -m gpu -q --no-header -x -k "attention" 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert" | head -30 >> $L
echo "=== attention bench" >> $L
for qt in 1 0; do
  VITOCM_ATTN_QUAD_TAILS=$qt TILES=175 TOKENS=785 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad_tails=$qt /" >> $L
  VITOCM_ATTN_QUAD_TAILS=$qt TILES=1225 TOKENS=785 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad_tails=$qt /" >> $L
done
VITOCM_ATTN_QUAD_TAILS=1 TILES=1225 TOKENS=768 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad_tails=1 /" >> $L
VITOCM_ATTN_QUAD_TAILS=1 TILES=175 TOKENS=820 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad_tails=1 /" >> $L
echo "=== parity" >> $L
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -s -k "test_vits8_tile_config1 or over_tiles" 2>&1 | grep -E "rel err|agreement|passed|failed|median|Error" >> $L
echo "=== suite" >> $L
timeout 1200 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:" | head -40 >> $L
echo "=== done" >> $L
cat $L
