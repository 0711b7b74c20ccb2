#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2o.log
: > $L
for qd in 1 0; do
  VITOCM_ATTN_QUAD=$qd TILES=175 TOKENS=768 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad=$qd /" >> $L
  VITOCM_ATTN_QUAD=$qd TILES=175 TOKENS=785 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad=$qd /" >> $L
done
VITOCM_ATTN_TL_ITEM=3 timeout 120 python tools/attn_quad_timeline.py 175 6 785 >> $L 2>&1
cat $L
