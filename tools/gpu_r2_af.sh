#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2af.log
: > $L
for dbg in 0 8 1 9; do
  echo "== VITOCM_TAIL_DEBUG=$dbg (1 = no MMAs, 8 = no weight loads)" >> $L
  VITOCM_TAIL_DEBUG=$dbg VITOCM_MLP_TL_ITEM=3 timeout 200 python tools/tail_timeline.py 175 2 1 2>&1 | grep -v "^ chunk\|ep1 steps" >> $L
done
cat $L
