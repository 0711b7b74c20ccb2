#!/bin/bash
# round 2, call S: block-tail boundary phases: finer stamps, ablations (no fp32 row loads / no stores)
mkdir -p gpurun_out
L=gpurun_out/r2s.log
: > $L
for dbg in 0 2 4 6; do
  echo "== debug $dbg" >> $L
  VITOCM_TAIL_DEBUG=$dbg VITOCM_MLP_TL_ITEM=3 timeout 200 python tools/tail_timeline.py 175 2 >> $L 2>&1
done
cat $L
