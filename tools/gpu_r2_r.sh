#!/bin/bash
# round 2, call R: block-tail start stagger sweep at the bench's shard size
mkdir -p gpurun_out
L=gpurun_out/r2r.log
: > $L
for stg in 0 30000 60000 100000 150000; do
  echo "== stagger $stg" >> $L
  VITOCM_TAIL_STAGGER=$stg VITOCM_MLP_TL_ITEM=20 timeout 200 python tools/tail_timeline.py 1225 2 >> $L 2>&1
done
cat $L
