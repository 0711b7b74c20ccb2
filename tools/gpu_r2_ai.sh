#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2ai.log
: > $L
for it in 3 4; do
VITOCM_TAIL_DEBUG=16 VITOCM_MLP_TL_ITEM=$it timeout 200 python tools/tail_timeline.py 175 2 1 2>&1 | grep -v "QKV chunks" >> $L
done
cat $L
