#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2ak.log
: > $L
VITOCM_TAIL_DEBUG=0 VITOCM_MLP_TL_ITEM=0 timeout 60 python tools/tail_timeline.py 1 2 0 2>&1 | grep "timeout" | sort | uniq -c | head -30 >> $L
cat $L
