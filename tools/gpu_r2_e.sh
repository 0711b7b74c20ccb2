#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2e.log
: > $L
for cl in 2 4; do for dbg in 0 1 2; do
VITOCM_MLP_DEBUG=$dbg VITOCM_FUSE_MLP=$cl timeout 120 python tools/mlp_timeline.py 2>&1 | tail -16 >> $L
done; done
cat $L
