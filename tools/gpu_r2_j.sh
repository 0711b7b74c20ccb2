#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2j.log
: > $L
for dbg in 0 8; do
  VITOCM_FUSE_MLP=2 PRECISION=2 VITOCM_MLP_DEBUG=$dbg timeout 120 python tools/mlp_timeline.py 2>&1 | tail -16 >> $L
  VITOCM_MLP_DEBUG=$dbg VITOCM_FUSE_MLP=2 PRECISION=2 timeout 120 python tools/mlp_bench.py 2>&1 | tail -2 | head -1 >> $L
done
cat $L
