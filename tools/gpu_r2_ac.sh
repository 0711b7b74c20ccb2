#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2ac.log
: > $L
for rep in 1 2; do for h in 1 0; do
  VITOCM_ATTN_HOIST=$h TILES=1225 TOKENS=785 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/hoist=$h /" >> $L
  VITOCM_ATTN_HOIST=$h TILES=1225 TOKENS=768 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/hoist=$h /" >> $L
done; done
VITOCM_ATTN_HOIST=0 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fp16.py -m gpu -q --no-header -x -k "attention" 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert" | head -30 >> $L
cat $L
