#!/bin/bash
# GPU box: forward attention with packed tail items -- correctness, then timing with and without packing (VITOCM_ATTN_PACK)
mkdir -p gpurun_out
: > gpurun_out/pack.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -k "attention" 2>&1 | grep -E "passed|failed|FAILED|vitocm:" | head -30 >> gpurun_out/pack.log
for pk in 0 1; do
  for t in 32 175; do
    VITOCM_ATTN_PACK=$pk TILES=$t timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/pack=$pk /" >> gpurun_out/pack.log
  done
done
cat gpurun_out/pack.log
