#!/bin/bash
# GPU box: forward attention with packed tail items -- correctness (both mask polarities on the first bring-up), then timing
mkdir -p gpurun_out
: > gpurun_out/pack.log
for inv in 0 1; do
  echo "=== mask_inv=$inv" >> gpurun_out/pack.log
  VITOCM_ATTN_MASK_INV=$inv timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -k "attention" 2>&1 | grep -E "passed|failed|FAILED|vitocm:" | head -30 >> gpurun_out/pack.log
done
for pk in 0 1; do
  for t in 32 175; do
    VITOCM_ATTN_PACK=$pk TILES=$t timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/pack=$pk /" >> gpurun_out/pack.log
  done
done
VITOCM_ATTN_PACK=0 TILES=32 TOKENS=3137 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/pack=0 /" >> gpurun_out/pack.log
VITOCM_ATTN_PACK=1 TILES=32 TOKENS=3137 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/pack=1 /" >> gpurun_out/pack.log
cat gpurun_out/pack.log
