#!/bin/bash
# GPU box: A/B of two builds of the forward attention kernel: the current libvitocm.so first, then libvitocm_prev.so
mkdir -p gpurun_out
: > gpurun_out/ab.log
P=vit-ocm-wmsegmentation_b200
cp $P/libvitocm.so $P/libvitocm_new.so
for v in new prev new; do
  cp $P/libvitocm_$v.so $P/libvitocm.so
  TILES=4 TOKENS=12545 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/$v /" >> gpurun_out/ab.log
  TILES=64 TOKENS=1024 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/$v /" >> gpurun_out/ab.log
  TILES=175 TOKENS=785 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/$v /" >> gpurun_out/ab.log
done
cp $P/libvitocm_new.so $P/libvitocm.so
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -k "attention" 2>&1 | grep -E "passed|failed|FAILED|vitocm:" | head -10 >> gpurun_out/ab.log
cat gpurun_out/ab.log
