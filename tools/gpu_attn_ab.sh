#!/bin/bash
# GPU box: A/B of two builds of the forward attention kernel: libvitocm_prev.so against the current libvitocm.so
mkdir -p gpurun_out
: > gpurun_out/ab.log
P=vit-ocm-wmsegmentation_b200
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train_kernels.py -m gpu -q --no-header -k "attention" 2>&1 | grep -E "passed|failed|FAILED|vitocm:" | head -30 >> gpurun_out/ab.log
cp $P/libvitocm.so $P/libvitocm_new.so
for rep in 1 2; do
for v in prev new; do
  cp $P/libvitocm_$v.so $P/libvitocm.so
  for t in 32 175; do
    TILES=$t timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/$v /" >> gpurun_out/ab.log
  done
done
done
for v in prev new; do
  cp $P/libvitocm_$v.so $P/libvitocm.so
  TILES=32 TOKENS=3137 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/$v /" >> gpurun_out/ab.log
  timeout 120 python tools/attn_timeline.py 175 6 785 > gpurun_out/timeline_$v.txt 2>&1
done
cp $P/libvitocm_new.so $P/libvitocm.so
cat gpurun_out/ab.log
for v in prev new; do echo "== $v"; sed -n 3,5p gpurun_out/timeline_$v.txt; tail -2 gpurun_out/timeline_$v.txt; done
