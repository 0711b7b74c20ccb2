#!/bin/bash
# fused MLP: kernel tests + timeline + micro-benchmark (short call)
mkdir -p gpurun_out
L=gpurun_out/r2d.log
: > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -k "mlp_fused" 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert" | head -30 >> $L
echo "=== mlp timeline (16 epilogue warps)" >> $L
VITOCM_FUSE_MLP=1 timeout 120 python tools/mlp_timeline.py 2>&1 | tail -16 >> $L
echo "=== mlp timeline (8 epilogue warps)" >> $L
VITOCM_FUSE_MLP=8 timeout 120 python tools/mlp_timeline.py 2>&1 | tail -16 >> $L
echo "=== mlp bench" >> $L
for ew in 1 8; do
  VITOCM_FUSE_MLP=$ew PRECISION=2 timeout 120 python tools/mlp_bench.py 2>&1 | tail -2 >> $L
done
echo "=== done" >> $L
cat $L
