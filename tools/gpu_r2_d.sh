#!/bin/bash
# fused MLP: kernel tests + timeline + micro-benchmark (short call)
mkdir -p gpurun_out
L=gpurun_out/r2d.log
: > $L
for cl in 4 2; do
VITOCM_FUSE_MLP=$cl timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -k "mlp_fused" 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert" | head -30 >> $L
done
echo "=== mlp timeline (cluster 4, multicast weights)" >> $L
VITOCM_FUSE_MLP=4 timeout 120 python tools/mlp_timeline.py 2>&1 | tail -16 >> $L
echo "=== mlp bench" >> $L
for cl in 4 2; do for sg in 40000 0; do
  VITOCM_FUSE_MLP=$cl VITOCM_MLP_STAGGER=$sg PRECISION=2 timeout 120 python tools/mlp_bench.py 2>&1 | tail -2 | head -1 >> $L
done; done
VITOCM_FUSE_MLP=4 VITOCM_MLP_STAGGER=80000 PRECISION=2 timeout 120 python tools/mlp_bench.py 2>&1 | tail -2 >> $L
echo "=== done" >> $L
cat $L
