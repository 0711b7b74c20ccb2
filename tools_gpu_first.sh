#!/bin/bash
# first bring-up run on the GPU box: each group in its own process so that a device fault in one
# kernel does not poison the others
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { echo "=== $*" >> gpurun_out/t.log; timeout 600 python -m pytest "$@" -m gpu -q --no-header -rA --maxfail=20 2>&1 | tail -60 >> gpurun_out/t.log; }
: > gpurun_out/t.log
run tests/test_gpu_kernels.py -k "gemm or layernorm or launch"
run tests/test_gpu_kernels.py -k "attention"
run tests/test_gpu_post.py
run tests/test_gpu_parity.py
timeout 300 python __graft_entry__.py smoke >> gpurun_out/t.log 2>&1
echo "=== done" >> gpurun_out/t.log
