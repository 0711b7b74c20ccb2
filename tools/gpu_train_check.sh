#!/bin/bash
# GPU box: training-step tests + MIM bench lines (batch 32 and 256 on one GPU)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_kernels.py -m gpu -q --no-header -s 2>&1 | grep -E "passed|failed|FAILED|step |ViT-S|Error|assert " | head -40 > gpurun_out/train_tests.log
timeout 600 python bench.py --workload mim_train --steps 5 --warmup 3 --batch-per-gpu 32 > gpurun_out/bench_mim_b32.json 2> gpurun_out/bench_mim_b32.err
timeout 900 python bench.py --workload mim_train --steps 3 --warmup 3 --batch-per-gpu 256 --no-cpu-baseline > gpurun_out/bench_mim_b256.json 2> gpurun_out/bench_mim_b256.err
echo done >> gpurun_out/train_tests.log
