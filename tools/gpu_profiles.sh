#!/bin/bash
# GPU box: ncu launch lists of the bench commands themselves + --set full captures of the dominant kernels (final state of the round)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1935 -c 700 --csv --log-file gpurun_out/launches_bench_seg.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_train.csv python tools/profile_train.py 256 1 > gpurun_out/ncu_b2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -o gpurun_out/prof_attn -f python tools/profile_step.py 64 > gpurun_out/ncu_b3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 2 -c 1 -o gpurun_out/prof_attn_bwd -f python tools/profile_train.py 64 1 > gpurun_out/ncu_b4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 9 -c 5 -o gpurun_out/prof_gemm -f python tools/profile_step.py 64 > gpurun_out/ncu_b5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad -s 6 -c 2 -o gpurun_out/prof_wgrad -f python tools/profile_train.py 64 1 > gpurun_out/ncu_b6.log 2>&1
echo done
