#!/bin/bash
# GPU box: ncu captures of the hot kernels on one 32-tile chunk (run only after gpu_check.sh exited 0)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 64 > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -o gpurun_out/prof_attn -f python tools/profile_step.py 64 > gpurun_out/ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 9 -c 5 -o gpurun_out/prof_gemm -f python tools/profile_step.py 64 > gpurun_out/ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:layernorm -s 4 -c 1 -o gpurun_out/prof_ln -f python tools/profile_step.py 64 > gpurun_out/ncu4.log 2>&1
echo done
