#!/bin/bash
# GPU box: correctness (each group in its own process), smoke, short bench, ncu launch list + full captures
mkdir -p gpurun_out
: > gpurun_out/t.log
i=0
run() { i=$((i+1)); echo "=== $*" >> gpurun_out/t.log; timeout 600 python -m pytest "$@" -m gpu -q --no-header --maxfail=20 > gpurun_out/pytest_$i.log 2>&1; grep -E "vitocm: mbarrier|^E  |passed|failed" gpurun_out/pytest_$i.log | head -40 >> gpurun_out/t.log; }
run tests/test_gpu_kernels.py -k "gemm or layernorm or launch"
run tests/test_gpu_kernels.py -k "attention"
run tests/test_gpu_post.py
run tests/test_gpu_parity.py -s
timeout 300 python __graft_entry__.py smoke >> gpurun_out/t.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "=== bench done" >> gpurun_out/t.log
if [ "$1" = "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 32 > gpurun_out/ncu1.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -o gpurun_out/prof_attn -f python tools/profile_step.py 32 > gpurun_out/ncu2.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 8 -c 4 -o gpurun_out/prof_gemm -f python tools/profile_step.py 32 > gpurun_out/ncu3.log 2>&1
fi
echo "=== done" >> gpurun_out/t.log
