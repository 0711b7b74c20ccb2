#!/bin/bash
# GPU box: the whole GPU test-suite, smoke, both bench workloads, ncu launch lists and --set full captures
mkdir -p gpurun_out
: > gpurun_out/full.log
timeout 900 python -m pytest tests -m gpu -q --no-header -x 2>&1 | grep -E "passed|failed|FAILED|Error|assert |vitocm:" | head -30 >> gpurun_out/full.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 >> gpurun_out/full.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_seg.json 2> gpurun_out/bench_seg.err
timeout 600 python bench.py --workload mim_train --steps 5 --warmup 3 --batch-per-gpu 32 > gpurun_out/bench_mim_b32.json 2> gpurun_out/bench_mim_b32.err
timeout 600 python bench.py --workload mim_train --steps 3 --warmup 3 --batch-per-gpu 256 --no-cpu-baseline > gpurun_out/bench_mim_b256.json 2> gpurun_out/bench_mim_b256.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "=== bench done" >> gpurun_out/full.log
if [ "$1" = "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_train.csv python tools/profile_train.py 256 1 > gpurun_out/ncu_t1.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 2 -c 1 -o gpurun_out/prof_attn_bwd -f python tools/profile_train.py 64 1 > gpurun_out/ncu_t2.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad -s 6 -c 2 -o gpurun_out/prof_wgrad -f python tools/profile_train.py 64 1 > gpurun_out/ncu_t3.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_bwd -s 2 -c 1 -o gpurun_out/prof_ln_bwd -f python tools/profile_train.py 64 1 > gpurun_out/ncu_t4.log 2>&1
fi
echo "=== done" >> gpurun_out/full.log
