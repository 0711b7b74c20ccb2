#!/bin/bash
# 8-GPU box: the scaling bench lines (segmentation weak scaling + MIM data parallel), launched as the driver does
mkdir -p gpurun_out
N=${1:-8}
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_seg_n$N.json 2> gpurun_out/bench_seg_n$N.err
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --workload mim_train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mim_n$N.json 2> gpurun_out/bench_mim_n$N.err
tail -c 600 gpurun_out/bench_seg_n$N.json; echo; tail -c 400 gpurun_out/bench_mim_n$N.json; echo; tail -3 gpurun_out/bench_seg_n$N.err; tail -3 gpurun_out/bench_mim_n$N.err
