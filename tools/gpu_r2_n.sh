#!/bin/bash
# N-GPU box: the driver's scaling line (default bench with extras: multi_gpu_bitwise_equal, cfg3 strong scaling, mim_train)
mkdir -p gpurun_out
N=${1:-2}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err ) 2>&1 | tail -3
tail -5 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
print("bench", d["n_gpus"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", d["e2e"], d["clocks"], d["config"])
print("bitwise", d.get("multi_gpu_bitwise_equal"))
print("cfg3", d.get("cfg3")); print("mim", d.get("mim_train"))
PY
