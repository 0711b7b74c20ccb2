#!/bin/bash
# round 2, call AN: spread of the timed steps with the in-process NVML sampler (six runs of the default command without extras)
mkdir -p gpurun_out
L=gpurun_out/r2an.log
: > $L
for run in 1 2 3 4 5 6; do
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2an_bench_$run.json 2> gpurun_out/r2an_bench_$run.err
python - $run >> $L <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2an_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), "+", d["extra_warmup_steps"], d["step_ms_rank0"], d["clocks"])
except Exception as e:
    print("bench ERR", repr(e)); print(open(f"gpurun_out/r2an_bench_{sys.argv[1]}.err").read()[-800:])
PY
done
cat $L
