#!/bin/bash
# round 2, call AM: the driver's bench commands (default flags; --steps 20 --warmup 5), twice, to see the spread of the timed steps
mkdir -p gpurun_out
L=gpurun_out/r2am.log
: > $L
for run in 1 2; do
( time timeout 1200 python bench.py --no-extras > gpurun_out/r2am_bench_$run.json 2> gpurun_out/r2am_bench_$run.err ) 2>&1 | grep real >> $L
python - $run >> $L <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2am_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), "steps", d["steps"], "warmup", d["warmup"], "+", d["extra_warmup_steps"], d["step_ms_rank0"], d["clocks"])
    print("cpu_baseline", d.get("cpu_baseline"))
except Exception as e:
    print("bench ERR", repr(e))
PY
done
( time timeout 1200 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2am_bench_3.json 2> gpurun_out/r2am_bench_3.err ) 2>&1 | grep real >> $L
python - 3 >> $L <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r2am_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("bench", round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), "steps", d["steps"], "warmup", d["warmup"], "+", d["extra_warmup_steps"], d["step_ms_rank0"], d["clocks"])
PY
cat $L
