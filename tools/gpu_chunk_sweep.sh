#!/bin/bash
# GPU box: bench.py over chunk sizes (tiles per kernel launch = tiles per engine call)
mkdir -p gpurun_out
: > gpurun_out/chunk_sweep.log
for c in 175 245 153 205 123; do
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --chunk-tiles $c --tile-batch $c > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
  python - "$c" >> gpurun_out/chunk_sweep.log <<'PY'
import json, sys
c = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_c{c}.json").read().strip().splitlines()[-1])
    print("chunk", c, round(d["value"], 2), "MP/s", round(d["ms_per_step"], 2), "ms", {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}, d["clocks"]["sm_mhz"])
except Exception as e:
    print("chunk", c, "ERR", e)
PY
done
cat gpurun_out/chunk_sweep.log
