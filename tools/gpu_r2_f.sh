#!/bin/bash
# chunk-size sweep with the fused MLP kernel (balanced engine calls)
mkdir -p gpurun_out
L=gpurun_out/r2f.log
: > $L
for ct in 175 205 245 307 409 613 1225; do
timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 4 --chunk-tiles $ct --tile-batch $ct > gpurun_out/r2f_$ct.json 2> gpurun_out/r2f_$ct.err
tail -2 gpurun_out/r2f_$ct.err >> $L
python - $ct >> $L <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2f_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("chunk", sys.argv[1], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"]["sm_mhz"])
except Exception as e:
    print("bench ERR", sys.argv[1], repr(e))
PY
done
cat $L
