#!/bin/bash
# round 2, call A: fp16 format + act-split MLP: new tests, the whole GPU suite, precision sweep, bench per precision
mkdir -p gpurun_out
L=gpurun_out/r2a.log
: > $L
timeout 600 python -m pytest tests/test_gpu_fp16.py -m gpu -q --no-header -x 2>&1 | tail -25 >> $L
echo "=== suite" >> $L
timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|vitocm:" | head -30 >> $L
echo "=== sweep" >> $L
timeout 900 python tools/precision_sweep.py --tiles 64 --mosaic 4096 --schedules fp32,bf16,fp16,fp16+mlp2,bf16+mlp2 --out gpurun_out/r2a_sweep.jsonl >> $L 2>&1
echo "=== bench" >> $L
for p in bf16 fp16 fp16+mlp2 fp32; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision $p > gpurun_out/r2a_bench_$p.json 2> gpurun_out/r2a_bench_$p.err
  python - "$p" >> $L <<'PY'
import json, sys
p = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2a_bench_{p}.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print(p, round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms", "e2e", round(d["e2e"]["value"], 1), kc, d["clocks"])
except Exception as e:
    print(p, "ERR", e, open(f"gpurun_out/r2a_bench_{p}.err").read()[-800:])
PY
done
echo "=== done" >> $L
cat $L
