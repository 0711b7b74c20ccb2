#!/bin/bash
# GPU box: whole GPU suite, smoke, bench lines (segmentation + MIM), then the ncu launch list and the --set full capture of the
# forward attention kernel (final state of the round, packed tail items)
mkdir -p gpurun_out
: > gpurun_out/full.log
timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|vitocm:" | head -30 >> gpurun_out/full.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 >> gpurun_out/full.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_seg.json 2> gpurun_out/bench_seg.err
timeout 600 python bench.py --workload mim_train --steps 5 --warmup 3 --batch-per-gpu 32 --no-cpu-baseline > gpurun_out/bench_mim_b32.json 2> gpurun_out/bench_mim_b32.err
timeout 600 python bench.py --workload mim_train --steps 3 --warmup 3 --batch-per-gpu 256 --no-cpu-baseline > gpurun_out/bench_mim_b256.json 2> gpurun_out/bench_mim_b256.err
timeout 600 python bench.py --arch vit_base --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vitb.json 2> gpurun_out/bench_vitb.err
echo "=== bench done" >> gpurun_out/full.log
if [ "$1" = "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1935 -c 700 --csv --log-file gpurun_out/launches_bench_seg.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b1.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -o gpurun_out/prof_attn -f python tools/profile_step.py 64 > gpurun_out/ncu_b3.log 2>&1
fi
echo "=== done" >> gpurun_out/full.log
cat gpurun_out/full.log
for f in bench_seg bench_mim_b32 bench_mim_b256 bench_vitb; do python - "$f" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms", d.get("roofline", {}).get("kernel"), round(d.get("roofline", {}).get("achieved", 0), 1), d["clocks"])
except Exception as e:
    print(f, "ERR", e)
PY
done
