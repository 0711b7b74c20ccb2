#!/bin/bash
# round 2, call I: fused MLP order swap at the item start, stagger sweep, timeline, bench
mkdir -p gpurun_out
L=gpurun_out/r2i.log
: > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -x -k "mlp_fused" 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert" | head -30 >> $L
echo "=== mlp bench (stagger sweep)" >> $L
for sg in 0 20000 40000 60000; do
  VITOCM_MLP_STAGGER=$sg VITOCM_FUSE_MLP=2 PRECISION=2 timeout 120 python tools/mlp_bench.py 2>&1 | tail -2 | head -1 >> $L
done
echo "=== timeline" >> $L
VITOCM_FUSE_MLP=2 PRECISION=2 VITOCM_MLP_DEBUG=0 timeout 120 python tools/mlp_timeline.py 2>&1 | tail -16 >> $L
echo "=== bench" >> $L
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
tail -3 gpurun_out/r2i_bench.err >> $L
python - >> $L <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2i_bench.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("bench", d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"])
except Exception as e:
    print("bench ERR", repr(e))
PY
echo "=== done" >> $L
cat $L
