#!/bin/bash
# round 2, call AQ: LayerNorm affine parameters folded into W1 / Wqkv for the block tail: tests, same-box A/B of the bench step
mkdir -p gpurun_out
L=gpurun_out/r2aq.log
: > $L
timeout 900 python -m pytest tests/test_gpu_fp16.py tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q --no-header -s 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert |rows vs oracle" | head -40 >> $L
for rep in 0 1; do
  for fold in 0 1; do
    VITOCM_TAIL_FOLD=$fold timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2aq_bench_${fold}_${rep}.json 2> gpurun_out/r2aq_bench_${fold}_${rep}.err
    python - $fold $rep >> $L <<'PY'
import json, sys
fold, rep = sys.argv[1:3]
try:
    d = json.loads(open(f"gpurun_out/r2aq_bench_{fold}_{rep}.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("fold", fold, "rep", rep, d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"]["sm_mhz"])
except Exception as e:
    print("bench ERR", fold, rep, repr(e))
PY
  done
done
cat $L
