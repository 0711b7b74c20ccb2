#!/bin/bash
# round 2, call C: fused MLP kernel (fc1 + GELU + fc2 + residual): kernel tests, micro-benchmark, parity suite, bench line
mkdir -p gpurun_out
L=gpurun_out/r2c.log
: > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -x -k "mlp_fused" 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert" | head -30 >> $L
echo "=== mlp bench" >> $L
for ew in 2 4; do for pr in 0 2; do
  VITOCM_FUSE_MLP=$ew PRECISION=$pr timeout 120 python tools/mlp_bench.py 2>&1 | tail -2 >> $L
done; done
echo "=== suite" >> $L
timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|agreement|rel err" | head -40 >> $L
echo "=== bench" >> $L
for fm in 2 4 0; do
VITOCM_FUSE_MLP=$fm timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2c_bench_$fm.json 2> gpurun_out/r2c_bench_$fm.err
tail -3 gpurun_out/r2c_bench_$fm.err >> $L
python - $fm >> $L <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2c_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("bench fuse_mlp=" + sys.argv[1], d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"])
except Exception as e:
    print("bench ERR", repr(e))
PY
done
echo "=== done" >> $L
cat $L
