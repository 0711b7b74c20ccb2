#!/bin/bash
# round 2, call U: accuracy of the block tail (TMEM residual accumulate) vs the separate kernels; parity test and mask agreement A/B
mkdir -p gpurun_out
L=gpurun_out/r2u.log
: > $L
timeout 300 python tools/tail_accuracy.py >> $L 2>&1
for ft in 1 0; do
echo "=== parity test, VITOCM_FUSE_TAIL=$ft" >> $L
VITOCM_FUSE_TAIL=$ft timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -s -k "test_vits8_tile_config1" 2>&1 | grep -E "rel err|agreement|passed|failed" >> $L
done
for ft in 1 0; do
echo "=== bench extras, VITOCM_FUSE_TAIL=$ft" >> $L
VITOCM_FUSE_TAIL=$ft timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2u_bench_$ft.json 2> gpurun_out/r2u_bench_$ft.err
python - $ft >> $L <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2u_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 1), "MP/s", json.dumps(d.get("mask_agreement")), json.dumps(d.get("precision_modes"))[:600])
except Exception as e:
    print("bench ERR", repr(e))
PY
done
cat $L
