#!/bin/bash
# round 2, call L: column-strip stitch kernel + block-parallel Otsu: post / full-size tests, bench, launch times of the post kernels
mkdir -p gpurun_out
L=gpurun_out/r2l.log
: > $L
timeout 900 python -m pytest tests/test_gpu_post.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:" | head -40 >> $L
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
tail -3 gpurun_out/r2l_bench.err >> $L
python - >> $L <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2l_bench.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("bench", d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), kc, d["clocks"])
    print("hbm", d.get("roofline_hbm"))
except Exception as e:
    print("bench ERR", repr(e))
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:stitch|otsu|head_mean" -s 12 -c 12 --csv --log-file gpurun_out/r2l_post_launches.csv python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2l_ncu.log 2>&1
python - >> $L <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2l_post_launches.csv")))
hdr = None
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        print(d["Kernel Name"][:40], d["Metric Name"], d["Metric Value"], d["Metric Unit"])
PY
echo "=== done" >> $L
cat $L
