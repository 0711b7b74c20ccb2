#!/bin/bash
# whole GPU suite, smoke, full default bench line (extras on), reference arm
mkdir -p gpurun_out
L=gpurun_out/r2g.log
: > $L
timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:" | head -40 >> $L
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 >> $L
echo "=== bench" >> $L
( time timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err ) 2>> $L
tail -5 gpurun_out/r2g_bench.err >> $L
python - >> $L <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
    kc = {k: round(v["ms"], 2) for k, v in d["kernel_classes"].items()}
    print("bench", d["dtype"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), d["e2e"].get("host_masks_match_device"), kc, d["clocks"], "launches", d["gpu_launches"])
    print("roofline", d["roofline"]); print("hbm", d.get("roofline_hbm")); print("step_tensor", d.get("step_tensor"))
    print("mask_agreement", json.dumps(d.get("mask_agreement")))
    for k, v in (d.get("precision_modes") or {}).items():
        print("mode", k, round(v["value"], 1), v["mask_agreement_mosaic"], v["mask_agreement_tiles_224"])
    c3 = d.get("cfg3"); print("cfg3", {k: c3[k] for k in ("value", "ms_per_step", "frac_of_peak", "e2e", "mask_sha256_16", "mask_sha256_16_host_path")} if c3 else None)
    mt = d.get("mim_train"); print("mim", {k: mt[k] for k in ("value", "ms_per_step", "e2e")} if mt else None)
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench ERR", repr(e))
PY
echo "=== done" >> $L
cat $L
