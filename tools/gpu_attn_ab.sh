#!/bin/bash
# GPU box: A/B of two builds of the forward attention kernel (current libvitocm.so, then libvitocm_prev.so), then the whole GPU suite,
# smoke and the bench line with the current build
mkdir -p gpurun_out
: > gpurun_out/ab.log
P=vit-ocm-wmsegmentation_b200
cp $P/libvitocm.so $P/libvitocm_new.so
for v in new prev; do
  cp $P/libvitocm_$v.so $P/libvitocm.so
  TILES=4 TOKENS=12545 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/$v /" >> gpurun_out/ab.log
  TILES=32 TOKENS=3137 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/$v /" >> gpurun_out/ab.log
  TILES=64 TOKENS=1024 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/$v /" >> gpurun_out/ab.log
  TILES=175 TOKENS=785 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/$v /" >> gpurun_out/ab.log
done
cp $P/libvitocm_new.so $P/libvitocm.so
rm -f $P/libvitocm_new.so $P/libvitocm_prev.so
timeout 600 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -E "passed|failed|FAILED|vitocm:" | head -10 >> gpurun_out/ab.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1 >> gpurun_out/ab.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_seg.json 2> gpurun_out/bench_seg.err
python - >> gpurun_out/ab.log <<'PY'
import json
d = json.loads(open("gpurun_out/bench_seg.json").read().strip().splitlines()[-1])
print("bench", round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms", d["step_ms_rank0"], d["clocks"], "attention", d["kernel_classes"]["attention"])
PY
cat gpurun_out/ab.log
