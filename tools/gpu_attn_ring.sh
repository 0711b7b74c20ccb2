#!/bin/bash
# GPU box: A/B of the K/V ring depth (libvitocm_ring{3,4}.so built with -DVITOCM_ATT_RING) and of tail packing
mkdir -p gpurun_out
: > gpurun_out/ring.log
P=vit-ocm-wmsegmentation_b200
cp $P/libvitocm.so $P/libvitocm_ring5.so
for rep in 1 2; do
for r in 3 4 5; do
  cp $P/libvitocm_ring$r.so $P/libvitocm.so
  for pk in 0 1; do
    VITOCM_ATTN_PACK=$pk TILES=175 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/ring=$r pack=$pk /" >> gpurun_out/ring.log
  done
done
done
cp $P/libvitocm_ring5.so $P/libvitocm.so
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pack1.json 2> gpurun_out/bench_pack1.err
VITOCM_ATTN_PACK=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pack0.json 2> gpurun_out/bench_pack0.err
cat gpurun_out/ring.log
python - <<'PY'
import json
for f in ("bench_pack1", "bench_pack0"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["kernel_classes"]["attention"], d["clocks"])
    except Exception as e:
        print(f, "ERR", e)
PY
