#!/bin/bash
# round 2, call W: cost of tail items inside the quad kernel: pairs per tile 1 / 2 / 4, timeline of a tail item and of a full item
mkdir -p gpurun_out
L=gpurun_out/r2w.log
: > $L
for pk in 1 2 4; do
  VITOCM_ATTN_QUAD_PACK=$pk TILES=1225 TOKENS=785 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 | sed "s/^/quad_pack=$pk /" >> $L
done
echo "== timeline, tail item (item 0), pack 2" >> $L
VITOCM_ATTN_TL_ITEM=0 timeout 120 python tools/attn_quad_timeline.py 175 6 785 2>&1 | head -15 >> $L
echo "== timeline, tail item (item 0), pack 1" >> $L
VITOCM_ATTN_QUAD_PACK=1 VITOCM_ATTN_TL_ITEM=0 timeout 120 python tools/attn_quad_timeline.py 175 6 785 2>&1 | head -15 >> $L
echo "== timeline, full item (item 5)" >> $L
VITOCM_ATTN_TL_ITEM=5 timeout 120 python tools/attn_quad_timeline.py 175 6 785 2>&1 | head -15 >> $L
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -k "test_vits8_tile_config1 or over_tiles" 2>&1 | grep -E "passed|failed|Error" >> $L
cat $L
