#!/bin/bash
# round 2, call AB: quad attention with PV_j / S_{j+1} prepared before the wait for P_j
mkdir -p gpurun_out
L=gpurun_out/r2ab.log
: > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fp16.py -m gpu -q --no-header -x -k "attention" 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert" | head -30 >> $L
TILES=1225 TOKENS=785 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 >> $L
TILES=175 TOKENS=785 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 >> $L
TILES=1225 TOKENS=768 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 >> $L
TILES=32 TOKENS=3137 PRECISION=2 timeout 120 python tools/attn_bench.py 2>&1 | tail -1 >> $L
VITOCM_ATTN_TL_ITEM=5 timeout 120 python tools/attn_quad_timeline.py 175 6 785 2>&1 | head -15 >> $L
cat $L
