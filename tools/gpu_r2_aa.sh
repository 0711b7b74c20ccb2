#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2aa.log
: > $L
VITOCM_MLP_TL_ITEM=3 timeout 200 python tools/tail_timeline.py 175 2 >> $L 2>&1
VITOCM_MLP_TL_ITEM=3 timeout 200 python tools/tail_timeline.py 175 0 2>&1 | head -1 >> $L
VITOCM_MLP_TL_ITEM=20 timeout 200 python tools/tail_timeline.py 1225 2 2>&1 | head -1 >> $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -x -k "block_tail" 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert|timeout" | head -30 >> $L
cat $L
