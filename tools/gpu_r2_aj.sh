#!/bin/bash
# round 2, call AJ: hand-overs with plain (.release.cta) remote arrives instead of .release.cluster: tests, timing A/B, fine stamps
mkdir -p gpurun_out
L=gpurun_out/r2aj.log
: > $L
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -x -k "block_tail" 2>&1 | grep -E "passed|failed|FAILED|Error|error|vitocm:|assert|timeout" | head -30 >> $L
for dbg in 0 32 0 32; do
  VITOCM_TAIL_DEBUG=$dbg VITOCM_MLP_TL_ITEM=4 timeout 200 python tools/tail_timeline.py 175 2 1 2>&1 | head -1 | sed "s/^/debug=$dbg /" >> $L
  VITOCM_TAIL_DEBUG=$dbg VITOCM_MLP_TL_ITEM=20 timeout 200 python tools/tail_timeline.py 1225 2 1 2>&1 | head -1 | sed "s/^/debug=$dbg /" >> $L
done
VITOCM_TAIL_DEBUG=16 VITOCM_MLP_TL_ITEM=4 timeout 200 python tools/tail_timeline.py 175 2 1 2>&1 | grep -v "QKV chunks" >> $L
cat $L
