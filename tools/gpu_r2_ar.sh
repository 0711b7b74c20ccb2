#!/bin/bash
# round 2, call AR: accuracy of the folded block tail against the plain one (rows, masks) over 96 tiles
mkdir -p gpurun_out
timeout 900 python tools/fold_accuracy.py 96 > gpurun_out/r2ar.log 2>&1
cat gpurun_out/r2ar.log
