#!/bin/bash
# round 2: ncu launch list of the bench command + --set full captures of the kernels of the final state (175-tile chunk so that
# one replayed launch stays short).  gpurun_out/ may carry at most 64 MiB back: every capture is exported to CSV on the box
# (details + raw pages, source page for the hot kernels) and only the attention / MLP reports travel.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_seg.csv $B > gpurun_out/ncu_l.log 2>&1
P="python tools/profile_step.py 175 vit_small fp16"
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -o /tmp/prof_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/prof_$name.ncu-rep --page details --csv > gpurun_out/prof_${name}_details.csv 2>/dev/null
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/prof_${name}_raw.csv 2>/dev/null
}
cap mlp mlp_fused 3 1 $P
cap attn attn_fwd 3 1 $P
cap gemm gemm_bf16 9 3 $P
cap ln layernorm 4 1 $P
cap post "stitch|head_mean|otsu" 6 6 $B
ncu -i /tmp/prof_attn.ncu-rep --page source --csv > gpurun_out/prof_attn_source.csv 2>/dev/null
ncu -i /tmp/prof_mlp.ncu-rep --page source --csv > gpurun_out/prof_mlp_source.csv 2>/dev/null
cp /tmp/prof_attn.ncu-rep /tmp/prof_mlp.ncu-rep gpurun_out/
du -sh gpurun_out
tail -2 gpurun_out/ncu_*.log
echo done
