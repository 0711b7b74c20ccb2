#!/bin/bash
# round 2, final kernels: ncu launch list of the bench command + --set full captures (175-tile chunk so that one replayed launch
# stays short).  gpurun_out/ may carry at most 64 MiB back: every capture is exported to CSV on the box.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline"
$B > gpurun_out/ncu2_bench_plain.json 2> gpurun_out/ncu2_bench_plain.err || { echo "plain bench failed"; tail -5 gpurun_out/ncu2_bench_plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_seg_final.csv $B > gpurun_out/ncu2_l.log 2>&1
P="python tools/profile_step.py 175 vit_small fp16"
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -o /tmp/prof_$name -f "$@" > gpurun_out/ncu2_$name.log 2>&1
  ncu -i /tmp/prof_$name.ncu-rep --page details --csv > gpurun_out/prof2_${name}_details.csv 2>/dev/null
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/prof2_${name}_raw.csv 2>/dev/null
}
cap tail block_tail 3 1 $P
cap attnq attn_fwd_quad 3 1 $P
cap gemm gemm_bf16 6 2 $P
ncu -i /tmp/prof_attnq.ncu-rep --page source --csv > gpurun_out/prof2_attnq_source.csv 2>/dev/null
ncu -i /tmp/prof_tail.ncu-rep --page source --csv > gpurun_out/prof2_tail_source.csv 2>/dev/null
python tools/ncu_summary.py /tmp/prof_tail.ncu-rep /tmp/prof_attnq.ncu-rep /tmp/prof_gemm.ncu-rep > gpurun_out/ncu2_summary.txt 2>&1
cp /tmp/prof_attnq.ncu-rep /tmp/prof_tail.ncu-rep gpurun_out/
du -sh gpurun_out
tail -2 gpurun_out/ncu2_*.log
cat gpurun_out/ncu2_summary.txt
echo done
