#!/bin/bash
# ncu --set full capture of one kernel of the resident pass, exported as CSV (raw metrics + per-instruction source page) so that
# only text travels back from the GPU box.  usage: scripts/ncu_capture.sh <tag> <kernel-regex> <skip> [stage_bench args...]
tag=$1; kern=$2; skip=$3; shift 3
rep=/tmp/prof_$tag
ncu --set full --clock-control none --import-source on -k "regex:$kern" -s "$skip" -c 1 -f -o $rep python scripts/stage_bench.py --steps 1 "$@" > /dev/null 2>&1
ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$tag.csv 2>/dev/null
ncu -i $rep.ncu-rep --page source --csv > gpurun_out/ncu_src_$tag.csv 2>/dev/null
ls -la $rep.ncu-rep gpurun_out/ncu_raw_$tag.csv gpurun_out/ncu_src_$tag.csv
