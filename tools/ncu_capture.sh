#!/bin/bash
# One `ncu --set full` capture of a kernel of the bench command, exported to CSV on the GPU box (the .ncu-rep embeds the whole
# cubin, ~40 MB per report, and gpurun_out/ may carry 64 MiB back).
#   tools/ncu_capture.sh <name> <kernel regex> <skip> <count> <bench args...>
name=$1; regex=$2; skip=$3; count=$4; shift 4
rep=/tmp/$name.ncu-rep
ncu --set full --clock-control none -k "regex:$regex" -s "$skip" -c "$count" -o /tmp/$name -f python bench.py "$@" > gpurun_out/${name}_ncu.log 2>&1 || exit 1
ncu -i $rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
ncu -i $rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
ls -la $rep gpurun_out/${name}_*.csv
