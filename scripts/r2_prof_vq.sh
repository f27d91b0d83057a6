#!/bin/bash
# ncu full capture of the float32 on-chip kernel on the headline workload, tag = $1, extra run_case args = $2..
TAG=${1:-r2d}; shift
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resident_v -s 2 -c 1 -o gpurun_out/resident_$TAG -f python scripts/run_case.py --steps 1 "$@" > gpurun_out/prof_$TAG.log 2>&1
ncu -i gpurun_out/resident_$TAG.ncu-rep --page raw --csv > gpurun_out/resident_${TAG}_raw.csv
ncu -i gpurun_out/resident_$TAG.ncu-rep --page source --csv > gpurun_out/resident_${TAG}_source.csv 2>/dev/null
python scripts/ncu_source_summary.py gpurun_out/resident_${TAG}_source.csv --phases > gpurun_out/resident_${TAG}_phases.txt 2>&1
rm -f gpurun_out/resident_$TAG.ncu-rep
tail -1 gpurun_out/prof_$TAG.log
