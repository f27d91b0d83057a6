#!/bin/bash
# ncu full capture of the on-chip erasure kernel on config 2 (131072 frames), tag = $1
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resident_bec -s 2 -c 1 -o gpurun_out/resident_bec_$TAG -f python scripts/run_case.py --channel bec --snr 0.4 --cw 0 --frames 131072 --steps 1 > gpurun_out/prof_bec_$TAG.log 2>&1
ncu -i gpurun_out/resident_bec_$TAG.ncu-rep --page raw --csv > gpurun_out/resident_bec_${TAG}_raw.csv
ncu -i gpurun_out/resident_bec_$TAG.ncu-rep --page source --csv > gpurun_out/resident_bec_${TAG}_source.csv 2>/dev/null
python scripts/ncu_source_summary.py gpurun_out/resident_bec_${TAG}_source.csv --phases > gpurun_out/resident_bec_${TAG}_phases.txt 2>&1
rm -f gpurun_out/resident_bec_$TAG.ncu-rep
tail -2 gpurun_out/prof_bec_$TAG.log
