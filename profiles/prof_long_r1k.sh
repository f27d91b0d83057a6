# ncu captures of the streaming sweeps on the synthetic (3,6) n = 64800 code (BASELINE config 5), run under gpurun.
TAG=${1:-r1k}
mkdir -p gpurun_out
CASE="python scripts/run_case.py --n 64800 --algo MSA --snr 2.5 --frames 2048 --steps 1 --warmup 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_long_$TAG.csv $CASE > gpurun_out/long_under_ncu.log 2>&1
for K in cn_sweep_tma vn_sweep; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 12 -c 1 -o gpurun_out/${K}_long_$TAG -f $CASE > /dev/null 2>&1
  ncu -i gpurun_out/${K}_long_$TAG.ncu-rep --page raw --csv > gpurun_out/${K}_long_${TAG}_raw.csv
done
$CASE
ls -la gpurun_out | grep long
