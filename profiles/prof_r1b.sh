# Round 1, second capture: resident on-chip kernel + bulk-async check-node sweep (run under gpurun, one B200).
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv $BENCH > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resident_bp -s 3 -c 1 -o gpurun_out/resident_bp_r1b -f $BENCH > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:cn_sweep_tma -s 12 -c 2 -o gpurun_out/cn_sweep_tma_r1b -f $BENCH > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:vn_sweep -s 12 -c 2 -o gpurun_out/vn_sweep_r1b -f $BENCH > /dev/null 2>&1
ls -la gpurun_out
