# Round 1, final capture of this round's kernels (run under gpurun, one B200):
#   launch list of the default bench (resident path, then the streaming path of the same workload),
#   --set full captures of the on-chip kernel and of the two streaming sweeps.
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1g.csv $BENCH > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resident_bp -s 3 -c 1 -o gpurun_out/resident_bp_r1g -f $BENCH > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:cn_sweep_tma -s 12 -c 2 -o gpurun_out/cn_sweep_tma_r1g -f $BENCH > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:vn_sweep -s 12 -c 2 -o gpurun_out/vn_sweep_r1g -f $BENCH > /dev/null 2>&1
ls -la gpurun_out
