mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cn_sweep -s 12 -c 2 -o gpurun_out/cn_sweep_r1 -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:vn_sweep -s 12 -c 2 -o gpurun_out/vn_sweep_r1 -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
