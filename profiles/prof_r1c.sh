mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:resident_bp -s 3 -c 1 -o gpurun_out/resident_bp_r1f -f $BENCH > /dev/null 2>&1
ls -la gpurun_out
