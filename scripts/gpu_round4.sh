# Final regression + profile pass of the round on one B200 (run under gpurun from the repo root), tag = $1 (default r1m):
# GPU parity tests, smoke(), both bench arms, ncu launch list of the default bench, full captures of the float32 and
# float64 on-chip kernels on the headline workload.
TAG=${1:-r1m}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'spa', d['spa']['value'], 'stream', d['roofline_streaming']['value'])
print('kernel', d['roofline']['kernel'], 'traffic', d['roofline']['traffic'], 'shared frac', d['roofline']['shared']['frac'], 'clocks', d['clocks'])
for e in d.get('extra',[]): print(e['workload'][:90], e['value'], e.get('mean_iters'), e.get('path'))
print(d.get('extra_error'))
r=json.load(open('gpurun_out/bench_ref.json')); print('ref', r['value'], r['cpu_baseline']['cores'])
P
BENCH="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resident_ -s 2 -c 1 -o gpurun_out/resident_msa_$TAG -f python scripts/run_case.py --algo MSA --steps 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resident_ -s 2 -c 1 -o gpurun_out/resident_msa_f64_$TAG -f python scripts/run_case.py --algo MSA --dtype f64 --steps 1 > /dev/null 2>&1
for f in msa msa_f64; do
  ncu -i gpurun_out/resident_${f}_$TAG.ncu-rep --page raw --csv > gpurun_out/resident_${f}_${TAG}_raw.csv
  ncu -i gpurun_out/resident_${f}_$TAG.ncu-rep --page source --csv > gpurun_out/resident_${f}_${TAG}_source.csv 2>/dev/null
done
ls -la gpurun_out | grep $TAG
