# Regression + profile pass on one B200 (run under gpurun from the repo root):
# GPU parity tests, smoke(), both bench arms, ncu launch list of the default bench, full captures of the on-chip
# kernel (min-sum and sum-product).  Tag = $1 (default r1i).
TAG=${1:-r1i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'spa', d['spa']['value'], 'stream', d['roofline_streaming']['value'])
print('clocks', d['clocks'])
for e in d.get('extra',[]): print(e['workload'][:70], e['value'], e.get('mean_iters'))
print(d.get('extra_error'))
r=json.load(open('gpurun_out/bench_ref.json')); print('ref', r['value'], r['cpu_baseline']['cores'])
P
BENCH="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/bench_under_ncu.log 2>&1
for A in MSA SPA; do
  a=$(echo $A | tr A-Z a-z); CW=1; [ $A = SPA ] && CW=0
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:resident_ -s 2 -c 1 -o gpurun_out/resident_${a}_$TAG -f python scripts/run_case.py --algo $A --cw $CW --steps 1 > /dev/null 2>&1
  ncu -i gpurun_out/resident_${a}_$TAG.ncu-rep --page raw --csv > gpurun_out/resident_${a}_${TAG}_raw.csv
  ncu -i gpurun_out/resident_${a}_$TAG.ncu-rep --page source --csv > gpurun_out/resident_${a}_${TAG}_source.csv 2>/dev/null
done
ls -la gpurun_out
