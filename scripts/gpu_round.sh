# Round-end style regression on one B200 (run under gpurun from the repo root):
# GPU parity tests, smoke(), both bench arms, then ncu launch list + full capture of the on-chip SPA kernel.
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
if [ "$1" = "prof" ]; then
ncu --set full --clock-control none --import-source on -k regex:resident_bp -s 2 -c 1 -o gpurun_out/resident_spa_r1h -f python scripts/run_case.py --algo SPA --cw 0 --steps 1 > /dev/null 2>&1
ncu -i gpurun_out/resident_spa_r1h.ncu-rep --page raw --csv > gpurun_out/resident_spa_r1h_raw.csv
ncu -i gpurun_out/resident_spa_r1h.ncu-rep --page source --csv > gpurun_out/resident_spa_r1h_source.csv 2>/dev/null
fi
ls -la gpurun_out
