# GPU regression: parity tests, then the bench line (run under gpurun from the repo root)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
if [ "$1" != "notbench" ]; then
timeout 400 python bench.py $BENCH_ARGS > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['e2e']['value'])
for e in d.get('extra',[]): print(e['workload'][:70], e['value'], e.get('mean_iters'))
print(d.get('extra_error'))
P
fi
