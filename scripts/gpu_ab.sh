# A/B of the resident kernels on one B200: GPU parity tests, then the bench headline with the variable-plane layout
# (default) and the check-major layout (LDPC_RESIDENT_LAYOUT=check).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
for lay in vp check; do
  LDPC_RESIDENT_LAYOUT=$lay timeout 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench_$lay.json 2> gpurun_out/bench_$lay.err
  python - <<P
import json
try:
    d=json.load(open('gpurun_out/bench_$lay.json'))
    print('$lay', 'value %.3fM' % (d['value']/1e6), 'e2e %.3fM' % (d['e2e']['value']/1e6), 'spa %.3fM' % (d['spa']['value']/1e6), 'stream %.3fM' % (d['roofline_streaming']['value']/1e6), d['roofline'].get('shared_memory_plan'))
except Exception as e:
    print('$lay failed', e); print(open('gpurun_out/bench_$lay.err').read()[-2000:])
P
done
