#!/bin/bash
# round-2 regression pass on one B200: GPU tests, smoke, both bench arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu.log
tail -6 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2_smoke.log
tail -2 gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench exit $?"
tail -5 gpurun_out/r2_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench.json'))
print('value %.3g e2e %.3g spa %.3g f64 %.3g stream %.3g' % (d['value'], d['e2e']['value'], d['spa']['value'], d['msa_f64']['value'], d['roofline_streaming']['value']))
print('roofline', {k: d['roofline'][k] for k in ('bound','kernel','achieved','peak','frac','traffic','peak_source')})
print('effective', d['effective_hbm']['over_hbm_peak'], 'f64 roof', d['msa_f64']['roofline']['frac'], d['msa_f64']['path'])
for k,v in d['mc'].items(): print('mc', k, '%.4g' % v['value'], v['wer'], v['mean_iters'], v.get('step_hbm_frac'))
print('e2e', d['e2e']['bytes_per_frame'], d['e2e']['h2d_GBps_by_rank'], d['e2e']['numa_node'])
for v in d['e2e_variants']: print('  var', v.get('workload','')[:70], '%.4g' % v.get('value',0), v.get('e2e_over_device'), v.get('error'))
for e in d.get('extra',[]): print(e['workload'][:90], '%.4g' % e['value'], e.get('mean_iters'), e.get('path'))
print(d.get('extra_error'), d['clocks'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
r=json.load(open('gpurun_out/r2_bench_ref.json')); print('ref', r['value'], r['steps'], r['config']['frames_per_step_per_gpu'], r['ms_per_step'])
P
