# Round-2 regression + profile pass on one B200 (run under gpurun from the repo root), tag = $1 (default r2f):
# GPU parity tests, smoke(), both bench arms, the ncu launch list of the default bench, full captures of the float32
# on-chip kernel (resident_vq) on the headline workload and of the erasure kernel (resident_bec) on config 2,
# the SPA error table, a short config-5 run.
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke_$TAG.log
tail -2 gpurun_out/smoke_$TAG.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
./tools/smem_peak > gpurun_out/smem_peak_$TAG.json
python scripts/spa_error_buckets.py > gpurun_out/spa_error_buckets_$TAG.txt 2>&1
timeout 300 python scripts/config5.py --frames 200000 --out gpurun_out/config5_n1_$TAG.json > /dev/null 2> gpurun_out/config5_n1_$TAG.err
BENCH="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --mc-rounds 8"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/bench_under_ncu_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resident_vq -s 2 -c 1 -o gpurun_out/resident_vq_$TAG -f python scripts/run_case.py --algo MSA --steps 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resident_bec -s 2 -c 1 -o gpurun_out/resident_bec_$TAG -f python scripts/run_case.py --channel bec --snr 0.4 --cw 0 --frames 131072 --steps 1 > /dev/null 2>&1
for f in resident_vq resident_bec; do
  ncu -i gpurun_out/${f}_$TAG.ncu-rep --page raw --csv > gpurun_out/${f}_${TAG}_raw.csv
  ncu -i gpurun_out/${f}_$TAG.ncu-rep --page source --csv > gpurun_out/${f}_${TAG}_source.csv 2>/dev/null
  python scripts/ncu_source_summary.py gpurun_out/${f}_${TAG}_source.csv --phases > gpurun_out/${f}_${TAG}_phases.txt 2>&1
  rm -f gpurun_out/${f}_$TAG.ncu-rep gpurun_out/${f}_${TAG}_source.csv
done
python - <<P
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print('value %.4g e2e %.4g spa %.4g f64 %.4g stream %.4g' % (d['value'], d['e2e']['value'], d['spa']['value'], d['msa_f64']['value'], d['roofline_streaming']['value']))
print('roofline', {k: d['roofline'][k] for k in ('bound','kernel','achieved','peak','frac','traffic')})
for k,v in d['mc'].items(): print('mc', k, '%.4g' % v['value'], v['wer'], v.get('step_hbm_frac'))
for v in d['e2e_variants']: print('  var', v.get('workload','')[:60], '%.4g' % v.get('value',0), v.get('e2e_over_device'), v.get('error'))
for e in d.get('extra',[]): print(e['workload'][:80], '%.4g' % e['value'], e.get('path'))
print(d.get('extra_error'), d['clocks'])
r=json.load(open('gpurun_out/bench_ref_$TAG.json')); print('ref', r['value'], r['steps'])
print(open('gpurun_out/smem_peak_$TAG.json').read())
print(open('gpurun_out/spa_error_buckets_$TAG.txt').read())
print(open('gpurun_out/config5_n1_$TAG.json').read()[:600])
P
ls -la gpurun_out | grep $TAG
