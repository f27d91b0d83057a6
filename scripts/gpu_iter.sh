# Quick GPU iteration: parity tests, MSA / SPA case timings, optional source-level ncu capture of the on-chip kernel
#   bash scripts/gpu_iter.sh [prof-msa] [prof-spa] [bench]
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
{
python scripts/run_case.py --algo MSA --steps 10
python scripts/run_case.py --algo SPA --cw 0 --steps 10
python scripts/run_case.py --algo SPA --cw 0 --snr 3.0 --steps 10
python scripts/run_case.py --algo SPA --cw 0 --streaming
python scripts/run_case.py --code 1200_rho_x5_rand_ldpc_1 --channel bsc --snr 0.06 --algo SPA --cw 0 --max-iter 100
} 2>&1 | tee gpurun_out/cases.txt
for a in "$@"; do
case $a in
prof-msa)
ncu --set full --clock-control none --import-source on -k regex:resident_bp -s 2 -c 1 -o gpurun_out/resident_msa -f python scripts/run_case.py --algo MSA --steps 1 > /dev/null 2>&1
ncu -i gpurun_out/resident_msa.ncu-rep --page raw --csv > gpurun_out/resident_msa_raw.csv
ncu -i gpurun_out/resident_msa.ncu-rep --page source --csv > gpurun_out/resident_msa_source.csv 2>/dev/null ;;
prof-spa)
ncu --set full --clock-control none --import-source on -k regex:resident_bp -s 2 -c 1 -o gpurun_out/resident_spa -f python scripts/run_case.py --algo SPA --cw 0 --steps 1 > /dev/null 2>&1
ncu -i gpurun_out/resident_spa.ncu-rep --page raw --csv > gpurun_out/resident_spa_raw.csv
ncu -i gpurun_out/resident_spa.ncu-rep --page source --csv > gpurun_out/resident_spa_source.csv 2>/dev/null ;;
bench)
timeout 500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc $?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'spa', d['spa']['value'], 'stream', d['roofline_streaming']['value'])" ;;
esac
done
rm -f gpurun_out/*.ncu-rep.tmp
