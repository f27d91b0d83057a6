# float64 on-chip min-sum (resident_vd): memcheck, parity tests, timings (run under gpurun).
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/run_case.py --algo MSA --dtype f64 --frames 2048 --steps 1 --warmup 0 > gpurun_out/memcheck6.txt 2>&1; echo "memcheck exit $?" >> gpurun_out/memcheck6.txt; tail -3 gpurun_out/memcheck6.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
{
python scripts/run_case.py --algo MSA --dtype f64 --steps 10
python scripts/run_case.py --algo MSA --dtype f64 --steps 5 --streaming
python scripts/run_case.py --algo MSA --dtype f64 --snr 3.0 --steps 10
python scripts/run_case.py --algo MSA --dtype f64 --channel bsc --snr 0.05 --steps 10
python scripts/run_case.py --code 512_3_6_rand_ldpc_1 --algo MSA --dtype f64 --cw 0
python scripts/run_case.py --code 1200_rho_x5_rand_ldpc_1 --algo MSA --dtype f64 --cw 0
python scripts/run_case.py --code 1200_rho_x5_rand_ldpc_1 --algo MSA --dtype f64 --cw 0 --streaming
python scripts/run_case.py --algo MSA --steps 10
} 2>&1 | tee gpurun_out/cases5.txt
timeout 500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "import json; d=json.load(open(\"gpurun_out/bench.json\")); print(d[\"value\"], d[\"e2e\"][\"value\"]); [print(e[\"workload\"][:80], e[\"value\"], e.get(\"path\")) for e in d[\"extra\"]]"
