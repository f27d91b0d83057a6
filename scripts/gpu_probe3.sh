# one-CTA-per-SM on-chip geometry (Margulis n = 2640): memcheck, parity tests, timings (run under gpurun).
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/run_case.py --code margulis --algo MSA --snr 2.0 --cw 0 --frames 1024 --steps 1 --warmup 0 > gpurun_out/memcheck3.txt 2>&1; echo "memcheck exit $?" >> gpurun_out/memcheck3.txt; tail -4 gpurun_out/memcheck3.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
{
python scripts/run_case.py --code margulis --algo MSA --snr 2.0 --cw 0 --frames 16384
python scripts/run_case.py --code margulis --algo MSA --snr 2.0 --cw 0 --frames 16384 --streaming
python scripts/run_case.py --code margulis --algo SPA --snr 2.0 --cw 0 --frames 16384
python scripts/run_case.py --code margulis --algo SPA --snr 2.0 --cw 0 --frames 16384 --streaming
python scripts/run_case.py --algo MSA --steps 10
} 2>&1 | tee gpurun_out/cases3.txt
