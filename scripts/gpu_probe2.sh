# irregular variable-plane kernel: memcheck on a small batch, GPU parity tests, case timings (run under gpurun).
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/run_case.py --code 1200_rho_x5_rand_ldpc_1 --algo SPA --channel bsc --snr 0.06 --cw 0 --frames 1024 --steps 1 --warmup 0 > gpurun_out/memcheck.txt 2>&1; echo "memcheck exit $?" >> gpurun_out/memcheck.txt; tail -4 gpurun_out/memcheck.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/run_case.py --code 7_4_hamming --algo MSA --cw 0 --frames 1024 --steps 1 --warmup 0 > gpurun_out/memcheck2.txt 2>&1; echo "memcheck exit $?" >> gpurun_out/memcheck2.txt; tail -3 gpurun_out/memcheck2.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
{
python scripts/run_case.py --algo MSA --steps 10
python scripts/run_case.py --code 1200_rho_x5_rand_ldpc_1 --channel bsc --snr 0.06 --algo SPA --cw 0 --max-iter 10
python scripts/run_case.py --code 1200_rho_x5_rand_ldpc_1 --channel bsc --snr 0.06 --algo SPA --cw 0 --max-iter 100
python scripts/run_case.py --code 1200_rho_x5_rand_ldpc_1 --channel bsc --snr 0.06 --algo MSA --cw 0 --max-iter 10
python scripts/run_case.py --code 1200_rho_x5_rand_ldpc_1 --channel biawgn --snr 2.0 --algo MSA --cw 0 --max-iter 10
python scripts/run_case.py --code 1200_rho_x5_rand_ldpc_1 --channel biawgn --snr 2.0 --algo SPA --cw 0 --max-iter 10
LDPC_RESIDENT_LAYOUT=check python scripts/run_case.py --code 1200_rho_x5_rand_ldpc_1 --channel biawgn --snr 2.0 --algo MSA --cw 0 --max-iter 10
} 2>&1 | tee gpurun_out/cases2.txt
