# L2 probe + named case timings on one B200 (run under gpurun from the repo root).
mkdir -p gpurun_out
timeout 300 python scripts/l2_probe.py > gpurun_out/l2_probe.txt 2>&1; tail -12 gpurun_out/l2_probe.txt
{
python scripts/run_case.py --algo MSA --steps 10
python scripts/run_case.py --algo SPA --cw 0 --steps 10
python scripts/run_case.py --code 1200_rho_x5_rand_ldpc_1 --channel bsc --snr 0.06 --algo SPA --cw 0 --max-iter 10
python scripts/run_case.py --code 1200_rho_x5_rand_ldpc_1 --channel bsc --snr 0.06 --algo SPA --cw 0 --max-iter 100
python scripts/run_case.py --code 1200_rho_x5_rand_ldpc_1 --channel biawgn --snr 2.0 --algo MSA --cw 0 --max-iter 10
python scripts/run_case.py --code 1200_rho_x5_rand_ldpc_1 --channel biawgn --snr 2.0 --algo MSA --cw 0 --max-iter 10 --streaming
python scripts/run_case.py --code margulis --algo MSA --snr 2.0 --cw 0 --frames 16384
python scripts/run_case.py --n 64800 --algo MSA --snr 2.5 --frames 2048
python scripts/run_case.py --n 64800 --algo SPA --snr 2.5 --frames 2048 --cw 0
python scripts/run_case.py --n 64800 --algo MSA --snr 2.5 --frames 512
} 2>&1 | tee gpurun_out/cases.txt
