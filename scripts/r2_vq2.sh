#!/bin/bash
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sim.py tests/test_gpu_round2.py -m gpu -q -x 2>&1 | tail -4
C=1200_rho_x5_rand_ldpc_1
for a in SPA MSA; do
  python scripts/run_case.py --code $C --algo $a --cw 0 --channel bsc --snr 0.06 --steps 10
  LDPC_RESIDENT_VP=1 python scripts/run_case.py --code $C --algo $a --cw 0 --channel bsc --snr 0.06 --steps 10
done
python scripts/run_case.py --code $C --algo SPA --cw 0 --channel bsc --snr 0.06 --steps 5 --max-iter 100
LDPC_RESIDENT_VP=1 python scripts/run_case.py --code $C --algo SPA --cw 0 --channel bsc --snr 0.06 --steps 5 --max-iter 100
python scripts/run_case.py --code $C --algo MSA --cw 0 --snr 2.0 --steps 10
LDPC_RESIDENT_VP=1 python scripts/run_case.py --code $C --algo MSA --cw 0 --snr 2.0 --steps 10
timeout 600 compute-sanitizer --tool racecheck python scripts/run_case.py --code $C --algo MSA --cw 0 --channel bsc --snr 0.06 --frames 600 --steps 1 --warmup 0 2>&1 | tail -2
timeout 600 compute-sanitizer --tool memcheck python scripts/run_case.py --code 7_4_hamming --algo SPA --cw 0 --frames 601 --steps 1 --warmup 0 2>&1 | tail -2
} > gpurun_out/r2_vq2.log 2>&1
cat gpurun_out/r2_vq2.log
