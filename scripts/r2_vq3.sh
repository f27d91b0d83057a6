#!/bin/bash
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sim.py tests/test_gpu_round2.py -m gpu -q -x 2>&1 | tail -4
python scripts/run_case.py --algo MSA --dtype f64 --steps 10
LDPC_RESIDENT_VP=1 python scripts/run_case.py --algo MSA --dtype f64 --steps 10
python scripts/run_case.py --algo MSA --dtype f64 --snr 3.0 --steps 10
LDPC_RESIDENT_VP=1 python scripts/run_case.py --algo MSA --dtype f64 --snr 3.0 --steps 10
python scripts/run_case.py --algo MSA --dtype f64 --channel bsc --snr 0.05 --steps 10
LDPC_RESIDENT_VP=1 python scripts/run_case.py --algo MSA --dtype f64 --channel bsc --snr 0.05 --steps 10
python scripts/run_case.py --algo MSA --dtype f64 --code 1200_rho_x5_rand_ldpc_1 --cw 0 --steps 10
LDPC_RESIDENT_VP=1 python scripts/run_case.py --algo MSA --dtype f64 --code 1200_rho_x5_rand_ldpc_1 --cw 0 --steps 10
python scripts/run_case.py --algo MSA --steps 10
python scripts/run_case.py --algo SPA --cw 0 --steps 10
timeout 600 compute-sanitizer --tool racecheck python scripts/run_case.py --algo MSA --dtype f64 --frames 600 --steps 1 --warmup 0 2>&1 | tail -2
timeout 600 compute-sanitizer --tool memcheck python scripts/run_case.py --algo MSA --dtype f64 --frames 601 --steps 1 --warmup 0 2>&1 | tail -2
} > gpurun_out/r2_vq3.log 2>&1
cat gpurun_out/r2_vq3.log
