#!/bin/bash
# round 2: on-chip erasure kernel — parity, A/B against the streaming sweeps, sanitizer
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "bec or golden_run or kat" 2>&1 | tail -15 > gpurun_out/r2_bec_tests.log
{
for f in 131072 32768 8192; do
  python scripts/run_case.py --channel bec --snr 0.4 --cw 0 --frames $f --steps 5
  python scripts/run_case.py --channel bec --snr 0.4 --cw 0 --frames $f --steps 5 --streaming
done
python scripts/run_case.py --channel bec --snr 0.4 --cw 0 --frames 131072 --steps 5 --max-iter 100
python scripts/run_case.py --channel bec --snr 0.4 --cw 0 --frames 131072 --steps 5 --max-iter 100 --streaming
python scripts/run_case.py --channel bec --snr 0.4 --cw 0 --frames 131072 --steps 5 --code 1200_rho_x5_rand_ldpc_1
python scripts/run_case.py --channel bec --snr 0.4 --cw 0 --frames 131072 --steps 5 --code 1200_rho_x5_rand_ldpc_1 --streaming
} > gpurun_out/r2_bec_perf.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python scripts/run_case.py --channel bec --snr 0.4 --cw 0 --frames 700 --steps 1 --warmup 0 > gpurun_out/r2_bec_race.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python scripts/run_case.py --channel bec --snr 0.4 --cw 0 --frames 700 --steps 1 --warmup 0 --code 1200_rho_x5_rand_ldpc_1 > gpurun_out/r2_bec_mem.log 2>&1
tail -15 gpurun_out/r2_bec_tests.log; cat gpurun_out/r2_bec_perf.log; tail -4 gpurun_out/r2_bec_race.log; tail -4 gpurun_out/r2_bec_mem.log
