#!/bin/bash
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x 2>&1 | tail -4
for a in MSA SPA; do
  python scripts/run_case.py --code margulis --algo $a --cw 0 --steps 5
  LDPC_RESIDENT_VP=1 python scripts/run_case.py --code margulis --algo $a --cw 0 --steps 5
done
python scripts/run_case.py --code margulis --algo MSA --cw 0 --channel bsc --snr 0.05 --steps 5
LDPC_RESIDENT_VP=1 python scripts/run_case.py --code margulis --algo MSA --cw 0 --channel bsc --snr 0.05 --steps 5
python scripts/run_case.py --n 2000 --algo MSA --steps 5
LDPC_RESIDENT_VP=1 python scripts/run_case.py --n 2000 --algo MSA --steps 5
timeout 600 compute-sanitizer --tool racecheck python scripts/run_case.py --code margulis --algo MSA --cw 0 --frames 400 --steps 1 --warmup 0 2>&1 | tail -2
timeout 600 compute-sanitizer --tool memcheck python scripts/run_case.py --code margulis --algo SPA --cw 0 --frames 401 --steps 1 --warmup 0 2>&1 | tail -2
} > gpurun_out/r2_vq4.log 2>&1
cat gpurun_out/r2_vq4.log
