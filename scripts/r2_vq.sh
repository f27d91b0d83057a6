#!/bin/bash
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sim.py -m gpu -q -x 2>&1 | tail -4
for a in MSA SPA; do
  cw=1; [ $a = SPA ] && cw=0
  python scripts/run_case.py --algo $a --cw $cw --steps 10
  LDPC_RESIDENT_VP=1 python scripts/run_case.py --algo $a --cw $cw --steps 10
done
python scripts/run_case.py --algo MSA --snr 3.0 --steps 10
LDPC_RESIDENT_VP=1 python scripts/run_case.py --algo MSA --snr 3.0 --steps 10
python scripts/run_case.py --algo MSA --snr 1.0 --steps 10
LDPC_RESIDENT_VP=1 python scripts/run_case.py --algo MSA --snr 1.0 --steps 10
python scripts/run_case.py --algo MSA --channel bsc --snr 0.05 --steps 10
LDPC_RESIDENT_VP=1 python scripts/run_case.py --algo MSA --channel bsc --snr 0.05 --steps 10
python scripts/run_case.py --algo MSA --code 512_3_6_rand_ldpc_1 --steps 10
LDPC_RESIDENT_VP=1 python scripts/run_case.py --algo MSA --code 512_3_6_rand_ldpc_1 --steps 10
timeout 600 compute-sanitizer --tool racecheck python scripts/run_case.py --algo MSA --frames 600 --steps 1 --warmup 0 2>&1 | tail -2
timeout 600 compute-sanitizer --tool memcheck python scripts/run_case.py --algo SPA --cw 0 --frames 601 --steps 1 --warmup 0 2>&1 | tail -2
} > gpurun_out/r2_vq.log 2>&1
cat gpurun_out/r2_vq.log
