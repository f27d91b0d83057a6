#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests -m gpu -q -x -k "bec" 2>&1 | tail -3
for f in 131072 32768; do
  python scripts/run_case.py --channel bec --snr 0.4 --cw 0 --frames $f --steps 10
  python scripts/run_case.py --channel bec --snr 0.4 --cw 0 --frames $f --steps 10
done
python scripts/run_case.py --channel bec --snr 0.4 --cw 0 --frames 131072 --steps 5 --max-iter 100
python scripts/run_case.py --channel bec --snr 0.35 --cw 0 --frames 131072 --steps 5 --max-iter 10
python scripts/run_case.py --channel bec --snr 0.5 --cw 1 --frames 131072 --steps 5 --max-iter 10
python scripts/run_case.py --channel bec --snr 0.4 --cw 0 --frames 131072 --steps 5 --code 512_3_6_rand_ldpc_1
./tools/smem_peak
} > gpurun_out/r2_bec2.log 2>&1
cat gpurun_out/r2_bec2.log
