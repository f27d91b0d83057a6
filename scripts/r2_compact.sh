#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "compaction" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "long or unlimited or polling or max_iter or stream" 2>&1 | tail -3
for env in "" "LDPC_NO_COMPACTION=1"; do
  env $env python scripts/run_case.py --n 4000 --snr 2.5 --max-iter 100 --frames 16384 --steps 3 --streaming
  env $env python scripts/run_case.py --n 64800 --snr 3.0 --max-iter 100 --frames 2048 --steps 3
  env $env python scripts/run_case.py --algo MSA --snr 2.6 --max-iter 100 --frames 32768 --steps 3 --streaming
done
bash scripts/r2_compaction_probe.sh 2.5 16384 4000
} > gpurun_out/r2_compact.log 2>&1
cat gpurun_out/r2_compact.log
