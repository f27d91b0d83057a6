#!/bin/bash
# round 2, two GPUs of one box: the Monte-Carlo path under NCCL (torchrun -m ldpc_decoders_b200.sim) and both bench arms at N = 2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_n2_gpus.txt
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "nccl" -rs > gpurun_out/r2_n2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_n2_pytest.log
tail -5 gpurun_out/r2_n2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 exit $?"
tail -3 gpurun_out/r2_bench_n2.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
print('N=2 value %.4g e2e %.4g' % (d['value'], d['e2e']['value']), d['e2e']['h2d_GBps_by_rank'], d['e2e']['numa_node'])
for k,v in d['mc'].items(): print('mc', k, '%.4g' % v['value'], v['frames'], v['ms'], v['wer'])
P
