#!/bin/bash
# round 2, eight GPUs of one box: host-fabric probe, both bench arms at N = 8, config 5 (n = 64800, 1e8 frames) over 8 GPUs
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_n8_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29701 scripts/h2d_topo_probe.py > gpurun_out/r2_n8_h2d_probe.txt 2> gpurun_out/r2_n8_h2d_probe.err
timeout 600 $TR --master-port 29702 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench n8 exit $?"
timeout 900 $TR --master-port 29703 scripts/config5.py --frames ${1:-100000000} --snr 2.5 --out gpurun_out/config5_n8.json > gpurun_out/r2_config5_n8.log 2> gpurun_out/r2_config5_n8.err; echo "config5 exit $?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_n8.json'))
print('N=8 value %.4g e2e %.4g' % (d['value'], d['e2e']['value']), [round(x,1) for x in d['e2e']['h2d_GBps_by_rank']], d['e2e']['numa_node'])
for k,v in d['mc'].items(): print('mc', k, '%.4g' % v['value'], v['frames'], v['ms'], v['wer'])
c=json.load(open('gpurun_out/config5_n8.json')); print({k:c[k] for k in ('frames','seconds','frames_per_s','wer','ber','mean_iters','hbm_GBps_per_gpu_algorithmic')}, c['progress_reports'][:4])
P
tail -20 gpurun_out/r2_n8_h2d_probe.txt
