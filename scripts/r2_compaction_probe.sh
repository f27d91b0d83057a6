#!/bin/bash
# How much DRAM traffic does the streaming path waste on frames that already converged?  n = 64800, BIAWGN 3 dB,
# max_iter 100: total dram bytes of ONE decode (all kernels, ncu) against B_iter * sum of iteration counts.
mkdir -p gpurun_out
SNR=${1:-3.0}; FR=${2:-2048}; N=${3:-64800}
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/compaction_${N}_${SNR}.csv \
    python scripts/run_case.py --n $N --snr $SNR --max-iter 100 --frames $FR --steps 1 --warmup 0 --iters-out gpurun_out/compaction_${N}_${SNR}_iters.npy > gpurun_out/compaction_${N}_${SNR}.log 2>&1
python - <<P
import csv, numpy as np
rows=[r for r in csv.reader(open('gpurun_out/compaction_${N}_${SNR}.csv')) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
tot=0; t=0; per={}
for r in rows[1:]:
    name=r[ix['Kernel Name']].split('<')[0].split('(')[0]; val=float(r[ix['Metric Value']].replace(',','')); m=r[ix['Metric Name']]; u=r[ix['Metric Unit']]
    if m.startswith('dram__bytes'):
        mult={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}[u]
        tot+=val*mult; per[name]=per.get(name,0)+val*mult
    elif m.startswith('gpu__time'):
        t+=val*{'ns':1e-9,'us':1e-6,'usecond':1e-6,'ms':1e-3,'msecond':1e-3,'nsecond':1e-9,'second':1}[u]
it=np.load('gpurun_out/compaction_${N}_${SNR}_iters.npy')
n=$N; E=3*n
alg=(4*E*4+n*4+(n+E)/8)*it.sum()
print('n=%d snr=${SNR} frames=%d: iters mean %.2f max %d; dram %.3f GB, algorithmic %.3f GB, ratio %.3f; kernel time %.2f ms' % (n, it.size, it.mean(), it.max(), tot/1e9, alg/1e9, tot/alg, t*1e3))
for k,v in sorted(per.items(), key=lambda kv:-kv[1])[:6]: print('   %-30s %.3f GB' % (k, v/1e9))
P
