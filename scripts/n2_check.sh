# 2-GPU run of bench.py under torchrun, with and without the NUMA binding (run under gpurun --gpus 2)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
python -c "
import torch,os
for i in range(torch.cuda.device_count()):
    p=torch.cuda.get_device_properties(i); b='%04x:%02x:%02x.0'%(p.pci_domain_id,p.pci_bus_id,p.pci_device_id)
    try: print(i,b,open('/sys/bus/pci/devices/%s/local_cpulist'%b).read().strip(), open('/sys/bus/pci/devices/%s/numa_node'%b).read().strip())
    except Exception as e: print(i,b,repr(e))
print('affinity',len(os.sched_getaffinity(0)),'cpus',os.cpu_count())
" > gpurun_out/numa.txt 2>&1
cat gpurun_out/numa.txt
for nb in 0 1; do
LDPC_NUMA_BIND=$nb timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$nb bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > gpurun_out/bench_n2_numa$nb.json 2> gpurun_out/bench_n2_numa$nb.err
echo "rc=$? bind=$nb"; head -c 100 gpurun_out/bench_n2_numa$nb.json; echo
python -c "
import json; d=json.loads(open('gpurun_out/bench_n2_numa$nb.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['e2e'])"
done
LDPC_NUMA_BIND=1 python bench.py --no-extras --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['e2e'])"
