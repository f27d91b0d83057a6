# ncu captures of the bit-plane erasure kernels (BASELINE config 2: 1200_3_6_rand_ldpc_1 on BEC), run under gpurun.
TAG=${1:-r1m}
mkdir -p gpurun_out
cat > /tmp/bec_case.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, _golden as G
from ldpc_decoders_b200 import Tables, _lib as lib, engine as E
tab = Tables(*G.code_tables("1200_3_6_rand_ldpc_1"))
eng = E.engine_for(tab)
frames = 131072
g = torch.Generator(device="cuda").manual_seed(4)
yb = torch.where(torch.rand((frames, tab.n), generator=g, device="cuda") < 0.4, 2, 0).to(torch.uint8)
res = {}
for _ in range(3):
    res["o"] = eng.decode_device_channel(lib.CH_BEC, lib.BEC, lib.F32, 0.0, yb, max_iter=10, out=res.get("o"))
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(5):
    res["o"] = eng.decode_device_channel(lib.CH_BEC, lib.BEC, lib.F32, 0.0, yb, max_iter=10, out=res.get("o"))
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / 5
print("BEC p=0.4 frames=%d: %.3f ms/step, %.2f M frames/s, mean iters %.2f" % (frames, ms, frames / ms / 1e3, res["o"]["iters"].float().mean().item()))
P
python /tmp/bec_case.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_bec_$TAG.csv python /tmp/bec_case.py > /dev/null 2>&1
for K in bec_cn bec_vn; do
  timeout 600 ncu --set full --clock-control none -k regex:$K -s 14 -c 1 -o gpurun_out/${K}_$TAG -f python /tmp/bec_case.py > /dev/null 2>&1
  ncu -i gpurun_out/${K}_$TAG.ncu-rep --page raw --csv > gpurun_out/${K}_${TAG}_raw.csv
done
ls -la gpurun_out | grep bec
