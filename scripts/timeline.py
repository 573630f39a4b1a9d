"""Dev tool: GPU kernel timeline of one CUDA-graph replay of the config-2 step (torch.profiler / CUPTI).
usage: python scripts/timeline.py [groups] -> gpurun_out/timeline_g<groups>.csv + a printed critical-path view"""
import os, sys, json
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200det
from b200det import _C, fused, workload

groups = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
B, K = 8, 8
w = workload.config2(B=B, K=K)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cls, reg = [T(c) for c in w["cls"]], [T(r) for r in w["reg"]]
feats = [T(f).contiguous(memory_format=torch.channels_last) for f in w["feats"]]
gt, gl = T(w["gt"]), T(w["gt_label"])
gcount = torch.full((B,), K, dtype=torch.int32, device=dev)
img_hw = torch.tensor([[800.0, 1333.0]] * B, device=dev)
hp = fused.TrainHotPath(B, w["grids"], dev, gt_ld=K, feat_channels=256, layout=1, overlap=True, groups=groups)
step = lambda: hp.step(cls, reg, feats, gt, gcount, gl, img_hw)
for _ in range(3): step()
torch.cuda.synchronize()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    step()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        step()
torch.cuda.current_stream().wait_stream(side)
for _ in range(5): g.replay()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
rows = sorted(((e.time_range.start, e.time_range.end, e.name) for e in ev), key=lambda r: r[0])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "timeline_g%d.csv" % groups), "w") as f:
    f.write("start_us,end_us,name\n")
    for s, e, n in rows: f.write("%.3f,%.3f,%s\n" % (s, e, n.replace(",", ";")))
# last replay only: split by gaps > 50 us
reps, cur = [], []
for r in rows:
    if cur and r[0] - max(x[1] for x in cur) > 30: reps.append(cur); cur = []
    cur.append(r)
reps.append(cur)
last = reps[-1]
t0 = last[0][0]
print("kernels in last replay:", len(last), "span %.1f us" % (max(x[1] for x in last) - t0))
for s, e, n in last:
    print("%7.1f %7.1f  %6.1f  %s" % (s - t0, e - t0, e - s, n[:60]))
