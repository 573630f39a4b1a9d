"""Dev probe: the RPN proposal stage alone (one CUDA graph), warm L2 (back-to-back replays) against cold L2 (a 512 MB write
between replays, as the RoIAlign of the previous step leaves it)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200det
from b200det import fused, workload

dev = torch.device("cuda:0")
B, K = 8, 8
w = workload.config2(B=B, K=K, with_feats=False)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cls, reg = [T(c) for c in w["cls"]], [T(r) for r in w["reg"]]
img_hw = torch.tensor([[800.0, 1333.0]] * B, device=dev)
if len(sys.argv) > 1:
    os.environ["B2D_DBG"] = sys.argv[1]; b200det._C.reload_knobs()
hp = fused.TrainHotPath(B, w["grids"], dev, gt_ld=K, feat_channels=16, layout=1)
step = lambda: hp.proposals(cls, reg, img_hw)
for _ in range(3): step()
torch.cuda.synchronize()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    step()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        step()
torch.cuda.current_stream().wait_stream(side)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for cold in (False, True, False, True):
    ts = []
    for _ in range(20):
        if cold: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print("proposal stage, %s L2: median %.1f us  min %.1f us" % ("cold" if cold else "warm", float(np.median(ts)), float(np.min(ts))))
