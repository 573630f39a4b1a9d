"""Dev benchmark: BASELINE config 3 (cascade_rcnn_r50_fpn) hot path, batch 8: 3 x (bbox_target -> RoIAlign fwd ->
refine) + 3 x RoIAlign bwd on the config-2 proposals / features (fused.CascadeHotPath)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200det
from b200det import fused, workload

dev = torch.device("cuda:0")
B, K = 8, 8
w = workload.config2(B=B, K=K)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cls, reg = [T(c) for c in w["cls"]], [T(r) for r in w["reg"]]
feats = [T(f).contiguous(memory_format=torch.channels_last) for f in w["feats"]]
gt, gl = T(w["gt"]), T(w["gt_label"])
gcount = torch.full((B,), K, dtype=torch.int32, device=dev)
img_hw = torch.tensor([[800.0, 1333.0]] * B, device=dev)
hp = fused.TrainHotPath(B, w["grids"], dev, gt_ld=K, feat_channels=256, layout=1)
out = hp.step(cls, reg, feats, gt, gcount, gl, img_hw)
props, pcount = out["props"].clone(), out["prop_count"].clone()
cp = fused.CascadeHotPath(B, props.shape[2], w["grids"], dev, gt_ld=K, feat_channels=256)
regs = [torch.randn((B, cp.m, 4 * 21), device=dev) * 0.5 for _ in range(3)]
gos = [torch.randn((B * cp.m, 256, 7, 7), device=dev) for _ in range(3)]
def fwd(): return cp.step(props, pcount, feats[:4], gt, gcount, gl, img_hw, regs)
def bwd(): return cp.backward(gos)
for _ in range(3): fwd(); bwd()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
n, tf, tb = 10, 0.0, 0.0
for _ in range(n):
    e[0].record(); fwd(); e[1].record(); bwd(); e[2].record(); torch.cuda.synchronize()
    tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
print("config 3, batch %d: 3-stage forward chain %.1f us, 3 x RoIAlign backward %.1f us -> %.0f images/s" % (
    B, tf / n * 1e3, tb / n * 1e3, B / ((tf + tb) / n * 1e-3)))
