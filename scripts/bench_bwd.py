"""Dev micro-benchmark: RoIAlign backward (K6) at config-2/3 sizes (8 images x 512 RoIs, 256 channels)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200det
from b200det import region, fused, workload

dev = torch.device("cuda:0")
B, K = 8, 8
w = workload.config2(B=B, K=K)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cls, reg = [T(c) for c in w["cls"]], [T(r) for r in w["reg"]]
feats = [T(f).contiguous(memory_format=torch.channels_last).requires_grad_(True) for f in w["feats"]]
gt, gl = T(w["gt"]), T(w["gt_label"])
gcount = torch.full((B,), K, dtype=torch.int32, device=dev)
img_hw = torch.tensor([[800.0, 1333.0]] * B, device=dev)
hp = fused.TrainHotPath(B, w["grids"], dev, gt_ld=K, feat_channels=256, layout=1)
out = hp.step(cls, reg, [f.detach() for f in feats], gt, gcount, gl, img_hw)
bt = out["rcnn"]
rois = torch.cat([bt.tar_box[b, :, :int(bt.n_chosen[b])] for b in range(B)], 1).contiguous()
roi_img = torch.cat([torch.full((int(bt.n_chosen[b]),), b, dtype=torch.int32, device=dev) for b in range(B)])
scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
np.save(os.path.join(ROOT, "gpurun_out", "bwd_rois.npy"), rois.cpu().numpy()); np.save(os.path.join(ROOT, "gpurun_out", "bwd_img.npy"), roi_img.cpu().numpy())
o = region.roi_align_levels(feats, rois, roi_img, scales)
go = torch.randn_like(o)
def fb():
    for f in feats: f.grad = None
    o = region.roi_align_levels(feats, rois, roi_img, scales)
    o.backward(go)
for _ in range(3): fb()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
n = 10
tf = tb = 0.0
for _ in range(n):
    for f in feats: f.grad = None
    e[0].record(); o = region.roi_align_levels(feats, rois, roi_img, scales); e[1].record(); o.backward(go); e[2].record()
    torch.cuda.synchronize()
    tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
R = rois.shape[1]
gbytes = R * 256 * 49 * 4 + sum(f.numel() * 4 for f in feats)
print("R=%d fwd %.1f us  bwd %.1f us  (bwd algorithmic %.0f MB -> %.0f GB/s)" % (R, tf / n * 1e3, tb / n * 1e3, gbytes / 1e6, gbytes / (tb / n * 1e-3) / 1e9))
