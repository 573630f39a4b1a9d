"""Dev experiment: does the processing ORDER of the RoIs matter for K5 (DRAM page / L2 locality)?
Times b2d_roi_align_fwd on the config-2 sampled RoIs in sampler order and in several spatial orders."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200det
from b200det import _C, fused, workload, region

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
torch.cuda.synchronize()
bt = out["rcnn"]
rois = bt.tar_box.permute(1, 0, 2).reshape(4, -1).contiguous()          # [4, B*512]
img = torch.arange(B, dtype=torch.int32, device=dev).repeat_interleave(512).contiguous()
r = rois.cpu().numpy(); im = img.cpu().numpy()
s = np.sqrt((r[2] - r[0] + 1) * (r[3] - r[1] + 1))
lv = np.clip(np.floor(np.log2(s / 56 + 1e-6)), 0, 3).astype(int)
stride = np.array([4, 8, 16, 32])[lv]
cx, cy = (r[0] + r[2]) / 2 / stride, (r[1] + r[3]) / 2 / stride
def morton(x, y):
    x = x.astype(np.uint32); y = y.astype(np.uint32); k = np.zeros_like(x, dtype=np.uint64)
    for b in range(12):
        k |= ((x >> b) & 1).astype(np.uint64) << (2 * b) | ((y >> b) & 1).astype(np.uint64) << (2 * b + 1)
    return k
orders = {
    "sampler order": np.arange(r.shape[1]),
    "img,level,y,x": np.lexsort((cx, cy.astype(int), lv, im)),
    "img,level,y-band8,x": np.lexsort((cx, (cy / 8).astype(int), lv, im)),
    "img,level,morton": np.lexsort((morton(cx, cy), lv, im)),
    "level,img,y,x": np.lexsort((cx, cy.astype(int), im, lv)),
    "random": np.random.default_rng(0).permutation(r.shape[1]),
    # longest-processing-time-first: the cost of a CTA grows with the window size, i.e. with the RoI's size in cells
    "cells desc (LPT)": np.argsort(-((r[2] - r[0]) / stride) * ((r[3] - r[1]) / stride), kind="stable"),
    "cells asc": np.argsort(((r[2] - r[0]) / stride) * ((r[3] - r[1]) / stride), kind="stable"),
    "img, cells desc": np.lexsort((-((r[2] - r[0]) / stride) * ((r[3] - r[1]) / stride), im)),
}
cfg = hp.roi_align.cfg
outbuf = torch.empty((r.shape[1], 256, 7, 7), device=dev)
for name, o in orders.items():
    ro = T(r[:, o].copy()); io = T(im[o].copy())
    call = lambda: _C.call("b2d_roi_align_fwd", _C.ptr(outbuf), fused._ptrs(feats), _C.ptr(ro), ro.shape[1], _C.ptr(io), None,
                           ro.shape[1], ctypes.byref(cfg), _C.stream())
    for _ in range(3): call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): call()
    e1.record(); torch.cuda.synchronize()
    print("%-24s %.1f us" % (name, e0.elapsed_time(e1) / 20 * 1e3))
