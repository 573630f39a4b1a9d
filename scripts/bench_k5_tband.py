"""Dev experiment: the tensor-map band kernel (B2D_ROI_TMA=2) against k_roi_align_win on the config-2 sampled RoIs, split by
the RoI's width in cells (<= 8, 9..16: TMA path with 8 / 16-cell boxes; > 16: window path inside the same kernel)."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200det
from b200det import _C, fused, workload

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
ro = bt.tar_box.permute(1, 0, 2).reshape(4, -1).contiguous()
io = torch.arange(B, dtype=torch.int32, device=dev).repeat_interleave(512).contiguous()
r = ro.cpu().numpy()
s = np.sqrt((r[2] - r[0] + 1) * (r[3] - r[1] + 1))
lv = np.clip(np.floor(np.log2(s / 56 + 1e-6)), 0, 3).astype(int)
stride = np.array([4, 8, 16, 32])[lv]
wcells = np.floor(r[2] / stride).astype(int) - np.floor(r[0] / stride).astype(int) + 2
print("RoIs: %d; width in cells: <=8 %d, 9..16 %d, >16 %d; levels %s" % (r.shape[1], (wcells <= 8).sum(), ((wcells > 8) & (wcells <= 16)).sum(),
                                                                       (wcells > 16).sum(), np.bincount(lv, minlength=4)))
cfg = hp.roi_align.cfg
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(rois, imgs):
    n = rois.shape[1]
    outbuf = torch.zeros((n, 256, 7, 7), device=dev)
    call = lambda: _C.call("b2d_roi_align_fwd", _C.ptr(outbuf), fused._ptrs(feats), _C.ptr(rois), n, _C.ptr(imgs), None, n,
                           ctypes.byref(cfg), _C.stream())
    for _ in range(3): call()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return outbuf, float(np.median(ts))


for name, m in (("all", np.ones_like(wcells, bool)), ("<=8 cells", wcells <= 8), ("9..16 cells", (wcells > 8) & (wcells <= 16)), (">16 cells", wcells > 16)):
    if m.sum() == 0: continue
    rs, ims = T(r[:, m].copy()), T(io.cpu().numpy()[m].copy())
    os.environ.pop("B2D_ROI_TMA", None); _C.reload_knobs()
    o0, t0 = run(rs, ims)
    os.environ["B2D_ROI_TMA"] = "2"; _C.reload_knobs()
    o1, t1 = run(rs, ims)
    print("%-12s n=%5d  window kernel %7.1f us  (%.1f ns/RoI)   tensor-map bands %7.1f us  (%.1f ns/RoI)  identical %s" % (
        name, int(m.sum()), t0, t0 * 1e3 / m.sum(), t1, t1 * 1e3 / m.sum(), bool(torch.equal(o0, o1))), flush=True)
