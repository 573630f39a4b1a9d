"""Dev experiment: K5 (k_roi_align_win) on the config-2 sampled RoIs under knob settings given as KEY=VAL[,KEY=VAL] args;
prints the time per launch and whether the output is bit-identical to the default build's.
usage: python scripts/bench_k5_knobs.py B2D_ROI_PF=-1 B2D_ROI_PF=-2 ..."""
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
cfg = hp.roi_align.cfg
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(tag):
    outbuf = torch.zeros((ro.shape[1], 256, 7, 7), device=dev)
    call = lambda: _C.call("b2d_roi_align_fwd", _C.ptr(outbuf), fused._ptrs(feats), _C.ptr(ro), ro.shape[1], _C.ptr(io), None,
                           ro.shape[1], ctypes.byref(cfg), _C.stream())
    for _ in range(3): call()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return outbuf, float(np.median(ts)), float(np.min(ts))


base, med, mn = run("default")
print("%-40s median %.1f us  min %.1f us" % ("default", med, mn), flush=True)
for spec in sys.argv[1:]:
    kv = dict(x.split("=") for x in spec.split(","))
    old = {k: os.environ.get(k) for k in kv}
    os.environ.update(kv); _C.reload_knobs()
    o, med, mn = run(spec)
    print("%-40s median %.1f us  min %.1f us  bit-identical %s" % (spec, med, mn, bool(torch.equal(o, base))), flush=True)
    for k, v in old.items():
        if v is None: os.environ.pop(k)
        else: os.environ[k] = v
    _C.reload_knobs()
