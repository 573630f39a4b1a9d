"""Dev tool: the RPN proposal stage (K3 + K4) alone in a CUDA graph on the config-2 workload.
usage: python scripts/bench_rpn.py "B2D_NMS_CUT=1.5" "B2D_NMS_CUT=0" ...   (one graph capture + timing + timeline per setting)"""
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
img_hw = torch.tensor([[800.0, 1333.0]] * B, device=dev)
hp = fused.TrainHotPath(B, w["grids"], dev, gt_ld=K, feat_channels=256, layout=1)
hp.proposals.ws.zero_()
step = lambda: hp.proposals(cls, reg, img_hw)
COLD = os.environ.get("RPN_COLD") == "1"            # a 512 MB write before every profiled replay: L2 as RoIAlign leaves it
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if COLD else None
ref = None
for setting in (sys.argv[1:] or [""]):
    kv = dict(x.split("=") for x in setting.split(";") if x)
    for k, v in kv.items(): os.environ[k] = v
    b200det._C.reload_knobs()
    for _ in range(3): step()
    torch.cuda.synchronize()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            out = step()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): g.replay()
    e1.record(); torch.cuda.synchronize()
    props = [o.clone() if torch.is_tensor(o) else o for o in (out if isinstance(out, (tuple, list)) else [out])]
    same = None
    if ref is None: ref = props
    else: same = all(torch.equal(a, b) for a, b in zip(ref, props) if torch.is_tensor(a))
    print("[%s] proposals graph: %.1f us/replay  same_as_first=%s" % (setting, e0.elapsed_time(e1) / 50 * 1e3, same))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            if COLD: flush.zero_(); torch.cuda.synchronize()
            g.replay()
        torch.cuda.synchronize()
    ev = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                 if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda r: r[0])
    reps, cur = [], []
    for r in ev:
        if cur and r[0] - max(x[1] for x in cur) > 30: reps.append(cur); cur = []
        cur.append(r)
    reps.append(cur)
    last = reps[-1]; t0 = last[0][0]
    for s, e, n in last: print("   %7.1f %7.1f  %6.1f  %s" % (s - t0, e - t0, e - s, n[:48]))
    if kv.get("B2D_DBG") == "10":
        import ctypes
        off = b200det._C.lib().b2d_rpn_proposals_debug_offset(ctypes.byref(hp.pyr.c), B, ctypes.byref(hp.proposals.cfg))
        tt = hp.proposals.ws[off:off + B * 64 * 16 * 8].view(torch.int64).view(B, 64, 16).cpu().numpy()
        t0 = tt[tt > 0].min()
        for b_ in (0, B - 1):
            for c in range(64):
                row = tt[b_, c]
                if (row > 0).any():
                    print("   img %d cta %2d: %s" % (b_, c, " ".join("%6.1f" % ((v - t0) / 1e3) if v > 0 else "     -" for v in row[:16])))
    for k in kv: os.environ.pop(k, None)
    b200det._C.reload_knobs()
