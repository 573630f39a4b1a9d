"""Dev tool: globaltimer phase stamps of the cluster kernels INSIDE the full config-2 step graph (B2D_DBG=10), to compare
with the proposal stage alone (scripts/bench_rpn.py B2D_DBG=10): which phase stretches when the RPN-target kernels and
the previous step's RoIAlign traffic are around.  usage: python scripts/bench_step_phases.py [order]"""
import ctypes, os, sys
os.environ["B2D_DBG"] = "10"
if len(sys.argv) > 1: os.environ["B2D_STEP_ORDER"] = sys.argv[1]
import numpy as np, torch
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
hp = fused.TrainHotPath(B, w["grids"], dev, gt_ld=K, feat_channels=256, layout=1, overlap=True)
hp.proposals.ws.zero_()
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
for _ in range(6): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): g.replay()
e1.record(); torch.cuda.synchronize()
print("step graph with stamps: %.1f us/replay" % (e0.elapsed_time(e1) / 30 * 1e3))
off = b200det._C.lib().b2d_rpn_proposals_debug_offset(ctypes.byref(hp.pyr.c), B, ctypes.byref(hp.proposals.cfg))
tt = hp.proposals.ws[off:off + B * 64 * 16 * 8].view(torch.int64).view(B, 64, 16).cpu().numpy()
t0 = tt[tt > 0].min()
print("front stamps (rows 0..): 0 after load, 1 hist, 2 select, 3 flag, 4 compaction, 5 push, 6 sync, 7 sort, 8 decode")
print("back stamps (rows 32..): order 0 start, 1 cut, 12 zero, 13 bucket, 2 fence, 14 sweep, 3 fence, 15 sync, 4 scan, 10 K, 11 merge")
order = [0, 1, 12, 13, 2, 14, 3, 15, 4, 10, 11]
for b_ in (0, 3, B - 1):
    for c in list(range(0, 14)) + [32, 33, 39]:
        row = tt[b_, c]
        if not (row > 0).any(): continue
        idx = order if c >= 32 else range(9)
        print("   img %d cta %2d: %s" % (b_, c, " ".join("%6.1f" % ((row[k] - t0) / 1e3) if row[k] > 0 else "     -" for k in idx)))
