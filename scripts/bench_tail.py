"""Dev tool: phase stamps of k_rpn_back with the RoI-target stage riding on it (B2D_FUSE_TARGETS=1, B2D_DBG=10)."""
import ctypes, os, sys
os.environ["B2D_DBG"] = "10"; os.environ["B2D_FUSE_TARGETS"] = "1"
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200det
from b200det import fused, workload
dev = torch.device("cuda:0")
B, K = 8, 8
w = workload.config2(B=B, K=K, channels=16)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cls, reg = [T(c) for c in w["cls"]], [T(r) for r in w["reg"]]
feats = [T(f).contiguous(memory_format=torch.channels_last) for f in w["feats"]]
gt, gl = T(w["gt"]), T(w["gt_label"])
gcount = torch.full((B,), K, dtype=torch.int32, device=dev)
img_hw = torch.tensor([[800.0, 1333.0]] * B, device=dev)
hp = fused.TrainHotPath(B, w["grids"], dev, gt_ld=K, feat_channels=16)
hp.proposals.ws.zero_()
for _ in range(3): hp.step(cls, reg, feats, gt, gcount, gl, img_hw)
torch.cuda.synchronize()
off = b200det._C.lib().b2d_rpn_proposals_debug_offset(ctypes.byref(hp.pyr.c), B, ctypes.byref(hp.proposals.cfg))
tt = hp.proposals.ws[off:off + B * 64 * 16 * 8].view(torch.int64).view(B, 64, 16).cpu().numpy()
t0 = tt[:, 32:40][tt[:, 32:40] > 0].min()
print("stamps: 0 start, 1 cut, 12 zero, 13 bucket, 2 fence, 14 sweep, 3 fence, 15 sync, 4 scan, 10 K, 11 merge, 5 T-start, 6 pass1, 7 pass2, 8 sync, 9 end")
order = [0, 1, 12, 13, 2, 14, 3, 15, 4, 10, 11, 5, 6, 7, 8, 9]
for c in (32, 33, 39):
    row = tt[0, c]
    print("img 0 cta %d: %s" % (c, " ".join("%5.1f" % ((row[k] - t0) / 1e3) if row[k] > 0 else "    -" for k in order)))
