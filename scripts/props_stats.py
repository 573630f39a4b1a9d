"""Dev tool: width / x1 statistics of the config-2 proposals and the sweep kernel's candidate counts (simulated)."""
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
props, scores, count = hp.proposals(cls, reg, img_hw)
torch.cuda.synchronize()
p = props[0].cpu().numpy()[:, :int(count[0])]          # [4, n]
x1, y1, x2, y2 = p
wd, ht = x2 - x1, y2 - y1
print("n", p.shape[1], "width pct 10/50/90/99:", np.percentile(wd, [10, 50, 90, 99]).round(1), "height:", np.percentile(ht, [10, 50, 90, 99]).round(1))
print("x1==0 frac %.3f  x2>=1332 frac %.3f" % ((x1 <= 0).mean(), (x2 >= 1332).mean()))
thr = 0.7
xhi = x1 + (1 - 0.9 * thr) * wd
order = np.argsort(x1); xs = x1[order]
cand = np.searchsorted(xs, xhi, side="right") - np.searchsorted(xs, x1, side="left")
print("candidates per box (x-window): mean %.1f  p50 %.0f  p90 %.0f  p99 %.0f  max %d" % (cand.mean(), *np.percentile(cand, [50, 90, 99]), cand.max()))
cells = np.clip(((x1 - x1.min()) * 256 / (x1.max() - x1.min())).astype(int), 0, 255)
cnt = np.bincount(cells, minlength=256)
print("cell counts: max %d, sum sq %d (n*192 = %d)" % (cnt.max(), (cnt.astype(np.int64) ** 2).sum(), p.shape[1] * 192))
c0, c1 = cells, np.clip(((xhi - x1.min()) * 256 / (x1.max() - x1.min())).astype(int), 0, 255)
cs = np.concatenate([[0], np.cumsum(cnt)])
vis = cs[c1 + 1] - cs[c0]
print("cell-window visits per box: mean %.1f p90 %.0f max %d" % (vis.mean(), np.percentile(vis, 90), vis.max()))
