import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200det
from b200det import fused
DEV = torch.device("cuda:0")
rng = np.random.default_rng(5)
grids = [(40, 56), (20, 28), (10, 14), (5, 7), (3, 4)]
pyr = fused.AnchorPyramid([4, 8, 16, 32, 64], grids)
B = 2
cls = [torch.from_numpy(rng.normal(0, 1, (B, 3) + g).astype(np.float32)).to(DEV) for g in grids]
reg = [torch.from_numpy(rng.normal(0, 0.5, (B, 12) + g).astype(np.float32)).to(DEV) for g in grids]
img_hw = torch.tensor([[160.0, 213.0]] * B, device=DEV)
for cfg in (dict(pre_nms=300, post_nms=300, max_num=500, nms_iou=0.7, min_bbox_size=0),
            dict(pre_nms=200, post_nms=100, max_num=1000, nms_iou=0.7, min_bbox_size=8),
            dict(pre_nms=0, post_nms=0, max_num=0, nms_iou=0.5, min_bbox_size=0)):
    rp = fused.RpnProposals(pyr, B, cfg, (0, 0, 0, 0), (1, 1, 1, 1), DEV)
    rp(cls, reg, img_hw)
    torch.cuda.synchronize()
    print(cfg, rp.count.tolist())
