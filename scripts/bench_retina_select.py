"""Dev probe: the selection stage of AnchorHead.predict_single_image at RetinaNet size (config 4: 9 anchors x 80 classes,
5 levels of a 800 x 1344 image, pre_nms 1000, no NMS in the stage): b2d_rpn_proposals with score_mode 2, do_nms 0."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200det
from b200det import fused

dev = torch.device("cuda:0")
B, A, C = 1, 9, 80
strides = (8, 16, 32, 64, 128)
grids = [(-(-800 // s), -(-1344 // s)) for s in strides]
pyr = fused.AnchorPyramid(strides, grids, tuple(2 ** (i / 3) * 4 for i in range(3)), (0.5, 1.0, 2.0))
cfg = dict(pre_nms=1000, post_nms=0, max_num=0, nms_iou=0.5, min_bbox_size=0)
rp = fused.RpnProposals(pyr, B, cfg, (0, 0, 0, 0), (1, 1, 1, 1), dev, score_mode=2, cls_channels=C, do_nms=False)
g = torch.Generator(device=dev).manual_seed(1)
cls = [torch.randn((B, A * C) + gr, device=dev, generator=g) for gr in grids]
reg = [torch.randn((B, A * 4) + gr, device=dev, generator=g) * 0.3 for gr in grids]
img_hw = torch.tensor([[800.0, 1333.0]] * B, device=dev)
for _ in range(3): rp(cls, reg, img_hw)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rp(cls, reg, img_hw); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("RetinaNet selection stage (%d anchors x %d classes): median %.1f us, count %d, checksum %.6f" % (
    pyr.total, C, float(np.median(ts)), int(rp.count[0]), float(rp.props[0].double().sum())))
