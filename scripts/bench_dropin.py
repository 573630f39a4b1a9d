"""Dev tool: where the reference call sequence (refpath.TrainCallSequence) spends its time, phase by phase."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200det
from b200det import refpath, workload, heads, bbox, utils

dev = torch.device("cuda:0")
B, K = 8, 8
w = workload.config2(B=B, K=K)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cls, reg = [T(c) for c in w["cls"]], [T(r) for r in w["reg"]]
feats = [T(f) for f in w["feats"]]
gt, gl = T(w["gt"]), T(w["gt_label"])
seq = refpath.TrainCallSequence(w["strides"], dev)
metas = [dict(img_shape=(800, 1333, 3), pad_shape=(800, 1344, 3), scale_factor=1.0) for _ in range(B)]
gtl, gll = [gt[i] for i in range(B)], [gl[i] for i in range(B)]
grid_sizes = [tuple(int(v) for v in c.shape[-2:]) for c in cls]

def phase(name, fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): out = fn()
    torch.cuda.synchronize()
    print("%-28s %.2f ms" % (name, (time.perf_counter() - t0) / n * 1e3))
    return out

la = phase("create_anchors", lambda: seq.create_anchors(grid_sizes))
phase("rpn targets x8", lambda: [seq.rpn_targets_single_image([c[i] for c in cls], [r[i] for r in reg], gtl[i], la, grid_sizes, metas[i]) for i in range(B)])
props = phase("rpn predict x8", lambda: [heads.rpn_predict_single_image(seq.head, [c[i] for c in cls], [r[i] for r in reg], la, metas[i], seq.rpn_proposal)[0] for i in range(B)])
tars = phase("bbox_target x8", lambda: utils.multi_apply(bbox.bbox_target, props, gtl, gll, seq.rcnn_assigner, seq.rcnn_sampler, (0., 0., 0., 0.), (0.1, 0.1, 0.2, 0.2)))
tp = [t[0] for t in tars]
phase("roi_extractor (NCHW)", lambda: seq.extractor(feats, tp))
fcl = [f.contiguous(memory_format=torch.channels_last) for f in feats]
phase("roi_extractor (NHWC)", lambda: seq.extractor(fcl, tp))
phase("whole step", lambda: seq.step(cls, reg, feats, gtl, gll, metas))

# ---- the loops rebound at the method level (batched.py)
from b200det import batched
seqb = refpath.BatchedCallSequence(w["strides"], dev)
ones = [torch.full_like(l, 1) for l in gll]
phase("batched anchor targets", lambda: batched.anchor_head_targets(seqb.head, cls, reg, gtl, ones, metas, seqb.rpn_train_cfg), 20)
bp = phase("batched rpn predict", lambda: batched.rpn_predict_bboxes_from_output(seqb.head, cls, reg, metas, refpath._Cfg(seqb.rpn_proposal))[0], 20)
bt = phase("batched bbox_targets", lambda: batched.bbox_head_bbox_targets(seqb.rcnn_head, bp, gtl, gll, seqb.rcnn_train_cfg), 20)
phase("roi_extractor (NCHW)", lambda: seqb.extractor(feats, bt[0]), 20)
phase("roi_extractor (NHWC)", lambda: seqb.extractor(fcl, bt[0]), 20)
phase("whole batched step (NCHW)", lambda: seqb.step(cls, reg, feats, gtl, gll, metas), 20)
phase("whole batched step (NHWC)", lambda: seqb.step(cls, reg, fcl, gtl, gll, metas), 20)
