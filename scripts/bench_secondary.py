"""Dev probe: device time of the paths of BASELINE configs 1 / 4 / 5 that bench.py does not time (batch 8 where the entry
point is batched): ATSS assignment, plain FCOS targets, RetinaNet dense assignment, RCNN test-time detection."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200det
from b200det import fused, heads as bheads, workload

dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


rng = np.random.default_rng(3)
B, K = 8, 16
strides = [8, 16, 32, 64, 128]
grids = [(-(-800 // s), -(-1344 // s)) for s in strides]
gt = np.zeros((B, 4, K), np.float32); gl = np.zeros((B, K), np.int64)
for b in range(B):
    gt[b], gl[b] = workload.synth_gt(rng, K, 800, 1333)
gtd, gld = T(gt), T(gl)
cnt = torch.full((B,), K, dtype=torch.int32, device=dev)
img_hw = torch.tensor([[800.0, 1333.0]] * B, device=dev)
print("ATSS assignment (config 5: 22 400 points, %d GT, batch %d): %.1f us" % (K, B, timeit(lambda: bheads.atss_assign(grids, strides, gtd, cnt, gld, img_hw))))
print("plain FCOS targets (same sizes): %.1f us" % timeit(lambda: bheads.fcos_targets(grids, strides, gtd, cnt, gld, img_hw)))
# RetinaNet dense assignment: 9 anchors per cell, pos 0.5 / neg 0.4 (+ a 256-sample draw: the batched object always samples)
pyr = fused.AnchorPyramid(strides, grids, tuple(2 ** (i / 3) * 4 for i in range(3)), (0.5, 1.0, 2.0))
bt = fused.BatchedTargets(B, pyr.total, K, dict(pos_iou=0.5, neg_iou=0.4, min_pos_iou=0.0), dict(max_num=256, pos_num=128), (0, 0, 0, 0), (1, 1, 1, 1), dev,
                          pyramid=pyr, border=-1)
print("RetinaNet dense assignment (config 4: %d anchors x %d GT, batch %d): %.1f us" % (pyr.total, K, B, timeit(lambda: bt(gtd, cnt, None, img_hw=img_hw))))
# RCNN test-time detection: 1000 proposals x 81 classes per image
n, C = 1000, 81
props = np.zeros((B, 4, n), np.float32)
for b in range(B):
    cx, cy = rng.uniform(0, 1333, n), rng.uniform(0, 800, n)
    w, h = rng.uniform(16, 400, n), rng.uniform(16, 400, n)
    props[b] = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2])
cls = rng.normal(0, 2.5, (B, n, C)).astype(np.float32); cls[..., 0] += 2.0
reg = rng.normal(0, 0.5, (B, n, 4 * C)).astype(np.float32)
pd, cd, rd = T(props), T(cls), T(reg)
cts = torch.full((B,), n, dtype=torch.int32, device=dev)
print("RCNN detection (1000 proposals x 81 classes, batch %d): %.1f us" % (B, timeit(
    lambda: bheads.rcnn_detect(pd, cd, rd, (800, 1333), [0, 0, 0, 0], [0.1, 0.1, 0.2, 0.2], 0.05, 0.5, 100, "official", counts=cts))))
