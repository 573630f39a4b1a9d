"""Calibration of bench.py's reference arm (authoring container only: needs /root/reference).
The GPU box has no /root/reference, so `bench.py --impl reference` times the C port of the path (oracle/).  This script
times, on THIS machine's cores and on the same config-2 inputs (one image), (a) the UNMODIFIED reference functions
imported through tests/golden/ref_shim.py -- RPNHead.predict_single_image, anchor_target, bbox_target,
BasicRoIExtractor -- wired as CascadeRCNN.forward_train wires them (lib/detectors/cascade_rcnn.py:106-131) and (b) the C
port (oracle.pipeline.ImagePath), so that the port's speed relative to the real reference is on record."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import ref_shim
lib = ref_shim.install()
from lib import anchor as ranchor, bbox as rbbox, region as rregion  # noqa: E402
from lib.builder import build_module  # noqa: E402
import oracle  # noqa: E402
from oracle import pipeline as opipe  # noqa: E402
import importlib.util
spec = importlib.util.spec_from_file_location("workload", os.path.join(ROOT, "pytorch-faster-rcnn_b200", "workload.py"))
workload = importlib.util.module_from_spec(spec); spec.loader.exec_module(workload)

torch.set_num_threads(os.cpu_count())
w = workload.config2(B=1, K=8)
grids, strides = w["grids"], w["strides"]
T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
head = build_module(dict(type="RPNHead", in_channels=8, feat_channels=8, anchor_scales=[8], anchor_ratios=[0.5, 1.0, 2.0],
                         anchor_strides=list(strides), target_means=[0.0] * 4, target_stds=[1.0] * 4,
                         loss_cls=dict(type="CrossEntropyLoss", use_sigmoid=True, loss_weight=1.0),
                         loss_bbox=dict(type="SmoothL1Loss", beta=1.0 / 9.0, loss_weight=1.0)))
ext = rregion.BasicRoIExtractor([dict(type="RoIAlign", spatial_scale=1.0 / s, sampling_ratio=2) for s in strides[:4]], output_size=(7, 7))
cls, reg = [T(c[0]) for c in w["cls"]], [T(r[0]) for r in w["reg"]]
feats = [T(f) for f in w["feats"]]
gt, gl = T(w["gt"][0]), T(w["gt_label"][0])
meta = dict(img_shape=w["img_shape"] + (3,), pad_shape=w["pad_shape"] + (3,), scale_factor=1.0)
anchors = head.create_anchors(grids)
test_cfg = ref_shim.AttrDict(pre_nms=2000, post_nms=2000, max_num=2000, nms_iou=0.7, min_bbox_size=0)
rpn_cfg = ref_shim.AttrDict(assigner=dict(type="MaxIoUAssigner", pos_iou=0.7, neg_iou=0.3, min_pos_iou=0.3),
                            sampler=dict(type="RandomSampler", max_num=256, pos_num=128), allowed_border=0)


def reference_step():
    with torch.no_grad():
        props, _, _ = head.predict_single_image(cls, reg, anchors, meta, test_cfg)
        head.single_image_targets(cls, reg, gt, None, anchors, w["pad_shape"], grids, meta, rpn_cfg)
        tp = rbbox.bbox_target(props, gt, gl, dict(type="MaxIoUAssigner", pos_iou=0.5, neg_iou=0.5, min_pos_iou=0.5),
                               dict(type="RandomSampler", max_num=512, pos_num=128), (0., 0., 0., 0.), (0.1, 0.1, 0.2, 0.2))
        return ext(feats, [tp[0]])


def timed(fn, n):
    fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


np.random.seed(2019)
t_ref = timed(reference_step, 5)
oracle.set_num_threads(os.cpu_count())
path = opipe.ImagePath(grids, strides, w["img_shape"])
a = ([c[0] for c in w["cls"]], [r[0] for r in w["reg"]], [f[0] for f in w["feats"]], w["gt"][0], w["gt_label"][0])
t_port = timed(lambda: path.run(*a), 5)
print("cores %d, torch %s" % (os.cpu_count(), torch.__version__))
print("unmodified reference (lib.*, torch CPU):  %.3f s/image = %.2f images/s" % (t_ref, 1 / t_ref))
print("C port (oracle/, OpenMP %d threads):      %.3f s/image = %.2f images/s" % (oracle.num_threads(), t_port, 1 / t_port))
print("port / reference speed: %.2fx (bench.py's reference arm is the faster of the two: its ratio is conservative)" % (t_ref / t_port))
