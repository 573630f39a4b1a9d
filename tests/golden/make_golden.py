"""Generate golden vectors by EXECUTING THE UNMODIFIED REFERENCE on CPU.

Run in the authoring container only:  python tests/golden/make_golden.py
Writes tests/golden/*.npz (small, committed).  The GPU box never runs this and
never reads /root/reference; the tests there only load the .npz files.

Versions are recorded in tests/golden/VERSIONS.json (the arithmetic of NMS /
RoIAlign / RoIPool lives in torchvision, un-pinned by the reference).
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

lib = ref_shim.install()
from lib import anchor as ranchor, bbox as rbbox, region as rregion, utils as rutils  # noqa: E402
from lib.builder import build_module  # noqa: E402

torch.set_num_threads(8)
SEED = 2019  # test/anchor_target_test.py:12-13


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: (v.shape, str(v.dtype)) for k, v in out.items()})


def sha(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def synth_gt(rng, K, img_h, img_w):
    """SURVEY 8(d) GT recipe."""
    x1 = rng.uniform(0, 0.75 * img_w, K)
    y1 = rng.uniform(0, 0.75 * img_h, K)
    w = rng.uniform(20, 0.4 * img_w, K)
    h = rng.uniform(20, 0.4 * img_h, K)
    x2 = np.minimum(x1 + w, img_w - 1)
    y2 = np.minimum(y1 + h, img_h - 1)
    return np.stack([x1, y1, x2, y2]).astype(np.float32), rng.integers(1, 21, K).astype(np.int64)


def rand_boxes(rng, n, img_h, img_w, smin=8, smax=300):
    cx = rng.uniform(0, img_w, n); cy = rng.uniform(0, img_h, n)
    w = rng.uniform(smin, smax, n); h = rng.uniform(smin, smax, n)
    return np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2]).astype(np.float32)


# --------------------------------------------------------------------------
def g_anchors():
    out = {}
    cases = [(4, (5, 7), [8], [0.5, 1.0, 2.0], False), (16, (3, 4), [8, 16, 32], [0.5, 1.0, 2.0], False),
             (8, (4, 6), [4 * 2 ** (i / 3) for i in range(3)], [0.5, 1.0, 2.0], False),
             (8, (3, 3), [8], [1.0], True), (64, (13, 21), [8], [0.5, 1.0, 2.0], False)]
    meta = []
    for i, (s, g, sc, ar, lt) in enumerate(cases):
        ac = ranchor.AnchorCreator(base=s, scales=sc, aspect_ratios=ar, center_lt=lt)
        out["a%d" % i] = ac(s, g)
        out["ws%d" % i] = ac.anchor_ws
        out["hs%d" % i] = ac.anchor_hs
        meta.append(json.dumps(dict(stride=s, grid=g, scales=sc, ratios=ar, center_lt=lt)))
    out["meta"] = np.array(meta)
    # full-size R50-FPN RPN pyramid at 800x1344: per-level sha1 + float64 sums
    grids = [(200, 336), (100, 168), (50, 84), (25, 42), (13, 21)]
    for l, (s, g) in enumerate(zip([4, 8, 16, 32, 64], grids)):
        a = ranchor.AnchorCreator(base=s, scales=[8], aspect_ratios=[0.5, 1.0, 2.0])(s, g).numpy()
        out["full_sha%d" % l] = sha(a)
        out["full_sum%d" % l] = a.astype(np.float64).reshape(4, -1).sum(1)
    # masks (lib/region.py:10-29) on the stride-64 level, img 800x1333
    a = ranchor.AnchorCreator(base=64, scales=[8], aspect_ratios=[0.5, 1.0, 2.0])(64, (13, 21))
    out["mask_img_b0"] = rregion.inside_anchor_mask(a.view(4, -1), (800, 1333), 0)
    out["mask_img_b16"] = rregion.inside_anchor_mask(a.view(4, -1), (800, 1333), 16)
    out["mask_grid"] = rregion.inside_grid_mask(3, (800, 1333), (13, 21), 64)
    out["mask_grid_small"] = rregion.inside_grid_mask(3, (500, 700), (13, 21), 64)
    save("anchors", **out)


def fpn_anchors(grids, strides=(4, 8, 16, 32, 64), scales=(8,), ratios=(0.5, 1.0, 2.0)):
    acs = [ranchor.AnchorCreator(base=s, scales=list(scales), aspect_ratios=list(ratios)) for s in strides]
    return [ac(s, g) for ac, s, g in zip(acs, strides, grids)]


def g_iou_assign():
    rng = np.random.default_rng(SEED)
    H, W = 800, 1333
    gt, _ = synth_gt(rng, 8, H, W)
    boxes = rand_boxes(rng, 1500, H, W)
    # make some boxes coincide with / jitter around GT so positives exist
    for j in range(8):
        boxes[:, 10 * j] = gt[:, j]
        for r in range(1, 10):
            boxes[:, 10 * j + r] = gt[:, j] + rng.uniform(-12, 12, 4).astype(np.float32)
    out = dict(boxes=boxes, gt=gt)
    out["iou"] = rutils.calc_iou(T(boxes), T(gt))
    out["elem_iou"] = rutils.elem_iou(T(boxes[:, :8]), T(gt))
    cfgs = [(0.7, 0.3, 0.3), (0.5, 0.5, 0.5), (0.5, 0.4, 0.0), (0.6, 0.6, 0.6)]
    out["cfgs"] = np.array(cfgs, np.float64)
    for i, c in enumerate(cfgs):
        lab, iou = rregion.MaxIoUAssigner(*c)(T(boxes), T(gt))
        out["labels%d" % i] = lab
        out["miou%d" % i] = iou
    # SURVEY A4 verified example: anchors {A,B,C,A} vs GTs {A, far-away}
    A = [10, 10, 50, 50]; B = [200, 200, 260, 280]; C = [400, 100, 450, 180]; far = [1000, 600, 1100, 700]
    ex_b = np.array([A, B, C, A], np.float32).T.copy(); ex_g = np.array([A, far], np.float32).T.copy()
    out["ex_boxes"], out["ex_gt"] = ex_b, ex_g
    for i, mp in enumerate([0.0, 0.3]):
        lab, iou = rregion.MaxIoUAssigner(0.5, 0.4, mp)(T(ex_b), T(ex_g))
        out["ex_labels%d" % i], out["ex_miou%d" % i] = lab, iou
    # test/bbox_test.py:27-31 style 3-box literal
    lit = np.array([[0, 0, 10, 10], [5, 5, 15, 15], [20, 20, 30, 30]], np.float32).T.copy()
    out["lit"] = lit
    out["lit_iou"] = rutils.calc_iou(T(lit), T(lit))
    save("iou_assign", **out)

    # full-size RPN assignment: 268 569 anchors -> valid mask -> assign, K=8 (config 2)
    grids = [(200, 336), (100, 168), (50, 84), (25, 42), (13, 21)]
    anchors = torch.cat([a.view(4, -1) for a in fpn_anchors(grids)], 1)
    in_img = rregion.inside_anchor_mask(anchors, (H, W), 0)
    in_grid = torch.cat([rregion.inside_grid_mask(3, (H, W), g, s) for g, s in zip(grids, [4, 8, 16, 32, 64])])
    mask = in_img & in_grid.bool()
    full = dict(gt=gt, mask_packed=np.packbits(mask.numpy()), n=np.int64(anchors.shape[1]))
    for i, c in enumerate([(0.7, 0.3, 0.3), (0.5, 0.4, 0.0)]):
        lab, iou = rregion.MaxIoUAssigner(*c)(anchors[:, mask], T(gt))
        full["labels%d" % i] = lab.numpy().astype(np.int8)
        nz = np.nonzero(iou.numpy())[0]
        full["iou_nz_idx%d" % i] = nz.astype(np.int32)
        full["iou_nz_val%d" % i] = iou.numpy()[nz]
        full["iou_sha%d" % i] = sha(iou.numpy() + np.float32(0.0))  # +0.0 folds -0.0 into +0.0
    full["cfgs"] = np.array([(0.7, 0.3, 0.3), (0.5, 0.4, 0.0)])
    save("assign_full", **full)


def g_deltas():
    rng = np.random.default_rng(SEED + 1)
    base = rand_boxes(rng, 512, 800, 1333)
    bbox = base + rng.uniform(-20, 20, base.shape).astype(np.float32)
    bbox[2:] = np.maximum(bbox[2:], bbox[:2] + 1)
    out = dict(base=base, bbox=bbox)
    stds = [0.1, 0.1, 0.2, 0.2]; means = [0.0, 0.0, 0.0, 0.0]
    out["enc_plain"] = rutils.bbox2param(T(base), T(bbox))
    out["enc_norm"] = rutils.bbox2param(T(base), T(bbox), means, stds)
    param = rng.normal(0, 0.5, base.shape).astype(np.float32)
    out["param"] = param
    out["dec_plain"] = rutils.param2bbox(T(base), T(param))
    out["dec_norm_clamp"] = rutils.param2bbox(T(base), T(param), means, stds, (800, 1333))
    out["dec_clamp3"] = rutils.param2bbox(T(base), T(param), [0.1, -0.1, 0.05, 0.0], [0.05, 0.05, 0.1, 0.1], (800, 1333, 3))
    # round trip of test/bbox_test.py:13-25
    out["roundtrip"] = rutils.param2bbox(T(base), rutils.bbox2param(T(base), T(bbox)))
    # batched_param2bbox: reg [4*cls, n] viewed (4, cls, n)  (lib/utils.py:96-106)
    bp = rng.normal(0, 0.5, (4 * 5, 64)).astype(np.float32)
    out["bparam"] = bp
    out["bdec"] = rutils.batched_param2bbox(T(base[:, :64].copy()), T(bp), means, stds, (800, 1333))
    out["clamp"] = rutils.clamp_bbox(T(bbox - 100), (600, 1000))
    save("deltas", **out)


def g_nms():
    rng = np.random.default_rng(SEED + 2)
    out = {}
    # clustered boxes, distinct scores
    ctr = rand_boxes(rng, 40, 800, 1333, 30, 250)
    b = np.repeat(ctr, 50, axis=1) + rng.normal(0, 6, (4, 2000)).astype(np.float32)
    b[2:] = np.maximum(b[2:], b[:2] + 1)
    s = rng.permutation(2000).astype(np.float32) / 2000
    out["b0"], out["s0"] = b.T.copy(), s
    for thr in (0.7, 0.5, 0.3):
        out["keep0_%d" % int(thr * 10)] = torchvision.ops.nms(T(b.T.copy()), T(s), thr)
    # ties in scores (stable: lower index first) + exact-threshold pair + degenerate boxes
    b1 = np.array([[0, 0, 10, 10], [0, 0, 3, 10], [0, 0, 10, 10], [5, 5, 5, 5], [0, 0, 7, 10],
                   [100, 100, 90, 90], [0, 0, 10, 10]], np.float32)
    s1 = np.array([0.5, 0.9, 0.5, 0.5, 0.9, 0.2, 0.1], np.float32)
    out["b1"], out["s1"] = b1, s1
    for thr in (0.3, 0.7, 0.5, 0.0):
        out["keep1_%d" % int(thr * 10)] = torchvision.ops.nms(T(b1), T(s1), thr)
    out["keep_empty"] = torchvision.ops.nms(torch.zeros(0, 4), torch.zeros(0), 0.5)
    # batched / multiclass (lib/utils.py:211-269)
    n, C = 300, 6
    bb = rand_boxes(rng, n, 600, 800, 20, 200).T.copy()
    sc = rng.uniform(0, 1, (n, C)).astype(np.float32) ** 3
    lab = rng.integers(0, C, n)
    kb, ks, kl = rutils.batched_nms(T(bb), T(sc[:, 0].copy()), T(lab), 0.5)
    out["bn_bbox"], out["bn_score"], out["bn_label"] = bb, sc, lab
    out["bn_kb"], out["bn_ks"], out["bn_kl"] = kb, ks, kl
    for mode in ("official", "strict"):
        kb, ks, kl = rutils.multiclass_nms(T(bb), T(sc), list(range(1, C)), 0.5, 0.05, 100, mode=mode)
        out["mc_%s_b" % mode], out["mc_%s_s" % mode], out["mc_%s_l" % mode] = kb, ks, kl
    # per-class boxes [n, 4*C] + score_factor
    bbc = np.repeat(bb[:, :, None], C, 2) + rng.normal(0, 3, (n, 4, C)).astype(np.float32)
    bbc = bbc.reshape(n, 4 * C)
    fac = rng.uniform(0.2, 1, n).astype(np.float32)
    out["mc_bbc"], out["mc_fac"] = bbc, fac
    kb, ks, kl = rutils.multiclass_nms(T(bbc), T(sc), list(range(1, C)), 0.5, 0.05, 50, score_factor=T(fac))
    out["mc_pc_b"], out["mc_pc_s"], out["mc_pc_l"] = kb, ks, kl
    kb, ks, kl = rutils.multiclass_nms(T(bbc), T(sc), list(range(1, C)), 0.5, 0.05, 50, mode="strict")
    out["mc_pcs_b"], out["mc_pcs_s"], out["mc_pcs_l"] = kb, ks, kl
    save("nms", **out)


def make_rpn_head():
    return build_module(dict(type="RPNHead", in_channels=8, feat_channels=8, anchor_scales=[8],
                             anchor_ratios=[0.5, 1.0, 2.0], anchor_strides=[4, 8, 16, 32, 64],
                             target_means=[0.0, 0.0, 0.0, 0.0], target_stds=[1.0, 1.0, 1.0, 1.0],
                             loss_cls=dict(type="CrossEntropyLoss", use_sigmoid=True, loss_weight=1.0),
                             loss_bbox=dict(type="SmoothL1Loss", beta=1.0 / 9.0, loss_weight=1.0)))


def g_rpn():
    rng = np.random.default_rng(SEED + 3)
    head = make_rpn_head()
    img_shape, pad_shape = (160, 213, 3), (160, 224, 3)
    grids = [(40, 56), (20, 28), (10, 14), (5, 7), (3, 4)]
    cls = [rng.normal(0, 1, (3,) + g).astype(np.float32) for g in grids]
    reg = [rng.normal(0, 0.5, (12,) + g).astype(np.float32) for g in grids]
    out = dict(img_shape=np.array(img_shape), pad_shape=np.array(pad_shape))
    for l in range(5):
        out["cls%d" % l], out["reg%d" % l] = cls[l], reg[l]
    anchors = head.create_anchors(grids)
    meta = dict(img_shape=img_shape, pad_shape=pad_shape, scale_factor=1.0)
    cfgs = [dict(pre_nms=300, post_nms=300, max_num=500, nms_iou=0.7, min_bbox_size=0),
            dict(pre_nms=200, post_nms=100, max_num=1000, nms_iou=0.7, min_bbox_size=8),
            dict(pre_nms=0, post_nms=0, max_num=0, nms_iou=0.5, min_bbox_size=0)]
    out["cfgs"] = np.array([json.dumps(c) for c in cfgs])
    for i, c in enumerate(cfgs):
        with torch.no_grad():
            b, s, _ = head.predict_single_image([T(x) for x in cls], [T(x) for x in reg], anchors, meta,
                                                ref_shim.AttrDict(c))
        out["props%d" % i], out["scores%d" % i] = b, s
    # legacy ProposalCreator (lib/region.py:175-209): 2-channel softmax
    A = 3
    pc_cls = rng.normal(0, 1, (1, 2 * A, 10, 14)).astype(np.float32)
    pc_reg = rng.normal(0, 0.3, (1, 4 * A, 10, 14)).astype(np.float32)
    pc_anchor = ranchor.AnchorCreator(base=16, scales=[2, 4, 8], aspect_ratios=[1.0])(16, (10, 14)).view(4, -1)
    pb, ps = rregion.ProposalCreator(200, 50, 0.7, 4)(T(pc_cls), T(pc_reg), pc_anchor, (160, 213))
    out.update(pc_cls=pc_cls, pc_reg=pc_reg, pc_anchor=pc_anchor, pc_props=pb, pc_scores=ps)
    save("rpn", **out)


def g_roi():
    rng = np.random.default_rng(SEED + 4)
    C = 8
    grids = [(40, 56), (20, 28), (10, 14), (5, 7)]
    feats = [rng.normal(0, 1, (1, C) + g).astype(np.float32) for g in grids]
    rois = rand_boxes(rng, 96, 160, 213, 4, 220)
    rois[[0, 2]] = np.clip(rois[[0, 2]], 0, 212); rois[[1, 3]] = np.clip(rois[[1, 3]], 0, 159)
    rois[2:] = np.maximum(rois[2:], rois[:2])
    # include out-of-image RoIs to exercise the (-1, H] sample rule
    rois[:, 0] = [-30, -30, 20, 20]; rois[:, 1] = [200, 150, 260, 200]; rois[:, 2] = [50, 50, 50, 50]
    ext = build_module(dict(type="BasicRoIExtractor",
                            roi_layers=[dict(type="RoIAlign", spatial_scale=1 / s, sampling_ratio=2)
                                        for s in (4, 8, 16, 32)], output_size=(7, 7)))
    tf = [T(f).requires_grad_(True) for f in feats]
    o = ext(tf, [T(rois)])[0]
    gout = rng.normal(0, 1, tuple(o.shape)).astype(np.float32)
    (o * T(gout)).sum().backward()
    out = dict(rois=rois, out=o, gout=gout, lvls=ext.map_rois_to_levels(T(rois), 4))
    for l in range(4):
        out["feat%d" % l] = feats[l]
        out["gfeat%d" % l] = tf[l].grad
    # level-map literal of test/map2level_test.py:13 (sides) -> square boxes
    sides = np.array([20, 48, 66, 127, 1000, 200000, 111, 112, 113, 223, 224, 225, 447, 448, 449], np.float32)
    lm = np.stack([np.zeros_like(sides), np.zeros_like(sides), sides - 1, sides - 1])
    out["lm_boxes"] = lm
    out["lm_lvls"] = ext.map_rois_to_levels(T(lm), 4)
    # single-level RoIAlign (adaptive sampling_ratio=0, aligned True) straight from torchvision
    r5 = torch.cat([torch.zeros(96, 1), T(rois).t()], 1)
    out["ra_adapt"] = torchvision.ops.roi_align(T(feats[1]), r5, (7, 7), 1 / 8, 0, False)
    out["ra_aligned"] = torchvision.ops.roi_align(T(feats[1]), r5, (5, 3), 1 / 8, 2, True)
    # RoIPool (C4 config; reference bug: configs/faster_rcnn_r50.py:26 passes sampling_ratio -> popped)
    ext_p = build_module(dict(type="BasicRoIExtractor", roi_layers=[dict(type="RoIPool", spatial_scale=1 / 16)],
                              output_size=(7, 7)))
    fp = T(feats[2]).requires_grad_(True)
    op = ext_p([fp], [T(rois)])[0]
    (op * T(gout)).sum().backward()
    out["pool_out"], out["pool_gfeat"] = op, fp.grad
    # SingleRoIExtractor + ScalableRoIAlign (registered, lib/region.py:212-239,309-375)
    sre = build_module(dict(type="SingleRoIExtractor", roi_layer="RoIAlign", output_size=7,
                            featmap_strides=[4, 8, 16, 32]))
    out["single_out"] = sre([T(f) for f in feats], [T(rois)])[0]
    sra = build_module(dict(type="ScalableRoIAlign", scale=1.5, output_size=(7, 7), spatial_scale=1 / 8,
                            sampling_ratio=2))
    out["scalable_out"] = sra(T(feats[1]), r5)
    save("roi", **out)


def g_targets():
    rng = np.random.default_rng(SEED + 5)
    H, W = 160, 213
    gt, gl = synth_gt(rng, 5, H, W)
    grids = [(40, 56), (20, 28), (10, 14), (5, 7), (3, 4)]
    head = make_rpn_head()
    cls = [rng.normal(0, 1, (1, 3) + g).astype(np.float32) for g in grids]
    reg = [rng.normal(0, 0.5, (1, 12) + g).astype(np.float32) for g in grids]
    meta = dict(img_shape=(H, W, 3), pad_shape=(160, 224, 3), scale_factor=1.0)
    tcfg = ref_shim.AttrDict(assigner=dict(type="MaxIoUAssigner", pos_iou=0.7, neg_iou=0.3, min_pos_iou=0.3),
                             sampler=dict(type="RandomSampler", max_num=64, pos_num=32), allowed_border=0)
    anchors = head.create_anchors(grids)
    np.random.seed(SEED)
    r = head.single_image_targets([T(c[0]) for c in cls], [T(x[0]) for x in reg], T(gt), None, anchors,
                                  (160, 224), grids, meta, tcfg)
    out = dict(gt=gt, gl=gl)
    for l in range(5):
        out["cls%d" % l], out["reg%d" % l] = cls[l], reg[l]
    for k, v in zip(("at_cls", "at_reg", "at_lab", "at_par"), r):
        out[k] = v
    # no-sampler + class labels (RetinaNet form, lib/anchor.py:60-64)
    tcfg2 = ref_shim.AttrDict(assigner=dict(type="MaxIoUAssigner", pos_iou=0.5, neg_iou=0.4, min_pos_iou=0.0),
                              allowed_border=-1)
    r2 = head.single_image_targets([T(c[0]) for c in cls], [T(x[0]) for x in reg], T(gt), T(gl), anchors,
                                   (160, 224), grids, meta, tcfg2)
    for k, v in zip(("at2_cls", "at2_reg", "at2_lab", "at2_par"), r2):
        out[k] = v
    # bbox_target (lib/bbox.py) with the host numpy RNG
    props = rand_boxes(rng, 400, H, W, 10, 120)
    for j in range(5):
        for q in range(12):
            props[:, 12 * j + q] = gt[:, j] + rng.uniform(-8, 8, 4).astype(np.float32)
    props[2:] = np.maximum(props[2:], props[:2] + 1)
    out["props"] = props
    np.random.seed(SEED)
    r3 = rbbox.bbox_target(T(props), T(gt), T(gl), dict(type="MaxIoUAssigner", pos_iou=0.5, neg_iou=0.5, min_pos_iou=0.5),
                           dict(type="RandomSampler", max_num=128, pos_num=32), (0., 0., 0., 0.), (0.1, 0.1, 0.2, 0.2))
    for k, v in zip(("bt_props", "bt_bbox", "bt_label", "bt_param", "bt_isgt"), r3):
        out[k] = v
    # samplers alone, on a labels vector (host numpy RNG stream)
    labels = T(rng.choice([-1, 0, 0, 0, 1, 2, 3], 3000).astype(np.int64))
    np.random.seed(SEED)
    out["rs_in"] = labels
    out["rs_out"] = rregion.RandomSampler(256, 64)(labels.clone())
    ious = T(rng.uniform(0, 1, 3000).astype(np.float32))
    np.random.seed(SEED)
    out["ib_iou"] = ious
    out["ib_out"] = rregion.IoUBalancedNegSampler(256, 64)(labels.clone(), ious, None, None)
    save("targets", **out)


# --------------------------------------------------------------------------
def _patched_fcos_module():
    """lib.heads.fcos_head with the reference's known bug worked around, documented in SURVEY 8(c):
    topk_by_center line 115 `k_inds / w` is true division on torch >= 1.5 (float indices ->
    IndexError at :342); the oracle is the same function with `//`.  The function source is
    re-executed from the reference file with that one token changed; nothing is copied."""
    import inspect
    import lib.heads.fcos_head as fh
    if getattr(fh, "_b2d_patched", False):
        return fh
    fh._b2d_patched = True
    src = inspect.getsource(fh.topk_by_center)
    assert "k_inds / w" in src
    exec(compile(src.replace("k_inds / w", "k_inds // w"), "<topk_by_center //>", "exec"), fh.__dict__)
    return fh


def g_atss():
    import types
    fh = _patched_fcos_module()
    rng = np.random.default_rng(SEED + 11)
    strides = [8, 16, 32, 64, 128]
    out = {}
    for tag, img_shape, pad, K in (("s", (500, 597, 3), (512, 640), 6), ("f", (800, 1333, 3), (800, 1344), 16)):
        grids = [(-(-pad[0] // s), -(-pad[1] // s)) for s in strides]
        gt, gl = synth_gt(rng, K, img_shape[0], img_shape[1])
        creators = [ranchor.AnchorCreator(base=s, scales=[8], aspect_ratios=[1.0]) for s in strides]
        lvl_anchors = tuple(creators[i](strides[i], grids[i]).squeeze() for i in range(5))
        me = types.SimpleNamespace(strides=strides, atss_cfg=ref_shim.AttrDict(topk=9, scale=8))
        dummy = [torch.zeros((20,) + g) for g in grids]
        with torch.no_grad():
            cls_t, reg_t, ctr_t = fh.FCOSHead.single_image_targets_atss(
                me, dummy, dummy, dummy, lvl_anchors, T(gt), T(gl), dict(img_shape=img_shape), None)
        out.update({"gt_" + tag: gt, "gl_" + tag: gl, "img_" + tag: np.array(img_shape), "grids_" + tag: np.array(grids),
                    "cls_" + tag: torch.cat([c.reshape(-1) for c in cls_t]),
                    "reg_" + tag: torch.cat([r.reshape(-1, 4) for r in reg_t]),
                    "ctr_" + tag: torch.cat([c.reshape(-1) for c in ctr_t])})
        print("atss", tag, "positives", int((out["cls_" + tag] > 0).sum()))
        # plain FCOS targets (lib/heads/fcos_head.py:371-416) on the same GTs
        me2 = types.SimpleNamespace(strides=strides, level_scale_thr=[0, 64, 128, 256, 512, 1e6])
        with torch.no_grad():
            cls_p, reg_p, ctr_p = fh.FCOSHead.single_image_targets(me2, dummy, dummy, dummy, T(gt), T(gl),
                                                                   dict(img_shape=img_shape), None)
        out.update({"pcls_" + tag: torch.cat([c.reshape(-1) for c in cls_p]),
                    "preg_" + tag: torch.cat([r.reshape(-1, 4) for r in reg_p]),
                    "pctr_" + tag: torch.cat([c.reshape(-1) for c in ctr_p])})
        print("fcos plain", tag, "positives", int((out["pcls_" + tag] > 0).sum()))
    save("atss", **out)


def g_heads():
    """a17 cascade refine (lib/heads/bbox_head.py:100-120) and a19 FCOS predict (lib/heads/fcos_head.py:570-631)."""
    import types
    import lib.heads.bbox_head as bh
    fh = _patched_fcos_module()
    rng = np.random.default_rng(SEED + 12)
    out = {}
    # ---- refine_bboxes_single_image: 21 classes, per-class deltas, first 5 columns are GT
    s, C = 200, 21
    props = rand_boxes(rng, s, 400, 600, 8, 200)
    label = rng.integers(0, C, s).astype(np.int64)
    reg_out = rng.normal(0, 1, (s, 4 * C)).astype(np.float32)
    is_gt = np.zeros(s, np.int64); is_gt[:5] = 1; is_gt[37] = 1
    me = types.SimpleNamespace(reg_class_agnostic=False, num_classes=C, target_means=[0.0, 0.0, 0.0, 0.0],
                               target_stds=[0.05, 0.05, 0.1, 0.1])
    with torch.no_grad():
        ref = bh.BBoxHead.refine_bboxes_single_image(me, T(props), T(label), T(reg_out), T(is_gt), dict(img_shape=(400, 600, 3)))
        me2 = types.SimpleNamespace(reg_class_agnostic=True, num_classes=C, target_means=[0.0, 0.0, 0.0, 0.0],
                                    target_stds=[0.1, 0.1, 0.2, 0.2])
        ref2 = bh.BBoxHead.refine_bboxes_single_image(me2, T(props), T(label), T(reg_out[:, :4].copy()), None, None)
    out.update(rf_props=props, rf_label=label, rf_reg=reg_out, rf_is_gt=is_gt, rf_out=ref, rf_out_agnostic=ref2)
    # ---- FCOS predict_single_image (centerness, strict multiclass NMS)
    strides = [8, 16, 32, 64, 128]
    img_shape, pad = (250, 317, 3), (256, 320)
    grids = [(-(-pad[0] // st) , -(-pad[1] // st)) for st in strides]
    cls = [rng.normal(-2, 2, (20,) + g).astype(np.float32) for g in grids]
    reg = [np.abs(rng.normal(0.15, 0.1, (4,) + g)).astype(np.float32) for g in grids]
    ctr = [rng.normal(0, 1, (1,) + g).astype(np.float32) for g in grids]
    me = types.SimpleNamespace(use_centerness=True, use_dfl=False, strides=strides, reg_std=300, reg_mean=0, cls_channels=20)
    for i, cfg in enumerate([dict(pre_nms=1000, min_bbox_size=0, min_score=0.05, nms_iou=0.6, nms_type="strict", max_per_img=100),
                             dict(pre_nms=50, min_bbox_size=40, min_score=0.3, nms_iou=0.5, nms_type="official", max_per_img=60)]):
        with torch.no_grad():
            b, sc, lab = fh.FCOSHead.predict_single_image(me, [T(x) for x in cls], [T(x.copy()) for x in reg], [T(x) for x in ctr],
                                                          dict(img_shape=img_shape, scale_factor=1.0), ref_shim.AttrDict(cfg))
        out.update({"fc_bbox%d" % i: b, "fc_score%d" % i: sc, "fc_label%d" % i: lab})
        print("fcos predict", i, "detections", int(sc.numel()))
    for l in range(5):
        out["fc_cls%d" % l], out["fc_reg%d" % l], out["fc_ctr%d" % l] = cls[l], reg[l], ctr[l]
    out["fc_img"] = np.array(img_shape)
    # ---- BBoxHead.predict_bboxes_single_image (lib/heads/bbox_head.py:122-146), 21 classes
    n, C = 300, 21
    dp = rand_boxes(rng, n, 400, 600, 16, 250)
    dcls = rng.normal(0, 2.5, (n, C)).astype(np.float32)
    dcls[:, 0] += 2.0                                       # most proposals are background
    dreg = rng.normal(0, 0.6, (n, 4 * C)).astype(np.float32)
    me = types.SimpleNamespace(use_sigmoid=False, reg_class_agnostic=False, num_classes=C,
                               target_means=[0.0, 0.0, 0.0, 0.0], target_stds=[0.1, 0.1, 0.2, 0.2])
    out.update(det_props=dp, det_cls=dcls, det_reg=dreg)
    for i, cfg in enumerate([dict(min_score=0.05, nms_iou=0.5, max_per_img=100, nms_type="official"),
                             dict(min_score=0.2, nms_iou=0.3, max_per_img=40, nms_type="strict")]):
        b, sc, lab = bh.BBoxHead.predict_bboxes_single_image(me, T(dp), T(dcls), T(dreg), (400, 600), ref_shim.AttrDict(cfg))
        out.update({"det_bbox%d" % i: b, "det_score%d" % i: sc, "det_label%d" % i: lab})
        print("rcnn detect", i, "detections", int(sc.numel()))
    save("heads", **out)


def g_loss():
    """AnchorHead.calc_loss pieces without sampler (RetinaNet form): anchor_target (lib/anchor.py:11-76)
    -> sigmoid_focal_loss (lib/losses.py:33-61) + smooth_l1_loss_v2 (:77-83), with torch autograd gradients
    w.r.t. the head maps."""
    from lib import losses as rlosses
    rng = np.random.default_rng(SEED + 21)
    strides, grids = [8, 16, 32], [(20, 28), (10, 14), (5, 7)]
    scales, ratios, C, K = [4.0, 4.0 * 2 ** (1 / 3), 4.0 * 2 ** (2 / 3)], [0.5, 1.0, 2.0], 5, 6
    A = 9
    H, W = 160, 224
    gt, _ = synth_gt(rng, K, H, W)
    gl = rng.integers(1, C + 1, K).astype(np.int64)
    creators = [ranchor.AnchorCreator(base=s, scales=scales, aspect_ratios=ratios) for s in strides]
    anchors = torch.cat([creators[i](strides[i], grids[i]).view(4, -1) for i in range(3)], 1)
    cls = [torch.from_numpy(rng.normal(-1.5, 1.5, (A * C,) + g).astype(np.float32)).requires_grad_(True) for g in grids]
    reg = [torch.from_numpy(rng.normal(0, 0.5, (A * 4,) + g).astype(np.float32)).requires_grad_(True) for g in grids]
    cls_out = torch.cat([c.view(C, -1) for c in cls], 1)
    reg_out = torch.cat([r.view(4, -1) for r in reg], 1)
    in_mask = torch.ones(anchors.shape[1], dtype=torch.bool)
    means, stds = [0.0, 0.0, 0.0, 0.0], [1.0, 1.0, 1.0, 1.0]
    tar_cls_out, tar_reg_out, tar_labels, _, _, tar_param = ranchor.anchor_target(
        cls_out, reg_out, C, anchors, in_mask, T(gt), T(gl),
        dict(type="MaxIoUAssigner", pos_iou=0.5, neg_iou=0.4, min_pos_iou=0.0), None, means, stds)
    pos = tar_labels > 0
    focal = rlosses.sigmoid_focal_loss(tar_cls_out.t(), tar_labels, alpha=0.25, gamma=2.0)
    sl1 = rlosses.smooth_l1_loss_v2(tar_reg_out[:, pos], tar_param[:, pos], 1.0 / 9.0)
    (2.0 * focal + 3.0 * sl1).backward()
    out = dict(gt=gt, gl=gl, focal=focal.detach(), sl1=sl1.detach(), npos=np.array(int(pos.sum())),
               scales=np.array(scales), tar_labels=tar_labels)
    for l in range(3):
        out["cls%d" % l], out["reg%d" % l] = cls[l].detach(), reg[l].detach()
        out["dcls%d" % l], out["dreg%d" % l] = cls[l].grad, reg[l].grad
    print("loss golden: focal %.4f sl1 %.4f npos %d kept %d" % (float(focal), float(sl1), int(pos.sum()), int(tar_labels.numel())))
    save("loss", **out)


def c4_inputs():
    """Seeded inputs of BASELINE config 1 (faster_rcnn_r50 C4 inference, 600x1000 -> 608x1024, one level of 38x64x12
    anchors, 1024-channel features); regenerated identically by tests/test_gpu_heads.py (the 10 MB feature map is not
    stored)."""
    rng = np.random.default_rng(SEED + 31)
    cls = rng.normal(0, 1, (12, 38, 64)).astype(np.float32)
    reg = rng.normal(0, 0.3, (48, 38, 64)).astype(np.float32)
    feat = rng.standard_normal((1, 1024, 38, 64), dtype=np.float32)
    cls_out = rng.normal(0, 2.5, (300, 21)).astype(np.float32)
    cls_out[:, 0] += 2.0
    reg_out = rng.normal(0, 0.5, (300, 84)).astype(np.float32)
    return cls, reg, feat, cls_out, reg_out


def g_c4():
    """BASELINE config 1 end to end behind the backbone / head convolutions: RPNHead.predict_single_image with
    test_cfg.rpn 6000/300/300/0.7 (configs/faster_rcnn_r50.py:96-102) -> BasicRoIExtractor(RoIPool 7x7 @1/16) ->
    BBoxHead.predict_bboxes_single_image with test_cfg.rcnn (min_score .05, nms .3, 100)."""
    import types
    import lib.heads.bbox_head as bh
    cls, reg, feat, cls_out, reg_out = c4_inputs()
    head = build_module(dict(type="RPNHead", in_channels=8, feat_channels=8, anchor_scales=[4, 8, 16, 32],
                             anchor_ratios=[0.5, 1.0, 2.0], anchor_strides=[16], target_means=[0.0] * 4, target_stds=[1.0] * 4,
                             loss_cls=dict(type="CrossEntropyLoss", use_sigmoid=True, loss_weight=1.0),
                             loss_bbox=dict(type="SmoothL1Loss", beta=1.0 / 9.0, loss_weight=1.0)))
    meta = dict(img_shape=(600, 1000, 3), pad_shape=(608, 1024, 3), scale_factor=1.0)
    anchors = head.create_anchors([(38, 64)])
    with torch.no_grad():
        props, scores, _ = head.predict_single_image([T(cls)], [T(reg)], anchors, meta,
                                                     ref_shim.AttrDict(pre_nms=6000, post_nms=300, max_num=300, nms_iou=0.7,
                                                                       min_bbox_size=0.0))
        ext = build_module(dict(type="BasicRoIExtractor", roi_layers=[dict(type="RoIPool", spatial_scale=1 / 16)],
                                output_size=(7, 7)))
        pooled = ext([T(feat)], [props])[0]
        n = props.shape[1]
        me = types.SimpleNamespace(use_sigmoid=False, reg_class_agnostic=False, num_classes=21,
                                   target_means=[0.0] * 4, target_stds=[0.1, 0.1, 0.2, 0.2])
        db, ds, dl = bh.BBoxHead.predict_bboxes_single_image(me, props, T(cls_out[:n]), T(reg_out[:n]), (600, 1000),
                                                             ref_shim.AttrDict(min_score=0.05, nms_iou=0.3, max_per_img=100))
    print("c4: proposals", n, "pooled", tuple(pooled.shape), "detections", int(ds.numel()))
    save("c4", props=props, scores=scores, pooled_sub=pooled[::7, ::37].contiguous(), pooled_sum=pooled.double().sum(),
         det_bbox=db, det_score=ds, det_label=dl, cls_sha=sha(cls), feat_sha=sha(feat))


def g_heads2():
    """Round 2: AnchorHead.predict_single_image at BASELINE config-4 sizes (lib/heads/anchor_head.py:207-258), the
    softmax RPN (lib/heads/rpn_head.py:83-86), Scalable RoI layers (lib/region.py:212-239), IoUBalancedNegSampler
    (lib/region.py:128-172, numpy stream)."""
    import types
    import hashlib
    import inputs as gin
    import lib.heads.anchor_head as ah
    import lib.heads.rpn_head as rh
    out = {}
    # ---- RetinaNet test path: 151 200 / 37 800 / 9 450 / 2 457 / 693 anchors x 20 classes (inputs regenerated by the test)
    cls, reg = gin.retina_inputs(20)
    acs = [ranchor.AnchorCreator(base=s, scales=gin.RETINA_SCALES, aspect_ratios=[0.5, 1.0, 2.0]) for s in gin.RETINA_STRIDES]
    anchors = [ac(s, g) for ac, s, g in zip(acs, gin.RETINA_STRIDES, gin.RETINA_GRIDS)]
    meta = dict(img_shape=(800, 1333, 3), pad_shape=(800, 1344, 3), scale_factor=1.0)
    me = types.SimpleNamespace(cls_channels=20, use_sigmoid=True, num_classes=21, target_means=[0.0] * 4, target_stds=[1.0] * 4)
    cfgs = [dict(pre_nms=1000, min_bbox_size=0, min_score=0.05, nms_iou=0.5, nms_type="strict", max_per_img=100),
            dict(pre_nms=300, min_bbox_size=24, min_score=0.9, nms_iou=0.4, nms_type="official", max_per_img=250)]
    out["ret_cfgs"] = np.array([json.dumps(c) for c in cfgs])
    out["ret_cls_sha"], out["ret_reg_sha"] = sha(cls[0]), sha(reg[4])
    for i, c in enumerate(cfgs):
        with torch.no_grad():
            b, sc, lab = ah.AnchorHead.predict_single_image(me, [T(x) for x in cls], [T(x) for x in reg], anchors, meta,
                                                            ref_shim.AttrDict(c))
        out.update({"ret_bbox%d" % i: b, "ret_score%d" % i: sc, "ret_label%d" % i: lab})
        print("retina predict", i, "detections", int(sc.numel()))
    # softmax AnchorHead (use_sigmoid False, 21 channels) on a small pyramid
    rng = np.random.default_rng(SEED + 42)
    sgrids = [(20, 28), (10, 14), (5, 7)]
    scls = [rng.normal(0, 2, (3 * 5,) + g).astype(np.float32) for g in sgrids]
    sreg = [rng.normal(0, 0.3, (12,) + g).astype(np.float32) for g in sgrids]
    sacs = [ranchor.AnchorCreator(base=s, scales=[8], aspect_ratios=[0.5, 1.0, 2.0]) for s in (8, 16, 32)]
    sanchors = [ac(s, g) for ac, s, g in zip(sacs, (8, 16, 32), sgrids)]
    me = types.SimpleNamespace(cls_channels=5, use_sigmoid=False, num_classes=5, target_means=[0.0] * 4, target_stds=[1.0] * 4)
    scfg = dict(pre_nms=200, min_bbox_size=0, min_score=0.3, nms_iou=0.5, nms_type="official", max_per_img=80)
    with torch.no_grad():
        b, sc, lab = ah.AnchorHead.predict_single_image(me, [T(x) for x in scls], [T(x) for x in sreg], sanchors,
                                                        dict(img_shape=(160, 213, 3), scale_factor=1.0), ref_shim.AttrDict(scfg))
    print("softmax anchor head", "detections", int(sc.numel()))
    out.update(sm_bbox=b, sm_score=sc, sm_label=lab, sm_cfg=np.array(json.dumps(scfg)))
    for l in range(3):
        out["sm_cls%d" % l], out["sm_reg%d" % l] = scls[l], sreg[l]
    # ---- RPN with a two-channel softmax classifier (use_sigmoid False, lib/heads/rpn_head.py:83-86)
    grids = [(40, 56), (20, 28), (10, 14), (5, 7), (3, 4)]
    rcls = [rng.normal(0, 1, (6,) + g).astype(np.float32) for g in grids]
    rreg = [rng.normal(0, 0.5, (12,) + g).astype(np.float32) for g in grids]
    racs = [ranchor.AnchorCreator(base=s, scales=[8], aspect_ratios=[0.5, 1.0, 2.0]) for s in (4, 8, 16, 32, 64)]
    ranchors = [ac(s, g) for ac, s, g in zip(racs, (4, 8, 16, 32, 64), grids)]
    me = types.SimpleNamespace(cls_channels=2, use_sigmoid=False, target_means=[0.0] * 4, target_stds=[1.0] * 4)
    rcfg = dict(pre_nms=300, post_nms=200, max_num=400, nms_iou=0.7, min_bbox_size=0)
    with torch.no_grad():
        b, sc, _ = rh.RPNHead.predict_single_image(me, [T(x) for x in rcls], [T(x) for x in rreg], ranchors,
                                                   dict(img_shape=(160, 213, 3), pad_shape=(160, 224, 3), scale_factor=1.0),
                                                   ref_shim.AttrDict(rcfg))
    out.update(rs_props=b, rs_scores=sc, rs_cfg=np.array(json.dumps(rcfg)))
    for l in range(5):
        out["rs_cls%d" % l], out["rs_reg%d" % l] = rcls[l], rreg[l]
    # ---- ScalableRoIPool / ScalableRoIAlign (lib/region.py:212-239)
    feat = rng.normal(0, 1, (2, 8, 20, 28)).astype(np.float32)
    bx = rand_boxes(rng, 40, 160, 213, 8, 120)
    rois5 = np.concatenate([rng.integers(0, 2, (40, 1)).astype(np.float32), bx.T], 1)
    with torch.no_grad():
        sp = rregion.ScalableRoIPool(scale=1.3, output_size=(7, 7), spatial_scale=1 / 8)(T(feat), T(rois5))
        sa = rregion.ScalableRoIAlign(scale=0.8, output_size=(7, 7), spatial_scale=1 / 8, sampling_ratio=2)(T(feat), T(rois5))
    out.update(sc_feat=feat, sc_rois=rois5, sc_pool=sp, sc_align=sa)
    # ---- IoUBalancedNegSampler, numpy stream
    n = 3000
    lab = np.where(rng.uniform(size=n) < 0.1, rng.integers(1, 9, n), 0).astype(np.int64)
    lab[rng.uniform(size=n) < 0.05] = -1
    iou = np.where(lab > 0, rng.uniform(0.5, 1.0, n), rng.uniform(0, 0.5, n) ** 2).astype(np.float32)
    out.update(ib_labels=lab, ib_iou=iou)
    for i, (mx, ps, nb) in enumerate([(512, 128, 3), (256, 400, 5)]):
        if ps > mx:
            ps = mx
        np.random.seed(7 + i)
        res = rregion.IoUBalancedNegSampler(mx, ps, num_bins=nb, max_iou=0.5)(T(lab), T(iou), None, None)
        out["ib_out%d" % i] = res
        out["ib_cfg%d" % i] = np.array([mx, ps, nb])
        print("iou-balanced", i, "kept", int((res >= 0).sum()), "pos", int((res > 0).sum()))
    save("heads2", **out)


def g_garpn():
    """GA-RPN call sites (lib/heads/guided_head.py:557-669): GARPNHead.rpn_target_single_image and
    GARPNHead.predict_bboxes_single_image, called as the reference defines them on a stand-in `self` that carries the two
    attributes they read (target_means / target_stds).  The module imports mmdet's DeformConv at import time (stubbed by
    ref_shim); neither function touches it.  Guided anchors are explicit per-location boxes with a location mask."""
    import importlib
    import types
    gh = importlib.import_module("lib.heads.guided_head")
    rng = np.random.default_rng(SEED + 31)
    H, W = 160, 213
    grids, strides = [(20, 28), (10, 14), (5, 7), (3, 4)], [8, 16, 32, 64]
    gt, gl = synth_gt(rng, 4, H, W)
    cls = [rng.normal(0, 1.5, (1,) + g).astype(np.float32) for g in grids]
    reg = [rng.normal(0, 0.3, (4,) + g).astype(np.float32) for g in grids]
    anchors, masks = [], []
    for g, s in zip(grids, strides):
        ys, xs = np.meshgrid(np.arange(g[0]), np.arange(g[1]), indexing="ij")
        cx, cy = xs * s + s / 2.0, ys * s + s / 2.0
        w = np.exp(rng.uniform(np.log(1.0 * s), np.log(8.0 * s), g))
        h = np.exp(rng.uniform(np.log(1.0 * s), np.log(8.0 * s), g))
        anchors.append(np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2]).astype(np.float32))
        masks.append(rng.random((1,) + g) > 0.35)
    out = dict(gt=gt, gl=gl, n_levels=np.array(len(grids)))
    for l in range(len(grids)):
        out["cls%d" % l], out["reg%d" % l], out["anc%d" % l], out["mask%d" % l] = cls[l], reg[l], anchors[l], masks[l]
    meta = dict(img_shape=(H, W, 3), pad_shape=(160, 224, 3), scale_factor=1.0)
    for i, (means, stds, cfg) in enumerate([
            ((0., 0., 0., 0.), (1., 1., 1., 1.), dict(pre_nms=200, post_nms=0, max_num=100, nms_iou=0.7, min_bbox_size=0)),
            ((0., 0., 0., 0.), (0.07, 0.07, 0.14, 0.14), dict(pre_nms=60, post_nms=0, max_num=0, nms_iou=0.5, min_bbox_size=12)),
            ((0., 0., 0., 0.), (1., 1., 1., 1.), dict(pre_nms=0, post_nms=0, max_num=300, nms_iou=0.8, min_bbox_size=0))]):
        me = types.SimpleNamespace(target_means=list(means), target_stds=list(stds))
        b, sc, _ = gh.GARPNHead.predict_bboxes_single_image(
            me, [T(c).clone() for c in cls], [T(r).clone() for r in reg], [T(a).clone() for a in anchors],
            [T(m).clone() for m in masks], meta, ref_shim.AttrDict(cfg))
        out["pred_box%d" % i], out["pred_score%d" % i] = b, sc
        out["pred_cfg%d" % i] = np.array(json.dumps(dict(cfg, means=means, stds=stds)))
        print("ga-rpn predict", i, tuple(b.shape))
    tcfg = ref_shim.AttrDict(assigner=dict(type="MaxIoUAssigner", pos_iou=0.5, neg_iou=0.3, min_pos_iou=0.3),
                             sampler=dict(type="RandomSampler", max_num=64, pos_num=32))
    me = types.SimpleNamespace(target_means=[0., 0., 0., 0.], target_stds=[0.07, 0.07, 0.14, 0.14])
    np.random.seed(SEED)
    r = gh.GARPNHead.rpn_target_single_image(
        me, [T(c).clone() for c in cls], [T(x).clone() for x in reg], [T(a).clone() for a in anchors],
        [T(m).clone() for m in masks], T(gt), None, meta, tcfg)
    for k, v in zip(("tar_cls", "tar_reg", "tar_lab", "tar_anc", "tar_box", "tar_par"), r):
        out[k] = v
    print("ga-rpn target: %d samples, %d positive" % (int(r[2].numel()), int((r[2] > 0).sum())))
    save("garpn", **out)


if __name__ == "__main__":
    only = sys.argv[1:]
    gens = dict(anchors=g_anchors, iou_assign=g_iou_assign, deltas=g_deltas, nms=g_nms, rpn=g_rpn, roi=g_roi,
                targets=g_targets, atss=g_atss, heads=g_heads, loss=g_loss, c4=g_c4, heads2=g_heads2, garpn=g_garpn)
    for k, fn in gens.items():
        if not only or k in only:
            fn()
    with open(os.path.join(HERE, "VERSIONS.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, torchvision=torchvision.__version__, numpy=np.__version__,
                       seed=SEED, reference="/root/reference (pengfeidip/pytorch-faster-rcnn, unmodified)"), f, indent=1)
