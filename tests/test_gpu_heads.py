"""GPU parity of the head-side rows and of BASELINE configs 3-5:
cascade refine / 3-stage targets, RetinaNet dense assignment + per-level top-k NMS, ATSS.
Compared with golden vectors from the unmodified reference and with the CPU oracle."""
import json
import types

import numpy as np
import pytest
import torch

import oracle
from conftest import c4_inputs, load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import b200det
    from b200det import heads as bheads, region as bregion, utils as butils, fused, workload
    DEV = torch.device("cuda:0")


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------ a17
def test_refine_bboxes_vs_reference():
    g = load_golden("heads")
    me = types.SimpleNamespace(reg_class_agnostic=False, num_classes=21, target_means=[0.0] * 4,
                               target_stds=[0.05, 0.05, 0.1, 0.1])
    out = bheads.refine_bboxes_single_image(me, T(g["rf_props"]), T(g["rf_label"]), T(g["rf_reg"]), T(g["rf_is_gt"]),
                                            dict(img_shape=(400, 600, 3)))
    assert out.shape == g["rf_out"].shape
    np.testing.assert_allclose(N(out), g["rf_out"], rtol=1e-5, atol=1e-3)       # expf: 1e-5 relative (north_star)
    me2 = types.SimpleNamespace(reg_class_agnostic=True, num_classes=21, target_means=[0.0] * 4, target_stds=[0.1, 0.1, 0.2, 0.2])
    out2 = bheads.refine_bboxes_single_image(me2, T(g["rf_props"]), T(g["rf_label"]), T(g["rf_reg"][:, :4].copy()), None, None)
    np.testing.assert_allclose(N(out2), g["rf_out_agnostic"], rtol=1e-5, atol=1e-3)
    # ragged / degenerate: everything is GT -> empty result
    all_gt = bheads.refine_bboxes(T(g["rf_props"]), T(g["rf_label"]), T(g["rf_reg"]), torch.ones(200, dtype=torch.int64, device=DEV),
                                  (400, 600), None, None, False, 21)
    assert all_gt.shape == (4, 0)


# ------------------------------------------------------------------ a18 / config 5
@pytest.mark.parametrize("tag", ["s", "f"])
def test_atss_targets_vs_reference(tag):
    g = load_golden("atss")
    grids = [tuple(int(v) for v in x) for x in g["grids_" + tag]]
    me = types.SimpleNamespace(strides=[8, 16, 32, 64, 128], atss_cfg=types.SimpleNamespace(topk=9, scale=8))
    dummy = [torch.zeros((20,) + gr, device=DEV) for gr in grids]
    cls_t, reg_t, ctr_t = bheads.single_image_targets_atss(me, dummy, dummy, dummy, None, T(g["gt_" + tag]), T(g["gl_" + tag]),
                                                           dict(img_shape=tuple(int(v) for v in g["img_" + tag])), None)
    assert [tuple(c.shape) for c in cls_t] == [gr + (1,) for gr in grids]
    cls = np.concatenate([N(c).reshape(-1) for c in cls_t])
    reg = np.concatenate([N(r).reshape(-1, 4) for r in reg_t])
    ctr = np.concatenate([N(c).reshape(-1) for c in ctr_t])
    assert np.array_equal(cls, g["cls_" + tag])                                  # labels: bit-exact
    assert np.array_equal(reg, g["reg_" + tag])                                  # ltrb: exact fp32 differences
    np.testing.assert_allclose(ctr, g["ctr_" + tag], rtol=1e-5, atol=1e-6)
    assert (cls > 0).sum() > 0


def test_atss_batched_vs_oracle_ragged():
    """config 5 size (22 400 points), ragged GT counts incl. a single GT, K up to 64."""
    rng = np.random.default_rng(7)
    strides, pad, img = [8, 16, 32, 64, 128], (800, 1344), (800, 1333)
    grids = [(-(-pad[0] // s), -(-pad[1] // s)) for s in strides]
    counts = [1, 16, 64, 5, 8]
    B, ld = len(counts), 64
    gt = np.zeros((B, 4, ld), np.float32)
    gl = np.zeros((B, ld), np.int64)
    for b, k in enumerate(counts):
        bb, ll = workload.synth_gt(rng, k, *img)
        gt[b, :, :k], gl[b, :k] = bb, ll
    # image 4: GTs whose centres sit in the corners / on the borders of the image (the candidate search scans a window
    # around the centre's cell: clipped windows must still hold the 9 nearest cells of every level)
    gt[4, :, :8] = np.array([[0, 0, 30, 30], [1300, 770, 1332, 799], [0, 760, 12, 799], [1320, 0, 1332, 9], [600, 0, 700, 6],
                             [0, 300, 5, 500], [1326, 100, 1332, 700], [2, 2, 1330, 797]], np.float32).T
    cls, reg, ctr = bheads.atss_assign(grids, strides, T(gt), torch.tensor(counts, dtype=torch.int32, device=DEV), T(gl),
                                       torch.tensor([[800.0, 1333.0]] * B, device=DEV))
    for b, k in enumerate(counts):
        oc, orr, octr = oracle.atss_assign(grids, strides, gt[b, :, :k], gl[b, :k], img)
        assert np.array_equal(N(cls[b]), oc), b
        assert np.array_equal(N(reg[b]), orr), b
        np.testing.assert_allclose(N(ctr[b]), octr, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ a19
@pytest.mark.parametrize("i,cfg", [(0, dict(pre_nms=1000, min_bbox_size=0, min_score=0.05, nms_iou=0.6, nms_type="strict", max_per_img=100)),
                                   (1, dict(pre_nms=50, min_bbox_size=40, min_score=0.3, nms_iou=0.5, nms_type="official", max_per_img=60))])
def test_fcos_predict_vs_reference(i, cfg):
    g = load_golden("heads")
    cls = [T(g["fc_cls%d" % l]) for l in range(5)]
    reg = [T(g["fc_reg%d" % l]) for l in range(5)]
    ctr = [T(g["fc_ctr%d" % l]) for l in range(5)]
    b, s, lab = bheads.fcos_predict_single_image(cls, reg, ctr, [8, 16, 32, 64, 128],
                                                 dict(img_shape=tuple(int(v) for v in g["fc_img"]), scale_factor=1.0), cfg,
                                                 reg_mean=0, reg_std=300, use_centerness=True)
    assert np.array_equal(N(lab), g["fc_label%d" % i])
    np.testing.assert_allclose(N(s), g["fc_score%d" % i], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(N(b), g["fc_bbox%d" % i], rtol=1e-5, atol=1e-3)


# ------------------------------------------------------------------ config 4: RetinaNet
def test_retinanet_dense_assign_full_size_vs_oracle():
    """201 600 anchors (A=9, strides 8-128, scales 4*2^{0,1/3,2/3}) x K GT, pos .5 / neg .4 / min_pos 0,
    allowed_border -1, no sampler (configs/retinanet_r50_fpn.py:34-43)."""
    rng = np.random.default_rng(11)
    strides, pad, img = [8, 16, 32, 64, 128], (800, 1344), (800, 1333)
    grids = [(-(-pad[0] // s), -(-pad[1] // s)) for s in strides]
    scales = [4 * 2 ** (k / 3) for k in range(3)]
    pyr = fused.AnchorPyramid(strides, grids, scales=scales, ratios=(0.5, 1.0, 2.0))
    assert pyr.total == 201600
    B, K = 2, 16
    gts = [workload.synth_gt(rng, K, *img)[0] for _ in range(B)]
    gt = np.stack(gts)
    bt = fused.BatchedTargets(B, pyr.total, K, dict(pos_iou=0.5, neg_iou=0.4, min_pos_iou=0.0), dict(max_num=256, pos_num=128),
                              None, None, DEV, pyramid=pyr, border=-1.0)
    import ctypes
    from b200det import _C
    img_hw = torch.tensor([[800.0, 1333.0]] * B, device=DEV)
    gcount = torch.full((B,), K, dtype=torch.int32, device=DEV)
    _C.call("b2d_assign_max_iou", _C.ptr(bt.labels), _C.ptr(bt.iou), bt.out_ld, None, 0, None, bt.N, ctypes.byref(pyr.c),
            _C.ptr(img_hw), -1.0, _C.ptr(T(gt)), K, _C.ptr(gcount), B, 0.5, 0.4, 0.0, 0, _C.ptr(bt.census), _C.ptr(bt.pos_list),
            bt.out_ld, _C.ptr(bt.colmax), bt.colmax.numel() * 4, _C.stream())
    anc = np.concatenate([oracle.anchor_grid(s, g, scales=scales).reshape(4, -1) for s, g in zip(strides, grids)], 1)
    mask = np.concatenate([oracle.valid_mask(oracle.anchor_grid(s, g, scales=scales), img, g, s, -1) for s, g in zip(strides, grids)])
    for b in range(B):
        lab, iou = oracle.assign_max_iou(np.ascontiguousarray(anc[:, mask]), gt[b], 0.5, 0.4, 0.0)
        full = np.full(anc.shape[1], -1, np.int64)
        full[mask] = lab
        assert np.array_equal(N(bt.labels[b]), full), b
        cen = N(bt.census[b])
        assert cen[0] == (full > 0).sum() and cen[1] == (full == 0).sum()


def test_retinanet_per_level_topk_batched_nms_vs_oracle():
    """test path of config 4: per-level top-1000 on max-over-classes score, decode, then class-aware NMS."""
    rng = np.random.default_rng(12)
    n, C = 3000, 20
    boxes = np.sort(rng.uniform(0, 800, (n, 2, 2)), axis=1).reshape(n, 4)[:, [0, 2, 1, 3]].astype(np.float32)
    score = rng.uniform(0, 1, n).astype(np.float32)
    label = rng.integers(0, C, n).astype(np.int64)
    kb, ks, kl = butils.batched_nms(T(boxes), T(score), T(label), 0.5)
    keep = oracle.batched_nms(boxes, score, label, 0.5)
    assert np.array_equal(N(kl), label[keep]) and np.array_equal(N(ks), score[keep]) and np.array_equal(N(kb), boxes[keep])
    idx = bregion.topk_desc(T(score), 1000)
    assert np.array_equal(N(idx), np.argsort(-score, kind="stable")[:1000])


# ------------------------------------------------------------------ config 3: cascade
def test_cascade_three_stage_targets_refine_roialign_fwd_bwd():
    """3 x (bbox_target -> RoIAlign fwd -> synthetic reg_out -> refine) + RoIAlign bwd, thresholds .5/.6/.7 and
    stage stds of configs/cascade_rcnn_r50_fpn.py; every stage is checked against the oracle on the GPU's own inputs."""
    from b200det import bbox as bbbox
    rng = np.random.default_rng(13)
    img, C = (320, 416), 16
    strides = (4, 8, 16, 32)
    grids = [(-(-img[0] // s), -(-img[1] // s)) for s in strides]
    feats_np = [rng.standard_normal((1, C) + g).astype(np.float32) for g in grids]
    feats = [T(f).requires_grad_(True) for f in feats_np]
    gt, gl = workload.synth_gt(rng, 5, *img)
    cx, cy = rng.uniform(0, img[1], 300), rng.uniform(0, img[0], 300)
    w, h = rng.uniform(16, 200, 300), rng.uniform(16, 200, 300)
    props = np.stack([np.clip(cx - w / 2, 0, img[1] - 1), np.clip(cy - h / 2, 0, img[0] - 1),
                      np.clip(cx + w / 2, 0, img[1] - 1), np.clip(cy + h / 2, 0, img[0] - 1)]).astype(np.float32)
    props = T(props)
    ext = bregion.BasicRoIExtractor([dict(type="RoIAlign", spatial_scale=1 / s, sampling_ratio=2) for s in strides], output_size=(7, 7))
    stage_stds = [(0.1, 0.1, 0.2, 0.2), (0.05, 0.05, 0.1, 0.1), (0.033, 0.033, 0.067, 0.067)]
    total = 0
    for st, (thr, stds) in enumerate(zip((0.5, 0.6, 0.7), stage_stds)):
        assigner = bregion.MaxIoUAssigner(thr, thr, thr)
        sampler = bregion.RandomSampler(128, 32, rng="numpy")
        np.random.seed(100 + st)
        tar_props, tar_bbox, tar_label, tar_param, tar_is_gt = bbbox.bbox_target(props, T(gt), T(gl), assigner, sampler, [0.0] * 4, list(stds))
        # stage-wise oracle: assignment of these proposals
        olab, _ = oracle.assign_max_iou(N(props), gt, thr, thr, thr)
        lab_gpu, _ = assigner(props, T(gt))
        assert np.array_equal(N(lab_gpu), olab)
        enc = oracle.bbox2param(N(tar_props), N(tar_bbox), [0.0] * 4, list(stds))
        np.testing.assert_allclose(N(tar_param), enc, rtol=1e-5, atol=1e-5)
        out = ext(feats, [tar_props])[0]
        ref = oracle.roi_extract([f[0] for f in feats_np], N(tar_props))
        np.testing.assert_allclose(N(out), ref, rtol=1e-5, atol=1e-6)
        total = total + (out * out).sum()
        reg_out = T(rng.standard_normal((tar_props.shape[1], 4)).astype(np.float32))
        refined = bheads.refine_bboxes(tar_props, tar_label, reg_out, tar_is_gt, img + (3,), [0.0] * 4, list(stds), True, 21)
        keep = N(tar_is_gt) == 0
        dec = oracle.param2bbox(N(tar_props)[:, keep], N(reg_out).T[:, keep], [0.0] * 4, list(stds), img)
        np.testing.assert_allclose(N(refined), dec, rtol=1e-5, atol=1e-3)
        props = refined
    total.backward()
    assert all(f.grad is not None and torch.isfinite(f.grad).all() for f in feats)
    assert sum(float(f.grad.abs().sum()) for f in feats) > 0


# ------------------------------------------------------------------ SURVEY 8(f-3)
_DET_CFGS = [dict(min_score=0.05, nms_iou=0.5, max_per_img=100, nms_type="official"),
             dict(min_score=0.2, nms_iou=0.3, max_per_img=40, nms_type="strict")]


@pytest.mark.parametrize("i", [0, 1])
def test_rcnn_predict_bboxes_single_image_vs_reference(i):
    """Method form with the reference's signature (lib/heads/bbox_head.py:122) against its golden output."""
    g = load_golden("heads")
    me = types.SimpleNamespace(use_sigmoid=False, reg_class_agnostic=False, num_classes=21,
                               target_means=[0.0, 0.0, 0.0, 0.0], target_stds=[0.1, 0.1, 0.2, 0.2])
    b, s, l = bheads.predict_bboxes_single_image(me, T(g["det_props"]), T(g["det_cls"]), T(g["det_reg"]), (400, 600),
                                                 dict(_DET_CFGS[i]))
    assert np.array_equal(N(l), g["det_label%d" % i])
    np.testing.assert_allclose(N(s), g["det_score%d" % i], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(N(b), g["det_bbox%d" % i], rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("mode,agnostic", [("official", False), ("strict", False), ("official", True)])
def test_rcnn_detect_batched_vs_oracle(mode, agnostic):
    """ragged batch (1000 / 37 / 0 proposals), 21 and 81 classes, class-agnostic regression, no clamp."""
    rng = np.random.default_rng(5 + len(mode) + agnostic)
    B, n = 3, 1000
    for C in (21, 81):
        props = np.zeros((B, 4, n), np.float32)
        for b in range(B):
            cx, cy = rng.uniform(0, 1333, n), rng.uniform(0, 800, n)
            w, h = rng.uniform(16, 400, n), rng.uniform(16, 400, n)
            props[b] = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2])
        cls = rng.normal(0, 2.5, (B, n, C)).astype(np.float32)
        cls[..., 0] += 2.0
        reg = rng.normal(0, 0.5, (B, n, 4 if agnostic else 4 * C)).astype(np.float32)
        counts = np.array([n, 37, 0], np.int32)
        img = (800, 1333) if C == 21 else None
        ob, os_, ol, oc, ovf = bheads.rcnn_detect(T(props), T(cls), T(reg), img, [0, 0, 0, 0], [0.1, 0.1, 0.2, 0.2], 0.05, 0.5,
                                                  100, mode, counts=T(counts))
        assert int(ovf[0]) == 0
        for b in range(B):
            m = int(counts[b])
            kb, ks, kl = oracle.rcnn_detect(np.ascontiguousarray(props[b][:, :m]), cls[b, :m], reg[b, :m], img, [0, 0, 0, 0],
                                            [0.1, 0.1, 0.2, 0.2], 0.05, 0.5, 100, mode)
            k = int(oc[b])
            assert k == ks.shape[0], (C, b, k, ks.shape)
            assert np.array_equal(N(ol[b, :k]), kl)
            np.testing.assert_allclose(N(os_[b, :k]), ks, rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(N(ob[b, :, :k]), kb, rtol=1e-5, atol=1e-3)


# ------------------------------------------------------------------ GA-RPN call sites
@pytest.mark.parametrize("i", [0, 1, 2])
def test_ga_rpn_predict_vs_reference(i):
    """GARPNHead.predict_bboxes_single_image (lib/heads/guided_head.py:621-669) on explicit guided anchors + location masks:
    golden from the unmodified reference function (tests/golden/make_golden.py g_garpn); cfg 0: top-k + NMS + max_num,
    cfg 1: target stds, min-size filter, no max_num, cfg 2: no per-level top-k."""
    g = load_golden("garpn")
    L = int(g["n_levels"])
    cfg = json.loads(str(g["pred_cfg%d" % i]))
    head = types.SimpleNamespace(target_means=cfg.pop("means"), target_stds=cfg.pop("stds"))
    b, s, extra = bheads.ga_rpn_predict_single_image(
        head, [T(g["cls%d" % l]) for l in range(L)], [T(g["reg%d" % l]) for l in range(L)], [T(g["anc%d" % l]) for l in range(L)],
        [T(g["mask%d" % l]) for l in range(L)], dict(img_shape=(160, 213, 3), pad_shape=(160, 224, 3), scale_factor=1.0), cfg)
    assert extra is None and tuple(b.shape) == g["pred_box%d" % i].shape
    np.testing.assert_allclose(N(s), g["pred_score%d" % i], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(N(b), g["pred_box%d" % i], rtol=1e-5, atol=1e-3)


def test_ga_rpn_target_vs_reference():
    """GARPNHead.rpn_target_single_image (lib/heads/guided_head.py:557-572) with the reference's host RNG stream."""
    g = load_golden("garpn")
    L = int(g["n_levels"])
    head = types.SimpleNamespace(target_means=[0., 0., 0., 0.], target_stds=[0.07, 0.07, 0.14, 0.14])
    cfg = dict(assigner=bregion.MaxIoUAssigner(0.5, 0.3, 0.3), sampler=bregion.RandomSampler(64, 32, rng="numpy"))
    np.random.seed(2019)
    r = bheads.ga_rpn_target_single_image(
        head, [T(g["cls%d" % l]) for l in range(L)], [T(g["reg%d" % l]) for l in range(L)], [T(g["anc%d" % l]) for l in range(L)],
        [T(g["mask%d" % l]) for l in range(L)], T(g["gt"]), None, dict(img_shape=(160, 213, 3)), cfg)
    for got, name in zip(r, ("tar_cls", "tar_reg", "tar_lab", "tar_anc", "tar_box")):
        assert np.array_equal(N(got), g[name]), name
    np.testing.assert_allclose(N(r[5]), g["tar_par"], rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------ SURVEY 8(f-4)
@pytest.mark.parametrize("tag", ["s", "f"])
def test_fcos_plain_targets_vs_reference(tag):
    g = load_golden("atss")
    grids = [tuple(int(v) for v in x) for x in g["grids_" + tag]]
    me = types.SimpleNamespace(strides=[8, 16, 32, 64, 128], level_scale_thr=[0, 64, 128, 256, 512, 1e6])
    dummy = [torch.zeros((20,) + gr, device=DEV) for gr in grids]
    cls_t, reg_t, ctr_t = bheads.single_image_targets(me, dummy, dummy, dummy, T(g["gt_" + tag]), T(g["gl_" + tag]),
                                                      dict(img_shape=tuple(int(v) for v in g["img_" + tag])), None)
    assert [tuple(c.shape) for c in cls_t] == [gr + (1,) for gr in grids]
    cls = np.concatenate([N(c).reshape(-1) for c in cls_t])
    reg = np.concatenate([N(r).reshape(-1, 4) for r in reg_t])
    ctr = np.concatenate([N(c).reshape(-1) for c in ctr_t])
    assert np.array_equal(cls, g["pcls_" + tag]) and np.array_equal(reg, g["preg_" + tag])
    np.testing.assert_allclose(ctr, g["pctr_" + tag], rtol=1e-5, atol=1e-6)


def test_fcos_plain_targets_batched_vs_oracle_ragged():
    rng = np.random.default_rng(21)
    strides = [8, 16, 32, 64, 128]
    grids = [(-(-800 // s), -(-1344 // s)) for s in strides]
    B, K = 3, 40
    gt = np.zeros((B, 4, K), np.float32)
    for b in range(B):
        x1 = rng.uniform(0, 1000, K); y1 = rng.uniform(0, 600, K)
        gt[b] = np.stack([x1, y1, np.minimum(x1 + rng.uniform(10, 700, K), 1332), np.minimum(y1 + rng.uniform(10, 500, K), 799)])
    gl = rng.integers(1, 21, (B, K)).astype(np.int64)
    cnt = np.array([K, 1, 7], np.int32)
    img_hw = torch.tensor([[800.0, 1333.0]] * B, device=DEV)
    cls, reg, ctr = bheads.fcos_targets(grids, strides, T(gt), T(cnt), T(gl), img_hw)
    for b in range(B):
        k = int(cnt[b])
        c, r, t = oracle.fcos_targets(grids, strides, gt[b][:, :k], gl[b, :k], (800, 1333))
        assert np.array_equal(N(cls[b]), c) and np.array_equal(N(reg[b]), r)
        np.testing.assert_allclose(N(ctr[b]), t, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ SURVEY 8(f-2)
def test_anchor_head_loss_fused_vs_reference():
    """b2d_anchor_loss_fwd / _bwd against the reference's losses + torch autograd (golden `loss`)."""
    g = load_golden("loss")
    strides, grids = [8, 16, 32], [(20, 28), (10, 14), (5, 7)]
    scales = [float(v) for v in g["scales"]]
    pyr = fused.AnchorPyramid(strides, grids, scales=scales)
    anc = np.concatenate([oracle.anchor_grid(s, gr, scales=scales).reshape(4, -1) for s, gr in zip(strides, grids)], 1)
    lab, _ = oracle.assign_max_iou(anc, g["gt"], 0.5, 0.4, 0.0)
    cls = [T(g["cls%d" % l][None]).requires_grad_(True) for l in range(3)]
    reg = [T(g["reg%d" % l][None]).requires_grad_(True) for l in range(3)]
    s = bheads.anchor_head_loss_sums(cls, reg, T(lab[None]), pyr, T(g["gt"][None]), T(g["gl"][None]))
    (2.0 * s[0] + 3.0 * s[1]).backward()
    assert int(s[2]) == int(g["npos"])
    np.testing.assert_allclose(float(s[0].detach()), float(g["focal"]), rtol=1e-5)
    np.testing.assert_allclose(float(s[1].detach()), float(g["sl1"]), rtol=1e-5)
    for l in range(3):
        np.testing.assert_allclose(N(cls[l].grad[0]), g["dcls%d" % l], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(N(reg[l].grad[0]), g["dreg%d" % l], rtol=1e-4, atol=1e-6)


def test_anchor_head_loss_fused_config4_size_vs_oracle():
    """RetinaNet sizes: 201 600 anchors x 20 classes per image, B = 2, labels from the GPU's own dense assignment."""
    rng = np.random.default_rng(41)
    strides = [8, 16, 32, 64, 128]
    grids = [(-(-800 // s), -(-1344 // s)) for s in strides]
    scales = [4.0, 4.0 * 2 ** (1 / 3), 4.0 * 2 ** (2 / 3)]
    pyr = fused.AnchorPyramid(strides, grids, scales=scales)
    B, K, C, A = 2, 12, 20, 9
    gt = np.zeros((B, 4, K), np.float32)
    for b in range(B):
        x1 = rng.uniform(0, 1000, K); y1 = rng.uniform(0, 600, K)
        gt[b] = np.stack([x1, y1, np.minimum(x1 + rng.uniform(20, 500, K), 1332), np.minimum(y1 + rng.uniform(20, 320, K), 799)])
    gl = rng.integers(1, C + 1, (B, K)).astype(np.int64)
    cls_np = [rng.normal(-2, 1.5, (B, A * C) + g).astype(np.float32) for g in grids]
    reg_np = [rng.normal(0, 0.4, (B, A * 4) + g).astype(np.float32) for g in grids]
    anc = np.concatenate([oracle.anchor_grid(s, gr, scales=scales).reshape(4, -1) for s, gr in zip(strides, grids)], 1)
    assert anc.shape[1] == 201600
    labs = np.stack([oracle.assign_max_iou(anc, gt[b], 0.5, 0.4, 0.0)[0] for b in range(B)])
    cls = [T(c).requires_grad_(True) for c in cls_np]
    reg = [T(r).requires_grad_(True) for r in reg_np]
    closs, rloss = bheads.anchor_head_calc_loss(cls, reg, T(labs), pyr, T(gt), T(gl), beta=1.0 / 9.0)
    (closs + rloss).backward()
    f = s1 = 0.0
    npos, dcs, drs = 0, [], []
    for b in range(B):
        fb, sb, nb, dc, dr = oracle.anchor_head_loss([c[b] for c in cls_np], [r[b] for r in reg_np], labs[b], anc, gt[b], gl[b])
        f += fb; s1 += sb; npos += nb; dcs.append(dc); drs.append(dr)
    assert npos > 50
    np.testing.assert_allclose(float(closs.detach()), f / npos, rtol=1e-5)
    np.testing.assert_allclose(float(rloss.detach()), s1 / npos, rtol=1e-5)
    for l in range(5):
        want_c = np.stack([dcs[b][l] for b in range(B)]) / npos
        want_r = np.stack([drs[b][l] for b in range(B)]) / npos
        np.testing.assert_allclose(N(cls[l].grad), want_c, rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(N(reg[l].grad), want_r, rtol=1e-4, atol=1e-7)


# ------------------------------------------------------------------ BASELINE config 1 (C4), end to end
def test_config1_c4_inference_path_vs_reference():
    """faster_rcnn_r50 (C4) inference behind the convolutions, at the config's own sizes (600x1000 -> 608x1024,
    29 184 anchors, test_cfg.rpn 6000/300/300/0.7 -> dense NMS path, RoIPool 7x7 @1/16 on 1024 channels, 21-class
    detections with min_score .05 / nms .3 / 100): every stage against the unmodified reference's output."""
    import hashlib
    g = load_golden("c4")
    cls, reg, feat, cls_out, reg_out = c4_inputs()
    sha = lambda a: np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(sha(cls), g["cls_sha"]) and np.array_equal(sha(feat), g["feat_sha"]), "seeded inputs differ"
    from b200det import anchor as banchor
    head = types.SimpleNamespace(anchor_strides=[16], anchor_scales=[4, 8, 16, 32], anchor_ratios=[0.5, 1.0, 2.0],
                                 target_means=[0.0] * 4, target_stds=[1.0] * 4, use_sigmoid=True, cls_channels=1,
                                 anchor_creators=[banchor.AnchorCreator(base=16, scales=[4, 8, 16, 32])])
    head.anchor_creators[0].to(DEV)
    anchors = [head.anchor_creators[0](16, (38, 64))]
    assert anchors[0].numel() == 4 * 29184
    meta = dict(img_shape=(600, 1000, 3), pad_shape=(608, 1024, 3), scale_factor=1.0)
    props, scores, _ = bheads.rpn_predict_single_image(head, [T(cls)], [T(reg)], anchors, meta,
                                                       dict(pre_nms=6000, post_nms=300, max_num=300, nms_iou=0.7, min_bbox_size=0.0))
    assert tuple(props.shape) == (4, 300)
    np.testing.assert_allclose(N(scores), g["scores"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(N(props), g["props"], rtol=1e-5, atol=1e-3)
    # RoIPool on the REFERENCE's proposals (so that a last-bit difference in a decoded box cannot move a bin edge)
    ext = bregion.BasicRoIExtractor([dict(type="RoIPool", spatial_scale=1 / 16, sampling_ratio=2)], output_size=(7, 7))
    pooled = ext([T(feat)], [T(g["props"])])[0]
    assert tuple(pooled.shape) == (300, 1024, 7, 7)
    assert np.array_equal(N(pooled[::7, ::37]), g["pooled_sub"])                     # max pooling: exact
    np.testing.assert_allclose(float(pooled.double().sum()), float(g["pooled_sum"]), rtol=1e-9)
    me = types.SimpleNamespace(use_sigmoid=False, reg_class_agnostic=False, num_classes=21,
                               target_means=[0.0] * 4, target_stds=[0.1, 0.1, 0.2, 0.2])
    db, ds, dl = bheads.predict_bboxes_single_image(me, T(g["props"]), T(cls_out), T(reg_out), (600, 1000),
                                                    dict(min_score=0.05, nms_iou=0.3, max_per_img=100))
    assert np.array_equal(N(dl), g["det_label"])
    np.testing.assert_allclose(N(ds), g["det_score"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(N(db), g["det_bbox"], rtol=1e-5, atol=1e-3)


# ------------------------------------------------------------------ BASELINE config 3, batched
def test_cascade_hot_path_batched_vs_oracle():
    """fused.CascadeHotPath at config-3 geometry (800x1344 pyramid, 2000 proposals, 512 samples, 3 stages; B = 2,
    64 channels to keep the oracle fast): every stage's assignment, sampling (device-sampler spec), encoded deltas, RoI
    features, refined boxes and the RoIAlign backward against the oracle."""
    from oracle import sampler_spec
    rng = np.random.default_rng(77)
    B, K, C, NC = 2, 6, 64, 21
    strides = (4, 8, 16, 32)
    grids = [(-(-800 // s), -(-1344 // s)) for s in strides]
    feats_np = [rng.standard_normal((B, C) + g).astype(np.float32) for g in grids]
    feats = [T(f).contiguous(memory_format=torch.channels_last) for f in feats_np]
    gt = np.zeros((B, 4, K), np.float32); gl = np.zeros((B, K), np.int64)
    for b in range(B):
        gt[b], gl[b] = workload.synth_gt(rng, K, 800, 1333)
    n = 2000
    props = np.zeros((B, 4, n), np.float32)
    for b in range(B):
        j = rng.integers(0, K, n)
        jit = rng.normal(0, 25, (4, n))
        far = rng.random(n) < 0.5
        cx, cy = rng.uniform(0, 1333, n), rng.uniform(0, 800, n)
        w, h = rng.uniform(16, 300, n), rng.uniform(16, 300, n)
        rnd = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2])
        bb = np.where(far[None], rnd, gt[b][:, j] + jit)
        props[b] = np.stack([np.clip(np.minimum(bb[0], bb[2]), 0, 1332), np.clip(np.minimum(bb[1], bb[3]), 0, 799),
                             np.clip(np.maximum(bb[0], bb[2]) + 1, 0, 1332), np.clip(np.maximum(bb[1], bb[3]) + 1, 0, 799)])
    cp = fused.CascadeHotPath(B, n, grids, DEV, strides=strides, gt_ld=K, feat_channels=C, num_classes=NC, seed=3)
    m = cp.m
    reg_np = [rng.normal(0, 1, (B, m, 4 * NC)).astype(np.float32) for _ in range(3)]
    gcount = torch.full((B,), K, dtype=torch.int32, device=DEV)
    pcount = torch.full((B,), n, dtype=torch.int32, device=DEV)
    img_hw = torch.tensor([[800.0, 1333.0]] * B, device=DEV)
    outs = cp.step(T(props), pcount, feats, T(gt), gcount, T(gl), img_hw, [T(r) for r in reg_np])
    torch.cuda.synchronize()
    thresholds = (0.5, 0.6, 0.7)
    stds = ((0.1, 0.1, 0.2, 0.2), (0.05, 0.05, 0.1, 0.1), (0.033, 0.033, 0.067, 0.067))
    cur = [np.ascontiguousarray(props[b]) for b in range(B)]
    for s, (bt, roi_feats, refined, rcount) in enumerate(outs):
        for b in range(B):
            olab, _ = oracle.assign_max_iou(cur[b], gt[b], thresholds[s], thresholds[s], thresholds[s])
            full = np.concatenate([np.arange(1, K + 1), olab]).astype(np.int64)
            nb = full.shape[0]
            assert np.array_equal(N(bt.labels[b, :nb]), full), (s, b)
            seed = ((3 + s) * 1000003 + 1) & 0xFFFFFFFFFFFFFFFF
            want = sampler_spec.sample(full, m, 128, seed, image_index=b)
            k = int(bt.n_chosen[b])
            assert np.array_equal(N(bt.chosen[b, :k]), want), (s, b)
            allb = np.concatenate([gt[b], cur[b]], 1)
            tb = allb[:, want]
            np.testing.assert_array_equal(N(bt.tar_box[b, :, :k]), tb)
            tg_ = gt[b][:, np.maximum(full[want] - 1, 0)]
            np.testing.assert_allclose(N(bt.tar_param[b, :, :k]), oracle.bbox2param(tb, tg_, [0.0] * 4, list(stds[s])), rtol=1e-5, atol=1e-5)
            ref = oracle.roi_extract([f[b] for f in feats_np], tb)
            assert np.array_equal(N(roi_feats[b * m:b * m + k]).view(np.uint32), ref.view(np.uint32)), (s, b)
            lab_cls = np.where(full[want] > 0, gl[b][np.maximum(full[want] - 1, 0)], 0)
            keep = want >= K                                       # GT columns (the first K candidates) are dropped
            reg = reg_np[s][b, :k].reshape(k, 4, NC)[np.arange(k), :, lab_cls].T       # [4, k]
            dec = oracle.param2bbox(np.ascontiguousarray(tb[:, keep]), np.ascontiguousarray(reg[:, keep]), [0.0] * 4, list(stds[s]), (800, 1333))
            kr = int(rcount[b])
            assert kr == int(keep.sum())
            np.testing.assert_allclose(N(refined[b, :, :kr]), dec, rtol=1e-5, atol=1e-3)
            cur[b] = np.ascontiguousarray(N(refined[b, :, :kr]))
    # backward of the three RoIAlign stages
    go = [rng.standard_normal((B * m, C, 7, 7)).astype(np.float32) for _ in range(3)]
    grads = cp.backward([T(g) for g in go])
    torch.cuda.synchronize()
    s = 2
    bt = outs[s][0]
    for b in range(B):
        k = int(bt.n_chosen[b])
        assert k == m
        tb = N(bt.tar_box[b, :, :k])
        lv = oracle.level_map(tb)
        for l, st in enumerate(strides):
            ref = oracle.roi_align_bwd(go[s][b * m:(b + 1) * m][lv == l], (C,) + grids[l], np.ascontiguousarray(tb[:, lv == l]), 1.0 / st)
            np.testing.assert_allclose(N(grads[s][l][b]), ref, rtol=1e-5, atol=2e-6)


# ------------------------------------------------------------------ round 2: RetinaNet test path, softmax heads, glue kernels
def _gin():
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import inputs as gin
    return gin


def _sha(a):
    import hashlib
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


@pytest.mark.parametrize("i", [0, 1])
def test_anchor_head_predict_retinanet_config4_vs_reference(i):
    """AnchorHead.predict_single_image (lib/heads/anchor_head.py:207-258) at BASELINE config-4 sizes -- 201 600 anchors x
    20 classes, per-level top-1000 on the best class score (score_mode 2, do_nms 0), decode, strict / official
    multiclass NMS, max_per_img -- through the reference-signature method against the reference's own output."""
    import json
    from b200det import anchor as banchor
    gin = _gin()
    g = load_golden("heads2")
    cls, reg = gin.retina_inputs(20)
    assert np.array_equal(_sha(cls[0]), g["ret_cls_sha"]) and np.array_equal(_sha(reg[4]), g["ret_reg_sha"]), "seeded inputs differ"
    head = types.SimpleNamespace(anchor_strides=list(gin.RETINA_STRIDES), anchor_scales=gin.RETINA_SCALES,
                                 anchor_ratios=[0.5, 1.0, 2.0], target_means=[0.0] * 4, target_stds=[1.0] * 4,
                                 use_sigmoid=True, cls_channels=20, num_classes=21, anchor_creators=None)
    anchors = [torch.empty((4, 9) + gr, device=DEV) for gr in gin.RETINA_GRIDS]       # only the grid sizes are read
    meta = dict(img_shape=(800, 1333, 3), pad_shape=(800, 1344, 3), scale_factor=1.0)
    cfg = json.loads(str(g["ret_cfgs"][i]))
    b, s, l = bheads.anchor_head_predict_single_image(head, [T(x) for x in cls], [T(x) for x in reg], anchors, meta, cfg)
    assert tuple(b.shape) == g["ret_bbox%d" % i].shape
    assert np.array_equal(N(l), g["ret_label%d" % i])
    np.testing.assert_allclose(N(s), g["ret_score%d" % i], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(N(b), g["ret_bbox%d" % i], rtol=1e-5, atol=1e-3)


def test_anchor_head_predict_softmax_head_vs_reference():
    """use_sigmoid False: per-level top-k on max_{c>=1} softmax_c (score_mode 3), softmax scores, classes 1..C-1."""
    import json
    g = load_golden("heads2")
    head = types.SimpleNamespace(anchor_strides=[8, 16, 32], anchor_scales=[8], anchor_ratios=[0.5, 1.0, 2.0],
                                 target_means=[0.0] * 4, target_stds=[1.0] * 4, use_sigmoid=False, cls_channels=5,
                                 num_classes=5, anchor_creators=None)
    grids = [(20, 28), (10, 14), (5, 7)]
    anchors = [torch.empty((4, 3) + gr, device=DEV) for gr in grids]
    cfg = json.loads(str(g["sm_cfg"]))
    b, s, l = bheads.anchor_head_predict_single_image(head, [T(g["sm_cls%d" % k]) for k in range(3)],
                                                      [T(g["sm_reg%d" % k]) for k in range(3)], anchors,
                                                      dict(img_shape=(160, 213, 3), scale_factor=1.0), cfg)
    assert tuple(b.shape) == g["sm_bbox"].shape
    assert np.array_equal(N(l), g["sm_label"])
    np.testing.assert_allclose(N(s), g["sm_score"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(N(b), g["sm_bbox"], rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("front", ["1", "0"])
def test_rpn_two_channel_softmax_vs_reference(front, setknob):
    """RPNHead with use_sigmoid False (lib/heads/rpn_head.py:83-86): score = softmax[1] = sigmoid(l1 - l0), score_mode 1,
    through the cluster kernel and the multi-kernel chain."""
    import json
    from b200det import anchor as banchor
    setknob(B2D_RPN_FRONT=front)
    g = load_golden("heads2")
    grids = [(40, 56), (20, 28), (10, 14), (5, 7), (3, 4)]
    head = types.SimpleNamespace(anchor_strides=[4, 8, 16, 32, 64], anchor_scales=[8], anchor_ratios=[0.5, 1.0, 2.0],
                                 target_means=[0.0] * 4, target_stds=[1.0] * 4, use_sigmoid=False, cls_channels=2,
                                 anchor_creators=None)
    anchors = [torch.empty((4, 3) + gr, device=DEV) for gr in grids]
    b, s, extra = bheads.rpn_predict_single_image(head, [T(g["rs_cls%d" % k]) for k in range(5)],
                                                  [T(g["rs_reg%d" % k]) for k in range(5)], anchors,
                                                  dict(img_shape=(160, 213, 3), pad_shape=(160, 224, 3), scale_factor=1.0),
                                                  json.loads(str(g["rs_cfg"])))
    assert extra is None and tuple(b.shape) == g["rs_props"].shape
    np.testing.assert_allclose(N(s), g["rs_scores"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(N(b), g["rs_props"], rtol=1e-5, atol=1e-3)


def test_scalable_roi_layers_vs_reference():
    """ScalableRoIPool / ScalableRoIAlign (lib/region.py:212-239): RoIs rescaled by b2d_scale_rois, then K7 / K5."""
    g = load_golden("heads2")
    feat, rois = T(g["sc_feat"]), T(g["sc_rois"])
    sp = bregion.ScalableRoIPool(scale=1.3, output_size=(7, 7), spatial_scale=1 / 8)(feat, rois)
    sa = bregion.ScalableRoIAlign(scale=0.8, output_size=(7, 7), spatial_scale=1 / 8, sampling_ratio=2)(feat, rois)
    np.testing.assert_allclose(N(sp), g["sc_pool"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(N(sa), g["sc_align"], rtol=1e-5, atol=1e-6)
    sa_nhwc = bregion.ScalableRoIAlign(scale=0.8, output_size=(7, 7), spatial_scale=1 / 8, sampling_ratio=2)(
        feat.contiguous(memory_format=torch.channels_last), rois)
    assert np.array_equal(N(sa_nhwc), N(sa))


def test_roi_align_nhwc_kernel_fixed_ratio_three():
    """k_roi_align_nhwc (channels_last, fixed sampling ratio != 2) against the generic NCHW kernel and the oracle."""
    rng = np.random.default_rng(11)
    feat = rng.standard_normal((2, 8, 24, 30)).astype(np.float32)
    n = 50
    x1, y1 = rng.uniform(0, 180, n), rng.uniform(0, 140, n)
    rois = np.stack([x1, y1, x1 + rng.uniform(4, 90, n), y1 + rng.uniform(4, 70, n)]).astype(np.float32)
    idx = rng.integers(0, 2, n).astype(np.int32)
    f = T(feat)
    a = bregion.roi_align_levels([f], T(rois), T(idx), [1 / 8], (5, 5), 3, False)
    b = bregion.roi_align_levels([f.contiguous(memory_format=torch.channels_last)], T(rois), T(idx), [1 / 8], (5, 5), 3, False)
    assert np.array_equal(N(a), N(b))
    for img in range(2):
        m = idx == img
        ref = oracle.roi_align(feat[img], np.ascontiguousarray(rois[:, m]), 1 / 8, (5, 5), 3, False)
        np.testing.assert_allclose(N(b)[m], ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("i", [0, 1])
def test_iou_balanced_sampler_numpy_stream_vs_reference(i):
    g = load_golden("heads2")
    mx, ps, nb = (int(v) for v in g["ib_cfg%d" % i])
    np.random.seed(7 + i)
    out = bregion.IoUBalancedNegSampler(mx, ps, num_bins=nb, max_iou=0.5, rng="numpy")(T(g["ib_labels"]), T(g["ib_iou"]), None, None)
    assert np.array_equal(N(out), g["ib_out%d" % i])


def test_iou_balanced_sampler_device_properties():
    """Device RNG mode (one kernel, no host sync): same cardinalities per class as the reference's rule, kept labels
    unchanged, everything else -1, deterministic for a seed, different for another."""
    g = load_golden("heads2")
    lab, iou = g["ib_labels"], g["ib_iou"]
    for (mx, ps, nb) in [(512, 128, 3), (256, 256, 5), (64, 0, 2)]:
        smp = bregion.IoUBalancedNegSampler(mx, ps, num_bins=nb, max_iou=0.5, seed=3)
        out = N(smp(T(lab), T(iou)))
        kept = out >= 0
        assert np.array_equal(out[kept], lab[kept]) and (out[~kept] == -1).all()
        npos = int((lab > 0).sum())
        kp = min(npos, ps)
        assert int((out > 0).sum()) == kp
        num_neg, per = mx - kp, int((mx - kp) / nb)
        width = 0.5 / nb
        lo = [np.float32(j * width) for j in range(nb)][::-1]
        hi = [np.float32(j * width + width) for j in range(nb)][::-1]
        taken = 0
        for j in range(nb):
            member = (lab == 0) & (iou >= lo[j]) & (iou < hi[j])
            quota = per if j < nb - 1 else num_neg - taken
            want = min(int(member.sum()), quota)
            assert int((kept & member).sum()) == want, (mx, ps, nb, j)
            taken += want
        out2 = N(bregion.IoUBalancedNegSampler(mx, ps, num_bins=nb, max_iou=0.5, seed=3)(T(lab), T(iou)))
        assert np.array_equal(out, out2)
        out3 = N(bregion.IoUBalancedNegSampler(mx, ps, num_bins=nb, max_iou=0.5, seed=4)(T(lab), T(iou)))
        assert not np.array_equal(out, out3)


@pytest.mark.parametrize("use_sigmoid,C", [(False, 21), (True, 1), (True, 6)])
@pytest.mark.parametrize("layout", ["rows", "transposed"])
def test_sampled_cross_entropy_vs_torch(use_sigmoid, C, layout):
    """CrossEntropyLoss.forward on sampled rows (lib/losses.py:129-156) against plain PyTorch fp32, forward and gradient;
    the transposed layout is the RPN's tar_cls_out.t() read in place.  Tolerance 1e-5 relative (expf / logf)."""
    import torch.nn.functional as F
    rng = np.random.default_rng(C)
    n = 777
    x = rng.normal(0, 2, (n, C)).astype(np.float32)
    lab = rng.integers(0, 2 if (use_sigmoid and C == 1) else (C + 1 if use_sigmoid else C), n).astype(np.int64)
    base = T(x) if layout == "rows" else T(np.ascontiguousarray(x.T))
    base.requires_grad_(True)
    pred = base if layout == "rows" else base.t()
    loss = bheads.sampled_cross_entropy(pred, T(lab), use_sigmoid, 0.5)
    loss.backward()
    ref_in = T(x).requires_grad_(True)
    if use_sigmoid and C == 1:
        ref = F.binary_cross_entropy_with_logits(ref_in, T(lab).view(-1, 1).float(), reduction="none").sum() * 0.5
    elif use_sigmoid:
        onehot = F.one_hot(T(lab), C + 1)[:, 1:].float()
        ref = F.binary_cross_entropy_with_logits(ref_in, onehot, reduction="none").sum() * 0.5
    else:
        ref = F.cross_entropy(ref_in, T(lab), reduction="none").sum() * 0.5
    ref.backward()
    np.testing.assert_allclose(float(loss), float(ref), rtol=1e-5)
    got = N(base.grad) if layout == "rows" else N(base.grad).T
    np.testing.assert_allclose(got, N(ref_in.grad), rtol=1e-4, atol=1e-6)


def test_sampler_step_counter_survives_cuda_graph_replay():
    """the device sampler's per-step counter lives in device memory, so two replays of a captured step on
    identical labels draw different samples (a by-value seed would be frozen by the capture)."""
    B, n, K = 2, 3000, 8
    rng = np.random.default_rng(3)
    gt, gl = workload.synth_gt(rng, K, 800, 1333)
    cx, cy = rng.uniform(0, 1333, n), rng.uniform(0, 800, n)
    w, h = rng.uniform(8, 400, n), rng.uniform(8, 400, n)
    boxes = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2]).astype(np.float32)
    bt = fused.BatchedTargets(B, n, K, dict(pos_iou=0.5, neg_iou=0.5, min_pos_iou=0.5), dict(max_num=256, pos_num=64),
                              [0, 0, 0, 0], [0.1, 0.1, 0.2, 0.2], DEV, prepend_gt=True, seed=9)
    bx = T(np.stack([boxes] * B)); g = T(np.stack([gt] * B)); lab = T(np.stack([gl] * B))
    cnt = torch.full((B,), n, dtype=torch.int32, device=DEV); gc = torch.full((B,), K, dtype=torch.int32, device=DEV)
    run = lambda: bt(g, gc, lab, boxes=bx, box_count=cnt)
    run(); torch.cuda.synchronize()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            run()
    torch.cuda.current_stream().wait_stream(side)
    graph.replay(); torch.cuda.synchronize()
    first = N(bt.chosen).copy()
    graph.replay(); torch.cuda.synchronize()
    second = N(bt.chosen).copy()
    assert (N(bt.n_chosen) == 256).all()
    assert not np.array_equal(first, second)
    bt.reset_step(); graph.replay(); torch.cuda.synchronize()
    third = N(bt.chosen).copy()
    bt.reset_step(); graph.replay(); torch.cuda.synchronize()
    assert np.array_equal(third, N(bt.chosen))            # same counter value -> same sample (reproducible)


@pytest.mark.parametrize("mode", ["official", "strict"])
@pytest.mark.parametrize("n,use_factor,per_class", [(1500, True, False), (400, False, True), (40, True, False)])
def test_multiclass_nms_kernel_vs_oracle_incl_more_than_16384_candidates(mode, n, use_factor, per_class):
    """utils.multiclass_nms (lib/utils.py:224-269) as one library call: candidates enumerated in numpy (row-major
    boolean-mask order), class-offset NMS by the oracle.  n = 1500 x 20 classes at min_score 0.05 gives ~20 000
    (box, class) candidates: beyond b2d_nms's 16384 boxes, the candidate kernel + large-n NMS path must agree too."""
    rng = np.random.default_rng(n)
    C = 20
    cx, cy = rng.uniform(0, 300, n), rng.uniform(0, 250, n)
    w, h = rng.uniform(8, 100, n), rng.uniform(8, 100, n)
    base = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1).astype(np.float32)
    if per_class:                                                      # [n, 4*C] viewed (n, 4, C)
        bbox = (base[:, :, None] + rng.normal(0, 3, (n, 4, C))).astype(np.float32).reshape(n, 4 * C)
    else:
        bbox = base
    score = (1 / (1 + np.exp(-rng.normal(-2, 2, (n, C))))).astype(np.float32)
    fac = (1 / (1 + np.exp(-rng.normal(0, 1, n)))).astype(np.float32) if use_factor else None
    chans = list(range(1, C))
    kb, ks, kl = butils.multiclass_nms(T(bbox), T(score), chans, 0.6, 0.05, 100, T(fac) if use_factor else None, mode=mode)
    # numpy restatement of the candidate enumeration
    bb3 = bbox.reshape(n, 4, C) if per_class else np.repeat(base[:, :, None], C, 2)
    if mode == "official":
        chosen = (score >= np.float32(0.05)) & np.isin(np.arange(C), chans)[None, :]
        sc = score * fac[:, None] if use_factor else score
        ii, cc = np.nonzero(chosen)
    else:
        cc_all = score.argmax(1)
        best = score[np.arange(n), cc_all]
        chosen = (best >= np.float32(0.05)) & np.isin(cc_all, chans)
        ii = np.nonzero(chosen)[0]
        cc = cc_all[ii]
        sc = np.zeros_like(score)
        sc[np.arange(n), cc_all] = best * fac if use_factor else best
    cb, cs = bb3[ii, :, cc].astype(np.float32), sc[ii, cc].astype(np.float32)
    keep = oracle.batched_nms(cb, cs, cc.astype(np.int64), 0.6)[:100]
    assert np.array_equal(N(kl), cc[keep])
    assert np.array_equal(N(ks), cs[keep]) and np.array_equal(N(kb), cb[keep])


def test_rpn_proposals_packed_records_match_props_and_scores():
    """b2d_rpn_cfg::records: the merge writes (x1, y1, x2, y2, score) rows, zero past count (the all-gather payload)."""
    B = 2
    w = workload.config2(B=B, K=4, channels=8, img_shape=(160, 213), pad_shape=(160, 224))
    pyr = fused.AnchorPyramid(w["strides"], w["grids"])
    cfg = dict(pre_nms=300, post_nms=300, max_num=300, nms_iou=0.7, min_bbox_size=0)
    rp = fused.RpnProposals(pyr, B, cfg, [0, 0, 0, 0], [1, 1, 1, 1], DEV)
    rec = torch.full((B, rp.P, 5), -7.0, device=DEV)
    props, scores, count = rp([T(c) for c in w["cls"]], [T(r) for r in w["reg"]], torch.tensor([[160.0, 213.0]] * B, device=DEV),
                              records=rec)
    torch.cuda.synchronize()
    for b in range(B):
        n = int(count[b])
        assert np.array_equal(N(rec[b, :n, :4]), N(props[b, :, :n]).T) and np.array_equal(N(rec[b, :n, 4]), N(scores[b, :n]))
        assert (N(rec[b, n:]) == 0).all()


_NCCL_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
import b200det
from b200det import dist as bdist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
n_img, M = 6, 5
mine = bdist.image_partition(n_img, rank, world)
recs, cnts = [], []
for g in mine:
    k = g %% (M + 1)
    boxes = (torch.arange(k * 4, dtype=torch.float32).view(k, 4) + g).to(dev)
    sc = (torch.linspace(1, 0.5, k) if k else torch.zeros(0)).to(dev)
    r, c = bdist.pack_detections(boxes, sc, torch.full((k,), g, device=dev), M)
    recs.append(r); cnts.append(c)
rec, cnt = bdist.gather_detections(torch.stack(recs), torch.cat(cnts))
assert rec.shape == (n_img, M, 6) and cnt.tolist() == [g %% (M + 1) for g in range(n_img)], cnt
for g in range(n_img):
    k = g %% (M + 1)
    assert torch.equal(rec[g, :k, 5].cpu(), torch.full((k,), float(g)))
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_nccl_world2_gather_detections(tmp_path):
    """SURVEY 8(e) on hardware: the detection all-gather over NCCL (two ranks, two GPUs).  Skipped on a 1-GPU box."""
    import os, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "w.py"
    script.write_text(_NCCL_WORKER % dict(root=root))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29583")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)


# ------------------------------------------------------------------ the per-image LOOPS rebound at the method level (batched.py)
def _loop_inputs(B=3, K=(3, 6, 1), img_shapes=((250, 310), (256, 320), (200, 333)), pad=(256, 336), channels=16, seed=11):
    from b200det import refpath
    rng = np.random.default_rng(seed)
    strides = workload.STRIDES
    w = workload.config2(B=B, K=1, seed=seed, channels=channels, img_shape=img_shapes[0], pad_shape=pad)
    cls, reg = [T(c) for c in w["cls"]], [T(r) for r in w["reg"]]
    feats = [T(f) for f in w["feats"]]
    gts = [workload.synth_gt(rng, K[i], *img_shapes[i]) for i in range(B)]
    gt_bboxes, gt_labels = [T(g[0]) for g in gts], [T(g[1]) for g in gts]
    metas = [dict(img_shape=(img_shapes[i][0], img_shapes[i][1], 3), pad_shape=(pad[0], pad[1], 3), scale_factor=1.0)
             for i in range(B)]
    return refpath, strides, cls, reg, feats, gt_bboxes, gt_labels, metas


def test_batched_predict_bboxes_from_output_equals_the_per_image_loop():
    """batched.rpn_predict_bboxes_from_output (what install() binds onto RPNHead.predict_bboxes_from_output) returns,
    image by image, exactly what the reference's loop over predict_single_image returns (ragged image sizes)."""
    from b200det import batched
    refpath, strides, cls, reg, feats, gt_bboxes, gt_labels, metas = _loop_inputs()
    seq = refpath.BatchedCallSequence(strides, DEV, rpn_proposal=dict(pre_nms=600, post_nms=300, max_num=500, nms_iou=0.7,
                                                                      min_bbox_size=0))
    cfg = refpath._Cfg(seq.rpn_proposal)
    out = batched.rpn_predict_fast(seq.head, cls, reg, metas, cfg)
    assert out is not None and len(out) == 3 and out[2] == [None] * 3
    grids = [tuple(int(v) for v in c.shape[-2:]) for c in cls]
    anchors = seq.create_anchors(grids)
    for i in range(3):
        b, s, _ = bheads.rpn_predict_single_image(seq.head, [c[i] for c in cls], [r[i] for r in reg], anchors, metas[i], cfg)
        assert out[0][i].shape == b.shape and b.shape[1] > 50
        assert torch.equal(out[0][i], b) and torch.equal(out[1][i], s)
    # not covered (per-image size filter with different scale factors): the caller keeps the reference loop
    metas2 = [dict(m, scale_factor=1.0 + 0.1 * i) for i, m in enumerate(metas)]
    assert batched.rpn_predict_fast(seq.head, cls, reg, metas2, refpath._Cfg(seq.rpn_proposal, min_bbox_size=4)) is None
    seq.head.predict_single_image = lambda *a: bheads.rpn_predict_single_image(seq.head, *a)     # the head's (rebound) method
    loop = batched.rpn_predict_bboxes_from_output(seq.head, cls, reg, metas2, refpath._Cfg(seq.rpn_proposal, min_bbox_size=4))
    assert len(loop[0]) == 3 and loop[0][0].shape[0] == 4


def _find_columns(cols, pool):
    """ascending indices j_0 < j_1 < ... with pool[:, j_k] == cols[:, k] (exact)."""
    idx, j = [], 0
    for k in range(cols.shape[1]):
        while j < pool.shape[1] and not np.array_equal(pool[:, j], cols[:, k]):
            j += 1
        assert j < pool.shape[1], "column %d is not a column of [gt; props] in ascending order" % k
        idx.append(j)
        j += 1
    return np.asarray(idx)


def test_batched_bbox_targets_equal_the_per_image_semantics():
    """batched.bbox_head_bbox_targets (bound onto BBoxHead.bbox_targets): per image the result is what bbox_target
    computes for the SAME sampled set -- columns of [gt; props] in ascending order, labels / GT boxes / deltas /
    is_gt of lib/bbox.py:27-80, the sampler's cardinalities -- for packed proposals (views of the batched predict) and
    for re-packed ragged lists; a second call draws a different sample."""
    from b200det import batched, bbox as bbbox
    refpath, strides, cls, reg, feats, gt_bboxes, gt_labels, metas = _loop_inputs(K=(3, 6, 1))
    seq = refpath.BatchedCallSequence(strides, DEV, rpn_proposal=dict(pre_nms=600, post_nms=300, max_num=500, nms_iou=0.7,
                                                                      min_bbox_size=0),
                                      rcnn_sampler=dict(max_num=128, pos_num=32))
    props = batched.rpn_predict_fast(seq.head, cls, reg, metas, refpath._Cfg(seq.rpn_proposal))[0]
    ragged = [p[:, :p.shape[1] - 7 * i].clone() for i, p in enumerate(props)]       # plain tensors: the re-packing branch
    assign = bregion.MaxIoUAssigner(0.5, 0.5, 0.5)
    first = None
    for plist in (props, ragged, props):
        res = batched.bbox_targets_fast(seq.rcnn_head, plist, gt_bboxes, gt_labels, seq.rcnn_train_cfg)
        assert res is not None and len(res) == 5
        for i in range(3):
            tp, tb, tl, tpar, tg = [N(r[i]) for r in res]
            K = gt_bboxes[i].shape[1]
            pool = np.concatenate([N(gt_bboxes[i]), N(plist[i])], axis=1)
            lab = np.concatenate([np.arange(1, K + 1), N(assign(plist[i], gt_bboxes[i])[0])])
            j = _find_columns(tp, pool)
            assert (lab[j] >= 0).all()
            npos, nneg = int((lab > 0).sum()), int((lab == 0).sum())
            kp = min(npos, 32)
            assert int((lab[j] > 0).sum()) == kp and len(j) == kp + min(128 - kp, nneg)
            gi = np.maximum(lab[j] - 1, 0)
            assert np.array_equal(tl, np.where(lab[j] > 0, N(gt_labels[i])[gi], 0))
            assert np.array_equal(tb, N(gt_bboxes[i])[:, gi]) and np.array_equal(tg, (j < K).astype(np.int64))
            want = N(butils.bbox2param(T(tp), T(tb), (0., 0., 0., 0.), (0.1, 0.1, 0.2, 0.2)))
            np.testing.assert_allclose(tpar, want, rtol=1e-5, atol=1e-6)
        if first is None:
            first = [N(t) for t in res[0]]
        elif plist is props:
            assert any(a.shape != N(b).shape or not np.array_equal(a, N(b)) for a, b in zip(first, res[0])), \
                "the sampler stream did not advance between two calls"
    # host-RNG sampler: not covered, the per-image loop is used
    cfg_np = refpath._Cfg(assigner=seq.rcnn_assigner, sampler=bregion.RandomSampler(128, 32, rng="numpy"))
    assert batched.bbox_targets_fast(seq.rcnn_head, props, gt_bboxes, gt_labels, cfg_np) is None
    loop = batched.bbox_head_bbox_targets(seq.rcnn_head, props, gt_bboxes, gt_labels, cfg_np)
    assert len(loop) == 5 and loop[0][0].shape[0] == 4


def test_batched_anchor_targets_are_differentiable_and_equal_the_per_image_semantics():
    """batched.anchor_head_targets (the target part of AnchorHead.loss): the gradient of tar_cls_out marks the sampled
    anchors; per image they are inside anchors with labels >= 0 of the reference assignment, in ascending order, with
    the sampler's cardinalities, and tar_label / tar_param / tar_reg_out are what anchor_target returns for them."""
    from b200det import batched, anchor as banchor
    refpath, strides, cls, reg, feats, gt_bboxes, gt_labels, metas = _loop_inputs(K=(3, 6, 1))
    seq = refpath.BatchedCallSequence(strides, DEV, rpn_sampler=dict(max_num=64, pos_num=16))
    cls = [c.clone().requires_grad_(True) for c in cls]
    reg = [r.clone().requires_grad_(True) for r in reg]
    ones = [torch.full_like(l, 1) for l in gt_labels]
    tc, tr, tl, tp = batched.anchor_head_targets(seq.head, cls, reg, gt_bboxes, ones, metas, seq.rpn_train_cfg)
    assert tc.shape[0] == 1 and tr.shape[0] == 4 and tc.shape[1] == tr.shape[1] == tl.shape[0] == tp.shape[1]
    (tc.sum() + 2.0 * tr.sum()).backward()
    grids = [tuple(int(v) for v in c.shape[-2:]) for c in cls]
    anchors = torch.cat([a.view(4, -1) for a in seq.create_anchors(grids)], dim=1)
    off = 0
    for i in range(3):
        g_cls = torch.cat([c.grad[i].reshape(1, -1) for c in cls], dim=1)[0]
        g_reg = torch.cat([r.grad[i].reshape(4, -1) for r in reg], dim=1)
        chosen = torch.nonzero(g_cls).view(-1)
        assert torch.equal(g_cls[chosen], torch.ones_like(g_cls[chosen])) and torch.equal(torch.nonzero(g_reg[0]).view(-1), chosen)
        hw = metas[i]['img_shape'][:2]
        in_mask = bregion.inside_anchor_mask(anchors, hw, 0) & torch.cat(
            [bregion.inside_grid_mask(3, hw, grids[l], st, DEV) for l, st in enumerate(strides)]).bool()
        lab_in, _ = seq.rpn_assigner(anchors[:, in_mask], gt_bboxes[i])
        lab = torch.full((anchors.shape[1],), -1, dtype=torch.int64, device=DEV)
        lab[in_mask] = lab_in
        lc = lab[chosen]
        assert bool((lc >= 0).all())
        npos, nneg = int((lab > 0).sum()), int((lab == 0).sum())
        kp = min(npos, 16)
        n = kp + min(64 - kp, nneg)
        assert int((lc > 0).sum()) == kp and chosen.numel() == n
        sl = slice(off, off + n)
        assert torch.equal(tl[sl], (lc > 0).to(torch.int64))
        gi = (lc - 1).clamp(min=0)
        want = butils.bbox2param(anchors[:, chosen], gt_bboxes[i][:, gi], (0., 0., 0., 0.), (1., 1., 1., 1.))
        np.testing.assert_allclose(N(tp[:, sl]), N(want), rtol=1e-5, atol=1e-6)
        reg_flat = torch.cat([r[i].reshape(4, -1) for r in reg], dim=1)
        assert torch.equal(tr[:, sl].detach(), reg_flat[:, chosen].detach())
        off += n
    assert off == tl.numel()
    # a RetinaNet-style call (no sampler) is not covered: None -> the reference loop
    assert batched.anchor_head_targets(seq.head, cls, reg, gt_bboxes, ones, metas, refpath._Cfg(assigner=seq.rpn_assigner)) is None


def test_batched_call_sequence_runs_the_whole_forward_train_path():
    from b200det import refpath
    rp, strides, cls, reg, feats, gt_bboxes, gt_labels, metas = _loop_inputs()
    seqb = refpath.BatchedCallSequence(strides, DEV)
    out = seqb.step(cls, reg, feats, gt_bboxes, gt_labels, metas)
    n = sum(int(t.shape[1]) for t in out["rcnn_targets"][0])
    rois = out["roi_feats"]
    rois = torch.cat(list(rois)) if isinstance(rois, (list, tuple)) else rois
    assert rois.shape[0] == n and rois.shape[1:] == (16, 7, 7) and bool(torch.isfinite(rois).all())


def test_batched_refine_bboxes_equals_the_per_image_loop():
    """batched.refine_bboxes_fast (bound onto BBoxHead.refine_bboxes, the cascade stage loop's last step): ragged
    per-image lists, GT columns dropped, per-class deltas -- image by image what refine_bboxes_single_image returns."""
    from b200det import batched
    rng = np.random.default_rng(31)
    me = types.SimpleNamespace(reg_class_agnostic=False, num_classes=21, target_means=[0.0] * 4, target_stds=[0.05, 0.05, 0.1, 0.1])
    ns = (512, 300, 417)
    props, labels, regs, gts, metas = [], [], [], [], []
    for i, n in enumerate(ns):
        xy = rng.uniform(0, 500, (2, n)); wh = rng.uniform(8, 200, (2, n))
        props.append(T(np.concatenate([xy, xy + wh]).astype(np.float32)))
        labels.append(T(rng.integers(0, 21, n).astype(np.int64)))
        regs.append(T((0.3 * rng.standard_normal((n, 84))).astype(np.float32)))
        g = np.zeros(n, np.int64); g[:5 + i] = 1
        gts.append(T(g)); metas.append(dict(img_shape=(600 + 10 * i, 800 - 7 * i, 3)))
    out = batched.refine_bboxes_fast(me, props, labels, regs, gts, metas)
    assert out is not None and len(out) == 3
    for i in range(3):
        want = bheads.refine_bboxes_single_image(me, props[i], labels[i], regs[i], gts[i], metas[i])
        assert out[i].shape == want.shape == (4, ns[i] - 5 - i) and torch.equal(out[i], want)
    # class-agnostic, no GT flags, no clamp
    me2 = types.SimpleNamespace(reg_class_agnostic=True, num_classes=21, target_means=[0.0] * 4, target_stds=[0.1, 0.1, 0.2, 0.2])
    regs4 = [r[:, :4].contiguous() for r in regs]
    out2 = batched.refine_bboxes_fast(me2, props, labels, regs4, None, None)
    for i in range(3):
        assert torch.equal(out2[i], bheads.refine_bboxes_single_image(me2, props[i], labels[i], regs4[i], None, None))
    # the refined proposals feed the next stage's batched bbox_targets without re-packing
    assert out[0]._b2d_batch[0] is out[1]._b2d_batch[0]


@pytest.mark.parametrize("use_sigmoid", [True, False])
def test_batched_dense_head_predict_equals_the_per_image_loop(use_sigmoid):
    """batched.anchor_head_predict_fast (bound onto AnchorHead.predict_bboxes_from_output for dense heads): image by image
    exactly what the loop over anchor_head_predict_single_image returns (3 images of different sizes, 9 anchors x 6
    classes, per-level top-300, official multiclass NMS)."""
    from b200det import batched, anchor as banchor
    rng = np.random.default_rng(41)
    strides, scales, ratios = [8, 16, 32], [4, 4 * 2 ** (1 / 3), 4 * 2 ** (2 / 3)], [0.5, 1.0, 2.0]
    grids = [(40, 52), (20, 26), (10, 13)]
    C, B = (6, 3) if use_sigmoid else (7, 3)
    creators = [banchor.AnchorCreator(base=s, scales=scales, aspect_ratios=ratios) for s in strides]
    head = types.SimpleNamespace(anchor_strides=strides, anchor_scales=scales, anchor_ratios=ratios, target_means=[0.0] * 4,
                                 target_stds=[1.0] * 4, use_sigmoid=use_sigmoid, cls_channels=C,
                                 num_classes=C + 1 if use_sigmoid else C, anchor_creators=creators)
    cls = [T(rng.standard_normal((B, 9 * C) + g).astype(np.float32) * 2.0 - 1.0) for g in grids]
    reg = [T((0.3 * rng.standard_normal((B, 36) + g)).astype(np.float32)) for g in grids]
    metas = [dict(img_shape=(300 + 9 * i, 400 - 11 * i, 3), scale_factor=1.0) for i in range(B)]
    cfg = dict(pre_nms=300, min_bbox_size=0, min_score=0.3, nms_iou=0.5, max_per_img=100)
    out = batched.anchor_head_predict_fast(head, cls, reg, metas, cfg)
    assert out is not None and len(out) == 3 and len(out[0]) == B
    anchors = [torch.empty((4, 9) + g, device=DEV) for g in grids]
    for i in range(B):
        b, s, l = bheads.anchor_head_predict_single_image(head, [c[i] for c in cls], [r[i] for r in reg], anchors, metas[i], cfg)
        assert b.shape[1] > 5
        assert torch.equal(out[0][i], b) and torch.equal(out[1][i], s) and torch.equal(out[2][i], l)
