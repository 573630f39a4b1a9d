"""GPU parity of the head-side rows and of BASELINE configs 3-5:
cascade refine / 3-stage targets, RetinaNet dense assignment + per-level top-k NMS, ATSS.
Compared with golden vectors from the unmodified reference and with the CPU oracle."""
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
    counts = [1, 16, 64, 5]
    B, ld = len(counts), 64
    gt = np.zeros((B, 4, ld), np.float32)
    gl = np.zeros((B, ld), np.int64)
    for b, k in enumerate(counts):
        bb, ll = workload.synth_gt(rng, k, *img)
        gt[b, :, :k], gl[b, :k] = bb, ll
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
