"""Pin the CPU oracle (oracle/) against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import hashlib
import json

import numpy as np
import pytest

import oracle
from conftest import fold0, load_golden


def sha(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def test_anchor_grid_bit_exact():
    g = load_golden("anchors")
    for i, m in enumerate(g["meta"]):
        m = json.loads(str(m))
        a = oracle.anchor_grid(m["stride"], tuple(m["grid"]), scales=m["scales"], ratios=m["ratios"],
                               center_lt=m["center_lt"])
        ws, hs = oracle.anchor_sizes(m["stride"], m["scales"], m["ratios"])
        assert np.array_equal(ws, g["ws%d" % i]) and np.array_equal(hs, g["hs%d" % i])
        assert np.array_equal(a, g["a%d" % i])
    grids = [(200, 336), (100, 168), (50, 84), (25, 42), (13, 21)]
    for l, (s, gr) in enumerate(zip([4, 8, 16, 32, 64], grids)):
        a = oracle.anchor_grid(s, gr, scales=[8])
        assert np.array_equal(sha(a), g["full_sha%d" % l])


def test_valid_masks():
    g = load_golden("anchors")
    a = oracle.anchor_grid(64, (13, 21), scales=[8])
    full = oracle.valid_mask(a, (800, 1333), (13, 21), 64, 0)
    assert np.array_equal(full, g["mask_img_b0"] & (g["mask_grid"] > 0))
    b16 = oracle.valid_mask(a, (800, 1333), (13, 21), 64, 16)
    assert np.array_equal(b16, g["mask_img_b16"] & (g["mask_grid"] > 0))
    small = oracle.valid_mask(a, (500, 700), (13, 21), 64, -1)
    assert np.array_equal(small, g["mask_grid_small"] > 0)


def test_calc_iou_bit_exact_including_signed_zero():
    g = load_golden("iou_assign")
    iou = oracle.calc_iou(g["boxes"], g["gt"])
    assert np.array_equal(iou.view(np.uint32), g["iou"].view(np.uint32))
    assert np.array_equal(oracle.calc_iou(g["lit"], g["lit"]), g["lit_iou"])
    e = oracle.elem_iou(g["boxes"][:, :8], g["gt"])
    assert np.array_equal(e, g["elem_iou"])


def test_assigner_bit_exact():
    g = load_golden("iou_assign")
    for i, c in enumerate(g["cfgs"]):
        lab, iou = oracle.assign_max_iou(g["boxes"], g["gt"], *c)
        assert np.array_equal(lab, g["labels%d" % i])
        assert np.array_equal(iou.view(np.uint32), g["miou%d" % i].view(np.uint32))
    for i, mp in enumerate([0.0, 0.3]):
        lab, iou = oracle.assign_max_iou(g["ex_boxes"], g["ex_gt"], 0.5, 0.4, mp)
        assert np.array_equal(lab, g["ex_labels%d" % i])
        assert np.array_equal(fold0(iou), fold0(g["ex_miou%d" % i]))
    assert g["ex_labels0"].tolist() == [1, 2, 2, 1] and g["ex_labels1"].tolist() == [1, 0, 0, 1]
    with pytest.raises(ValueError):
        oracle.assign_max_iou(g["boxes"], np.zeros((4, 0), np.float32), 0.7, 0.3, 0.3)


def full_rpn_anchors():
    grids = [(200, 336), (100, 168), (50, 84), (25, 42), (13, 21)]
    anc = [oracle.anchor_grid(s, gr, scales=[8]) for s, gr in zip([4, 8, 16, 32, 64], grids)]
    mask = np.concatenate([oracle.valid_mask(a, (800, 1333), gr, s, 0)
                           for a, gr, s in zip(anc, grids, [4, 8, 16, 32, 64])])
    return np.concatenate([a.reshape(4, -1) for a in anc], 1), mask


def test_assigner_full_size_config2():
    g = load_golden("assign_full")
    anchors, mask = full_rpn_anchors()
    assert anchors.shape[1] == int(g["n"]) == 268569
    assert np.array_equal(np.packbits(mask), g["mask_packed"])
    inb = np.ascontiguousarray(anchors[:, mask])
    assert inb.shape[1] == 240700
    for i, c in enumerate(g["cfgs"]):
        lab, iou = oracle.assign_max_iou(inb, g["gt"], *c)
        assert np.array_equal(lab, g["labels%d" % i].astype(np.int64))
        assert np.array_equal(sha(fold0(iou)), g["iou_sha%d" % i])


def test_deltas():
    g = load_golden("deltas")
    st = [0.1, 0.1, 0.2, 0.2]
    # encode: logf differs from torch's vectorised log by <= 1 ulp -> 1e-5 relative (north_star tolerance)
    np.testing.assert_allclose(oracle.bbox2param(g["base"], g["bbox"]), g["enc_plain"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(oracle.bbox2param(g["base"], g["bbox"], [0, 0, 0, 0], st), g["enc_norm"],
                               rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(oracle.param2bbox(g["base"], g["param"]), g["dec_plain"], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(oracle.param2bbox(g["base"], g["param"], [0, 0, 0, 0], st, (800, 1333)),
                               g["dec_norm_clamp"], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(
        oracle.param2bbox(g["base"], g["param"], [0.1, -0.1, 0.05, 0.0], [0.05, 0.05, 0.1, 0.1], (800, 1333, 3)),
        g["dec_clamp3"], rtol=1e-5, atol=1e-3)
    rt = oracle.param2bbox(g["base"], oracle.bbox2param(g["base"], g["bbox"]))
    np.testing.assert_allclose(rt, g["roundtrip"], rtol=1e-5, atol=1e-3)
    # batched decode: channel = coord*cls + c (lib/utils.py:101)
    bp = g["bparam"].reshape(4, 5, 64)
    base = np.ascontiguousarray(g["base"][:, :64])
    dec = np.stack([oracle.param2bbox(base, np.ascontiguousarray(bp[:, c]), [0, 0, 0, 0], st, (800, 1333))
                    for c in range(5)], 1).reshape(20, 64)
    np.testing.assert_allclose(dec, g["bdec"], rtol=1e-5, atol=1e-3)


def test_nms_bit_exact():
    g = load_golden("nms")
    for thr in (0.7, 0.5, 0.3):
        assert np.array_equal(oracle.nms(g["b0"], g["s0"], thr), g["keep0_%d" % int(thr * 10)])
    for thr in (0.3, 0.7, 0.5, 0.0):
        assert np.array_equal(oracle.nms(g["b1"], g["s1"], thr), g["keep1_%d" % int(thr * 10)])
    assert oracle.nms(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), 0.5).shape == (0,)
    keep = oracle.batched_nms(g["bn_bbox"], g["bn_score"][:, 0], g["bn_label"], 0.5)
    assert np.array_equal(g["bn_bbox"][keep], g["bn_kb"]) and np.array_equal(g["bn_label"][keep], g["bn_kl"])


def test_level_map_and_roi_align():
    g = load_golden("roi")
    assert np.array_equal(oracle.level_map(g["lm_boxes"]), g["lm_lvls"])
    assert np.array_equal(oracle.level_map(g["rois"]), g["lvls"])
    feats = [g["feat%d" % l][0] for l in range(4)]
    out = oracle.roi_extract(feats, g["rois"])
    np.testing.assert_allclose(out, g["out"], rtol=1e-5, atol=1e-6)
    r = g["rois"]
    np.testing.assert_allclose(oracle.roi_align(feats[1], r, 1 / 8, (7, 7), 0, False), g["ra_adapt"],
                               rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(oracle.roi_align(feats[1], r, 1 / 8, (5, 3), 2, True), g["ra_aligned"],
                               rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(oracle.roi_extract(feats, r), g["single_out"], rtol=1e-5, atol=1e-6)


def test_roi_align_backward():
    g = load_golden("roi")
    lv = g["lvls"]
    for l, s in enumerate((4, 8, 16, 32)):
        m = lv == l
        C, H, W = g["feat%d" % l][0].shape
        gf = oracle.roi_align_bwd(g["gout"][m], (C, H, W), np.ascontiguousarray(g["rois"][:, m]), 1.0 / s)
        np.testing.assert_allclose(gf, g["gfeat%d" % l][0], rtol=1e-4, atol=1e-5)


def test_roi_pool():
    g = load_golden("roi")
    out, arg = oracle.roi_pool(g["feat2"][0], g["rois"], 1 / 16)
    assert np.array_equal(out, g["pool_out"])


def test_rpn_proposals():
    g = load_golden("rpn")
    grids = [(40, 56), (20, 28), (10, 14), (5, 7), (3, 4)]
    anc = [oracle.anchor_grid(s, gr, scales=[8]).reshape(4, -1) for s, gr in zip([4, 8, 16, 32, 64], grids)]
    lg = [g["cls%d" % l].reshape(-1) for l in range(5)]
    dl = [g["reg%d" % l].reshape(4, -1) for l in range(5)]
    img = tuple(g["img_shape"][:2])
    for i, c in enumerate(g["cfgs"]):
        c = json.loads(str(c))
        b, s, lv, ix = oracle.rpn_proposals(lg, dl, anc, c, [0, 0, 0, 0], [1, 1, 1, 1], img)
        assert b.shape == g["props%d" % i].shape, (i, b.shape)
        np.testing.assert_allclose(s, g["scores%d" % i], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(b, g["props%d" % i], rtol=1e-5, atol=1e-3)


def test_atss_oracle_vs_reference():
    """a18: FCOSHead.single_image_targets_atss (with the `//` fix) -- labels and ltrb bit-exact,
    centerness within 1e-6 (sqrt/divide of the vectorised torch kernel)."""
    g = load_golden("atss")
    for tag in "sf":
        grids = [tuple(int(v) for v in x) for x in g["grids_" + tag]]
        cls, reg, ctr = oracle.atss_assign(grids, [8, 16, 32, 64, 128], g["gt_" + tag], g["gl_" + tag], g["img_" + tag][:2])
        assert np.array_equal(cls, g["cls_" + tag])
        assert np.array_equal(reg, g["reg_" + tag])
        np.testing.assert_allclose(ctr, g["ctr_" + tag], rtol=1e-5, atol=1e-6)


def test_rcnn_detect_oracle_vs_reference():
    """BBoxHead.predict_bboxes_single_image (lib/heads/bbox_head.py:122-146)."""
    g = load_golden("heads")
    for i, cfg in enumerate([dict(min_score=0.05, nms_iou=0.5, max_per_img=100, mode="official"),
                             dict(min_score=0.2, nms_iou=0.3, max_per_img=40, mode="strict")]):
        kb, ks, kl = oracle.rcnn_detect(g["det_props"], g["det_cls"], g["det_reg"], (400, 600), [0, 0, 0, 0],
                                        [0.1, 0.1, 0.2, 0.2], **cfg)
        assert np.array_equal(kl, g["det_label%d" % i])
        np.testing.assert_allclose(ks, g["det_score%d" % i], rtol=1e-5, atol=1e-7)      # expf: libm vs Sleef
        np.testing.assert_allclose(kb, g["det_bbox%d" % i], rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("tag", ["s", "f"])
def test_fcos_plain_targets_oracle_vs_reference(tag):
    """FCOSHead.single_image_targets (lib/heads/fcos_head.py:371-416)."""
    g = load_golden("atss")
    grids = [tuple(int(v) for v in x) for x in g["grids_" + tag]]
    c, r, t = oracle.fcos_targets(grids, [8, 16, 32, 64, 128], g["gt_" + tag], g["gl_" + tag], tuple(g["img_" + tag][:2]))
    assert np.array_equal(c, g["pcls_" + tag]) and np.array_equal(r, g["preg_" + tag])
    np.testing.assert_allclose(t, g["pctr_" + tag], rtol=1e-5, atol=1e-6)


def test_anchor_head_loss_oracle_vs_reference():
    """focal + smooth-L1 sums and their autograd gradients (lib/losses.py:33-61, 77-83)."""
    g = load_golden("loss")
    strides, grids = [8, 16, 32], [(20, 28), (10, 14), (5, 7)]
    anc = np.concatenate([oracle.anchor_grid(s, gr, scales=list(g["scales"])).reshape(4, -1) for s, gr in zip(strides, grids)], 1)
    lab, _ = oracle.assign_max_iou(anc, g["gt"], 0.5, 0.4, 0.0)
    f, s1, npos, dc, dr = oracle.anchor_head_loss([g["cls%d" % l] for l in range(3)], [g["reg%d" % l] for l in range(3)], lab, anc,
                                                  g["gt"], g["gl"])
    assert npos == int(g["npos"])
    np.testing.assert_allclose(f, float(g["focal"]), rtol=1e-6)
    np.testing.assert_allclose(s1, float(g["sl1"]), rtol=1e-6)
    for l in range(3):                                   # the golden backward was (2 focal + 3 sl1).backward()
        np.testing.assert_allclose(2 * dc[l], g["dcls%d" % l], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(3 * dr[l], g["dreg%d" % l], rtol=1e-5, atol=1e-6)


def test_config1_c4_oracle_vs_reference():
    """BASELINE config 1 sizes through the oracle: proposals (29 184 anchors, 6000/300), RoIPool, detections."""
    from conftest import c4_inputs
    g = load_golden("c4")
    cls, reg, feat, cls_out, reg_out = c4_inputs()
    anc = oracle.anchor_grid(16, (38, 64), scales=[4, 8, 16, 32]).reshape(4, -1)
    cfg = dict(pre_nms=6000, post_nms=300, max_num=300, nms_iou=0.7, min_bbox_size=0.0)
    pb, ps, _, _ = oracle.rpn_proposals([cls.reshape(-1)], [reg.reshape(4, -1)], [anc], cfg, [0] * 4, [1] * 4, (600, 1000))
    np.testing.assert_allclose(ps, g["scores"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(pb, g["props"], rtol=1e-5, atol=1e-3)
    pooled, _ = oracle.roi_pool(feat[0], np.ascontiguousarray(g["props"][:, ::7]), 1 / 16)
    assert np.array_equal(pooled[:, ::37], g["pooled_sub"])
    kb, ks, kl = oracle.rcnn_detect(g["props"], cls_out, reg_out, (600, 1000), [0] * 4, [0.1, 0.1, 0.2, 0.2], 0.05, 0.3, 100)
    assert np.array_equal(kl, g["det_label"])
    np.testing.assert_allclose(ks, g["det_score"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(kb, g["det_bbox"], rtol=1e-5, atol=1e-3)
