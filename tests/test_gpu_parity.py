"""Parity of the CUDA path (through the package -> ctypes -> C-ABI of libb200det.so)
against (1) golden vectors produced by the unmodified reference and (2) the CPU oracle
on seeded inputs.  Bit-exact for labels / indices / IoU / NMS / RoIAlign; 1e-5 relative
(north_star) where expf/logf are involved.  GPU only."""
import json

import numpy as np
import pytest
import torch

import oracle
from conftest import fold0, load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import b200det
    from b200det import anchor as banchor, bbox as bbbox, region as bregion, utils as butils, fused, workload
    DEV = torch.device("cuda:0")


def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def N(t):
    return t.detach().cpu().numpy()


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


# ------------------------------------------------------------------ K1 / a2
def test_anchor_grid_bit_exact():
    g = load_golden("anchors")
    for i, m in enumerate(g["meta"]):
        m = json.loads(str(m))
        ac = banchor.AnchorCreator(base=m["stride"], scales=m["scales"], aspect_ratios=m["ratios"],
                                   center_lt=m["center_lt"], device=DEV)
        a = N(ac(m["stride"], tuple(m["grid"])))
        assert np.array_equal(bits(a), bits(g["a%d" % i])), m
        assert np.array_equal(N(ac.anchor_ws), g["ws%d" % i])
    import hashlib
    grids = [(200, 336), (100, 168), (50, 84), (25, 42), (13, 21)]
    for l, (s, gr) in enumerate(zip([4, 8, 16, 32, 64], grids)):
        a = N(banchor.AnchorCreator(base=s, scales=[8], device=DEV)(s, gr))
        sha = np.frombuffer(hashlib.sha1(a.tobytes()).digest(), dtype=np.uint8)
        assert np.array_equal(sha, g["full_sha%d" % l])


def test_inside_masks():
    g = load_golden("anchors")
    a = banchor.AnchorCreator(base=64, scales=[8], device=DEV)(64, (13, 21)).view(4, -1)
    assert np.array_equal(N(bregion.inside_anchor_mask(a, (800, 1333), 0)), g["mask_img_b0"])
    assert np.array_equal(N(bregion.inside_anchor_mask(a, (800, 1333), 16)), g["mask_img_b16"])
    assert N(bregion.inside_anchor_mask(a, (800, 1333), -1)).all()
    assert np.array_equal(N(bregion.inside_grid_mask(3, (800, 1333), (13, 21), 64, DEV)), g["mask_grid"])
    assert np.array_equal(N(bregion.inside_grid_mask(3, (500, 700), (13, 21), 64, DEV)), g["mask_grid_small"])


# ------------------------------------------------------------------ a3 / K2
def test_calc_iou_bit_exact():
    g = load_golden("iou_assign")
    iou = N(butils.calc_iou(T(g["boxes"]), T(g["gt"])))
    assert np.array_equal(bits(iou), bits(g["iou"]))          # including -0.0
    assert np.array_equal(N(butils.calc_iou(T(g["lit"]), T(g["lit"]))), g["lit_iou"])
    assert np.array_equal(N(butils.elem_iou(T(g["boxes"][:, :8].copy()), T(g["gt"]))), g["elem_iou"])


def test_assigner_bit_exact_vs_reference():
    g = load_golden("iou_assign")
    for i, c in enumerate(g["cfgs"]):
        lab, iou = bregion.MaxIoUAssigner(*c)(T(g["boxes"]), T(g["gt"]))
        assert np.array_equal(N(lab), g["labels%d" % i]), i
        assert np.array_equal(bits(N(iou)), bits(g["miou%d" % i])), i
    for i, mp in enumerate([0.0, 0.3]):
        lab, iou = bregion.MaxIoUAssigner(0.5, 0.4, mp)(T(g["ex_boxes"]), T(g["ex_gt"]))
        assert np.array_equal(N(lab), g["ex_labels%d" % i])
        assert np.array_equal(fold0(N(iou)), fold0(g["ex_miou%d" % i]))
    with pytest.raises(ValueError):
        bregion.MaxIoUAssigner(0.7, 0.3, 0.3)(T(g["boxes"]), torch.zeros((4, 0), device=DEV))


@pytest.mark.parametrize("K", [1, 3, 64, 700])
def test_assigner_vs_oracle_many_gt(K):
    rng = np.random.default_rng(K)
    gt, _ = workload.synth_gt(rng, K, 800, 1333)
    cx, cy = rng.uniform(0, 1333, 5000), rng.uniform(0, 800, 5000)
    w, h = rng.uniform(8, 400, 5000), rng.uniform(8, 400, 5000)
    boxes = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2]).astype(np.float32)
    boxes[:, :K] = gt                                            # exact hits
    for cfg in [(0.7, 0.3, 0.3), (0.5, 0.4, 0.0)]:
        lab, iou = bregion.MaxIoUAssigner(*cfg)(T(boxes), T(gt))
        olab, oiou = oracle.assign_max_iou(boxes, gt, *cfg)
        assert np.array_equal(N(lab), olab)
        assert np.array_equal(bits(N(iou)), bits(oiou))


def test_assigner_full_size_fused_pyramid_config2():
    """268 569 in-register anchors x 8 GT (BASELINE config 2) == reference labels."""
    g = load_golden("assign_full")
    grids = [(200, 336), (100, 168), (50, 84), (25, 42), (13, 21)]
    pyr = fused.AnchorPyramid((4, 8, 16, 32, 64), grids)
    mask = np.unpackbits(g["mask_packed"])[: int(g["n"])].astype(bool)
    gt = torch.zeros((2, 4, 8), device=DEV)
    gt[:] = T(g["gt"])
    gcount = torch.full((2,), 8, dtype=torch.int32, device=DEV)
    img_hw = torch.tensor([[800.0, 1333.0]] * 2, device=DEV)
    for i, c in enumerate(g["cfgs"]):
        bt = fused.BatchedTargets(2, pyr.total, 8, dict(pos_iou=c[0], neg_iou=c[1], min_pos_iou=c[2]),
                                  dict(max_num=256, pos_num=128), [0, 0, 0, 0], [1, 1, 1, 1], DEV, pyramid=pyr,
                                  border=0.0 if i == 0 else 0.0)
        bt(gt, gcount, None, img_hw=img_hw)
        for b in range(2):
            lab = N(bt.labels[b])
            assert (lab[~mask] == -1).all()
            assert np.array_equal(lab[mask], g["labels%d" % i].astype(np.int64))
            iou = N(bt.iou[b])[mask]
            ref = np.zeros_like(iou)
            ref[g["iou_nz_idx%d" % i]] = g["iou_nz_val%d" % i]
            assert np.array_equal(fold0(iou), ref)
            cen = N(bt.census[b])
            assert cen[0] == (lab > 0).sum() and cen[1] == (lab == 0).sum()


# ------------------------------------------------------------------ K8
def test_deltas_vs_reference():
    g = load_golden("deltas")
    st, z = [0.1, 0.1, 0.2, 0.2], [0, 0, 0, 0]
    base, bbox, param = T(g["base"]), T(g["bbox"]), T(g["param"])
    np.testing.assert_allclose(N(butils.bbox2param(base, bbox)), g["enc_plain"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(N(butils.bbox2param(base, bbox, z, st)), g["enc_norm"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(N(butils.param2bbox(base, param)), g["dec_plain"], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(N(butils.param2bbox(base, param, z, st, (800, 1333))), g["dec_norm_clamp"],
                               rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(N(butils.param2bbox(base, param, [0.1, -0.1, 0.05, 0.0], [0.05, 0.05, 0.1, 0.1],
                                                   (800, 1333, 3))), g["dec_clamp3"], rtol=1e-5, atol=1e-3)
    rt = butils.param2bbox(base, butils.bbox2param(base, bbox))
    np.testing.assert_allclose(N(rt), g["roundtrip"], rtol=1e-5, atol=1e-3)
    bd = butils.batched_param2bbox(T(g["base"][:, :64].copy()), T(g["bparam"]), z, st, (800, 1333))
    np.testing.assert_allclose(N(bd), g["bdec"], rtol=1e-5, atol=1e-3)
    assert np.array_equal(N(butils.clamp_bbox(T(g["bbox"] - 100), (600, 1000))), g["clamp"])


# ------------------------------------------------------------------ K4
def test_nms_bit_exact_vs_torchvision():
    g = load_golden("nms")
    for thr in (0.7, 0.5, 0.3):
        assert np.array_equal(N(butils.nms(T(g["b0"]), T(g["s0"]), thr)), g["keep0_%d" % int(thr * 10)])
    for thr in (0.3, 0.7, 0.5, 0.0):
        assert np.array_equal(N(butils.nms(T(g["b1"]), T(g["s1"]), thr)), g["keep1_%d" % int(thr * 10)]), thr
    assert butils.nms(torch.zeros((0, 4), device=DEV), torch.zeros(0, device=DEV), 0.5).numel() == 0


@pytest.mark.parametrize("n", [1, 63, 64, 65, 2000, 6000])
def test_nms_vs_oracle_sizes_and_idempotence(n):
    rng = np.random.default_rng(n)
    ctr = rng.uniform(0, 1000, (2, max(n // 30, 1)))
    idx = rng.integers(0, ctr.shape[1], n)
    c = ctr[:, idx] + rng.normal(0, 8, (2, n))
    wh = rng.uniform(20, 200, (2, n))
    b = np.concatenate([c - wh / 2, c + wh / 2]).T.astype(np.float32).copy()
    s = rng.permutation(n).astype(np.float32)
    for thr in (0.7, 0.5):
        keep = N(butils.nms(T(b), T(s), thr))
        assert np.array_equal(keep, oracle.nms(b, s, thr))
        again = N(butils.nms(T(b[keep]), T(s[keep]), thr))       # NMS of its own output keeps everything
        assert np.array_equal(again, np.arange(keep.size))


def test_batched_and_multiclass_nms_vs_reference():
    g = load_golden("nms")
    kb, ks, kl = butils.batched_nms(T(g["bn_bbox"]), T(g["bn_score"][:, 0].copy()), T(g["bn_label"]), 0.5)
    assert np.array_equal(N(kb), g["bn_kb"]) and np.array_equal(N(ks), g["bn_ks"]) and np.array_equal(N(kl), g["bn_kl"])
    C = g["bn_score"].shape[1]
    for mode in ("official", "strict"):
        kb, ks, kl = butils.multiclass_nms(T(g["bn_bbox"]), T(g["bn_score"]), list(range(1, C)), 0.5, 0.05, 100, mode=mode)
        assert np.array_equal(N(kb), g["mc_%s_b" % mode])
        assert np.array_equal(N(ks), g["mc_%s_s" % mode]) and np.array_equal(N(kl), g["mc_%s_l" % mode])
    kb, ks, kl = butils.multiclass_nms(T(g["mc_bbc"]), T(g["bn_score"]), list(range(1, C)), 0.5, 0.05, 50,
                                       score_factor=T(g["mc_fac"]))
    assert np.array_equal(N(kb), g["mc_pc_b"]) and np.array_equal(N(kl), g["mc_pc_l"])
    np.testing.assert_allclose(N(ks), g["mc_pc_s"], rtol=1e-6)
    kb, ks, kl = butils.multiclass_nms(T(g["mc_bbc"]), T(g["bn_score"]), list(range(1, C)), 0.5, 0.05, 50, mode="strict")
    assert np.array_equal(N(kb), g["mc_pcs_b"]) and np.array_equal(N(kl), g["mc_pcs_l"])


# ------------------------------------------------------------------ K3 (+K4)
def _rpn_run(g, cfg, B=2):
    grids = [(40, 56), (20, 28), (10, 14), (5, 7), (3, 4)]
    pyr = fused.AnchorPyramid((4, 8, 16, 32, 64), grids)
    rp = fused.RpnProposals(pyr, B, cfg, [0, 0, 0, 0], [1, 1, 1, 1], DEV)
    cls = [T(np.stack([g["cls%d" % l]] * B)) for l in range(5)]
    reg = [T(np.stack([g["reg%d" % l]] * B)) for l in range(5)]
    img_hw = torch.tensor([[float(g["img_shape"][0]), float(g["img_shape"][1])]] * B, device=DEV)
    props, scores, count = rp(cls, reg, img_hw)
    torch.cuda.synchronize()
    return N(props), N(scores), N(count), N(rp.prov)


def test_rpn_proposals_vs_reference():
    g = load_golden("rpn")
    for i, c in enumerate(g["cfgs"]):
        c = json.loads(str(c))
        props, scores, count, prov = _rpn_run(g, c)
        ref_b, ref_s = g["props%d" % i], g["scores%d" % i]
        for b in range(2):
            n = int(count[b])
            assert n == ref_b.shape[1], (i, n, ref_b.shape)
            np.testing.assert_allclose(scores[b, :n], ref_s, rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(props[b, :, :n], ref_b, rtol=1e-5, atol=1e-3)


def test_rpn_head_predict_single_image_method_form_vs_reference():
    """heads.rpn_predict_single_image with the reference's own signature (lib/heads/rpn_head.py:68) on a stand-in
    head object carrying the attributes RPNHead defines (lib/heads/anchor_head.py:29-50)."""
    import types
    from b200det import heads as bheads
    g = load_golden("rpn")
    grids = [(40, 56), (20, 28), (10, 14), (5, 7), (3, 4)]
    head = types.SimpleNamespace(anchor_strides=[4, 8, 16, 32, 64], anchor_scales=[8], anchor_ratios=[0.5, 1.0, 2.0],
                                 target_means=[0.0] * 4, target_stds=[1.0] * 4, use_sigmoid=True, cls_channels=1,
                                 anchor_creators=[banchor.AnchorCreator(base=s, scales=[8]) for s in (4, 8, 16, 32, 64)])
    for ac in head.anchor_creators:
        ac.to(DEV)                                                   # (returns None, like the reference's)
    anchors = [ac(s, gr) for ac, s, gr in zip(head.anchor_creators, head.anchor_strides, grids)]
    meta = dict(img_shape=tuple(int(v) for v in g["img_shape"]), pad_shape=tuple(int(v) for v in g["pad_shape"]), scale_factor=1.0)
    cls, reg = [T(g["cls%d" % l]) for l in range(5)], [T(g["reg%d" % l]) for l in range(5)]
    for i, c in enumerate(g["cfgs"]):
        b, s_, extra = bheads.rpn_predict_single_image(head, cls, reg, anchors, meta, json.loads(str(c)))
        assert extra is None and tuple(b.shape) == g["props%d" % i].shape
        np.testing.assert_allclose(N(s_), g["scores%d" % i], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(N(b), g["props%d" % i], rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("cut", ["0", "1.0", "1.3", "2"])
def test_rpn_score_cut_nms_is_exact(cut, setknob):
    """Score-cut NMS (csrc/nms.cu k_nms_cut): pass 1 on the cut * max_num best boxes + conditional full pass must
    give the selection of the plain per-level NMS for every cut factor (1.0: pass 1 always falls short -> the
    fallback pass does the work; 2: the default; 0: disabled), on clustered boxes (low survival) and at config-2
    sizes, against the oracle's provenance (level, index) of every proposal."""
    setknob(B2D_NMS_CUT=cut)
    rng = np.random.default_rng(17)
    grids = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
    strides = (4, 8, 16, 32, 64)
    pyr = fused.AnchorPyramid(strides, grids)
    B = 2
    cls = [rng.normal(0, 1, (B, 3) + g).astype(np.float32) for g in grids]
    reg = [rng.normal(0, 0.15, (B, 12) + g).astype(np.float32) for g in grids]         # small deltas: heavy overlap
    img_hw = torch.tensor([[400.0, 666.0]] * B, device=DEV)
    anc = [oracle.anchor_grid(s, gr, scales=[8]).reshape(4, -1) for s, gr in zip(strides, grids)]
    offs = np.cumsum([0] + [a.shape[1] for a in anc])
    for cfg in (dict(pre_nms=1000, post_nms=1000, max_num=1000, nms_iou=0.7, min_bbox_size=0),
                dict(pre_nms=1000, post_nms=250, max_num=700, nms_iou=0.7, min_bbox_size=0)):   # per-level cap binds
        rp = fused.RpnProposals(pyr, B, cfg, [0, 0, 0, 0], [1, 1, 1, 1], DEV)
        props, scores, count = rp([T(c) for c in cls], [T(r) for r in reg], img_hw)
        torch.cuda.synchronize()
        for b in range(B):
            _, _, lv, ix = oracle.rpn_proposals([c[b].reshape(-1) for c in cls], [r[b].reshape(4, -1) for r in reg], anc, cfg,
                                                [0, 0, 0, 0], [1, 1, 1, 1], (400, 666))
            n = int(count[b])
            assert n == lv.shape[0]
            assert np.array_equal(N(rp.prov[b, :n]), offs[lv] + ix), (cut, cfg)


@pytest.mark.gpu
@pytest.mark.parametrize("sweep", ["1", "0"])
def test_rpn_sweep_nms_concentrated_boxes(sweep, setknob):
    """The x-sweep mask kernel of NMS pass 1 (csrc/nms.cu k_nms_sweep) gives up when the x1 values of a segment are
    concentrated (here: huge deltas, every box clips to the full image or to the same left edge) and the dense pass
    must then produce the reference selection; same check with the sweep disabled."""
    setknob(B2D_NMS_SWEEP=sweep)
    rng = np.random.default_rng(23)
    grids = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
    strides = (4, 8, 16, 32, 64)
    pyr = fused.AnchorPyramid(strides, grids)
    B = 2
    cls = [rng.normal(0, 1, (B, 3) + g).astype(np.float32) for g in grids]
    anc = [oracle.anchor_grid(s, gr, scales=[8]).reshape(4, -1) for s, gr in zip(strides, grids)]
    offs = np.cumsum([0] + [a.shape[1] for a in anc])
    cfg = dict(pre_nms=1000, post_nms=1000, max_num=1000, nms_iou=0.7, min_bbox_size=0)
    for shift in (4.0, 1.5):                      # 4: every box is the whole image; 1.5: wide boxes sharing clipped edges
        reg = [(rng.normal(0, 0.05, (B, 12) + g) + shift).astype(np.float32) for g in grids]
        rp = fused.RpnProposals(pyr, B, cfg, [0, 0, 0, 0], [1, 1, 1, 1], DEV)
        props, scores, count = rp([T(c) for c in cls], [T(r) for r in reg], torch.tensor([[400.0, 666.0]] * B, device=DEV))
        torch.cuda.synchronize()
        for b in range(B):
            _, _, lv, ix = oracle.rpn_proposals([c[b].reshape(-1) for c in cls], [r[b].reshape(4, -1) for r in reg], anc, cfg,
                                                [0, 0, 0, 0], [1, 1, 1, 1], (400, 666))
            n = int(count[b])
            assert n == lv.shape[0], (sweep, shift)
            assert np.array_equal(N(rp.prov[b, :n]), offs[lv] + ix), (sweep, shift)


def test_rpn_proposals_selection_bit_exact_vs_oracle():
    g = load_golden("rpn")
    grids = [(40, 56), (20, 28), (10, 14), (5, 7), (3, 4)]
    anc = [oracle.anchor_grid(s, gr, scales=[8]).reshape(4, -1) for s, gr in zip([4, 8, 16, 32, 64], grids)]
    offs = np.cumsum([0] + [a.shape[1] for a in anc])
    lg = [g["cls%d" % l].reshape(-1) for l in range(5)]
    dl = [g["reg%d" % l].reshape(4, -1) for l in range(5)]
    for i, c in enumerate(g["cfgs"]):
        c = json.loads(str(c))
        _, _, count, prov = _rpn_run(g, c, B=1)
        _, _, lv, ix = oracle.rpn_proposals(lg, dl, anc, c, [0, 0, 0, 0], [1, 1, 1, 1], tuple(g["img_shape"][:2]))
        assert np.array_equal(prov[0, :count[0]], offs[lv] + ix), i



@pytest.mark.parametrize("knobs", [{}, {"B2D_NMS_CUT": "0"}, {"B2D_NMS_SWEEP": "0"}, {"B2D_RPN_FRONT": "0", "B2D_RPN_BACK": "0"}])
def test_rpn_proposals_benchmarked_config2_full_size_vs_oracle(knobs, setknob):
    """The configuration bench.py times -- workload.config2(B=8): 200x336 ... 13x21 grids, 268 569 anchors per image,
    pre/post/max 2000, thr 0.7 -- through fused.RpnProposals against oracle.rpn_proposals for EVERY image: provenance
    (level, anchor index) of every proposal bit-exact, boxes / scores within 1e-5.  Default knobs (fused cluster
    kernels, score-cut + sweep NMS), score cut off, sweep off, and the round-1 multi-kernel chain."""
    setknob(**knobs)
    B = 8
    w = workload.config2(B=B, K=8, with_feats=False)
    grids, strides = w["grids"], w["strides"]
    pyr = fused.AnchorPyramid(strides, grids)
    cfg = dict(pre_nms=2000, post_nms=2000, max_num=2000, nms_iou=0.7, min_bbox_size=0)
    rp = fused.RpnProposals(pyr, B, cfg, [0, 0, 0, 0], [1, 1, 1, 1], DEV)
    img_hw = torch.tensor([[800.0, 1333.0]] * B, device=DEV)
    props, scores, count = rp([T(c) for c in w["cls"]], [T(r) for r in w["reg"]], img_hw)
    torch.cuda.synchronize()
    anc = [oracle.anchor_grid(s, gr, scales=[8]).reshape(4, -1) for s, gr in zip(strides, grids)]
    offs = np.cumsum([0] + [a.shape[1] for a in anc])
    for b in range(B):
        ob, osc, lv, ix = oracle.rpn_proposals([c[b].reshape(-1) for c in w["cls"]], [r[b].reshape(4, -1) for r in w["reg"]],
                                               anc, cfg, [0, 0, 0, 0], [1, 1, 1, 1], (800, 1333))
        n = int(count[b])
        assert n == lv.shape[0] == 2000, (b, n)
        assert np.array_equal(N(rp.prov[b, :n]), offs[lv] + ix), (knobs, b)
        np.testing.assert_allclose(N(scores[b, :n]), osc, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(N(props[b, :, :n]), ob, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("front", ["1", "0"])
@pytest.mark.parametrize("kind", ["all_equal", "few_values", "near_constant", "normal_small_k", "keep_all"])
def test_rpn_front_degenerate_score_maps_vs_oracle(kind, front, setknob):
    """The cluster selection kernel (csrc/rpn_front.cu) on score maps that defeat the 12-bit radix pass: every score
    equal (zero-initialised head: the lowest indices must win), five distinct values (exact ties beyond the sort
    capacity), scores inside one float octave (narrowing passes resolve them), plus a small top-k and a level that
    keeps everything -- against the oracle's provenance; same inputs through the multi-kernel path (front=0)."""
    setknob(B2D_RPN_FRONT=front)
    rng = np.random.default_rng(5)
    grids = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
    strides = (4, 8, 16, 32, 64)
    pyr = fused.AnchorPyramid(strides, grids)
    B = 2
    shape = lambda g: (B, 3) + g
    if kind == "all_equal":
        cls = [np.zeros(shape(g), np.float32) for g in grids]
    elif kind == "few_values":
        cls = [rng.integers(0, 5, shape(g)).astype(np.float32) for g in grids]
    elif kind == "near_constant":
        cls = [(1.0 + rng.uniform(0, 1e-3, shape(g))).astype(np.float32) for g in grids]
    else:
        cls = [rng.normal(0, 1, shape(g)).astype(np.float32) for g in grids]
    reg = [rng.normal(0, 0.5, (B, 12) + g).astype(np.float32) for g in grids]
    cfg = dict(pre_nms=1000, post_nms=1000, max_num=1000, nms_iou=0.7, min_bbox_size=0)
    if kind == "normal_small_k":
        cfg = dict(pre_nms=37, post_nms=20, max_num=60, nms_iou=0.7, min_bbox_size=0)
    if kind == "keep_all":
        grids = grids[2:]; strides = strides[2:]; cls = cls[2:]; reg = reg[2:]
        pyr = fused.AnchorPyramid(strides, grids)
        cfg = dict(pre_nms=4000, post_nms=4000, max_num=3000, nms_iou=0.7, min_bbox_size=8)
    anc = [oracle.anchor_grid(s, gr, scales=[8]).reshape(4, -1) for s, gr in zip(strides, grids)]
    offs = np.cumsum([0] + [a.shape[1] for a in anc])
    rp = fused.RpnProposals(pyr, B, cfg, [0, 0, 0, 0], [1, 1, 1, 1], DEV)
    props, scores, count = rp([T(c) for c in cls], [T(r) for r in reg], torch.tensor([[400.0, 666.0]] * B, device=DEV))
    torch.cuda.synchronize()
    for b in range(B):
        ob, osc, lv, ix = oracle.rpn_proposals([c[b].reshape(-1) for c in cls], [r[b].reshape(4, -1) for r in reg], anc, cfg,
                                               [0, 0, 0, 0], [1, 1, 1, 1], (400, 666))
        n = int(count[b])
        assert n == lv.shape[0], (kind, front, n, lv.shape)
        assert np.array_equal(N(rp.prov[b, :n]), offs[lv] + ix), (kind, front)
        np.testing.assert_allclose(N(props[b, :, :n]), ob, rtol=1e-5, atol=1e-3)


def test_topk_large_and_ties():
    rng = np.random.default_rng(7)
    v = rng.standard_normal(201600).astype(np.float32)
    for k in (1, 2000, 12000):
        assert np.array_equal(N(bregion.topk_desc(T(v), k)), oracle.topk_desc(v, k))
    tied = np.zeros(50000, np.float32)                            # all equal -> lowest indices, in order
    assert np.array_equal(N(bregion.topk_desc(T(tied), 3000)), np.arange(3000))
    few = rng.integers(0, 5, 40000).astype(np.float32)            # heavy ties across the radix bins
    assert np.array_equal(N(bregion.topk_desc(T(few), 9000)), oracle.topk_desc(few, 9000))
    small = rng.standard_normal(1000).astype(np.float32)
    assert np.array_equal(N(bregion.topk_desc(T(small), 2000)), oracle.topk_desc(small, 1000))


def test_proposal_creator_legacy():
    g = load_golden("rpn")
    pb, ps = bregion.ProposalCreator(200, 50, 0.7, 4)(T(g["pc_cls"]), T(g["pc_reg"]), T(g["pc_anchor"]), (160, 213))
    np.testing.assert_allclose(N(ps), g["pc_scores"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(N(pb), g["pc_props"], rtol=1e-5, atol=1e-3)


# ------------------------------------------------------------------ K5 / K6 / K7
def _roi_feats(g, channels_last):
    fs = [T(g["feat%d" % l]) for l in range(4)]
    if channels_last:
        fs = [f.contiguous(memory_format=torch.channels_last) for f in fs]
    return fs


@pytest.mark.parametrize("channels_last", [False, True])
def test_roi_extractor_vs_reference(channels_last):
    g = load_golden("roi")
    ext = bregion.BasicRoIExtractor([dict(type="RoIAlign", spatial_scale=1 / s, sampling_ratio=2) for s in (4, 8, 16, 32)],
                                    output_size=(7, 7))
    fs = [f.requires_grad_(True) for f in _roi_feats(g, channels_last)]
    rois = T(g["rois"])
    assert np.array_equal(N(ext.map_rois_to_levels(rois, 4)), g["lvls"])
    assert np.array_equal(N(ext.map_rois_to_levels(T(g["lm_boxes"]), 4)), g["lm_lvls"])
    out = ext(fs, [rois])[0]
    np.testing.assert_allclose(N(out), g["out"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(bits(N(out)), bits(g["out"])), "RoIAlign is expected to be bit-identical to the CPU reference"
    (out * T(g["gout"])).sum().backward()
    for l in range(4):
        np.testing.assert_allclose(N(fs[l].grad), g["gfeat%d" % l], rtol=1e-4, atol=1e-5)


def test_roi_align_variants_vs_torchvision():
    g = load_golden("roi")
    f1 = T(g["feat1"])
    r5 = torch.cat([torch.zeros((96, 1), device=DEV), T(g["rois"]).t()], 1)
    np.testing.assert_allclose(N(bregion.roi_align(f1, r5, (7, 7), 1 / 8, 0, False)), g["ra_adapt"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(N(bregion.roi_align(f1, r5, (5, 3), 1 / 8, 2, True)), g["ra_aligned"], rtol=1e-5, atol=1e-6)
    fs = _roi_feats(g, False)
    sre = bregion.SingleRoIExtractor("RoIAlign", 7, [4, 8, 16, 32])
    np.testing.assert_allclose(N(sre(fs, [T(g["rois"])])[0]), g["single_out"], rtol=1e-5, atol=1e-6)
    sra = bregion.ScalableRoIAlign(scale=1.5, output_size=(7, 7), spatial_scale=1 / 8, sampling_ratio=2)
    np.testing.assert_allclose(N(sra(f1, r5)), g["scalable_out"], rtol=1e-5, atol=1e-6)


def test_roi_align_bf16_nhwc_close():
    g = load_golden("roi")
    fs = [f.to(torch.bfloat16).contiguous(memory_format=torch.channels_last) for f in _roi_feats(g, False)]
    out = bregion.roi_align_levels(fs, T(g["rois"]), None, [1 / 4, 1 / 8, 1 / 16, 1 / 32])
    ref = oracle.roi_extract([N(f[0].float()) for f in fs], g["rois"])
    np.testing.assert_allclose(N(out), ref, rtol=1e-5, atol=1e-6)      # same bf16-rounded inputs, fp32 math


def _ring_case(seed, C, K, B=2):
    """RoIs that exercise every branch of the TMA-ring kernel (csrc/roi_align_tma.cu): level-mapped sizes,
    RoIs wider than the ring holds (generic in-kernel path), sub-cell RoIs, RoIs on / beyond the borders."""
    rng = np.random.default_rng(seed)
    grids = [(48, 80), (24, 40), (12, 20), (6, 10)]                       # 192 x 320 image, strides 4..32
    feats = [rng.standard_normal((B, C) + gsz).astype(np.float32) for gsz in grids]
    Wimg, Himg = 320.0, 192.0
    w = np.exp(rng.uniform(np.log(2.0), np.log(400.0), K)); h = np.exp(rng.uniform(np.log(2.0), np.log(300.0), K))
    cx = rng.uniform(-20, Wimg + 20, K); cy = rng.uniform(-20, Himg + 20, K)
    rois = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2]).astype(np.float32)
    rois[:, :8] = np.array([[0, 0, 319, 191], [-50, -50, 400, 300], [100, 80, 100.2, 80.1], [310, 185, 330, 200],
                            [-300, -300, -200, -250], [0, 0, 170, 30], [5, 5, 60, 190], [318.5, 190.5, 319, 191]], np.float32).T
    img = rng.integers(0, B, K).astype(np.int32)
    return grids, feats, rois, img


@pytest.mark.parametrize("mode", ["1", "2", "3"])
@pytest.mark.parametrize("C,K", [(256, 700), (128, 333)])
def test_roi_align_tma_ring_bit_exact_vs_oracle(C, K, mode, setknob):
    """The two opt-in TMA forms of K5 -- 1: bulk-copy row ring (roi_align_tma.cu), 2: tensor-map band kernel
    (roi_align_tband.cu, cp.async.bulk.tensor.4d / UTMALDG; RoIs wider than 16 cells take its in-kernel window path),
    3: heterogeneous launch (a persistent band CTA per SM takes every m-th RoI beside the window kernel)."""
    setknob(B2D_ROI_TMA=mode)                                             # opt-in kernels
    grids, feats, rois, img = _ring_case(11, C, K)
    fs = [T(f).contiguous(memory_format=torch.channels_last) for f in feats]
    out = N(bregion.roi_align_levels(fs, T(rois), T(img), [1 / 4, 1 / 8, 1 / 16, 1 / 32]))
    for b in range(feats[0].shape[0]):
        m = img == b
        ref = oracle.roi_extract([f[b] for f in feats], np.ascontiguousarray(rois[:, m]))
        assert np.array_equal(bits(out[m]), bits(ref)), "ring kernel must be bit-identical to the oracle"
    # single level, one image, aligned=True (bin sizes may be ~0 / negative -> generic path), 5x3 bins
    r1 = np.ascontiguousarray(rois[:, img == 0])
    o1 = N(bregion.roi_align_levels([fs[1][:1]], T(r1), None, [1 / 8], out_size=(5, 3), aligned=True))
    ref1 = oracle.roi_align(feats[1][0], r1, 1 / 8, (5, 3), 2, True)
    assert np.array_equal(bits(o1), bits(ref1))


def test_roi_align_tma_ring_bf16_and_l1_path_agree(setknob):
    grids, feats, rois, img = _ring_case(12, 256, 500)
    fs = [T(f).to(torch.bfloat16).contiguous(memory_format=torch.channels_last) for f in feats]
    setknob(B2D_ROI_TMA=1)
    out = N(bregion.roi_align_levels(fs, T(rois), T(img), [1 / 4, 1 / 8, 1 / 16, 1 / 32]))
    setknob(B2D_ROI_TMA=0)                                                # L1-path kernel (k_roi_align_win)
    out_l1 = N(bregion.roi_align_levels(fs, T(rois), T(img), [1 / 4, 1 / 8, 1 / 16, 1 / 32]))
    assert np.array_equal(bits(out), bits(out_l1))
    for b in range(2):
        m = img == b
        ref = oracle.roi_extract([N(f[b].float()) for f in fs], np.ascontiguousarray(rois[:, m]))
        assert np.array_equal(bits(out[m]), bits(ref))


@pytest.mark.parametrize("form", ["2", "1"])
@pytest.mark.parametrize("C,K", [(128, 300), (256, 500)])
def test_roi_align_backward_tile_kernel_vs_oracle_and_generic(C, K, form, setknob):
    """K6 in its two NHWC forms -- 2: per-RoI patches + ordered merge (csrc/roi_align_bwd_patch.cu, default), 1: tile
    gather (csrc/roi_align_bwd_tile.cu) -- against the oracle's sequential backward (1e-5 relative: the summation order
    inside a cell differs from torchvision's), against the generic torchvision-ordered kernel, and run-to-run
    bit-identical; RoIs incl. borders / outside / sub-cell / wider than the map."""
    setknob(B2D_ROI_BWD_TILE=form)
    grids, feats, rois, img = _ring_case(13, C, K)
    fs = [T(f).contiguous(memory_format=torch.channels_last).requires_grad_(True) for f in feats]
    rng = np.random.default_rng(5)
    go = rng.standard_normal((K, C, 7, 7)).astype(np.float32)
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]

    def run():
        for f in fs:
            f.grad = None
        (bregion.roi_align_levels(fs, T(rois), T(img), scales) * T(go)).sum().backward()
        return [N(f.grad).copy() for f in fs]

    g1 = run()
    g2 = run()
    for l in range(4):
        assert np.array_equal(bits(g1[l]), bits(g2[l])), "backward must be run-to-run deterministic"
    setknob(B2D_ROI_BWD_TILE=0)                                           # generic kernel (torchvision's order)
    g0 = run()
    lv = oracle.level_map(rois)
    # 1e-5 relative plus an absolute term for cancelling sums: a cell adds up to ~100 tap terms of magnitude ~1, so two
    # summation orders differ by a few 2^-24 of the LARGEST partial sum, whatever the result (torchvision's own CUDA
    # backward, atomics in arrival order, has the same spread).  The patch form also merges tap weights per cell.
    atol = 2e-6 if form == "1" else 5e-6
    for l, s_ in enumerate((4, 8, 16, 32)):
        np.testing.assert_allclose(g1[l], g0[l], rtol=1e-5, atol=atol)
        for b in range(feats[0].shape[0]):
            m = (lv == l) & (img == b)
            ref = oracle.roi_align_bwd(go[m], (C,) + grids[l], np.ascontiguousarray(rois[:, m]), 1.0 / s_)
            np.testing.assert_allclose(g1[l][b], ref, rtol=1e-5, atol=atol)


def test_roi_align_backward_patch_form_blocks_and_fallback(setknob):
    """Patch form of K6 beyond its common case: (a) patches larger than one 64 x 64 block of weight lists (a RoI forced
    onto the finest level, thin RoIs along the border) against the oracle and the tile form; (b) patches that do not fit
    the scratch raise the device-side fallback flag -- the patch kernels return at once and the guarded tile kernel
    behind them produces the gradient, bit-identical to the tile form alone, without a host synchronisation."""
    rng = np.random.default_rng(3)
    B, C, H, W = 2, 128, 80, 100
    feat = rng.standard_normal((B, C, H, W)).astype(np.float32)
    img = np.array([0, 1, 1, 0, 0, 1], np.int32)
    go = rng.standard_normal((6, C, 7, 7)).astype(np.float32)

    def grads_of(rois, form):
        setknob(B2D_ROI_BWD_TILE=form)
        f = T(feat).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        (bregion.roi_align_levels([f], T(rois), T(img), [1 / 4]) * T(go)).sum().backward()
        return N(f.grad).copy()

    # (a) 98 x 78, 1 x 100 and 80 x 2 cell patches next to ordinary ones; total area below the 16 K-cell scratch floor
    rois = np.array([[10, 10, 60, 50], [0, 0, 390, 310], [100, 40, 180, 120], [30, 200, 90, 260], [0, 318, 399, 319],
                     [398, 0, 399, 319]], np.float32).T
    g2, g1 = grads_of(rois, "2"), grads_of(rois, "1")
    np.testing.assert_allclose(g2, g1, rtol=1e-5, atol=5e-6)
    ref = np.zeros_like(feat)
    for b in range(B):
        m = img == b
        ref[b] = oracle.roi_align_bwd(go[m], (C, H, W), np.ascontiguousarray(rois[:, m]), 1 / 4)
    np.testing.assert_allclose(g2, ref, rtol=1e-5, atol=5e-6)
    assert np.array_equal(bits(g2), bits(grads_of(rois, "2"))), "run-to-run deterministic"
    # (b) three patches of ~7 600 cells: 22 K cells > 16 K -> fallback
    big = rois.copy()
    big[:, 0] = [0, 0, 388, 312]; big[:, 2] = [4, 2, 392, 316]
    assert np.array_equal(bits(grads_of(big, "2")), bits(grads_of(big, "1"))), "fallback must be the tile kernel's result"


def test_roi_align_backward_full_size_properties(setknob):
    """K6 at BASELINE config-2 / 3 geometry (the sampled RoIs of a 2-image step, 256 channels, full pyramid): the patch
    form against the tile form (rounding only), run-to-run bit-identical, exactly linear under a power-of-two scaling of
    the incoming gradient, and the size-independent checksum of the operator: every counted sample's four bilinear
    weights sum to one, so sum over all cells of grad_feat[:, c] == sum over (RoI, bin) of grad_out[:, c] x (fraction of
    the bin's 4 samples that lie on the map)."""
    B, K, C = 2, 8, 256
    w = workload.config2(B=B, K=K, channels=C)
    hp = fused.TrainHotPath(B, w["grids"], DEV, gt_ld=K, feat_channels=C)
    cls, reg = [T(c) for c in w["cls"]], [T(r) for r in w["reg"]]
    feats = [T(f).contiguous(memory_format=torch.channels_last) for f in w["feats"]]
    gcount = torch.full((B,), K, dtype=torch.int32, device=DEV)
    img_hw = torch.tensor([[800.0, 1333.0]] * B, device=DEV)
    out = hp.step(cls, reg, feats, T(w["gt"]), gcount, T(w["gt_label"]), img_hw)
    bt = out["rcnn"]
    rois = torch.cat([bt.tar_box[b, :, :int(bt.n_chosen[b])] for b in range(B)], 1).contiguous()
    roi_img = torch.cat([torch.full((int(bt.n_chosen[b]),), b, dtype=torch.int32, device=DEV) for b in range(B)])
    R = rois.shape[1]
    assert R == B * 512
    go = torch.randn((R, C, 7, 7), device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]

    def grads(form, g):
        setknob(B2D_ROI_BWD_TILE=form)
        fs = [f.detach().clone().requires_grad_(True) for f in feats]
        bregion.roi_align_levels(fs, rois, roi_img, scales).backward(g)
        return [f.grad for f in fs]

    gp, gt_ = grads("2", go), grads("1", go)
    for a, b in zip(gp, gt_):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=5e-6)
    for a, b in zip(gp, grads("2", go)):
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), "run-to-run deterministic"
    for a, b in zip(gp, grads("2", go * 4.0)):
        assert torch.equal((a * 4.0).view(torch.int32), b.view(torch.int32)), "linear under exact scaling"
    # checksum: a sample is dropped only if it lies beyond the map (y < -1 or y > H: thin RoIs at the image border, whose
    # size is raised to one cell); vy / vx = fraction of a bin's two samples per axis that count
    r = N(rois)
    lv = oracle.level_map(r)
    sc = np.array([1 / 4, 1 / 8, 1 / 16, 1 / 32], np.float32)[lv]
    Hs, Ws = np.array([g[0] for g in w["grids"]])[lv], np.array([g[1] for g in w["grids"]])[lv]
    frac = []
    for lo, hi, size in ((r[1], r[3], Hs), (r[0], r[2], Ws)):
        st = lo * sc
        b = np.maximum(hi * sc - st, np.float32(1)) / np.float32(7)
        pos = st[:, None, None] + np.arange(7, dtype=np.float32)[None, :, None] * b[:, None, None] + \
            (np.arange(2, dtype=np.float32)[None, None, :] + np.float32(0.5)) * b[:, None, None] / np.float32(2)
        frac.append(((pos >= -1) & (pos <= size[:, None, None])).mean(2))                 # [R, 7]
    wgt = torch.from_numpy(frac[0][:, :, None] * frac[1][:, None, :]).to(DEV)              # [R, 7, 7]
    total = sum(g.double().sum(dim=(0, 2, 3)) for g in gp)                # [C]
    want = (go.double() * wgt[:, None].double()).sum(dim=(0, 2, 3))
    torch.testing.assert_close(total, want, rtol=1e-6, atol=1e-3)


def test_roi_align_reference_layout_takes_fast_kernels():
    """fp32 NCHW features (the reference's FPN output layout, lib/necks.py) are transposed once and go through the same
    K5 / K6 kernels as channels_last inputs: identical forward bits, identical gradients (returned for the NCHW input)."""
    grids, feats, rois, img = _ring_case(14, 64, 200)
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    go = np.random.default_rng(2).standard_normal((200, 64, 7, 7)).astype(np.float32)
    outs, grads = [], []
    for cl in (False, True):
        fs = [T(f) for f in feats]
        if cl:
            fs = [f.contiguous(memory_format=torch.channels_last) for f in fs]
        fs = [f.requires_grad_(True) for f in fs]
        o = bregion.roi_align_levels(fs, T(rois), T(img), scales)
        (o * T(go)).sum().backward()
        outs.append(N(o)); grads.append([N(f.grad) for f in fs])
    assert np.array_equal(bits(outs[0]), bits(outs[1]))
    for l in range(4):
        assert np.array_equal(bits(grads[0][l]), bits(grads[1][l]))
    for b in range(feats[0].shape[0]):
        m = img == b
        assert np.array_equal(bits(outs[0][m]), bits(oracle.roi_extract([f[b] for f in feats], np.ascontiguousarray(rois[:, m]))))


def test_roi_pool_vs_torchvision():
    g = load_golden("roi")
    ext = bregion.BasicRoIExtractor([dict(type="RoIPool", spatial_scale=1 / 16, sampling_ratio=2)], output_size=(7, 7))
    f = T(g["feat2"]).requires_grad_(True)
    out = ext([f], [T(g["rois"])])[0]
    assert np.array_equal(N(out), g["pool_out"])
    (out * T(g["gout"])).sum().backward()
    np.testing.assert_allclose(N(f.grad), g["pool_gfeat"], rtol=1e-5, atol=1e-6)


def test_roi_align_multi_image_batch_and_determinism():
    rng = np.random.default_rng(3)
    B, C = 3, 32
    grids = [(40, 56), (20, 28), (10, 14), (5, 7)]
    feats = [rng.standard_normal((B, C) + gsz).astype(np.float32) for gsz in grids]
    rois = [np.sort(rng.uniform(0, 160, (2, 2, 40)), axis=1).reshape(4, 40)[[0, 2, 1, 3]].astype(np.float32) for _ in range(B)]
    ext = bregion.BasicRoIExtractor([dict(type="RoIAlign", spatial_scale=1 / s, sampling_ratio=2) for s in (4, 8, 16, 32)])
    fs = [T(f).contiguous(memory_format=torch.channels_last).requires_grad_(True) for f in feats]
    outs = ext(fs, [T(r) for r in rois])
    for b in range(B):
        ref = oracle.roi_extract([f[b] for f in feats], rois[b])
        assert np.array_equal(bits(N(outs[b])), bits(ref))
    go = rng.standard_normal((B * 40, C, 7, 7)).astype(np.float32)
    (torch.cat(outs) * T(go)).sum().backward()
    g1 = [N(f.grad).copy() for f in fs]
    lv = [oracle.level_map(r) for r in rois]
    for l, s in enumerate((4, 8, 16, 32)):
        for b in range(B):
            m = lv[b] == l
            ref = oracle.roi_align_bwd(go[b * 40:(b + 1) * 40][m], (C,) + grids[l], np.ascontiguousarray(rois[b][:, m]), 1.0 / s)
            np.testing.assert_allclose(g1[l][b], ref, rtol=1e-4, atol=1e-5)
    for f in fs:
        f.grad = None
    (torch.cat(ext(fs, [T(r) for r in rois])) * T(go)).sum().backward()
    for l in range(4):
        assert np.array_equal(bits(N(fs[l].grad)), bits(g1[l])), "backward must be run-to-run deterministic"


# ------------------------------------------------------------------ a5 / a6 / a13
def test_targets_numpy_rng_mode_vs_reference():
    g = load_golden("targets")
    grids = [(40, 56), (20, 28), (10, 14), (5, 7), (3, 4)]
    H, W = 160, 213
    acs = [banchor.AnchorCreator(base=s, scales=[8], device=DEV) for s in (4, 8, 16, 32, 64)]
    anchors = torch.cat([ac(s, gr).view(4, -1) for ac, s, gr in zip(acs, (4, 8, 16, 32, 64), grids)], 1)
    in_mask = bregion.inside_anchor_mask(anchors, (H, W), 0) & torch.cat(
        [bregion.inside_grid_mask(3, (H, W), gr, s, DEV) for gr, s in zip(grids, (4, 8, 16, 32, 64))]).bool()
    cls = torch.cat([T(g["cls%d" % l][0]).view(1, -1) for l in range(5)], 1)
    reg = torch.cat([T(g["reg%d" % l][0]).view(4, -1) for l in range(5)], 1)
    np.random.seed(2019)
    r = banchor.anchor_target(cls, reg, 1, anchors[:, in_mask], in_mask, T(g["gt"]), None,
                              dict(type="MaxIoUAssigner", pos_iou=0.7, neg_iou=0.3, min_pos_iou=0.3),
                              bregion.RandomSampler(64, 32, rng="numpy"), [0, 0, 0, 0], [1, 1, 1, 1])
    assert np.array_equal(N(r[0]), g["at_cls"]) and np.array_equal(N(r[1]), g["at_reg"])
    assert np.array_equal(N(r[2]), g["at_lab"])
    np.testing.assert_allclose(N(r[5]), g["at_par"], rtol=1e-5, atol=1e-6)
    allm = torch.ones_like(in_mask)
    grid_only = torch.cat([bregion.inside_grid_mask(3, (H, W), gr, s, DEV) for gr, s in zip(grids, (4, 8, 16, 32, 64))]).bool()
    r2 = banchor.anchor_target(cls, reg, 1, anchors[:, grid_only & allm], grid_only, T(g["gt"]), T(g["gl"]),
                               dict(type="MaxIoUAssigner", pos_iou=0.5, neg_iou=0.4, min_pos_iou=0.0), None,
                               [0, 0, 0, 0], [1, 1, 1, 1])
    assert np.array_equal(N(r2[2]), g["at2_lab"]) and np.array_equal(N(r2[0]), g["at2_cls"])
    np.testing.assert_allclose(N(r2[5]), g["at2_par"], rtol=1e-5, atol=1e-6)
    np.random.seed(2019)
    r3 = bbbox.bbox_target(T(g["props"]), T(g["gt"]), T(g["gl"]),
                           dict(type="MaxIoUAssigner", pos_iou=0.5, neg_iou=0.5, min_pos_iou=0.5),
                           bregion.RandomSampler(128, 32, rng="numpy"), (0., 0., 0., 0.), (0.1, 0.1, 0.2, 0.2))
    assert np.array_equal(N(r3[0]), g["bt_props"]) and np.array_equal(N(r3[1]), g["bt_bbox"])
    assert np.array_equal(N(r3[2]), g["bt_label"]) and np.array_equal(N(r3[4]), g["bt_isgt"])
    np.testing.assert_allclose(N(r3[3]), g["bt_param"], rtol=1e-5, atol=1e-5)
    np.random.seed(2019)
    assert np.array_equal(N(bregion.RandomSampler(256, 64, rng="numpy")(T(g["rs_in"]))), g["rs_out"])
    np.random.seed(2019)
    assert np.array_equal(N(bregion.IoUBalancedNegSampler(256, 64, rng="numpy")(T(g["rs_in"]), T(g["ib_iou"]), None, None)), g["ib_out"])


@pytest.mark.parametrize("threads", ["1024", "128"])
@pytest.mark.parametrize("n,max_num,pos_num", [(3000, 256, 64), (3000, 256, 2000), (100, 256, 128), (20000, 512, 128), (268569, 256, 128)])
def test_device_sampler_matches_its_spec_and_properties(n, max_num, pos_num, threads, setknob):
    from oracle import sampler_spec
    setknob(B2D_SAMPLE_THREADS=threads)                  # 128: the slim CTA form (4 permutation steps per thread and round)
    rng = np.random.default_rng(n + max_num)
    labels = rng.choice([-1, 0, 0, 0, 1, 2, 3], n).astype(np.int64)
    pos_num = min(pos_num, max_num)
    s = bregion.RandomSampler(max_num, pos_num, rng="device", seed=5)
    out = N(s(T(labels)))
    chosen = np.nonzero(out >= 0)[0]
    spec = sampler_spec.sample(labels, max_num, pos_num, (5 * 1000003 + 1) & ((1 << 64) - 1))
    assert np.array_equal(chosen, spec)
    assert np.array_equal(out[chosen], labels[chosen]) and (labels[chosen] >= 0).all()
    npos, nneg = (labels > 0).sum(), (labels == 0).sum()
    kp = min(npos, pos_num)
    assert (out > 0).sum() == kp and (out == 0).sum() == min(max_num - kp, nneg)
    out2 = N(bregion.RandomSampler(max_num, pos_num, rng="device", seed=5)(T(labels)))
    assert np.array_equal(out, out2)


@pytest.mark.parametrize("n,K,max_num,pos_num,spread,prepend", [
    (2000, 8, 512, 128, 1.0, True),        # config 2 (few positives)
    (2000, 24, 512, 128, 0.15, True),      # proposals hug the GTs: #pos > pos_num -> threshold bisection
    (4096, 64, 1024, 256, 0.3, True),      # largest supported problem, 16384-step permutation domain
    (300, 3, 64, 16, 0.5, False),          # no GT prepend, fewer negatives than wanted
    (40, 2, 512, 128, 0.2, True),          # fewer candidates than max_num
])
def test_roi_targets_fused_equals_three_kernel_path_and_spec(n, K, max_num, pos_num, spread, prepend):
    """b2d_roi_targets_fused (one launch) == b2d_assign_max_iou + b2d_sample_labels + b2d_encode_targets,
    and its `chosen` == the sampler specification (oracle/sampler_spec.py) on the oracle's labels."""
    from oracle import sampler_spec
    rng = np.random.default_rng(n + K)
    B = 3
    gt = np.zeros((B, 4, K), np.float32)
    for b in range(B):
        x1 = rng.uniform(0, 900, K); y1 = rng.uniform(0, 500, K)
        gt[b] = np.stack([x1, y1, x1 + rng.uniform(30, 400, K), y1 + rng.uniform(30, 300, K)])
    boxes = np.zeros((B, 4, n), np.float32)
    for b in range(B):
        j = rng.integers(0, K, n)
        jit = rng.normal(0, spread * 120, (4, n))
        far = rng.random(n) < 0.3
        bb = gt[b][:, j] + jit
        rnd = np.stack([rng.uniform(0, 900, n), rng.uniform(0, 500, n), rng.uniform(900, 1300, n), rng.uniform(500, 790, n)])
        bb = np.where(far[None], rnd, bb)
        boxes[b] = np.stack([np.minimum(bb[0], bb[2]), np.minimum(bb[1], bb[3]), np.maximum(bb[0], bb[2]) + 1, np.maximum(bb[1], bb[3]) + 1])
    counts = np.array([n, max(n - 7, 1), n], np.int32)
    gl = rng.integers(1, 21, (B, K)).astype(np.int64)
    gcount = np.array([K, max(K - 1, 1), K], np.int32)
    kw = dict(assigner=dict(pos_iou=0.5, neg_iou=0.5, min_pos_iou=0.5), sampler=dict(max_num=max_num, pos_num=pos_num),
              means=(0, 0, 0, 0), stds=(0.1, 0.1, 0.2, 0.2), device=DEV, prepend_gt=prepend, seed=5)
    a = fused.BatchedTargets(B, n, K, **kw)
    c = fused.BatchedTargets(B, n, K, **kw)
    assert a.fused
    c.fused = False                                                       # three-kernel path
    args = (T(gt), T(gcount), T(gl))
    kws = dict(boxes=T(boxes), box_count=T(counts))
    a(*args, **kws); c(*args, **kws)
    torch.cuda.synchronize()
    for name in ("census", "n_chosen", "chosen", "tar_box", "tar_gt", "tar_param", "tar_label", "tar_is_gt"):
        x, y = N(getattr(a, name)), N(getattr(c, name))
        if name == "census":
            x, y = x[:, :2], y[:, :2]
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), name
    seed = (5 * 1000003 + 1) & 0xFFFFFFFFFFFFFFFF
    saw_bisect = False
    for b in range(B):
        lead = int(gcount[b]) if prepend else 0
        nb = int(counts[b])
        olab, _ = oracle.assign_max_iou(np.ascontiguousarray(boxes[b][:, :nb]), np.ascontiguousarray(gt[b][:, :gcount[b]]), 0.5, 0.5, 0.5)
        full = np.concatenate([np.arange(1, lead + 1), olab]).astype(np.int64)
        assert np.array_equal(N(a.labels[b, :lead + nb]), full)
        saw_bisect |= int((full > 0).sum()) > pos_num
        want = sampler_spec.sample(full, max_num, pos_num, seed, image_index=b)
        m = int(a.n_chosen[b])
        assert np.array_equal(N(a.chosen[b, :m]), want)
    if spread <= 0.15:
        assert saw_bisect, "case meant to exercise the positive-threshold bisection"


# ------------------------------------------------------------------ fused train path vs oracle
def test_fused_train_path_small_vs_oracle():
    import __graft_entry__
    __graft_entry__.smoke()


def test_fused_train_path_full_size_properties():
    """BASELINE config 2 sizes (B=2 here): size-independent properties + stage-wise oracle checks."""
    B, K = 2, 8
    w = workload.config2(B=B, K=K, channels=32)
    hp = fused.TrainHotPath(B, w["grids"], DEV, gt_ld=K, feat_channels=32)
    cls, reg = [T(c) for c in w["cls"]], [T(r) for r in w["reg"]]
    feats = [T(f).contiguous(memory_format=torch.channels_last) for f in w["feats"]]
    gt, gl = T(w["gt"]), T(w["gt_label"])
    gcount = torch.full((B,), K, dtype=torch.int32, device=DEV)
    img_hw = torch.tensor([[800.0, 1333.0]] * B, device=DEV)
    out = hp.step(cls, reg, feats, gt, gcount, gl, img_hw)
    torch.cuda.synchronize()
    for b in range(B):
        n = int(out["prop_count"][b])
        assert n == 2000
        sc = N(out["scores"][b, :n])
        assert (np.diff(sc) <= 0).all()                              # global top-k is score-descending
        pb = N(out["props"][b, :, :n])
        assert (pb[0] >= 0).all() and (pb[2] <= 1332).all() and (pb[1] >= 0).all() and (pb[3] <= 799).all()
        # RPN targets: counts and label/param consistency
        rt = out["rpn"]
        m = int(rt.n_chosen[b])
        ch = N(rt.chosen[b, :m])
        assert m == 256 and (np.diff(ch) > 0).all()
        lab = N(rt.labels[b])
        assert (lab[ch] >= 0).all() and (lab[ch] > 0).sum() <= 128
        assert np.array_equal(N(rt.tar_label[b, :m]), (lab[ch] > 0).astype(np.int64))
        enc = oracle.bbox2param(N(rt.tar_box[b, :, :m]), N(rt.tar_gt[b, :, :m]))
        np.testing.assert_allclose(N(rt.tar_param[b, :, :m]), enc, rtol=1e-5, atol=1e-6)
        flat_cls = np.concatenate([c[b].reshape(-1) for c in w["cls"]])
        assert np.array_equal(N(out["rpn_tar_cls"][b, 0, :m]), flat_cls[ch])
        # RoI targets: assignment of the GPU's own proposals == oracle; GT rows first
        bt = out["rcnn"]
        olab, oiou = oracle.assign_max_iou(pb, w["gt"][b], 0.5, 0.5, 0.5)
        assert np.array_equal(N(bt.labels[b, K:K + n]), olab)
        assert np.array_equal(N(bt.labels[b, :K]), np.arange(1, K + 1))
        mm = int(bt.n_chosen[b])
        assert mm == 512 and N(bt.tar_is_gt[b, :mm]).sum() <= K
        rois = N(bt.tar_box[b, :, :mm])
        ref = oracle.roi_extract([f[b] for f in w["feats"]], rois)
        assert np.array_equal(bits(N(out["roi_feats"][b * 512:b * 512 + mm])), bits(ref))


def test_fused_step_is_run_to_run_deterministic():
    """Five replays of the config-2-sized step (per-level chains on internal streams, score-cut NMS, fused targets)
    give bit-identical proposals, samples, targets and RoI features -- a race between the chains would show here."""
    B, K = 4, 8
    w = workload.config2(B=B, K=K, channels=64)
    hp = fused.TrainHotPath(B, w["grids"], DEV, gt_ld=K, feat_channels=64, overlap=True)
    cls, reg = [T(c) for c in w["cls"]], [T(r) for r in w["reg"]]
    feats = [T(f).contiguous(memory_format=torch.channels_last) for f in w["feats"]]
    gt, gl = T(w["gt"]), T(w["gt_label"])
    gcount = torch.full((B,), K, dtype=torch.int32, device=DEV)
    img_hw = torch.tensor([[800.0, 1333.0]] * B, device=DEV)
    ref = None
    for it in range(5):
        hp.reset_step()                                             # same sampler stream every replay
        out = hp.step(cls, reg, feats, gt, gcount, gl, img_hw)
        torch.cuda.synchronize()
        snap = [N(out["props"]).copy(), N(out["scores"]).copy(), N(out["prop_count"]).copy(), N(out["rpn"].chosen).copy(),
                N(out["rpn"].tar_param).copy(), N(out["rcnn"].chosen).copy(), N(out["rcnn"].tar_param).copy(),
                N(out["roi_feats"]).copy()]
        if ref is None:
            ref = snap
        else:
            for a, b in zip(ref, snap):
                assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), it


def _tap_cells(rois, img, grids, strides, B, finest=56.0):
    """Exact set of (level, image, y, x) cells under the 2x2-per-bin bilinear taps of 7x7 RoIAlign (numpy restatement of
    torchvision's pre_calc_for_bilinear_interpolate index arithmetic, fp32)."""
    masks = [np.zeros((B,) + tuple(g), bool) for g in grids]
    r = rois.astype(np.float32)
    s = np.sqrt(((r[2] - r[0]) + np.float32(1)) * ((r[3] - r[1]) + np.float32(1)))
    lv = np.clip(np.floor(np.log2(s / np.float32(finest) + np.float32(1e-6))), 0, len(grids) - 1).astype(int)
    for i in range(r.shape[1]):
        l = lv[i]
        H, W = grids[l]
        sc = np.float32(1.0 / strides[l])
        sx, sy = r[0, i] * sc, r[1, i] * sc
        bw = np.maximum(r[2, i] * sc - sx, np.float32(1)) / np.float32(7)
        bh = np.maximum(r[3, i] * sc - sy, np.float32(1)) / np.float32(7)
        def axis(st, b, n):
            out = set()
            for p in range(7):
                for k in range(2):
                    v = st + np.float32(p) * b + (np.float32(k) + np.float32(0.5)) * b / np.float32(2)
                    v = max(v, np.float32(0))
                    lo = int(v)
                    if lo >= n - 1:
                        lo = hi = n - 1
                    else:
                        hi = lo + 1
                    out.update((lo, hi))
            return sorted(out)
        ys, xs = axis(sy, bh, H), axis(sx, bw, W)
        masks[l][img[i]][np.ix_(ys, xs)] = True
    return masks


def test_mark_and_fetch_touched_cells_cover_every_tap(setknob):
    """b2d_roi_mark_cells / b2d_fetch_marked_cells (sparse host -> device transfer of the RoI extractor's inputs): the
    bitmap holds exactly the cells under a bilinear tap (numpy restatement of the tap index arithmetic);
    exactly the marked cells are copied from pinned channels_last host maps; RoIAlign on a NaN-poisoned device pyramid
    that received only those cells is bit-identical to RoIAlign on the whole pyramid."""
    import ctypes
    grids, feats, rois, img = _ring_case(21, 64, 600)
    B, strides = feats[0].shape[0], [4, 8, 16, 32]
    order = np.argsort(img, kind="stable")
    rois, img = rois[:, order], img[order]
    ld = int(np.bincount(img, minlength=B).max())
    rb = np.zeros((B, 4, ld), np.float32)
    counts = np.bincount(img, minlength=B).astype(np.int32)
    for b in range(B):
        rb[b, :, :counts[b]] = rois[:, img == b]
    ra = fused.BatchedRoIAlign(B, ld, [(64,) + tuple(g) for g in grids], strides, DEV, layout=1)
    h_feats = [fused.pinned_channels_last(torch.from_numpy(f)) for f in feats]
    full = [h.to(DEV) for h in h_feats]
    want = N(ra(full, T(rb), T(counts))).copy()
    poisoned = [torch.full_like(f, float("nan")) for f in full]
    moved = ra.fetch_touched(poisoned, h_feats, T(rb), T(counts))
    got = N(ra(poisoned, T(rb), T(counts)))
    for b in range(B):
        sl = slice(b * ld, b * ld + counts[b])
        assert np.array_equal(bits(got[sl]), bits(want[sl]))
    # bitmap against the exact tap set
    words = N(ra._bitmap).view(np.uint32)
    exact = _tap_cells(rois, img, grids, strides, B)
    w0, n_marked, n_exact = 0, 0, 0
    for l, g in enumerate(grids):
        cells = B * g[0] * g[1]
        nw = (cells + 31) // 32
        bitsl = np.unpackbits(words[w0:w0 + nw].view(np.uint8), bitorder="little")[:cells].astype(bool).reshape((B,) + tuple(g))
        assert np.array_equal(bitsl, exact[l]), "marked cells != cells under a tap (level %d)" % l
        copied = ~torch.isnan(poisoned[l]).any(1).cpu().numpy()            # [B,H,W]: cells that received data
        assert np.array_equal(copied, bitsl), "copied cells != marked cells (level %d)" % l
        assert np.array_equal(N(poisoned[l]).transpose(0, 2, 3, 1)[bitsl], feats[l].transpose(0, 2, 3, 1)[bitsl])
        n_marked += int(bitsl.sum()); n_exact += int(exact[l].sum()); w0 += nw
    assert int(moved[0]) == n_marked
    assert n_marked == n_exact


def test_mark_and_fetch_ragged_counts_bf16_single_level():
    """Sparse transfer at its edges: an image without RoIs, a single-level extractor, bf16 channels_last maps (a cell is
    C * 2 bytes), C = 8 (one 16-byte chunk per cell): RoIAlign on the fetched cells == RoIAlign on the whole map."""
    rng = np.random.default_rng(11)
    B, C, H, W, ld = 3, 8, 40, 60, 16
    feat = rng.standard_normal((B, C, H, W)).astype(np.float32)
    counts = np.array([16, 0, 5], np.int32)
    rb = np.zeros((B, 4, ld), np.float32)
    for b in range(B):
        x1 = rng.uniform(-10, 400, ld); y1 = rng.uniform(-10, 280, ld)
        rb[b] = np.stack([x1, y1, x1 + rng.uniform(1, 200, ld), y1 + rng.uniform(1, 150, ld)])
    for layout, dt in ((1, torch.float32), (2, torch.bfloat16)):
        ra = fused.BatchedRoIAlign(B, ld, [(C, H, W)], [8], DEV, layout=layout)
        h = fused.pinned_channels_last(torch.from_numpy(feat).to(dt))
        full = [h.to(DEV)]
        want = N(ra([full[0]], T(rb), T(counts))).copy()
        part = [torch.full_like(full[0], float("nan"))]
        moved = ra.fetch_touched(part, [h], T(rb), T(counts))
        got = N(ra(part, T(rb), T(counts)))
        for b in range(B):
            sl = slice(b * ld, b * ld + counts[b])
            assert np.array_equal(bits(got[sl]), bits(want[sl]))
        touched = ~torch.isnan(part[0].float()).any(1).cpu().numpy()
        assert not touched[1].any(), "image without RoIs: nothing fetched"
        assert int(moved[0]) == int(touched.sum()) and 0 < int(moved[0]) < B * H * W


def test_step_from_host_sparse_fetch_equals_full_copy():
    """TrainHotPath.step_from_host: channels_last pinned host feature maps (sparse fetch of the touched cells) and NCHW
    pinned host maps (full copy + transposition) give bit-identical results, RoI features included."""
    B, K = 2, 8
    w = workload.config2(B=B, K=K, channels=32)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_cls, h_reg = [pin(c) for c in w["cls"]], [pin(r) for r in w["reg"]]
    h_gt, h_gl = pin(w["gt"]), pin(w["gt_label"])
    h_nchw = [pin(f) for f in w["feats"]]
    h_nhwc = [fused.pinned_channels_last(torch.from_numpy(f)) for f in w["feats"]]
    gcount = torch.full((B,), K, dtype=torch.int32, device=DEV)
    img_hw = torch.tensor([[800.0, 1333.0]] * B, device=DEV)
    snaps = []
    for h_feats, zc_reg in ((h_nchw, False), (h_nhwc, True), (h_nhwc, False)):
        hp = fused.TrainHotPath(B, w["grids"], DEV, gt_ld=K, feat_channels=32, overlap=True)
        hp.reg_zero_copy = zc_reg                                         # regression maps read in place from the host / copied
        for it in range(2):                                               # second pass: stale cells of pass 1 in the device maps
            hp.reset_step()
            out = hp.step_from_host(h_cls, h_reg, h_feats, h_gt, h_gl, gcount, img_hw, with_roi_feats=True)
            torch.cuda.synchronize()
        snaps.append({k: v.clone() for k, v in out.items()})
        if h_feats is h_nhwc:
            assert hasattr(hp, "_stage_sparse") and not hasattr(hp, "_stage")
            cells = int(hp.roi_align.cells_moved[0]) / 2
            total = B * sum(g[0] * g[1] for g in w["grids"][:4])
            assert 0.3 * total < cells < 0.8 * total                      # config 2: ~0.6 of the pyramid
            assert hp.last_reg_zero_copy == zc_reg
    for other in snaps[1:]:
        assert set(snaps[0]) == set(other)
        for k in snaps[0]:
            assert torch.equal(snaps[0][k].view(torch.uint8), other[k].view(torch.uint8)), k


def test_roi_targets_as_tail_of_the_proposal_kernel_equal_the_separate_launch(monkeypatch):
    """b2d_rpn_proposals_targets: bbox_target riding on k_rpn_back (assignment spread over the image's cluster, sampler
    + encode in its first CTA) gives bit-identical labels / IoUs / samples / encoded targets to the separate
    b2d_roi_targets_fused launch, at config-2 sizes."""
    B, K = 4, 8
    w = workload.config2(B=B, K=K, channels=16)
    cls, reg = [T(c) for c in w["cls"]], [T(r) for r in w["reg"]]
    feats = [T(f).contiguous(memory_format=torch.channels_last) for f in w["feats"]]
    gt, gl = T(w["gt"]), T(w["gt_label"])
    gcount = torch.tensor([8, 5, 8, 1], dtype=torch.int32, device=DEV)           # ragged GT counts
    img_hw = torch.tensor([[800.0, 1333.0]] * B, device=DEV)
    snaps = []
    for fuse in ("1", "0"):  # riding on k_rpn_back, separate launch
        monkeypatch.setenv("B2D_FUSE_TARGETS", fuse)
        hp = fused.TrainHotPath(B, w["grids"], DEV, gt_ld=K, feat_channels=16, overlap=True)
        out = hp.step(cls, reg, feats, gt, gcount, gl, img_hw)
        torch.cuda.synchronize()
        bt = out["rcnn"]
        n = N(out["prop_count"])
        snap = [N(out["props"]), N(bt.n_chosen), N(bt.chosen), N(bt.tar_box), N(bt.tar_gt), N(bt.tar_param), N(bt.tar_label),
                N(bt.tar_is_gt), N(bt.census)[:, :2]]
        for b in range(B):
            kb = int(gcount[b])
            snap.append(N(bt.labels[b, :kb + n[b]]).copy()); snap.append(N(bt.iou[b, :kb + n[b]]).copy())
        snaps.append([s.copy() for s in snap])
    for a, b in zip(*snaps):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


# ------------------------------------------------------------------ support limits, AT the limits
def test_topk_at_the_16384_cap_and_one_above():
    rng = np.random.default_rng(5)
    v = rng.standard_normal(50000).astype(np.float32)
    idx = N(bregion.topk_desc(T(v), 16384))
    want = np.argsort(-v, kind="stable")[:16384]
    assert np.array_equal(idx, want)
    from b200det import _C
    with pytest.raises(_C.B200DetError):
        bregion.topk_desc(T(v), 16385)


def test_nms_at_the_16384_box_cap_and_one_above():
    """utils.nms: the kernel up to NMS_MAX_BOXES = 16384 boxes, torchvision above it -- both equal to the oracle."""
    rng = np.random.default_rng(6)
    for n in (butils.NMS_MAX_BOXES, butils.NMS_MAX_BOXES + 1):
        xy = rng.uniform(0, 900, (n, 2)).astype(np.float32)
        wh = rng.uniform(8, 120, (n, 2)).astype(np.float32)
        boxes = np.concatenate([xy, xy + wh], axis=1)
        scores = rng.permutation(n).astype(np.float32) / n
        keep = N(butils.nms(T(boxes), T(scores), 0.5))
        want = oracle.nms(boxes, scores, 0.5)
        assert np.array_equal(keep, want), n


@pytest.mark.parametrize("B", [256, 257])
@pytest.mark.parametrize("knobs", [{}, {"B2D_RPN_FRONT": "0", "B2D_RPN_BACK": "0"}])
def test_rpn_proposals_at_the_1024_segment_cap(B, knobs, setknob):
    """B * levels = 1024 is the largest segment table of the item-walking NMS kernels (kItemSegs); 1028 takes the other
    route.  Tiny 4-level pyramid, every image its own scores; images 0, B/2, B-1 against the oracle."""
    setknob(**knobs)
    rng = np.random.default_rng(B)
    grids, strides = [(12, 16), (6, 8), (3, 4), (2, 2)], (4, 8, 16, 32)
    pyr = fused.AnchorPyramid(strides, grids)
    cls = [rng.normal(0, 1, (B, 3) + g).astype(np.float32) for g in grids]
    reg = [rng.normal(0, 0.2, (B, 12) + g).astype(np.float32) for g in grids]
    cfg = dict(pre_nms=200, post_nms=100, max_num=150, nms_iou=0.7, min_bbox_size=0)
    rp = fused.RpnProposals(pyr, B, cfg, [0, 0, 0, 0], [1, 1, 1, 1], DEV)
    props, scores, count = rp([T(c) for c in cls], [T(r) for r in reg], torch.tensor([[48.0, 64.0]] * B, device=DEV))
    torch.cuda.synchronize()
    anc = [oracle.anchor_grid(s, gr, scales=[8]).reshape(4, -1) for s, gr in zip(strides, grids)]
    offs = np.cumsum([0] + [a.shape[1] for a in anc])
    for b in (0, B // 2, B - 1):
        _, _, lv, ix = oracle.rpn_proposals([c[b].reshape(-1) for c in cls], [r[b].reshape(4, -1) for r in reg], anc, cfg,
                                            [0, 0, 0, 0], [1, 1, 1, 1], (48, 64))
        n = int(count[b])
        assert n == lv.shape[0], (B, b)
        assert np.array_equal(N(rp.prov[b, :n]), offs[lv] + ix), (B, b)
