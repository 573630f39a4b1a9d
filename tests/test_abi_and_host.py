"""CPU-side checks: the C-ABI library loads and exports every symbol include/b200det.h
declares, host logic (pyramid description, threshold rounding, partitioning, the
gloo all-gather of detection records) and loud failure without a GPU."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import b200det
    from b200det import _C
    hdr = open(os.path.join(ROOT, "include", "b200det.h")).read()
    declared = sorted(set(re.findall(r"\b(b2d_[a-z0-9_]+)\(", hdr)))
    assert len(declared) >= 25
    if not os.path.exists(_C.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_C.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libb200det.so does not export %s" % name
    assert sorted(_C.EXPORTS) == declared, "ctypes binding table and header disagree"
    assert _C.lib().b2d_version() >= 100


def test_no_cpu_fallback():
    from b200det import utils, region, _C
    a = torch.zeros((4, 3))
    with pytest.raises(_C.B200DetError):
        utils.calc_iou(a, a)
    with pytest.raises(_C.B200DetError):
        region.MaxIoUAssigner(0.7, 0.3, 0.3)(a, a)
    with pytest.raises(_C.B200DetError):
        utils.nms(torch.zeros((3, 4)), torch.zeros(3), 0.5)


def test_argument_errors_are_reported_not_crashed():
    from b200det import _C
    rc = _C.lib().b2d_calc_iou(None, None, 0, None, 0, None)
    assert rc == -1 and b"calc_iou" in _C.lib().b2d_last_error_string()
    assert _C.lib().b2d_nms_workspace_bytes(2000, 5) > 5 * 2000 * 32 * 8


def test_threshold_rounding_matches_cpu_double_compare():
    from b200det import _C
    for thr in (0.3, 0.5, 0.7, 0.05, 0.0, 1.0):
        f = np.float32(_C.floor_f32(thr))
        assert float(f) <= thr < float(np.nextafter(f, np.float32(np.inf)))
        for x in (f, np.nextafter(f, np.float32(np.inf)), np.nextafter(f, np.float32(-np.inf))):
            assert (float(x) > thr) == (x > f)


def test_pyramid_description():
    from b200det import fused, workload
    grids = workload.fpn_grids()
    assert grids == [(200, 336), (100, 168), (50, 84), (25, 42), (13, 21)]
    pyr = fused.AnchorPyramid(workload.STRIDES, grids)
    assert pyr.total == 268569 and pyr.level_sizes == [201600, 50400, 12600, 3150, 819]
    assert [pyr.c.lv[i].offset for i in range(5)] == [0, 201600, 252000, 264600, 267750]
    import oracle
    ws, hs = oracle.anchor_sizes(4, [8], [0.5, 1.0, 2.0])
    assert np.array_equal(np.array(list(pyr.c.lv[0].ws)[:3], np.float32), ws)


def test_workload_is_deterministic():
    from b200det import workload
    a = workload.config2(B=1, K=3, channels=4, img_shape=(64, 80), pad_shape=(64, 96))
    b = workload.config2(B=1, K=3, channels=4, img_shape=(64, 80), pad_shape=(64, 96))
    assert all(np.array_equal(x, y) for x, y in zip(a["cls"], b["cls"])) and np.array_equal(a["gt"], b["gt"])


def test_sampler_spec_properties():
    from oracle import sampler_spec
    rng = np.random.default_rng(0)
    labels = rng.choice([-1, 0, 0, 1, 2], 5000).astype(np.int64)
    ch = sampler_spec.sample(labels, 256, 64, 12345)
    assert ch.size == 256 and (np.diff(ch) > 0).all()
    assert (labels[ch] > 0).sum() == 64 and (labels[ch] == 0).sum() == 192
    assert not np.array_equal(ch, sampler_spec.sample(labels, 256, 64, 12346))


def test_image_partition():
    from b200det import dist as bdist
    parts = [bdist.image_partition(8, r, 4) for r in range(4)]
    assert parts == [[0, 4], [1, 5], [2, 6], [3, 7]]
    assert sorted(sum((bdist.image_partition(10, r, 3) for r in range(3)), [])) == list(range(10))


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
import b200det
from b200det import dist as bdist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
n_img, M = 6, 5
mine = bdist.image_partition(n_img, rank, world)
recs, cnts = [], []
for g in mine:                       # deterministic per-image "detections"
    k = g %% (M + 1)
    boxes = torch.arange(k * 4, dtype=torch.float32).view(k, 4) + g
    r, c = bdist.pack_detections(boxes, torch.linspace(1, 0.5, k) if k else torch.zeros(0), torch.full((k,), g), M)
    recs.append(r); cnts.append(c)
rec, cnt = bdist.gather_detections(torch.stack(recs), torch.cat(cnts))
assert rec.shape == (n_img, M, 6) and cnt.tolist() == [g %% (M + 1) for g in range(n_img)], cnt
for g in range(n_img):
    k = g %% (M + 1)
    assert torch.equal(rec[g, :k, 5], torch.full((k,), float(g)))
    if k: assert rec[g, 0, 0].item() == float(g)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_gloo_world2_partition_and_gather(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER % dict(root=ROOT))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29581")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)


def test_dropin_install_rebinds_reference_names():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_shim
    if not ref_shim.available():
        pytest.skip("reference checkout not present on this machine")
    lib = ref_shim.install()
    import b200det
    import lib.builder as rb
    import lib.heads.anchor_head as ah
    import lib.heads.rpn_head as rh
    import lib.heads.bbox_head as bh
    import lib.heads.guided_head as gh                                      # (mmdet's DeformConv stubbed by ref_shim)
    ref_ga_pred = gh.GARPNHead.predict_bboxes_single_image
    orig = rb.MODULES["MaxIoUAssigner"]
    ref_loss, ref_tars, ref_pred = ah.AnchorHead.loss, bh.BBoxHead.bbox_targets, rh.RPNHead.predict_bboxes_from_output
    b200det.install(lib)
    try:
        # the per-image loops themselves (batched.py)
        assert ah.AnchorHead.loss is not ref_loss and bh.BBoxHead.bbox_targets is not ref_tars
        assert bh.BBoxHead.refine_bboxes.__name__ == "_refine"
        assert rh.RPNHead.predict_bboxes_from_output is not ref_pred
        assert ah.AnchorHead.predict_bboxes_from_output is not ref_pred      # dense heads: batched selection
        assert rb.MODULES["MaxIoUAssigner"] is b200det.region.MaxIoUAssigner
        assert rb.MODULES["RoIAlign"] is b200det.region.RoIAlign
        assert ah.anchor_target is b200det.anchor.anchor_target
        assert ah.AnchorCreator is b200det.anchor.AnchorCreator
        assert rh.tvops.nms is b200det.utils.nms
        assert rh.RPNHead.predict_single_image is b200det.heads.rpn_predict_single_image
        assert sys.modules["lib.utils"].calc_iou.b2d_fast is b200det.utils.calc_iou         # grad-safe wrapper
        assert ah.AnchorHead.predict_single_image is b200det.heads.anchor_head_predict_single_image
        assert sys.modules["lib.utils"].tv.ops.nms is b200det.utils.nms
        assert sys.modules["lib.bbox"].bbox_target is b200det.bbox.bbox_target
        assert gh.GARPNHead.predict_bboxes_single_image is b200det.heads.ga_rpn_predict_single_image
        assert gh.GARPNHead.rpn_target_single_image is b200det.heads.ga_rpn_target_single_image
        assert gh.tvops.nms is b200det.utils.nms
        built = rb.build_module(dict(type="MaxIoUAssigner", pos_iou=0.7, neg_iou=0.3, min_pos_iou=0.3))
        assert isinstance(built, b200det.region.MaxIoUAssigner)
    finally:
        b200det.uninstall()
    assert rb.MODULES["MaxIoUAssigner"] is orig
    assert ah.AnchorHead.loss is ref_loss and bh.BBoxHead.bbox_targets is ref_tars
    assert rh.RPNHead.predict_single_image is not b200det.heads.rpn_predict_single_image
    assert gh.GARPNHead.predict_bboxes_single_image is ref_ga_pred


def test_dropin_keeps_gradients_and_cpu_tensors_on_the_reference_path():
    """install() must not cut autograd.  IoULoss is -utils.elem_iou(a, b).log() (lib/losses.py:7-10); after
    install() its gradient is still there, and CPU tensors still work (both routed to the saved reference code)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_shim
    if not ref_shim.available():
        pytest.skip("reference checkout not present on this machine")
    import torch
    lib = ref_shim.install()
    import b200det
    import lib.losses as rl
    import lib.utils as ru
    b200det.install(lib)
    try:
        a = torch.tensor([[10.0, 20.0], [10.0, 30.0], [50.0, 80.0], [60.0, 90.0]], requires_grad=True)
        b = torch.tensor([[12.0, 25.0], [11.0, 28.0], [55.0, 70.0], [58.0, 95.0]])
        loss = rl.iou_loss(a, b).sum() if hasattr(rl, "iou_loss") else (-ru.elem_iou(a, b).log()).sum()
        loss.backward()
        assert a.grad is not None and float(a.grad.abs().sum()) > 0.0
        with torch.no_grad():
            p = ru.param2bbox(b, torch.zeros_like(b))            # CPU tensors: reference implementation, no raise
        assert tuple(p.shape) == (4, 2)
        ce = rl.CrossEntropyLoss(use_sigmoid=False)
        x = torch.randn(5, 3, requires_grad=True)
        ce(x, torch.tensor([0, 1, 2, 1, 0])).backward()           # CPU: the reference's forward
        assert x.grad is not None
    finally:
        b200det.uninstall()


def test_install_channels_last_makes_the_reference_fpn_emit_nhwc():
    """install(channels_last=True) -> lib.necks.FPN outputs are channels_last and numerically the same."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_shim
    if not ref_shim.available():
        pytest.skip("reference checkout not present on this machine")
    import torch
    lib = ref_shim.install()
    import b200det
    import lib.necks as rn
    torch.manual_seed(0)
    fpn = rn.FPN(in_channels=[8, 16, 32, 64], out_channels=8, num_outs=5)
    feats = [torch.randn(2, c, 64 // s, 96 // s) for c, s in zip([8, 16, 32, 64], [1, 2, 4, 8])]
    with torch.no_grad():
        ref = fpn(feats)
    b200det.install(lib, channels_last=True)
    try:
        with torch.no_grad():
            out = fpn(feats)
        for a, b in zip(ref, out):
            assert b.is_contiguous(memory_format=torch.channels_last) or b.shape[1] == 1 or b.shape[2] * b.shape[3] == 1
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)
    finally:
        b200det.uninstall()
