"""install(): rebind the reference's hot-path names to the B200 kernels without editing
a single reference file: module attributes of lib.anchor / lib.bbox /
lib.region / lib.utils, the names the head modules imported with `from .. import x`,
`torchvision.ops.nms` as seen by lib.heads.rpn_head / lib.region / lib.utils, and the
lib.builder.MODULES registry entries."""
import sys
import types

_saved = []


def _grad_safe(fast, reference):
    """The kernels have no autograd graph and no CPU path.  A rebinding of a reference function that the reference
    also uses on differentiable tensors (IoULoss -> utils.elem_iou, lib/losses.py:7-10; GuidedAnchor ->
    utils.param2bbox, lib/heads/guided_head.py:129) or on CPU tensors therefore keeps the reference implementation
    for exactly those calls: any tensor argument that requires grad while grad mode is on, or that is not on a CUDA
    device, routes the call to the saved original."""
    import functools

    import torch

    @functools.wraps(reference)
    def call(*args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and ((a.requires_grad and torch.is_grad_enabled()) or not a.is_cuda):
                return reference(*args, **kwargs)
        return fast(*args, **kwargs)

    call.b2d_fast, call.b2d_reference = fast, reference
    return call


def _set(obj, name, value):
    if hasattr(obj, name) or isinstance(obj, dict):
        old = obj[name] if isinstance(obj, dict) else getattr(obj, name)
        _saved.append((obj, name, old))
        if isinstance(obj, dict):
            obj[name] = value
        else:
            setattr(obj, name, value)


def _channels_last_fpn(necks):
    """make the reference FPN (lib/necks.py:7-90) EMIT channels_last feature maps.  Its convolutions
    are switched to channels_last weights and fed channels_last inputs, so cuDNN runs NHWC kernels and the outputs
    arrive in the layout the RoIAlign kernels read (one contiguous channel vector per cell) -- no transposition
    kernel anywhere; lateral adds, nearest upsampling and max pooling preserve the memory format."""
    import torch
    ref_forward = necks.FPN.forward

    def forward(self, feats):
        if not getattr(self, "_b2d_channels_last", False):
            self.to(memory_format=torch.channels_last)
            self._b2d_channels_last = True
        return ref_forward(self, [f.contiguous(memory_format=torch.channels_last) for f in feats])

    _set(necks.FPN, "forward", forward)


def install(lib=None, channels_last=False):
    """lib: the reference's imported `lib` package (default: sys.modules['lib']).  channels_last=True additionally
    makes the FPN neck produce channels_last (NHWC) feature maps, the layout K5 / K6 read without a transposition."""
    from . import anchor, bbox, heads, region, utils
    lib = lib or sys.modules.get("lib")
    if lib is None:
        raise RuntimeError("import the reference package `lib` before calling install()")
    mods = {n: sys.modules.get("lib." + n) for n in
            ("anchor", "bbox", "region", "utils", "builder", "heads.anchor_head", "heads.rpn_head",
             "heads.bbox_head", "heads.fcos_head", "heads.retina_head", "heads.rcnn_head", "heads.guided_head",
             "detectors.cascade_rcnn")}
    util_names = ["calc_iou", "elem_iou", "bbox2param", "param2bbox", "batched_param2bbox", "clamp_bbox",
                  "batched_nms", "multiclass_nms"]
    if mods["utils"]:
        for n in util_names:
            if hasattr(mods["utils"], n):
                _set(mods["utils"], n, _grad_safe(getattr(utils, n), getattr(mods["utils"], n)))
        # lib.utils calls tv.ops.nms: give it a module-like shim whose .ops.nms is ours
        tvshim = types.SimpleNamespace(ops=types.SimpleNamespace(nms=utils.nms), transforms=mods["utils"].tv.transforms)
        _set(mods["utils"], "tv", tvshim)
    region_names = ["inside_grid_mask", "inside_anchor_mask", "MaxIoUAssigner", "RandomSampler",
                    "IoUBalancedNegSampler", "ProposalCreator", "ScalableRoIPool", "ScalableRoIAlign",
                    "BasicRoIExtractor", "SingleRoIExtractor"]
    if mods["region"]:
        for n in region_names:
            _set(mods["region"], n, getattr(region, n))
    if mods["anchor"]:
        _set(mods["anchor"], "AnchorCreator", anchor.AnchorCreator)
        _set(mods["anchor"], "anchor_target", anchor.anchor_target)
    if mods["bbox"]:
        _set(mods["bbox"], "bbox_target", bbox.bbox_target)
    # names bound with `from .. import x` inside the heads
    ah = mods["heads.anchor_head"]
    if ah:
        if hasattr(ah, "AnchorHead"):                    # RetinaNet test path (BASELINE config 4): per-level top-k + multiclass NMS
            _set(ah.AnchorHead, "predict_single_image", heads.anchor_head_predict_single_image)
        _set(ah, "AnchorCreator", anchor.AnchorCreator)
        _set(ah, "anchor_target", anchor.anchor_target)
        _set(ah, "inside_grid_mask", region.inside_grid_mask)
        _set(ah, "inside_anchor_mask", region.inside_anchor_mask)
    if mods["heads.bbox_head"]:
        _set(mods["heads.bbox_head"], "bbox_target", bbox.bbox_target)
    if mods["heads.rpn_head"]:
        _set(mods["heads.rpn_head"], "tvops", types.SimpleNamespace(nms=utils.nms))
        if hasattr(mods["heads.rpn_head"], "RPNHead"):   # a9: the whole per-level loop as one fused K3 + K4 call
            _set(mods["heads.rpn_head"].RPNHead, "predict_single_image", heads.rpn_predict_single_image)
    gh = mods["heads.guided_head"]                      # GA-RPN (imported only where mmdet's DeformConv exists)
    if gh:
        _set(gh, "tvops", types.SimpleNamespace(nms=utils.nms))
        if hasattr(gh, "anchor_target"):
            _set(gh, "anchor_target", anchor.anchor_target)
        if hasattr(gh, "GARPNHead"):                     # explicit guided anchors + location masks
            _set(gh.GARPNHead, "predict_bboxes_single_image", heads.ga_rpn_predict_single_image)
            _set(gh.GARPNHead, "rpn_target_single_image", heads.ga_rpn_target_single_image)
    if mods["heads.fcos_head"] and hasattr(mods["heads.fcos_head"], "AnchorCreator"):
        _set(mods["heads.fcos_head"], "AnchorCreator", anchor.AnchorCreator)
    # head methods that are part of the path: rebind on the classes
    bh, fh = mods["heads.bbox_head"], mods["heads.fcos_head"]
    if bh and hasattr(bh, "BBoxHead"):
        _set(bh.BBoxHead, "refine_bboxes_single_image", heads.refine_bboxes_single_image)
        _set(bh.BBoxHead, "predict_bboxes_single_image", heads.predict_bboxes_single_image)
    if fh and hasattr(fh, "FCOSHead"):
        _set(fh.FCOSHead, "single_image_targets_atss", heads.single_image_targets_atss)
        _set(fh.FCOSHead, "single_image_targets", heads.single_image_targets)
        ref_predict = fh.FCOSHead.predict_single_image

        def _predict(self, cls_outs, reg_outs, ctr_outs, img_meta, test_cfg):
            if getattr(self, "use_dfl", False):          # DFL decode stays with the reference
                return ref_predict(self, cls_outs, reg_outs, ctr_outs, img_meta, test_cfg)
            return heads.predict_single_image(self, cls_outs, reg_outs, ctr_outs, img_meta, test_cfg)

        _set(fh.FCOSHead, "predict_single_image", _predict)
    # the per-image LOOPS themselves: one batched pass when the call is covered, else the reference's loop
    from . import batched
    rh = mods["heads.rpn_head"]
    if rh and hasattr(rh, "RPNHead"):
        ref_predict_all = rh.RPNHead.predict_bboxes_from_output

        def _predict_all(self, cls_outs, reg_outs, img_metas, test_cfg):
            out = batched.rpn_predict_fast(self, cls_outs, reg_outs, img_metas, test_cfg)
            return out if out is not None else ref_predict_all(self, cls_outs, reg_outs, img_metas, test_cfg)

        _set(rh.RPNHead, "predict_bboxes_from_output", _predict_all)
    if ah and hasattr(ah, "AnchorHead"):
        ref_predict_dense = ah.AnchorHead.predict_bboxes_from_output

        def _predict_dense(self, cls_outs, reg_outs, img_metas, test_cfg):
            # RPNHead has its own binding above; GA-RPN / other subclasses with their own predict_single_image keep the loop
            out = None
            if type(self).predict_single_image is heads.anchor_head_predict_single_image:
                out = batched.anchor_head_predict_fast(self, cls_outs, reg_outs, img_metas, test_cfg)
            return out if out is not None else ref_predict_dense(self, cls_outs, reg_outs, img_metas, test_cfg)

        _set(ah.AnchorHead, "predict_bboxes_from_output", _predict_dense)
        ref_loss = ah.AnchorHead.loss

        def _loss(self, cls_outs, reg_outs, gt_bboxes, gt_labels, img_metas, train_cfg):
            return batched.anchor_head_loss(self, cls_outs, reg_outs, gt_bboxes, gt_labels, img_metas, train_cfg,
                                            reference_loss=ref_loss)

        _set(ah.AnchorHead, "loss", _loss)
    if bh and hasattr(bh, "BBoxHead"):
        ref_bbox_targets = bh.BBoxHead.bbox_targets

        def _bbox_targets(self, img_props, gt_bboxes, gt_labels, train_cfg):
            out = batched.bbox_targets_fast(self, img_props, gt_bboxes, gt_labels, train_cfg)
            return out if out is not None else ref_bbox_targets(self, img_props, gt_bboxes, gt_labels, train_cfg)

        _set(bh.BBoxHead, "bbox_targets", _bbox_targets)
        ref_refine = bh.BBoxHead.refine_bboxes

        def _refine(self, props, labels, reg_outs, is_gts=None, img_metas=None):
            out = batched.refine_bboxes_fast(self, props, labels, reg_outs, is_gts, img_metas)
            return out if out is not None else ref_refine(self, props, labels, reg_outs, is_gts, img_metas)

        _set(bh.BBoxHead, "refine_bboxes", _refine)
    necks = sys.modules.get("lib.necks")
    if channels_last and necks is not None and hasattr(necks, "FPN"):
        _channels_last_fpn(necks)
    losses = sys.modules.get("lib.losses")
    if losses is not None and hasattr(losses, "CrossEntropyLoss"):      # CE / BCE on the sampled rows
        ref_ce = losses.CrossEntropyLoss.forward

        def _ce_forward(self, pred, label):
            if not pred.is_cuda:
                return ref_ce(self, pred, label)
            return heads.cross_entropy_loss_forward(self, pred, label)

        _set(losses.CrossEntropyLoss, "forward", _ce_forward)
    if mods["builder"]:
        reg = mods["builder"].MODULES
        for n in ("MaxIoUAssigner", "RandomSampler", "IoUBalancedNegSampler", "BasicRoIExtractor",
                  "SingleRoIExtractor", "RoIAlign", "RoIPool", "ScalableRoIPool", "ScalableRoIAlign"):
            _set(reg, n, getattr(region, n))
    return lib


def uninstall():
    while _saved:
        obj, name, old = _saved.pop()
        if isinstance(obj, dict):
            obj[name] = old
        else:
            setattr(obj, name, old)
