"""Batch-level forms of the reference's per-image Python loops, behind the reference's own METHOD signatures, so that
dropin.install() can rebind the loops themselves and not only the functions they call:

  rpn_predict_bboxes_from_output   AnchorHead.predict_bboxes_from_output  lib/heads/anchor_head.py:268-289  (on RPNHead)
  bbox_head_bbox_targets           BBoxHead.bbox_targets                  lib/heads/bbox_head.py:47-52
  anchor_head_targets / _loss      AnchorHead.loss (the target part)      lib/heads/anchor_head.py:152-199
  bbox_head_refine_bboxes          BBoxHead.refine_bboxes                 lib/heads/bbox_head.py:94-96   (cascade stage loop)
  anchor_head_predict_fast         AnchorHead.predict_bboxes_from_output  lib/heads/anchor_head.py:268-289  (dense heads, test path)

Each one packs its per-image arguments into the image-major batch layout of fused.py, runs the batched kernels once
(one host synchronisation per call, for the ragged result sizes) and hands back exactly the per-image lists the
reference method returns.  A call the batched kernels do not cover (other assigner / sampler types, host-RNG sampler,
per-image min sizes, CPU tensors) returns None from the `*_fast` function and the caller keeps the reference's loop,
which still runs on the per-function drop-ins."""
import numpy as np
import torch
import torch.nn.functional as F

from . import _C, fused, region

_KEEP = 8            # cached stage objects per head (distinct batch shapes)


def _get(cfg, key, default=None):
    if cfg is None:
        return default
    if hasattr(cfg, "get"):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def _spec(obj, cls, fields):
    """Assigner / sampler given as a built module of OUR class or as its config dict -> dict of its fields, else None."""
    if isinstance(obj, cls):
        return {f: getattr(obj, f) for f in fields if hasattr(obj, f)}
    if obj is not None and not isinstance(obj, (torch.nn.Module,)) and hasattr(obj, "get") and obj.get("type") == cls.__name__:
        return {f: obj.get(f) for f in fields if obj.get(f) is not None}
    return None


def _assigner_spec(a):
    s = _spec(a, region.MaxIoUAssigner, ("pos_iou", "neg_iou", "min_pos_iou"))
    return s if s is not None and len(s) == 3 else None


def _sampler_spec(s):
    d = _spec(s, region.RandomSampler, ("max_num", "pos_num", "rng", "seed"))
    if d is None or "max_num" not in d or "pos_num" not in d:
        return None
    if (d.get("rng") or region.RandomSampler.default_rng) != "device":
        return None                                      # the host-RNG parity mode is per image by construction
    if d.get("seed") is None:
        d["seed"] = region._auto["seed"]                 # the cached stage object keeps its own device step counter
    return d


def _cache(owner, name):
    c = owner.__dict__.setdefault(name, {})
    while len(c) > _KEEP:
        c.pop(next(iter(c)))
    return c


def _all_cuda_f32(*lists):
    for ts in lists:
        for t in ts:
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32):
                return False
    return True


def _closed_form(head, n_levels):
    """The head's anchors as (strides, scales, ratios, center_lt) when they are the closed form the kernels evaluate
    in registers (lib/heads/anchor_head.py:29-36: one AnchorCreator per level with base == stride and shared
    scales / ratios), else None."""
    acs = getattr(head, "anchor_creators", None)
    strides = list(getattr(head, "anchor_strides", []))
    if not acs or len(acs) != n_levels or len(strides) != n_levels or n_levels > _C.MAX_LEVELS:
        return None
    s0, r0, c0 = list(acs[0].scales), list(acs[0].aspect_ratios), bool(getattr(acs[0], "center_lt", False))
    for ac, st in zip(acs, strides):
        if ac.base != st or list(ac.scales) != s0 or list(ac.aspect_ratios) != r0 or bool(getattr(ac, "center_lt", False)) != c0:
            return None
    return strides, tuple(s0), tuple(r0), c0


_hw_cache = {}


def _upload_i32(values, dev):
    # pageable -> device with non_blocking: the driver stages the bytes before returning and the stream is not
    # synchronised (torch.tensor(..., device=cuda) would synchronise it)
    return torch.tensor(values, dtype=torch.int32).to(dev, non_blocking=True)


def _img_hw(img_metas, dev):
    key = (tuple((float(m['img_shape'][0]), float(m['img_shape'][1])) for m in img_metas), str(dev))
    t = _hw_cache.get(key)
    if t is None:
        if len(_hw_cache) > 64:
            _hw_cache.clear()
        t = _hw_cache[key] = torch.tensor([list(hw) for hw in key[0]], dtype=torch.float32, device=dev)
    return t


def _pack_labels(gt_labels, K):
    if all(int(l.numel()) == K for l in gt_labels):
        return torch.stack([l.view(-1) for l in gt_labels]).to(torch.int64)
    return torch.stack([F.pad(l.to(torch.int64).view(-1), (0, K - int(l.numel()))) for l in gt_labels]).contiguous()


def _pack_cols(cols, ld=None, pad16=False):
    """list of [4, n_i] -> ([B, 4, ld] fp32 zero-padded, host counts)."""
    ns = [int(c.shape[1]) for c in cols]
    ld = ld or max(max(ns), 1)
    if all(n == ld for n in ns):                          # equal sizes: one stack, no padding
        return torch.stack([c.to(torch.float32) for c in cols]), ns
    if pad16:
        ld = (ld + 15) // 16 * 16
    return torch.stack([F.pad(c.to(torch.float32), (0, ld - n)) for c, n in zip(cols, ns)]).contiguous(), ns


# ------------------------------------------------------------------------------------------------ proposals
def rpn_predict_fast(self, cls_outs, reg_outs, img_metas, test_cfg):
    B = len(img_metas)
    cf = _closed_form(self, len(cls_outs))
    if cf is None or B == 0 or not _all_cuda_f32(cls_outs, reg_outs) or int(cls_outs[0].shape[0]) != B:
        return None
    sfs = [float(m.get('scale_factor', 1.0)) for m in img_metas]
    min_bbox = float(_get(test_cfg, 'min_bbox_size', 0) or 0)
    if min_bbox > 0 and any(s != sfs[0] for s in sfs):
        return None                                      # the size filter is per image (scale_factor * min_bbox_size)
    strides, scales, ratios, center_lt = cf
    dev = cls_outs[0].device
    grids = tuple(tuple(int(v) for v in c.shape[-2:]) for c in cls_outs)
    key = (B, grids, int(_get(test_cfg, 'pre_nms', 0)), int(_get(test_cfg, 'post_nms', 0)), int(_get(test_cfg, 'max_num', 0)),
           float(_get(test_cfg, 'nms_iou', 0.7)), min_bbox, sfs[0], str(dev), bool(self.use_sigmoid))
    cache = _cache(self, '_b2d_rpn_batch_cache')
    rp = cache.get(key)
    if rp is None:
        pyr = fused.AnchorPyramid(strides, grids, scales, ratios, center_lt)
        try:
            rp = fused.RpnProposals(pyr, B, test_cfg, self.target_means, self.target_stds, dev,
                                    score_mode=0 if self.use_sigmoid else 1, cls_channels=int(self.cls_channels),
                                    scale_factor=sfs[0])
        except _C.B200DetError:
            return None
        cache[key] = rp
    cls, reg = [c.contiguous() for c in cls_outs], [r.contiguous() for r in reg_outs]
    with torch.no_grad():
        props, scores, count = rp(cls, reg, _img_hw(img_metas, dev))
        props, scores, count = props.clone(), scores.clone(), count.clone()     # the stage object's buffers are reused by the next call
        ns = count.tolist()                               # the one synchronisation of this call
    bboxes, scs = [], []
    for i, n in enumerate(ns):
        b = props[i][:, :n]
        b._b2d_batch = (props, count, i, B)               # lets bbox_head_bbox_targets skip the re-packing
        bboxes.append(b)
        scs.append(scores[i][:n])
    return [bboxes, scs, [None] * B]


def rpn_predict_bboxes_from_output(self, cls_outs, reg_outs, img_metas, test_cfg):
    """Method form of AnchorHead.predict_bboxes_from_output for an RPN head: the per-image loop over
    predict_single_image (lib/heads/anchor_head.py:281-288) as one batched K3 + K4 call.  Returns what
    utils.unpack_multi_result(preds) returns: [[bbox [4,k_i]]*B, [score [k_i]]*B, [None]*B]."""
    out = rpn_predict_fast(self, cls_outs, reg_outs, img_metas, test_cfg)
    if out is not None:
        return out
    return _reference_loop_predict(self, cls_outs, reg_outs, img_metas, test_cfg)


def _reference_loop_predict(self, cls_outs, reg_outs, img_metas, test_cfg):
    grid_sizes = [c.shape[-2:] for c in cls_outs]
    level_anchors = [ac(self.anchor_strides[i], grid_sizes[i]) for i, ac in enumerate(self.anchor_creators)]   # create_anchors
    preds = []
    for i, img_meta in enumerate(img_metas):
        preds.append(list(self.predict_single_image([c[i] for c in cls_outs], [r[i] for r in reg_outs], level_anchors,
                                                    img_meta, test_cfg)))
    return [[p[k] for p in preds] for k in range(3)]


# ------------------------------------------------------------------------------------------------ RoI targets
def bbox_targets_fast(self, img_props, gt_bboxes, gt_labels, train_cfg):
    B = len(img_props)
    a, s = _assigner_spec(_get(train_cfg, 'assigner')), _sampler_spec(_get(train_cfg, 'sampler'))
    if a is None or s is None or B == 0 or len(gt_bboxes) != B or not _all_cuda_f32(img_props):
        return None
    if not all(g.is_cuda for g in gt_bboxes) or any(int(g.shape[1]) == 0 for g in gt_bboxes):
        return None
    dev = img_props[0].device
    src = getattr(img_props[0], '_b2d_batch', None)
    if src is not None and src[3] == B and all(getattr(p, '_b2d_batch', (None,))[0] is src[0] and p._b2d_batch[2] == i
                                               for i, p in enumerate(img_props)):
        props, count = src[0], src[1]                     # still packed: views of one [B, 4, P] buffer
    else:
        props, ns = _pack_cols(img_props, pad16=True)
        count = _upload_i32(ns, dev)
    gt, ks = _pack_cols(gt_bboxes, pad16=True)
    K = int(gt.shape[2])
    gt_count = _upload_i32(ks, dev)
    gl = _pack_labels(gt_labels, K)
    N = int(props.shape[2])
    key = (B, N, K, a['pos_iou'], a['neg_iou'], a['min_pos_iou'], s['max_num'], s['pos_num'], s['seed'],
           tuple(self.target_means), tuple(self.target_stds), str(dev))
    cache = _cache(self, '_b2d_roi_target_cache')
    bt = cache.get(key)
    if bt is None:
        bt = fused.BatchedTargets(B, N, K, a, s, self.target_means, self.target_stds, dev, prepend_gt=True, seed=int(s['seed']))
        cache[key] = bt
    with torch.no_grad():
        bt(gt, gt_count, gl, boxes=props, box_count=count)
        outs = [t.clone() for t in (bt.tar_box, bt.tar_gt, bt.tar_label, bt.tar_param, bt.tar_is_gt)]
        ns = bt.n_chosen.tolist()                         # the one synchronisation of this call
    res = [[], [], [], [], []]
    for i, n in enumerate(ns):
        for k, t in enumerate(outs):
            res[k].append(t[i][..., :n])
    return res


def bbox_head_bbox_targets(self, img_props, gt_bboxes, gt_labels, train_cfg):
    """Method form of BBoxHead.bbox_targets (multi_apply(bbox_target, ...) over the images, lib/heads/bbox_head.py:47-52)
    as one b2d_roi_targets_fused launch.  Returns [tar_props, tar_bbox, tar_label, tar_param, tar_is_gt], each a list
    over the images."""
    out = bbox_targets_fast(self, img_props, gt_bboxes, gt_labels, train_cfg)
    if out is not None:
        return out
    from . import bbox
    res = [bbox.bbox_target(p, g, l, _get(train_cfg, 'assigner'), _get(train_cfg, 'sampler'), tuple(self.target_means),
                            tuple(self.target_stds)) for p, g, l in zip(img_props, gt_bboxes, gt_labels)]
    return [[r[k] for r in res] for k in range(5)]


# ------------------------------------------------------------------------------------------------ anchor targets
def anchor_head_targets(self, cls_outs, reg_outs, gt_bboxes, gt_labels, img_metas, train_cfg):
    """The target part of AnchorHead.loss for a head that samples (the RPN): inside masks, assignment, sampling, delta
    encoding and the gathers of the head outputs at the sampled anchors for ALL images in one batched pass.  Returns
    the four concatenated tensors the reference hands to calc_loss (lib/heads/anchor_head.py:187-199): tar_cls_out
    [C, n], tar_reg_out [4, n], tar_label [n], tar_param [4, n] -- the two gathers are differentiable w.r.t. the head
    outputs --, or None when the call is not covered."""
    B = len(img_metas)
    a, s = _assigner_spec(_get(train_cfg, 'assigner')), _sampler_spec(_get(train_cfg, 'sampler'))
    cf = _closed_form(self, len(cls_outs))
    if a is None or s is None or cf is None or B == 0 or not _all_cuda_f32(cls_outs, reg_outs):
        return None
    if int(cls_outs[0].shape[0]) != B or not all(g.is_cuda and int(g.shape[1]) > 0 for g in gt_bboxes):
        return None
    strides, scales, ratios, center_lt = cf
    dev = cls_outs[0].device
    C = int(self.cls_channels)
    grids = tuple(tuple(int(v) for v in c.shape[-2:]) for c in cls_outs)
    gt, ks = _pack_cols(gt_bboxes, pad16=True)
    K = int(gt.shape[2])
    border = float(_get(train_cfg, 'allowed_border', 0) or 0)
    key = (B, grids, K, a['pos_iou'], a['neg_iou'], a['min_pos_iou'], s['max_num'], s['pos_num'], s['seed'], border,
           tuple(self.target_means), tuple(self.target_stds), str(dev))
    cache = _cache(self, '_b2d_anchor_target_cache')
    ent = cache.get(key)
    if ent is None:
        pyr = fused.AnchorPyramid(strides, grids, scales, ratios, center_lt)
        bt = fused.BatchedTargets(B, pyr.total, K, a, s, self.target_means, self.target_stds, dev, pyramid=pyr,
                                  border=border, seed=int(s['seed']))
        ent = cache[key] = (pyr, bt)
    pyr, bt = ent
    gt_count = _upload_i32(ks, dev)
    gl = None
    if gt_labels is not None:
        gl = _pack_labels(gt_labels, K)
    with torch.no_grad():
        bt(gt, gt_count, gl, img_hw=_img_hw(img_metas, dev))
        valid = torch.arange(bt.max_num, device=dev).view(1, -1) < bt.n_chosen.view(-1, 1)
        bidx, jidx = torch.nonzero(valid, as_tuple=True)  # image-major, ascending inside an image (synchronises once)
        aidx = bt.chosen.to(torch.int64)[bidx, jidx]      # anchor index in the level-major concatenation
        tar_label = bt.tar_label[bidx, jidx]
        tar_param = bt.tar_param[bidx, :, jidx].t()
    cls_flat = torch.cat([c.reshape(B, C, -1) for c in cls_outs], dim=2)      # [B, C, total], as single_image_targets :83-88
    reg_flat = torch.cat([r.reshape(B, 4, -1) for r in reg_outs], dim=2)
    tar_cls_out = cls_flat[bidx, :, aidx].t()
    tar_reg_out = reg_flat[bidx, :, aidx].t()
    return tar_cls_out, tar_reg_out, tar_label, tar_param


def anchor_head_loss(self, cls_outs, reg_outs, gt_bboxes, gt_labels, img_metas, train_cfg, reference_loss=None):
    """Method form of AnchorHead.loss: batched targets (above), then the head's own calc_loss."""
    tars = anchor_head_targets(self, cls_outs, reg_outs, gt_bboxes, gt_labels, img_metas, train_cfg)
    if tars is None:
        if reference_loss is None:
            raise _C.B200DetError("anchor_head_loss: call not covered by the batched path and no reference loop given")
        return reference_loss(self, cls_outs, reg_outs, gt_bboxes, gt_labels, img_metas, train_cfg)
    return self.calc_loss(*tars, train_cfg)


# ------------------------------------------------------------------------------------------------ cascade refinement
def refine_bboxes_fast(self, props, labels, reg_outs, is_gts=None, img_metas=None):
    """BBoxHead.refine_bboxes (multi_apply over refine_bboxes_single_image, lib/heads/bbox_head.py:94-120) for all
    images in one b2d_refine_bboxes launch: per image the non-GT columns, decoded with the stage's means / stds and
    clamped to img_shape.  Returns the list of [4, s_i - #gt_i] tensors, or None when the call is not covered."""
    B = len(props)
    if B == 0 or not _all_cuda_f32(props, reg_outs) or len(labels) != B or len(reg_outs) != B:
        return None
    if img_metas is not None and (not isinstance(img_metas, list) or len(img_metas) != B):
        return None
    if is_gts is not None and (not isinstance(is_gts, list) or len(is_gts) != B):
        return None
    dev = props[0].device
    C = 1 if self.reg_class_agnostic else int(self.num_classes)
    if any(int(r.shape[1]) != 4 * C or int(r.shape[0]) != int(p.shape[1]) for r, p in zip(reg_outs, props)):
        return None
    pk, ns = _pack_cols(props)
    ld = int(pk.shape[2])
    pad = lambda t, n: t if int(t.shape[0]) == ld else F.pad(t, (0, 0) * (t.dim() - 1) + (0, ld - n))
    lab = torch.stack([pad(l.to(torch.int64).view(-1), n) for l, n in zip(labels, ns)]).contiguous()
    reg = torch.stack([pad(r, n) for r, n in zip(reg_outs, ns)]).contiguous()
    gtf = None
    if is_gts is not None:
        gtf = torch.stack([pad(g.to(torch.int64).view(-1), n) for g, n in zip(is_gts, ns)]).contiguous()
    counts = _upload_i32(ns, dev)
    out = torch.empty((B, 4, ld), dtype=torch.float32, device=dev)
    out_count = torch.empty(B, dtype=torch.int32, device=dev)
    clamp = img_metas is not None
    hw = _img_hw(img_metas, dev) if clamp else None
    with torch.no_grad():
        _C.call("b2d_refine_bboxes", _C.ptr(out), _C.ptr(out_count), _C.ptr(pk), ld, _C.ptr(counts), ld, _C.ptr(lab),
                _C.ptr(reg), C, _C.ptr(gtf), _C.host_f4(self.target_means, [0, 0, 0, 0]),
                _C.host_f4(self.target_stds, [1, 1, 1, 1]), int(clamp), _C.ptr(hw), B, _C.stream())
        no = out_count.tolist() if gtf is not None else ns   # the one synchronisation (only when GT columns are dropped)
    res = []
    for i, n in enumerate(no):
        r = out[i][:, :n]
        r._b2d_batch = (out, out_count if gtf is not None else counts, i, B)
        res.append(r)
    return res


def bbox_head_refine_bboxes(self, props, labels, reg_outs, is_gts=None, img_metas=None):
    """Method form of BBoxHead.refine_bboxes."""
    out = refine_bboxes_fast(self, props, labels, reg_outs, is_gts, img_metas)
    if out is not None:
        return out
    from . import heads
    nb = len(props)
    is_gts = is_gts if isinstance(is_gts, list) else [is_gts] * nb
    img_metas = img_metas if isinstance(img_metas, list) else [img_metas] * nb
    return [heads.refine_bboxes_single_image(self, p, l, r, g, m) for p, l, r, g, m in zip(props, labels, reg_outs, is_gts, img_metas)]


# ------------------------------------------------------------------------------------------------ dense-head test path
def anchor_head_predict_fast(self, cls_outs, reg_outs, img_metas, test_cfg):
    """AnchorHead.predict_bboxes_from_output for a dense (RetinaNet-style) head: the per-level best-class top-k, decode,
    clamp and size filter of ALL images as one batched K3 call (b2d_rpn_proposals, score_mode 2 / 3, do_nms 0); the class
    scores of the selected anchors are gathered per image and go through utils.multiclass_nms (one library call per
    image, as in heads.anchor_head_predict_single_image).  Returns [[bbox [4,k]]*B, [score [k]]*B, [label [k]]*B] or None."""
    from . import utils
    B = len(img_metas)
    cf = _closed_form(self, len(cls_outs))
    if cf is None or B == 0 or not _all_cuda_f32(cls_outs, reg_outs) or int(cls_outs[0].shape[0]) != B:
        return None
    sfs = [float(m.get('scale_factor', 1.0)) for m in img_metas]
    min_bbox = float(_get(test_cfg, 'min_bbox_size', 0) or 0)
    if min_bbox > 0 and any(s != sfs[0] for s in sfs):
        return None
    strides, scales, ratios, center_lt = cf
    dev = cls_outs[0].device
    C = int(self.cls_channels)
    grids = tuple(tuple(int(v) for v in c.shape[-2:]) for c in cls_outs)
    key = (B, grids, int(_get(test_cfg, 'pre_nms', 0)), min_bbox, sfs[0], str(dev), C, bool(self.use_sigmoid))
    cache = _cache(self, '_b2d_pred_batch_cache')
    rp = cache.get(key)
    if rp is None:
        pyr = fused.AnchorPyramid(strides, grids, scales, ratios, center_lt)
        sel_cfg = dict(pre_nms=int(_get(test_cfg, 'pre_nms', 0)), post_nms=0, max_num=0, nms_iou=0.5, min_bbox_size=min_bbox)
        try:
            rp = fused.RpnProposals(pyr, B, sel_cfg, self.target_means, self.target_stds, dev,
                                    score_mode=2 if self.use_sigmoid else 3, cls_channels=C, do_nms=False, scale_factor=sfs[0])
        except _C.B200DetError:
            return None
        cache[key] = rp
    cls, reg = [c.contiguous() for c in cls_outs], [r.contiguous() for r in reg_outs]
    with torch.no_grad():
        props, _, count = rp(cls, reg, _img_hw(img_metas, dev))
        ns = count.tolist()                               # the one synchronisation of the selection
        flat = torch.cat([c.reshape(B, C, -1) for c in cls], dim=2)              # [B, C, total] logits
    if self.use_sigmoid:
        label_set, adjust = list(range(0, self.num_classes - 1)), 1
    else:
        label_set, adjust = list(range(1, self.num_classes)), 0
    out = [[], [], []]
    for i, n in enumerate(ns):
        boxes = props[i][:, :n]
        picked = flat[i].index_select(1, rp.prov[i, :n].to(torch.int64))
        score = picked.sigmoid() if self.use_sigmoid else picked.softmax(dim=0)
        kb, ks, kl = utils.multiclass_nms(boxes.t().contiguous(), score.t().contiguous(), label_set, _get(test_cfg, 'nms_iou'),
                                          _get(test_cfg, 'min_score'), _get(test_cfg, 'max_per_img'),
                                          mode=_get(test_cfg, 'nms_type', 'official'))
        out[0].append(kb.t()); out[1].append(ks); out[2].append(kl + adjust)
    return out
