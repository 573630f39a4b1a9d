"""Head-side glue of the hot path, same call signatures as the
reference methods they replace so dropin.install() can rebind them:

  refine_bboxes_single_image   BBoxHead.refine_bboxes_single_image   lib/heads/bbox_head.py:100-120
  single_image_targets_atss    FCOSHead.single_image_targets_atss    lib/heads/fcos_head.py:283-368
  fcos_predict_single_image    FCOSHead.predict_single_image         lib/heads/fcos_head.py:570-631

The arithmetic runs in csrc/heads.cu (b2d_refine_bboxes, b2d_atss_assign, b2d_fcos_decode)
plus the K3 top-k and K4 NMS; torch is only used for views, concatenation and indexing."""
import ctypes

import numpy as np
import torch

from . import _C, region, utils


def _ptrs(tensors):
    arr = (_C.c_void_p * _C.MAX_LEVELS)()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def _img_hw(img_size, device, B=1):
    return torch.tensor([[float(img_size[0]), float(img_size[1])]] * B, dtype=torch.float32, device=device)


# ------------------------------------------------------------------ a17
def refine_bboxes(props, label, reg_out, is_gt=None, img_shape=None, target_means=None, target_stds=None,
                  reg_class_agnostic=False, num_classes=21):
    """props [4,s], label int64[s], reg_out [s, 4] or [s, 4*num_classes] -> refined [4, s - #gt]."""
    _C.require_cuda(props, reg_out)
    s = int(props.shape[1])
    C = 1 if reg_class_agnostic else int(num_classes)
    if s == 0:
        return props.new_zeros((4, 0))
    assert reg_out.shape[0] == s and reg_out.shape[1] == 4 * C
    pr, rg = _C.f32c(props), _C.f32c(reg_out)
    lab = label.to(torch.int64).contiguous() if label is not None else None
    gt = is_gt.to(torch.int64).contiguous() if is_gt is not None else None
    out = torch.empty((4, s), dtype=torch.float32, device=props.device)
    cnt = torch.empty(1, dtype=torch.int32, device=props.device)
    clamp = img_shape is not None
    hw = _img_hw(img_shape[:2], props.device) if clamp else None
    _C.call("b2d_refine_bboxes", _C.ptr(out), _C.ptr(cnt), _C.ptr(pr), s, None, s, _C.ptr(lab), _C.ptr(rg), C,
            _C.ptr(gt), _C.host_f4(target_means, [0, 0, 0, 0]), _C.host_f4(target_stds, [1, 1, 1, 1]), int(clamp),
            _C.ptr(hw), 1, _C.stream())
    n_out = s if gt is None else int(cnt.item())
    return out[:, :n_out]


def refine_bboxes_single_image(self, props, label, reg_out, is_gt=None, img_meta=None):
    """Method form (bound onto the reference's BBoxHead by dropin.install)."""
    assert props.shape[1] == reg_out.shape[0]
    return refine_bboxes(props, label, reg_out, is_gt, img_meta['img_shape'] if img_meta is not None else None,
                         self.target_means, self.target_stds, self.reg_class_agnostic, self.num_classes)


# ------------------------------------------------------------------ a18 (K9)
def _point_pyramid(grids, strides, scale):
    levels = [dict(stride=s, H=int(g[0]), W=int(g[1]), ws=[np.float32(s * scale * np.sqrt(1.0))],
                   hs=[np.float32(s * scale / np.sqrt(1.0))], center_lt=False) for s, g in zip(strides, grids)]
    return _C.make_pyramid(levels)


def atss_assign(grids, strides, gt, gt_count, gt_label, img_hw, topk=9, scale=8):
    """Batched ATSS: gt [B,4,K], gt_count int32[B], gt_label int64[B,K], img_hw fp32[B,2] ->
    (cls int64 [B,total], reg fp32 [B,total,4], ctr fp32 [B,total]), cells level-major."""
    _C.require_cuda(gt)
    pyr = _point_pyramid(grids, strides, scale)
    B, dev = int(gt.shape[0]), gt.device
    total = int(pyr.total)
    cls = torch.empty((B, total), dtype=torch.int64, device=dev)
    reg = torch.empty((B, total, 4), dtype=torch.float32, device=dev)
    ctr = torch.empty((B, total), dtype=torch.float32, device=dev)
    wsb = _C.lib().b2d_atss_workspace_bytes(ctypes.byref(pyr), B)
    ws = utils._workspace(wsb, dev, "atss")
    gt_c, gl_c = _C.f32c(gt), gt_label.to(torch.int64).contiguous()      # locals: temporaries must outlive the launch
    _C.call("b2d_atss_assign", _C.ptr(cls), _C.ptr(reg), _C.ptr(ctr), ctypes.byref(pyr), _C.ptr(gt_c),
            int(gt.shape[2]), _C.ptr(gt_count), _C.ptr(gl_c), _C.ptr(img_hw), B, int(topk),
            _C.ptr(ws), ws.numel(), _C.stream())
    return cls, reg, ctr


def single_image_targets_atss(self, cls_outs, reg_outs, ctr_outs, lvl_anchors, gt_bboxes, gt_labels, img_meta,
                              train_cfg):
    """Method form of FCOSHead.single_image_targets_atss: per-level lists of [H,W,1] int64,
    [H,W,4] fp32, [H,W,1] fp32.  The level anchors are implied by (stride, atss_cfg.scale)."""
    grids = [tuple(int(v) for v in x.shape[-2:]) for x in cls_outs]
    dev = cls_outs[0].device
    K = int(gt_bboxes.shape[1])
    gt = _C.f32c(gt_bboxes).view(1, 4, K)
    cnt = torch.tensor([K], dtype=torch.int32, device=dev)
    cls, reg, ctr = atss_assign(grids, self.strides, gt, cnt, gt_labels.view(1, K), _img_hw(img_meta['img_shape'][:2], dev),
                                self.atss_cfg.topk, self.atss_cfg.scale)
    cls_t, reg_t, ctr_t, off = [], [], [], 0
    for (h, w) in grids:
        n = h * w
        cls_t.append(cls[0, off:off + n].view(h, w, 1))
        reg_t.append(reg[0, off:off + n].view(h, w, 4))
        ctr_t.append(ctr[0, off:off + n].view(h, w, 1))
        off += n
    return cls_t, reg_t, ctr_t


# ------------------------------------------------------------------ a19
def fcos_predict_single_image(cls_outs, reg_outs, ctr_outs, strides, img_meta, test_cfg, reg_mean=0.0, reg_std=300.0,
                              use_centerness=True):
    """cls_outs[l] [C,H,W] logits, reg_outs[l] [4,H,W] ltrb (before *std+mean), ctr_outs[l] [1,H,W]
    -> (bbox [4,k], score [k], label int64 [k], 1-based) like FCOSHead.predict_single_image."""
    get = (lambda k, d=None: test_cfg.get(k, d)) if hasattr(test_cfg, "get") else (lambda k, d=None: getattr(test_cfg, k, d))
    dev = cls_outs[0].device
    _C.require_cuda(*cls_outs)
    grids = [tuple(int(v) for v in x.shape[-2:]) for x in cls_outs]
    C = int(cls_outs[0].shape[0])
    pyr = _point_pyramid(grids, strides, 8)
    total = int(pyr.total)
    boxes = torch.empty((1, 4, total), dtype=torch.float32, device=dev)
    key = torch.empty((1, total), dtype=torch.float32, device=dev)
    score = torch.empty((1, C, total), dtype=torch.float32, device=dev)
    ctrs = torch.empty((1, total), dtype=torch.float32, device=dev) if use_centerness else None
    cl, rg = [_C.f32c(x) for x in cls_outs], [_C.f32c(x) for x in reg_outs]
    ct = [_C.f32c(x) for x in ctr_outs] if use_centerness else None
    min_size = float(np.float32(img_meta['scale_factor'] * get('min_bbox_size', 0)))
    hw_t = _img_hw(img_meta['img_shape'][:2], dev)
    _C.call("b2d_fcos_decode", _C.ptr(boxes), _C.ptr(key), _C.ptr(score), _C.ptr(ctrs), _ptrs(cl), _ptrs(rg),
            _ptrs(ct) if ct is not None else None, ctypes.byref(pyr), C, float(reg_mean), float(reg_std), min_size,
            _C.ptr(hw_t), 1, _C.stream())
    pre_nms = int(get('pre_nms', 0))
    sel, off = [], 0
    for (h, w) in grids:                                  # per-level filter + top-k (lib/heads/fcos_head.py:602-613)
        n = h * w
        k_l = key[0, off:off + n]
        valid = torch.nonzero(k_l > float("-inf")).view(-1)
        if pre_nms > 0 and pre_nms < int(valid.numel()):
            valid = region.topk_desc(k_l, pre_nms)        # -inf keys can never enter: #valid > pre_nms
        sel.append(valid + off)
        off += n
    sel = torch.cat(sel)
    mlvl_bbox, mlvl_score = boxes[0][:, sel], score[0][:, sel]
    mlvl_ctr = ctrs[0][sel] if use_centerness else None
    kb, ks, kl = utils.multiclass_nms(mlvl_bbox.t(), mlvl_score.t(), list(range(0, C)), get('nms_iou'), get('min_score'),
                                      get('max_per_img'), mlvl_ctr, mode=get('nms_type', 'official'))
    return kb.t(), ks, kl + 1


def predict_single_image(self, cls_outs, reg_outs, ctr_outs, img_meta, test_cfg):
    """Method form of FCOSHead.predict_single_image (DFL heads keep the reference path)."""
    return fcos_predict_single_image(cls_outs, reg_outs, ctr_outs, self.strides, img_meta, test_cfg, self.reg_mean,
                                     self.reg_std, self.use_centerness)


# ---------------------------------------------------------------------------------- a9
def rpn_predict_single_image(self, level_cls_outs, level_reg_outs, level_anchors, img_meta, test_cfg):
    """Method form of RPNHead.predict_single_image (lib/heads/rpn_head.py:68-120): the per-level
    sigmoid / top-k / decode / clamp / min-size / NMS / post-NMS loop and the final top-max_num as ONE
    call of the fused K3 + K4 path (b2d_rpn_proposals, B = 1).  Same arguments and return value
    `(bbox [4,k], score [k], None)`.  The anchors are regenerated in registers from the head's own
    anchor parameters (anchor_strides / anchor_scales / anchor_ratios, lib/heads/anchor_head.py:29-36);
    `level_anchors` is only used to read the grid sizes.  Selection happens on the logit (sigmoid and
    softmax[1] are monotone in it), ties go to the lowest index."""
    from . import fused
    dev = level_cls_outs[0].device
    _C.require_cuda(*level_cls_outs)
    grids = tuple(tuple(int(v) for v in a.shape[-2:]) for a in level_anchors)
    get = (lambda k, d=0: test_cfg.get(k, d)) if hasattr(test_cfg, "get") else (lambda k, d=0: getattr(test_cfg, k, d))
    sf = float(img_meta.get('scale_factor', 1.0))
    key = (grids, int(get('pre_nms')), int(get('post_nms')), int(get('max_num')), float(get('nms_iou', 0.7)),
           float(get('min_bbox_size', 0)), sf, str(dev))
    cache = self.__dict__.setdefault('_b2d_rpn_cache', {})
    rp = cache.get(key)
    if rp is None:
        center_lt = bool(getattr(self.anchor_creators[0], 'center_lt', False)) if getattr(self, 'anchor_creators', None) else False
        pyr = fused.AnchorPyramid(self.anchor_strides, grids, tuple(self.anchor_scales), tuple(self.anchor_ratios), center_lt)
        rp = fused.RpnProposals(pyr, 1, test_cfg, self.target_means, self.target_stds, dev,
                                score_mode=0 if self.use_sigmoid else 1, cls_channels=int(self.cls_channels),
                                scale_factor=sf)
        cache[key] = rp
    cls = [_C.f32c(x).view(1, *x.shape[-3:]) for x in level_cls_outs]
    reg = [_C.f32c(x).view(1, *x.shape[-3:]) for x in level_reg_outs]
    props, scores, count = rp(cls, reg, _img_hw(img_meta['img_shape'][:2], dev))
    k = int(count[0])
    return props[0][:, :k].clone(), scores[0][:k].clone(), None


# ---------------------------------------------------------------------------------- GA-RPN call sites
def ga_rpn_predict_single_image(self, cls_outs, reg_outs, anchors, in_masks, img_meta, cfg):
    """Method form of GARPNHead.predict_bboxes_single_image (lib/heads/guided_head.py:621-669): guided anchors are
    explicit per-location boxes `anchors[l]` [4,H,W] with a location mask `in_masks[l]` [1,H,W]; per level the unmasked
    places with the best objectness (top `pre_nms`), delta decode + clamp, min-size filter, NMS, `post_nms` cut, then
    the best `max_num` over all levels.  Returns `(bbox [4,k], score [k], None)`.
    All levels go through the kernels TOGETHER: one packed [L, ld] score array (masked places -inf), ONE segmented top-k,
    ONE gather + decode launch, ONE segmented NMS (levels = segments), one host read of the survivor counts.  Selection
    happens on the logit (sigmoid is monotone), ties go to the lowest index; where the reference slices with the
    misspelt `cfg.pos_nms` (:656-657, an AttributeError as soon as it is reached) this uses `post_nms`."""
    from . import utils
    dev = cls_outs[0].device
    _C.require_cuda(*cls_outs)
    L = len(cls_outs)
    get = (lambda k, d=0: cfg.get(k, d)) if hasattr(cfg, "get") else (lambda k, d=0: getattr(cfg, k, d))
    pre, post, max_num = int(get('pre_nms') or 0), int(get('post_nms') or 0), int(get('max_num') or 0)
    img_h, img_w = float(img_meta['img_shape'][0]), float(img_meta['img_shape'][1])
    min_size = float(np.float32(float(img_meta.get('scale_factor', 1.0)) * float(get('min_bbox_size', 0) or 0)))
    cls = [_C.f32c(c).reshape(-1) for c in cls_outs]
    reg = [_C.f32c(r).reshape(4, -1) for r in reg_outs]
    anc = [_C.f32c(a).reshape(4, -1) for a in anchors]
    msk = [m.reshape(-1).to(torch.bool).contiguous() for m in in_masks]
    ns = [int(c.numel()) for c in cls]
    ld = max(ns)
    k = min(pre, ld) if pre > 0 else ld
    if k > 16384:
        raise _C.B200DetError("ga_rpn_predict: at most 16384 candidates per level (pre_nms = %d, level of %d places)" % (pre, ld))
    n_host = (_C.c_int * _C.MAX_LEVELS)(*ns)
    packed = torch.empty((L, ld), dtype=torch.float32, device=dev)
    _C.call("b2d_ga_pack_scores", _C.ptr(packed), ld, _ptrs(cls), _ptrs(msk), n_host, L, _C.stream())
    idx = torch.empty((L, k), dtype=torch.int32, device=dev)
    cnt = torch.empty(L, dtype=torch.int32, device=dev)
    lens = torch.tensor(ns, dtype=torch.int32).to(dev, non_blocking=True)
    ws = utils._workspace(_C.lib().b2d_topk_workspace_bytes(ld, L, k), dev, "topk")
    _C.call("b2d_topk", _C.ptr(idx), _C.ptr(cnt), _C.ptr(packed), ld, _C.ptr(lens), ld, L, k, _C.ptr(ws), ws.numel(), _C.stream())
    box = torch.empty((L, k, 4), dtype=torch.float32, device=dev)
    score = torch.empty((L, k), dtype=torch.float32, device=dev)
    nvalid = torch.empty(L, dtype=torch.int32, device=dev)
    _C.call("b2d_ga_decode", _C.ptr(box), _C.ptr(score), _C.ptr(nvalid), _C.ptr(idx), _C.ptr(packed), ld, k, _ptrs(cls), _ptrs(msk),
            _ptrs(anc), _ptrs(reg), n_host, L, _C.host_f4(self.target_means, [0, 0, 0, 0]), _C.host_f4(self.target_stds, [1, 1, 1, 1]),
            img_h, img_w, min_size, _C.stream())
    keep = torch.empty((L, k), dtype=torch.int64, device=dev)
    kept = torch.empty(L, dtype=torch.int32, device=dev)
    ws2 = utils._workspace(_C.lib().b2d_nms_workspace_bytes(k, L), dev, "nms")
    _C.call("b2d_nms", _C.ptr(keep), _C.ptr(kept), _C.ptr(box), _C.ptr(score), k, _C.ptr(nvalid), k, L,
            _C.floor_f32(float(get('nms_iou', 0.7))), 0, 0, _C.ptr(ws2), ws2.numel(), _C.stream())
    # survivors in (level, score) order = the reference's concatenation; boxes below min_size carry score -inf and end
    # their level's list
    slot = torch.arange(k, device=dev)[None, :]
    kept_score = torch.gather(score, 1, keep.clamp(0, k - 1))          # (rows past a level's count are not written)
    ok = (slot < kept[:, None]) & (kept_score > float("-inf"))
    if post > 0:
        ok &= slot < post
    lv, js = torch.nonzero(ok, as_tuple=True)                # the one host synchronisation of the call
    src = keep[lv, js]
    out_score = score[lv, src]
    out_box = box[lv, src].t().contiguous()
    if max_num > 0 and int(out_score.numel()) > max_num:
        from . import region
        top = region.topk_desc(out_score, max_num).to(torch.int64)
        out_score, out_box = out_score[top], out_box[:, top].contiguous()
    return out_box, out_score, None


def ga_rpn_target_single_image(self, cls_outs, reg_outs, anchors, in_masks, gt_bbox, gt_label, img_meta, cfg):
    """Method form of GARPNHead.rpn_target_single_image (lib/heads/guided_head.py:557-572): anchor_target on the guided
    anchors of all levels, restricted to the location mask (K2 with explicit boxes + sampler + K8, anchor.anchor_target)."""
    from . import anchor as banchor
    cls_out = torch.cat([c.reshape(1, -1) for c in cls_outs], dim=1)
    reg_out = torch.cat([r.reshape(4, -1) for r in reg_outs], dim=1)
    all_anchors = torch.cat([a.reshape(4, -1) for a in anchors], dim=1)
    in_mask = torch.cat([m.reshape(-1) for m in in_masks]).to(torch.bool)
    get = (lambda k, d=None: cfg.get(k, d)) if hasattr(cfg, "get") else (lambda k, d=None: getattr(cfg, k, d))
    return banchor.anchor_target(cls_out, reg_out, 1, all_anchors[:, in_mask], in_mask, gt_bbox, None,
                                 assigner=get('assigner'), sampler=get('sampler'),
                                 target_means=self.target_means, target_stds=self.target_stds)


# ---------------------------------------------------------------------------------- RetinaNet test path (BASELINE config 4)
def anchor_head_predict_single_image(self, level_cls_outs, level_reg_outs, level_anchors, img_meta, test_cfg):
    """Method form of AnchorHead.predict_single_image (lib/heads/anchor_head.py:207-258), same arguments and return
    value `(bbox [4,k], score [k], label int64 [k])`.  Per level the best class score of every anchor, the top
    `pre_nms` of them, delta decode + clamp and the min-size filter are ONE call of the fused K3 path
    (b2d_rpn_proposals with score_mode 2 / 3 and do_nms 0; the anchors are regenerated in registers from the head's
    anchor parameters, `level_anchors` only gives the grid sizes); the class scores of the selected anchors are then
    gathered and handed to utils.multiclass_nms (one library call).  Selection happens on the best class LOGIT for a
    sigmoid head (sigmoid is monotone) and on max_{c>=1} softmax_c for a softmax head; ties go to the lowest index."""
    from . import fused
    dev = level_cls_outs[0].device
    _C.require_cuda(*level_cls_outs)
    grids = tuple(tuple(int(v) for v in a.shape[-2:]) for a in level_anchors)
    get = (lambda k, d=0: test_cfg.get(k, d)) if hasattr(test_cfg, "get") else (lambda k, d=0: getattr(test_cfg, k, d))
    sf = float(img_meta.get('scale_factor', 1.0))
    C = int(self.cls_channels)
    key = (grids, int(get('pre_nms')), float(get('min_bbox_size', 0)), sf, str(dev), C, bool(self.use_sigmoid))
    cache = self.__dict__.setdefault('_b2d_pred_cache', {})
    rp = cache.get(key)
    if rp is None:
        center_lt = bool(getattr(self.anchor_creators[0], 'center_lt', False)) if getattr(self, 'anchor_creators', None) else False
        pyr = fused.AnchorPyramid(self.anchor_strides, grids, tuple(self.anchor_scales), tuple(self.anchor_ratios), center_lt)
        sel_cfg = dict(pre_nms=int(get('pre_nms')), post_nms=0, max_num=0, nms_iou=0.5, min_bbox_size=get('min_bbox_size', 0))
        rp = fused.RpnProposals(pyr, 1, sel_cfg, self.target_means, self.target_stds, dev,
                                score_mode=2 if self.use_sigmoid else 3, cls_channels=C, do_nms=False, scale_factor=sf)
        cache[key] = rp
    cls = [_C.f32c(x).view(1, *x.shape[-3:]) for x in level_cls_outs]
    reg = [_C.f32c(x).view(1, *x.shape[-3:]) for x in level_reg_outs]
    props, _, count = rp(cls, reg, _img_hw(img_meta['img_shape'][:2], dev))
    k = int(count[0])
    boxes = props[0][:, :k]
    prov = rp.prov[0, :k].to(torch.int64)                 # flattened anchor index (level-major concat)
    flat = torch.cat([c.view(C, -1) for c in cls], dim=1)  # [C, total] logits
    picked = flat.index_select(1, prov)
    score = picked.sigmoid() if self.use_sigmoid else picked.softmax(dim=0)
    if self.use_sigmoid:
        nms_label_set, label_adjust = list(range(0, self.num_classes - 1)), 1
    else:
        nms_label_set, label_adjust = list(range(1, self.num_classes)), 0
    kb, ks, kl = utils.multiclass_nms(boxes.t().contiguous(), score.t().contiguous(), nms_label_set, get('nms_iou'),
                                      get('min_score'), get('max_per_img'), mode=get('nms_type', 'official'))
    return kb.t(), ks, kl + label_adjust


# ---------------------------------------------------------------------------------- CE / BCE on sampled rows
class _SampledCEFn(torch.autograd.Function):
    """sum over the rows of cross_entropy(pred, label) (softmax) or binary_cross_entropy_with_logits (sigmoid)."""

    @staticmethod
    def forward(ctx, pred, label, use_sigmoid):
        n, C = int(pred.shape[0]), int(pred.shape[1])
        # two layouts are read in place: contiguous [n, C], and the transposed view of a contiguous [C, n] tensor
        # (the reference's tar_cls_out.t(), lib/heads/anchor_head.py:129); anything else is made contiguous first
        if pred.is_contiguous():
            row_major, ld, src = 1, C, pred
        elif pred.t().is_contiguous():
            row_major, ld, src = 0, n, pred
        else:
            row_major, ld, src = 1, C, pred.contiguous()
        lab = label.to(torch.int64).contiguous()
        partial = torch.empty(int(_C.lib().b2d_sampled_ce_workspace_bytes()) // 4, dtype=torch.float32, device=pred.device)
        _C.call("b2d_sampled_ce_fwd", _C.ptr(partial), _C.ptr(src), ld, C, row_major, int(bool(use_sigmoid)), _C.ptr(lab), n,
                _C.stream())
        ctx.save_for_backward(src, lab)
        ctx.meta = (ld, C, row_major, int(bool(use_sigmoid)), n)
        return partial.view(-1, 2)[:, 0].sum()

    @staticmethod
    def backward(ctx, gout):
        src, lab = ctx.saved_tensors
        ld, C, row_major, sig, n = ctx.meta
        dev = src.device
        grad = torch.empty((n, C), dtype=torch.float32, device=dev) if row_major else \
            torch.empty((C, n), dtype=torch.float32, device=dev).t()
        scale = gout.detach().to(torch.float32).reshape(1).contiguous()
        _C.call("b2d_sampled_ce_bwd", _C.ptr(grad), _C.ptr(scale), _C.ptr(src), ld, C, row_major, sig, _C.ptr(lab), n, _C.stream())
        return grad, None, None


def sampled_cross_entropy(pred, label, use_sigmoid=False, loss_weight=1.0):
    """CrossEntropyLoss.forward (lib/losses.py:129-156) on the sampled rows: pred [n, C] (any strides: the RPN's
    `tar_cls_out.t()` is read in place), label int64 [n] -> loss_weight * sum of the row losses, differentiable in
    pred.  One kernel forward, one backward; callers divide by avg_factor like the reference
    (lib/heads/anchor_head.py:127-131, lib/heads/bbox_head.py:62-71)."""
    _C.require_cuda(pred, label)
    if pred.dtype != torch.float32:
        pred = pred.float()
    return _SampledCEFn.apply(pred, label, use_sigmoid) * loss_weight


def cross_entropy_loss_forward(self, pred, label):
    """Method form bound onto the reference's losses.CrossEntropyLoss by dropin.install()."""
    return sampled_cross_entropy(pred, label, self.use_sigmoid, self.loss_weight)


# ---------------------------------------------------------------------------------- SURVEY 8(f-3)
def rcnn_detect(props, cls_out, reg_out, img_size=None, target_means=None, target_stds=None, min_score=0.05,
                nms_iou=0.5, max_per_img=100, nms_type='official', counts=None, cap=None):
    """Batched RCNN test-time detections (b2d_rcnn_detect): props [B,4,n], cls_out [B,n,C] logits,
    reg_out [B,n,4C] (or [B,n,4], class-agnostic) -> (boxes [B,4,max_per_img], scores [B,max_per_img],
    labels int64 [B,max_per_img], count int32 [B]); softmax, per-class decode + clamp, class-offset NMS
    over classes 1..C-1 and the top max_per_img in one candidate kernel + K4.  No host sync."""
    _C.require_cuda(props, cls_out, reg_out)
    props, cls_out, reg_out = _C.f32c(props), _C.f32c(cls_out), _C.f32c(reg_out)
    B, _, n = props.shape
    C = int(cls_out.shape[-1])
    reg_c = int(reg_out.shape[-1]) // 4
    dev = props.device
    if cap is None:
        cap = max(1, min(16384, n * ((C - 1) if nms_type == 'official' else 1)))
    max_keep = max(1, min(int(max_per_img) if max_per_img is not None else cap, cap))
    out_box = torch.zeros((B, 4, max_keep), dtype=torch.float32, device=dev)
    out_score = torch.zeros((B, max_keep), dtype=torch.float32, device=dev)
    out_label = torch.zeros((B, max_keep), dtype=torch.int64, device=dev)
    out_count = torch.zeros(B, dtype=torch.int32, device=dev)
    overflow = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = utils._workspace(_C.lib().b2d_rcnn_detect_workspace_bytes(cap, B), dev, "rcnn_detect")
    hw = None
    if img_size is not None:
        hw = img_size if torch.is_tensor(img_size) else _img_hw(img_size[:2], dev, B)
    _C.call("b2d_rcnn_detect", _C.ptr(out_box), _C.ptr(out_score), _C.ptr(out_label), _C.ptr(out_count), _C.ptr(props),
            n, _C.ptr(counts), n, _C.ptr(cls_out), _C.ptr(reg_out), C, reg_c,
            _C.host_f4(target_means, [0, 0, 0, 0]), _C.host_f4(target_stds, [1, 1, 1, 1]), _C.ptr(hw),
            float(np.float32(min_score)), _C.floor_f32(float(nms_iou)), max_keep, int(nms_type == 'strict'), cap, B,
            _C.ptr(overflow), _C.ptr(ws), ws.numel(), _C.stream())
    return out_box, out_score, out_label, out_count, overflow


def predict_bboxes_single_image(self, props, cls_out, reg_out, img_size=None, cfg=None):
    """Method form of BBoxHead.predict_bboxes_single_image (lib/heads/bbox_head.py:122-146): same
    arguments, returns `(preds [4,k], score [k], label int64 [k])`."""
    if getattr(self, 'use_sigmoid', False):
        raise NotImplementedError('Need to be implemented')            # as the reference (bbox_head.py:132)
    get = (lambda k, d=None: cfg.get(k, d)) if hasattr(cfg, "get") else (lambda k, d=None: getattr(cfg, k, d))
    with torch.no_grad():
        b, s, l, cnt, ovf = rcnn_detect(props.unsqueeze(0), cls_out.unsqueeze(0), reg_out.unsqueeze(0), img_size,
                                        self.target_means, self.target_stds, get('min_score'), get('nms_iou'),
                                        get('max_per_img'), get('nms_type', 'official'))
        k, o = int(cnt[0]), int(ovf[0])
        if o:
            raise _C.B200DetError("rcnn_detect: more than 16384 (box, class) candidates; raise min_score")
    return b[0][:, :k], s[0][:k], l[0][:k]


# ---------------------------------------------------------------------------------- SURVEY 8(f-4)
def fcos_targets(grids, strides, gt, gt_count, gt_label, img_hw, level_scale_thr=(0, 64, 128, 256, 512, 1e6)):
    """Batched plain FCOS targets (b2d_fcos_targets): gt [B,4,K], gt_count int32[B], gt_label int64[B,K],
    img_hw fp32[B,2] -> (cls int64 [B,total], reg fp32 [B,total,4], ctr fp32 [B,total]), cells level-major."""
    _C.require_cuda(gt)
    pyr = _point_pyramid(grids, strides, 8)
    B, dev = int(gt.shape[0]), gt.device
    total = int(pyr.total)
    cls = torch.empty((B, total), dtype=torch.int64, device=dev)
    reg = torch.empty((B, total, 4), dtype=torch.float32, device=dev)
    ctr = torch.empty((B, total), dtype=torch.float32, device=dev)
    thr = (ctypes.c_float * (len(grids) + 1))(*[float(v) for v in level_scale_thr[:len(grids) + 1]])
    gt_c, gl_c = _C.f32c(gt), gt_label.to(torch.int64).contiguous()      # locals: temporaries must outlive the launch
    _C.call("b2d_fcos_targets", _C.ptr(cls), _C.ptr(reg), _C.ptr(ctr), ctypes.byref(pyr), _C.ptr(gt_c),
            int(gt.shape[2]), _C.ptr(gt_count), _C.ptr(gl_c), _C.ptr(img_hw), thr, B,
            _C.stream())
    return cls, reg, ctr


def single_image_targets(self, cls_outs, reg_outs, ctr_outs, gt_bboxes, gt_labels, img_meta, train_cfg):
    """Method form of FCOSHead.single_image_targets (lib/heads/fcos_head.py:371-416): per-level lists of
    [H,W,1] int64, [H,W,4] fp32, [H,W,1] fp32."""
    grids = [tuple(int(v) for v in x.shape[-2:]) for x in cls_outs]
    dev = cls_outs[0].device
    K = int(gt_bboxes.shape[1])
    gt = _C.f32c(gt_bboxes).view(1, 4, K)
    cnt = torch.tensor([K], dtype=torch.int32, device=dev)
    cls, reg, ctr = fcos_targets(grids, self.strides, gt, cnt, gt_labels.view(1, K), _img_hw(img_meta['img_shape'][:2], dev),
                                 self.level_scale_thr)
    cls_t, reg_t, ctr_t, off = [], [], [], 0
    for (h, w) in grids:
        n = h * w
        cls_t.append(cls[0, off:off + n].view(h, w, 1))
        reg_t.append(reg[0, off:off + n].view(h, w, 4))
        ctr_t.append(ctr[0, off:off + n].view(h, w, 1))
        off += n
    return cls_t, reg_t, ctr_t


# ---------------------------------------------------------------------------------- SURVEY 8(f-2)
class _AnchorLossFn(torch.autograd.Function):
    """{sum focal, sum smooth-L1, #pos} of an anchor head without sampler, differentiable in the head maps."""

    @staticmethod
    def forward(ctx, pyr, labels, gt, gt_label, meta, n, *maps):
        cls, reg = maps[:n], maps[n:]
        C, alpha, gamma, beta, means, stds = meta
        B, dev = int(cls[0].shape[0]), cls[0].device
        out = torch.empty(3, dtype=torch.float32, device=dev)
        ws = utils._workspace(_C.lib().b2d_anchor_loss_workspace_bytes(ctypes.byref(pyr), B), dev, "anchor_loss")
        _C.call("b2d_anchor_loss_fwd", _C.ptr(out), _ptrs(cls), _ptrs(reg), ctypes.byref(pyr), _C.ptr(labels),
                int(labels.shape[1]), _C.ptr(gt), int(gt.shape[2]), _C.ptr(gt_label), C, alpha, gamma, beta, means, stds, B,
                _C.ptr(ws), ws.numel(), _C.stream())
        ctx.save_for_backward(labels, gt, gt_label, *maps)
        ctx.misc = (pyr, meta, n)
        return out

    @staticmethod
    def backward(ctx, gout):
        labels, gt, gt_label = ctx.saved_tensors[:3]
        maps = ctx.saved_tensors[3:]
        pyr, (C, alpha, gamma, beta, means, stds), n = ctx.misc
        cls, reg = maps[:n], maps[n:]
        scale = gout[:2].detach().to(torch.float32).contiguous()
        dcls, dreg = [torch.empty_like(c) for c in cls], [torch.empty_like(r) for r in reg]
        _C.call("b2d_anchor_loss_bwd", _ptrs(dcls), _ptrs(dreg), _C.ptr(scale), _ptrs(cls), _ptrs(reg), ctypes.byref(pyr),
                _C.ptr(labels), int(labels.shape[1]), _C.ptr(gt), int(gt.shape[2]), _C.ptr(gt_label), C, alpha, gamma, beta,
                means, stds, int(cls[0].shape[0]), _C.stream())
        return (None,) * 6 + tuple(dcls) + tuple(dreg)


def anchor_head_loss_sums(cls_outs, reg_outs, labels, pyramid, gt, gt_label, target_means=None, target_stds=None,
                          alpha=0.25, gamma=2.0, beta=1.0 / 9.0):
    """Fused AnchorHead.calc_loss pieces for a head without sampler (lib/heads/anchor_head.py:113-139;
    sigmoid_focal_loss lib/losses.py:33-61, smooth_l1_loss_v2 :77-83): cls_outs[l] [B, A*C, H, W] logits,
    reg_outs[l] [B, 4A, H, W], labels int64 [B, total] from the batched assignment (gt index + 1 / 0 / -1),
    pyramid = fused.AnchorPyramid, gt [B,4,K], gt_label int64 [B,K] -> fp32[3] = {sum focal, sum smooth-L1,
    #pos}, differentiable w.r.t. the maps.  No gathers, no host sync."""
    _C.require_cuda(*cls_outs)
    cls = [_C.f32c(x) for x in cls_outs]
    reg = [_C.f32c(x) for x in reg_outs]
    A = int(pyramid.num_anchors)
    C = int(cls[0].shape[1]) // A
    meta = (C, float(alpha), float(gamma), float(beta), _C.host_f4(target_means, [0, 0, 0, 0]),
            _C.host_f4(target_stds, [1, 1, 1, 1]))
    return _AnchorLossFn.apply(pyramid.c, labels.contiguous(), _C.f32c(gt), gt_label.to(torch.int64).contiguous(), meta,
                               len(cls), *cls, *reg)


def anchor_head_calc_loss(cls_outs, reg_outs, labels, pyramid, gt, gt_label, target_means=None, target_stds=None,
                          alpha=0.25, gamma=2.0, beta=1.0 / 9.0, cls_weight=1.0, reg_weight=1.0):
    """(cls_loss, reg_loss) of AnchorHead.calc_loss without sampler: both sums divided by #pos
    (`avg_factor = pos_tars.sum()`, lib/heads/anchor_head.py:127-128)."""
    s = anchor_head_loss_sums(cls_outs, reg_outs, labels, pyramid, gt, gt_label, target_means, target_stds, alpha, gamma, beta)
    return cls_weight * s[0] / s[2], reg_weight * s[1] / s[2]
