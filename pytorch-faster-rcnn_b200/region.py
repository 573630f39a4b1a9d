"""Region ops with the signatures of the reference's lib/region.py, on sm_100a kernels:
inside masks (:10-29), MaxIoUAssigner (:60-107), RandomSampler (:112-126),
IoUBalancedNegSampler (:128-172), ProposalCreator (:175-209), ScalableRoIPool/Align
(:212-239), BasicRoIExtractor (:243-306), SingleRoIExtractor (:309-375), plus
RoIAlign / RoIPool modules standing in for the torchvision classes the reference
registers directly (lib/builder.py:9,22).
"""
import logging

import numpy as np
import torch
import torch.nn as nn

from . import _C
from . import utils


# ---- a2 ---------------------------------------------------------------------------
def inside_grid_mask(num_anchors, img_size, grid_size, stride, device=None):
    """lib/region.py:10-16 -> fp32 flags [num_anchors*H*W]."""
    device = torch.device("cuda") if device is None else torch.device(device)
    h_ratio, w_ratio = 1.0 / stride, 1.0 / stride
    in_h = min(grid_size[0], int(img_size[0] * h_ratio) + 1)
    in_w = min(grid_size[1], int(img_size[1] * w_ratio) + 1)
    flags = torch.empty((num_anchors * grid_size[0] * grid_size[1],), dtype=torch.float32, device=device)
    _C.require_cuda(flags)
    _C.call("b2d_inside_grid_mask", _C.ptr(flags), int(num_anchors), int(grid_size[0]), int(grid_size[1]),
            in_h, in_w, _C.stream())
    return flags


def inside_anchor_mask(anchors, img_size, allowed_border=0):
    """lib/region.py:19-29 -> bool[n]."""
    _C.require_cuda(anchors)
    a = _C.f32c(anchors.reshape(4, -1))
    n = a.shape[1]
    mask = torch.empty(n, dtype=torch.uint8, device=a.device)
    H, W = img_size[:2]
    _C.call("b2d_inside_anchor_mask", _C.ptr(mask), _C.ptr(a), n, float(H), float(W), float(allowed_border),
            _C.stream())
    return mask.view(torch.bool) if hasattr(mask, "view") else mask.bool()


# ---- a4 ---------------------------------------------------------------------------
def _assign(boxes, gt, pos_iou, neg_iou, min_pos_iou, want_census=False):
    """Single-image front-end of b2d_assign_max_iou (explicit boxes)."""
    _C.require_cuda(boxes, gt)
    b, g = _C.f32c(boxes.reshape(4, -1)), _C.f32c(gt.reshape(4, -1))
    N, K = b.shape[1], g.shape[1]
    if K < 1:
        raise ValueError("MaxIoUAssigner needs at least one GT box (the reference raises in torch.max)")
    dev = b.device
    labels = torch.empty(N, dtype=torch.int64, device=dev)
    iou = torch.empty(N, dtype=torch.float32, device=dev)
    census = torch.empty(4, dtype=torch.int32, device=dev)
    pos_list = torch.empty(max(N, 1), dtype=torch.int32, device=dev) if want_census else None
    gcount = torch.full((1,), K, dtype=torch.int32, device=dev)
    ws = torch.empty(K, dtype=torch.int32, device=dev)
    f32 = np.float32
    _C.call("b2d_assign_max_iou", _C.ptr(labels), _C.ptr(iou), N, _C.ptr(b), N, None, N, None, None, 0.0,
            _C.ptr(g), K, _C.ptr(gcount), 1, float(f32(pos_iou)), float(f32(neg_iou)), float(f32(min_pos_iou)), 0,
            _C.ptr(census), _C.ptr(pos_list), max(N, 1), _C.ptr(ws), K * 4, _C.stream())
    if want_census:
        return labels, iou, census, pos_list
    return labels, iou


class MaxIoUAssigner(object):
    """lib/region.py:60-107: labels -1 ignore / 0 negative / (gt index + 1) positive,
    and the IoU with the assigned GT.  Bit-exact in fp32 (K2)."""

    def __init__(self, pos_iou, neg_iou, min_pos_iou):
        self.pos_iou = pos_iou
        self.neg_iou = neg_iou
        self.min_pos_iou = min_pos_iou

    def __call__(self, bboxes, gt_bboxes):
        assert bboxes.shape[0] == 4 and gt_bboxes.shape[0] == 4
        with torch.no_grad():
            return _assign(bboxes, gt_bboxes, self.pos_iou, self.neg_iou, self.min_pos_iou)


# ---- a5 ---------------------------------------------------------------------------
def _np_discard(idx_tensor, n_discard):
    """np.random.choice over the ascending index list, host RNG (lib/region.py:47-48)."""
    return np.random.choice(idx_tensor.cpu().numpy(), size=n_discard, replace=False)


# The reference builds its sampler from the config dict on EVERY call (lib/anchor.py:29-31, lib/bbox.py:19-20), so a
# per-instance call counter would restart with every image and the device sampler would walk the same permutation
# each time.  Samplers without an explicit seed therefore draw their stream index from one process-wide counter
# (restartable with manual_seed, like torch.manual_seed); an explicit seed= keeps the per-instance counter (tests).
_auto = {"seed": 0, "calls": 0}


def manual_seed(seed):
    """Restart the process-wide stream of the device samplers that were built without seed=."""
    _auto["seed"], _auto["calls"] = int(seed), 0


def _stream_key(sampler):
    if sampler.seed is None:
        _auto["calls"] += 1
        return (_auto["seed"] * 1000003 + 0x9E3779B97F4A7C15 + _auto["calls"]) & 0xFFFFFFFFFFFFFFFF
    sampler._calls += 1
    return (int(sampler.seed) * 1000003 + sampler._calls) & 0xFFFFFFFFFFFFFFFF


class RandomSampler(object):
    """lib/region.py:112-126.  rng='device' (default): the library's counter-based
    device sampler (no host sync); rng='numpy': the reference's exact host procedure,
    consuming numpy's global RNG stream identically (parity mode)."""
    default_rng = "device"

    def __init__(self, max_num, pos_num, rng=None, seed=None):
        assert pos_num <= max_num
        self.max_num = max_num
        self.pos_num = pos_num
        self.rng = rng or self.default_rng
        self.seed = seed
        self._calls = 0

    def __call__(self, labels, overlaps_iou=None, props_bbox=None, gt_bbox=None):
        _C.require_cuda(labels)
        if self.rng == "numpy":
            return self._numpy(labels)
        lab = labels.contiguous()
        n = lab.numel()
        dev = lab.device
        census = torch.empty(4, dtype=torch.int32, device=dev)
        pos_list = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        _C.call("b2d_label_census", _C.ptr(census), _C.ptr(pos_list), max(n, 1), _C.ptr(lab), n, None, n, 1,
                _C.stream())
        chosen = torch.empty(self.max_num, dtype=torch.int32, device=dev)
        n_chosen = torch.empty(1, dtype=torch.int32, device=dev)
        _C.call("b2d_sample_labels", _C.ptr(chosen), _C.ptr(n_chosen), _C.ptr(lab), n, None, None, n, _C.ptr(census),
                _C.ptr(pos_list), max(n, 1), 1, self.max_num, self.pos_num, _stream_key(self), None, _C.stream())
        out = torch.full_like(lab, -1)
        _C.call("b2d_scatter_sampled", _C.ptr(out), _C.ptr(lab), n, n, _C.ptr(chosen), _C.ptr(n_chosen),
                self.max_num, 1, _C.stream())
        return out

    def _numpy(self, labels):
        out = labels.clone().detach()
        pos = torch.nonzero(labels > 0).view(-1)
        if pos.numel() > self.pos_num:
            out[torch.as_tensor(_np_discard(pos, pos.numel() - self.pos_num), device=labels.device)] = -1
        n_negs = self.max_num - min(pos.numel(), self.pos_num)
        neg = torch.nonzero(labels == 0).view(-1)
        if neg.numel() > n_negs:
            out[torch.as_tensor(_np_discard(neg, neg.numel() - n_negs), device=labels.device)] = -1
        return out


class IoUBalancedNegSampler(object):
    """lib/region.py:128-172: positives capped at pos_num, negatives drawn per IoU bin of [0, max_iou) from the highest
    bin down (int(num_neg / num_bins) each, the lowest bin takes what is left, no back-fill).  floor_thr /
    floor_fraction are accepted and ignored exactly as there.

    rng='device' (default): one kernel (b2d_sample_iou_balanced), no host sync; a class over its quota keeps the
    members with the smallest keyed hash, the RandomSampler device rule.  rng='numpy': the reference's host stream
    -- the class of every element comes from b2d_iou_bin_ids, one copy to the host, then np.random.choice with the
    reference's arguments in the reference's order, one upload, b2d_scatter_sampled."""
    default_rng = "device"

    def __init__(self, max_num, pos_num, num_bins=3, max_iou=0.5, floor_thr=-1, floor_fraction=0, rng=None, seed=None):
        assert max_num >= pos_num
        assert max_iou > 0 and max_iou <= 1
        self.max_num, self.pos_num, self.num_bins, self.max_iou = max_num, pos_num, num_bins, max_iou
        self.rng, self.seed, self._calls = rng or self.default_rng, seed, 0
        # bin bounds as torch compares them: python doubles i * bin_size and s + bin_size, rounded to the fp32 of
        # `overlaps` (tensor-vs-scalar comparison); handed over highest bin first, the order of the reference's walk
        width = self.max_iou / self.num_bins
        lo = [np.float32(i * width) for i in range(self.num_bins)][::-1]
        hi = [np.float32(i * width + width) for i in range(self.num_bins)][::-1]
        self._lo = (_C.c_float * self.num_bins)(*[float(v) for v in lo])
        self._hi = (_C.c_float * self.num_bins)(*[float(v) for v in hi])

    def __call__(self, labels, overlaps, props_bbox=None, gt_bbox=None):
        _C.require_cuda(labels, overlaps)
        lab, iou = labels.to(torch.int64).contiguous(), _C.f32c(overlaps)
        n, dev = int(lab.numel()), lab.device
        if self.rng != "numpy":
            out = torch.empty_like(lab)
            _C.call("b2d_sample_iou_balanced", _C.ptr(out), _C.ptr(lab), _C.ptr(iou), n, self.max_num, self.pos_num,
                    self.num_bins, self._lo, self._hi, _stream_key(self), _C.stream())
            return out
        ids = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        _C.call("b2d_iou_bin_ids", _C.ptr(ids), _C.ptr(lab), _C.ptr(iou), n, self.num_bins, self._lo, self._hi, _C.stream())
        cls = ids[:n].cpu().numpy()
        members = [np.flatnonzero(cls == c) for c in range(self.num_bins + 1)]      # ascending, like nonzero()
        picked = members[0]
        if picked.shape[0] > self.pos_num:
            picked = picked[np.random.choice(picked.shape[0], self.pos_num, False)]
        budget = self.max_num - picked.shape[0]
        share, taken, parts = int(budget / self.num_bins), 0, [picked]
        for j in range(self.num_bins):
            m = members[1 + j]
            quota = share if j < self.num_bins - 1 else budget - taken
            if m.shape[0] > quota:
                m = m[np.random.choice(m.shape[0], quota, False)]
            taken += m.shape[0]
            parts.append(m)
        chosen = np.concatenate(parts).astype(np.int32)
        if chosen.shape[0] < self.max_num:
            logging.warning('Sampler can not sample max number of samples, instead: {}'.format(chosen.shape[0]))
        out = torch.empty_like(lab)
        cap = max(int(chosen.shape[0]), 1)
        d_chosen = torch.zeros(cap, dtype=torch.int32, device=dev)
        d_chosen[:chosen.shape[0]] = torch.from_numpy(chosen).to(dev)
        d_n = torch.tensor([chosen.shape[0]], dtype=torch.int32, device=dev)
        out.fill_(-1)
        if n:
            _C.call("b2d_scatter_sampled", _C.ptr(out), _C.ptr(lab), n, n, _C.ptr(d_chosen), _C.ptr(d_n), cap, 1, _C.stream())
        return out


def topk_desc(values, k):
    """Indices of the k largest values, descending, ties -> lowest index (segmented K3 top-k)."""
    _C.require_cuda(values)
    v = _C.f32c(values.reshape(1, -1))
    n = v.shape[1]
    k = min(int(k), n)
    if k <= 0:
        return torch.zeros(0, dtype=torch.int64, device=v.device)
    idx = torch.empty((1, k), dtype=torch.int32, device=v.device)
    cnt = torch.empty(1, dtype=torch.int32, device=v.device)
    wsb = _C.lib().b2d_topk_workspace_bytes(n, 1, k)
    if wsb == 0:
        raise _C.B200DetError("topk: k must be <= 16384")
    ws = utils._workspace(wsb, v.device, "topk")
    _C.call("b2d_topk", _C.ptr(idx), _C.ptr(cnt), _C.ptr(v), n, None, n, 1, k, _C.ptr(ws), ws.numel(), _C.stream())
    return idx.view(-1).long()


class ProposalCreator(object):
    """Legacy single-level proposal layer (lib/region.py:175-209; unregistered in the
    reference, kept for signature completeness): 2-channel softmax score, decode ALL
    anchors, drop boxes smaller than min_size (no +1), top pre_nms, NMS, top post_nms."""

    def __init__(self, pre_nms, post_nms, nms_iou, min_size):
        self.pre_nms, self.post_nms, self.nms_iou, self.min_size = pre_nms, post_nms, nms_iou, min_size

    def __call__(self, rpn_cls_out, rpn_reg_out, anchors, img_size, scale=1.0):
        assert anchors.shape[0] == 4 and len(anchors.shape) == 2
        min_size = scale * self.min_size
        with torch.no_grad():
            cls_out = rpn_cls_out.reshape(2, -1)
            reg_out = rpn_reg_out.reshape(4, -1)
            scores = torch.softmax(cls_out, 0)[1]
            props = utils.param2bbox(anchors, reg_out, img_size=img_size)
            small = (props[2] - props[0] < min_size) | (props[3] - props[1] < min_size)
            n_ok = int(scores.numel() - small.sum().item())     # the reference syncs here too (:201)
            masked = torch.where(small, torch.full_like(scores, float("-inf")), scores)
            top = topk_desc(masked, min(self.pre_nms, n_ok))
            props, top_scores = props[:, top], scores[top]
            keep = utils.nms(props.t().contiguous(), top_scores, self.nms_iou)[:self.post_nms]
        return props[:, keep], top_scores[keep]


# ---- K5-K7 front-ends -----------------------------------------------------------------
_LAYOUT_NCHW, _LAYOUT_NHWC, _LAYOUT_NHWC_BF16 = 0, 1, 2


def _feat_layout(feats):
    """0 = fp32 NCHW-contiguous, 1 = fp32 channels_last, 2 = bf16 channels_last (all levels alike)."""
    cl = all(f.dim() == 4 and f.is_contiguous(memory_format=torch.channels_last) and not f.is_contiguous()
             for f in feats)
    if cl:
        if all(f.dtype == torch.bfloat16 for f in feats):
            return _LAYOUT_NHWC_BF16, list(feats)
        if all(f.dtype == torch.float32 for f in feats):
            return _LAYOUT_NHWC, list(feats)
    return _LAYOUT_NCHW, [_C.f32c(f) for f in feats]


_nhwc_pool = {}


def _to_nhwc(feats):
    """fp32 NCHW-contiguous level maps -> channels_last copies (b2d_nchw_to_nhwc), so that reference-layout
    features (lib/necks.py FPN output) take the channel-vectorised K5 / K6 kernels instead of the generic ones.
    The copies are call-local temporaries (the kernels consume them, backward only needs the RoIs), so they live in a
    per-(level shape, device, stream) pool: re-allocating 0.7 GB per call cost 6 ms in the caching allocator (r2)."""
    out = []
    for l, f in enumerate(feats):
        key = (l, tuple(f.shape), f.device, torch.cuda.current_stream(f.device).cuda_stream)
        d = _nhwc_pool.get(key)
        if d is None:
            d = torch.empty(f.shape, dtype=torch.float32, device=f.device, memory_format=torch.channels_last)
            _nhwc_pool[key] = d
        _C.call("b2d_nchw_to_nhwc", _C.ptr(d), _C.ptr(f), f.shape[0], f.shape[1], f.shape[2], f.shape[3], _C.stream())
        out.append(d)
    return out


def _roi_cfg(feats, strides_or_scales, out_size, sampling_ratio, aligned, layout, finest_scale, scales=True):
    cfg = _C.RoiCfg()
    cfg.num_levels = len(feats)
    cfg.C = int(feats[0].shape[1])
    cfg.PH, cfg.PW = int(out_size[0]), int(out_size[1])
    cfg.sampling_ratio, cfg.aligned, cfg.layout = int(sampling_ratio), int(bool(aligned)), int(layout)
    cfg.finest_scale = float(finest_scale)
    for l, f in enumerate(feats):
        cfg.H[l], cfg.W[l] = int(f.shape[2]), int(f.shape[3])
        cfg.spatial_scale[l] = float(strides_or_scales[l]) if scales else 1.0 / float(strides_or_scales[l])
    return cfg


def _ptr_array(tensors):
    arr = (_C.c_void_p * _C.MAX_LEVELS)()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


class _RoIAlignFn(torch.autograd.Function):
    """Multi-level RoIAlign: forward K5, backward K6 (deterministic, atomic-free)."""

    @staticmethod
    def forward(ctx, rois, roi_img, levels, meta, *feats):
        scales, out_size, sr, aligned, finest = meta
        if any(ctx.needs_input_grad[4:]) and (sr <= 0 or out_size[0] * sr > 16 or out_size[1] * sr > 16):
            # fail at the forward call, not at loss.backward(): K6 has no adaptive-sampling / large-grid form
            raise _C.B200DetError("roi_align: gradients w.r.t. the features need a fixed sampling_ratio with "
                                  "output_size * sampling_ratio <= 16 per axis (got output_size=%s, sampling_ratio=%d)"
                                  % (tuple(out_size), sr))
        layout, fl = _feat_layout(feats)
        if layout == _LAYOUT_NCHW and sr == 2 and int(fl[0].shape[1]) % 4 == 0 and fl[0].shape[1] >= 16:
            layout, fl = _LAYOUT_NHWC, _to_nhwc(fl)      # reference layout: transpose once, then the fast kernels
        cfg = _roi_cfg(fl, scales, out_size, sr, aligned, layout, finest)
        R = rois.shape[1]
        out = torch.empty((R, cfg.C, cfg.PH, cfg.PW), dtype=torch.float32, device=rois.device)
        if R > 0:
            _C.call("b2d_roi_align_fwd", _C.ptr(out), _ptr_array(fl), _C.ptr(rois), rois.shape[1], _C.ptr(roi_img),
                    _C.ptr(levels), R, _C.ctypes.byref(cfg), _C.stream())
        ctx.save_for_backward(rois, roi_img, levels)
        ctx.meta = (meta, layout, [tuple(f.shape) for f in feats], [f.dtype for f in feats])
        return out

    @staticmethod
    def backward(ctx, gout):
        rois, roi_img, levels = ctx.saved_tensors
        (scales, out_size, sr, aligned, finest), layout, shapes, dtypes = ctx.meta
        dev = gout.device
        B = shapes[0][0]
        lay = _LAYOUT_NHWC if layout != _LAYOUT_NCHW else _LAYOUT_NCHW
        mf = torch.channels_last if lay == _LAYOUT_NHWC else torch.contiguous_format
        grads = [torch.empty(s, dtype=torch.float32, device=dev, memory_format=mf) for s in shapes]
        cfg = _roi_cfg(grads, scales, out_size, sr, aligned, lay, finest)
        R = rois.shape[1]
        if R == 0:
            return (None, None, None, None) + tuple(torch.zeros_like(g) for g in grads)
        wsb = _C.lib().b2d_roi_align_bwd_workspace_bytes(R, B, _C.ctypes.byref(cfg))
        ws = utils._workspace(wsb, dev, "roi_bwd")
        gout_c = _C.f32c(gout)
        _C.call("b2d_roi_align_bwd", _ptr_array(grads), _C.ptr(gout_c), _C.ptr(rois), rois.shape[1],
                _C.ptr(roi_img), _C.ptr(levels), R, B, _C.ctypes.byref(cfg), _C.ptr(ws), ws.numel(), _C.stream())
        grads = [g if d == torch.float32 else g.to(d) for g, d in zip(grads, dtypes)]
        return (None, None, None, None) + tuple(grads)


def roi_align_levels(feats, rois, roi_img, scales, out_size=(7, 7), sampling_ratio=2, aligned=False,
                     finest_scale=56, levels=None):
    """feats: list of [B,C,H,W]; rois [4,R]; roi_img int32[R] or None -> [R,C,PH,PW]."""
    _C.require_cuda(rois, *feats)
    rois = _C.f32c(rois.reshape(4, -1)).detach()
    meta = (tuple(float(s) for s in scales), utils.to_pair(out_size), int(sampling_ratio), bool(aligned),
            float(finest_scale))
    return _RoIAlignFn.apply(rois, roi_img, levels, meta, *feats)


class _RoIPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, rois, roi_img, scale, out_size):
        f = _C.f32c(feat)
        B, C, H, W = f.shape
        R = rois.shape[1]
        out = torch.empty((R, C, out_size[0], out_size[1]), dtype=torch.float32, device=f.device)
        arg = torch.empty((R, C, out_size[0], out_size[1]), dtype=torch.int32, device=f.device)
        if R > 0:
            _C.call("b2d_roi_pool_fwd", _C.ptr(out), _C.ptr(arg), _C.ptr(f), B, C, H, W, _C.ptr(rois), R,
                    _C.ptr(roi_img), R, float(scale), out_size[0], out_size[1], _C.stream())
        ctx.save_for_backward(rois, roi_img, arg)
        ctx.meta = (tuple(f.shape), float(scale), out_size)
        return out

    @staticmethod
    def backward(ctx, gout):
        rois, roi_img, arg = ctx.saved_tensors
        (B, C, H, W), scale, out_size = ctx.meta
        g = torch.empty((B, C, H, W), dtype=torch.float32, device=gout.device)
        R = rois.shape[1]
        if R == 0:
            return torch.zeros_like(g), None, None, None, None
        gout_c = _C.f32c(gout)
        _C.call("b2d_roi_pool_bwd", _C.ptr(g), _C.ptr(gout_c), _C.ptr(arg), B, C, H, W, _C.ptr(rois), R,
                _C.ptr(roi_img), R, scale, out_size[0], out_size[1], None, 0, _C.stream())
        return g, None, None, None, None


def _split_rois5(rois):
    """torchvision format [K,5] (batch idx, x1, y1, x2, y2) or list of [L,4] -> ([4,K], int32[K])."""
    if isinstance(rois, (list, tuple)):
        idx = torch.cat([torch.full((r.shape[0],), i, dtype=torch.int32, device=r.device) for i, r in enumerate(rois)])
        return torch.cat(list(rois), 0).t().contiguous().float(), idx
    return rois[:, 1:].t().contiguous().float(), rois[:, 0].to(torch.int32).contiguous()


class RoIAlign(nn.Module):
    """Signature of torchvision.ops.RoIAlign (registered by lib/builder.py:9,22)."""

    def __init__(self, output_size, spatial_scale, sampling_ratio, aligned=False):
        super(RoIAlign, self).__init__()
        self.output_size = utils.to_pair(output_size)
        self.spatial_scale, self.sampling_ratio, self.aligned = spatial_scale, sampling_ratio, aligned

    def crop(self, input, r4, idx):
        return roi_align_levels([input], r4, idx, [self.spatial_scale], self.output_size, self.sampling_ratio,
                                self.aligned)

    def forward(self, input, rois):
        return self.crop(input, *_split_rois5(rois))


class RoIPool(nn.Module):
    """Signature of torchvision.ops.RoIPool."""

    def __init__(self, output_size, spatial_scale):
        super(RoIPool, self).__init__()
        self.output_size = utils.to_pair(output_size)
        self.spatial_scale = spatial_scale

    def crop(self, input, r4, idx):
        _C.require_cuda(input)
        return _RoIPoolFn.apply(input, r4.detach(), idx, self.spatial_scale, self.output_size)

    def forward(self, input, rois):
        return self.crop(input, *_split_rois5(rois))


def roi_align(input, boxes, output_size, spatial_scale=1.0, sampling_ratio=-1, aligned=False):
    return RoIAlign(output_size, spatial_scale, sampling_ratio, aligned)(input, boxes)


def roi_pool(input, boxes, output_size, spatial_scale=1.0):
    return RoIPool(output_size, spatial_scale)(input, boxes)


class ScalableRoICrop(nn.Module):
    """lib/region.py:212-233: every RoI is rescaled about its centre (half sizes with the +1 convention, times
    `scale`) by b2d_scale_rois, then cropped by the wrapped RoI layer.  rois: torchvision's [n,5] or a list of [L,4]."""
    ROI_OP = None

    def __init__(self, scale=1.0, **kwargs):
        super(ScalableRoICrop, self).__init__()
        self.roi_op = self.ROI_OP(**kwargs)
        self.scale = scale
        self.kwargs = kwargs

    def forward(self, feats, rois):
        r4, idx = _split_rois5(rois)                      # [4,n] + image index
        n = int(r4.shape[1])
        scaled = torch.empty_like(r4)
        _C.require_cuda(r4)
        _C.call("b2d_scale_rois", _C.ptr(scaled), _C.ptr(r4), n, n, float(self.scale), _C.stream())
        return self.roi_op.crop(feats, scaled, idx)


class ScalableRoIPool(ScalableRoICrop):
    ROI_OP = RoIPool


class ScalableRoIAlign(ScalableRoICrop):
    ROI_OP = RoIAlign


_ROI_LAYER_TYPES = {"RoIAlign": RoIAlign, "RoIPool": RoIPool, "ScalableRoIPool": ScalableRoIPool,
                    "ScalableRoIAlign": ScalableRoIAlign}


def _level_map(rois, finest_scale, num_lvls):
    r = _C.f32c(rois.reshape(4, -1))
    lv = torch.empty(r.shape[1], dtype=torch.int32, device=r.device)
    if r.shape[1]:
        _C.call("b2d_roi_levels", _C.ptr(lv), _C.ptr(r), r.shape[1], r.shape[1], float(finest_scale), int(num_lvls),
                _C.stream())
    return lv


class BasicRoIExtractor(nn.Module):
    """lib/region.py:243-306.  With RoIAlign layers on every level (the FPN configs) all
    images and levels run as ONE fused launch; other layer mixes go level by level."""

    def __init__(self, roi_layers, output_size=(7, 7), finest_scale=56):
        assert isinstance(roi_layers, list)
        super(BasicRoIExtractor, self).__init__()
        self.output_size = utils.to_pair(output_size)
        self.finest_scale = finest_scale
        built = []
        for cfg in roi_layers:
            if isinstance(cfg, dict):
                cfg = dict(cfg)
                cfg['output_size'] = output_size
                typ = cfg.pop('type')
                if typ not in _ROI_LAYER_TYPES:
                    raise ValueError("'{}' is not registered".format(typ))
                if typ == "RoIPool":
                    cfg.pop("sampling_ratio", None)   # configs/faster_rcnn_r50.py:26 passes it (reference bug)
                cfg = _ROI_LAYER_TYPES[typ](**cfg)
            built.append(cfg)
        self.roi_layers = built
        self._fusable = all(type(l) is RoIAlign for l in built) and \
            len({(l.sampling_ratio, l.aligned) for l in built}) == 1

    def map_rois_to_levels(self, rois, num_lvls):
        with torch.no_grad():
            return _level_map(rois, self.finest_scale, num_lvls).long()

    def forward_single_level(self, feat, rois, i):
        assert feat.dim() in (3, 4)
        if feat.dim() == 3:
            feat = feat.unsqueeze(0)
        idx = rois.new_zeros((1, rois.shape[1]))
        return self.roi_layers[i](feat, torch.cat([idx, rois], dim=0).t())

    def forward_single_image(self, feats, rois):
        feats = [f.unsqueeze(0) if f.dim() == 3 else f for f in feats]
        return self.forward(feats, [rois])[0]

    def forward(self, level_feats, rois_list):
        n_lvls = len(self.roi_layers)
        assert n_lvls > 0 and n_lvls <= len(level_feats)
        feats = list(level_feats[:n_lvls])
        counts = [int(r.shape[1]) for r in rois_list]
        if self._fusable:
            rois = torch.cat([r.reshape(4, -1) for r in rois_list], dim=1) if len(rois_list) > 1 else rois_list[0]
            roi_img = torch.from_numpy(np.repeat(np.arange(len(counts), dtype=np.int32), counts)).to(rois.device, non_blocking=True) \
                if len(rois_list) > 1 else None           # one small upload instead of a fill per image
            l0 = self.roi_layers[0]
            out = roi_align_levels(feats, rois, roi_img, [l.spatial_scale for l in self.roi_layers], self.output_size,
                                   l0.sampling_ratio, l0.aligned, self.finest_scale)
            return list(out.split(counts, dim=0)) if len(rois_list) > 1 else [out]
        outs = []
        for i, rois in enumerate(rois_list):
            fi = [f[i:i + 1] for f in feats]
            if n_lvls == 1:
                outs.append(self.forward_single_level(fi[0], rois, 0))
                continue
            lv = self.map_rois_to_levels(rois, n_lvls)
            out = fi[0].new_zeros((rois.shape[1], fi[0].shape[1]) + tuple(self.output_size))
            for l in range(n_lvls):
                m = lv == l
                out[m] = self.forward_single_level(fi[l], rois[:, m], l)
            outs.append(out)
        return outs


class SingleRoIExtractor(nn.Module):
    """lib/region.py:309-375."""

    def __init__(self, roi_layer='RoIPool', output_size=7, featmap_strides=[16], finest_scale=56):
        super(SingleRoIExtractor, self).__init__()
        assert roi_layer in ['RoIPool', 'RoIAlign'], 'Unknown roi_layer type: {}'.format(roi_layer)
        self.roi_layer = roi_layer
        self.output_size = utils.to_pair(output_size)
        self.featmap_strides = featmap_strides
        self.finest_scale = finest_scale

    def map_props_to_levels(self, props, num_lvls):
        with torch.no_grad():
            return _level_map(props, self.finest_scale, num_lvls).long()

    def forward(self, level_feats, props_list):
        n = len(self.featmap_strides)
        assert len(level_feats) >= n
        feats = list(level_feats[:n])
        counts = [int(p.shape[1]) for p in props_list]
        rois = torch.cat([p.reshape(4, -1) for p in props_list], dim=1)
        roi_img = torch.cat([torch.full((c,), i, dtype=torch.int32, device=rois.device) for i, c in enumerate(counts)])
        scales = [1.0 / s for s in self.featmap_strides]
        if self.roi_layer == 'RoIAlign':
            out = roi_align_levels(feats, rois, roi_img, scales, self.output_size, 2, False, self.finest_scale)
            return list(out.split(counts, dim=0))
        if n == 1:
            out = _RoIPoolFn.apply(feats[0], _C.f32c(rois).detach(), roi_img, scales[0], self.output_size)
            return list(out.split(counts, dim=0))
        lv = self.map_props_to_levels(rois, n)
        out = feats[0].new_zeros((rois.shape[1], feats[0].shape[1]) + tuple(self.output_size))
        for l in range(n):
            m = lv == l
            out[m] = _RoIPoolFn.apply(feats[l], _C.f32c(rois[:, m]).detach(), roi_img[m].contiguous(), scales[l],
                                      self.output_size)
        return list(out.split(counts, dim=0))

    def forward_single_image(self, feats, props):
        feats = [f.unsqueeze(0) if f.dim() == 3 else f for f in feats]
        return self.forward(feats, [props])[0]
