"""Box math + NMS front-ends with the signatures of the reference's lib/utils.py
(box part, lib/utils.py:40-269), executing on the sm_100a kernels of libb200det.so.

Layout: boxes are [4, n] fp32 (x1, y1, x2, y2), +1 width/height convention
(lib/utils.py:40-45).  CUDA tensors only -- there is no CPU path.
"""
import torch

from . import _C

_Z4 = [0.0, 0.0, 0.0, 0.0]
_O4 = [1.0, 1.0, 1.0, 1.0]


# ---- pure-python glue kept for API completeness (lib/utils.py:278-301) ----------
def multi_apply(func, *args):
    list_args = [a for a in args if isinstance(a, list)]
    n = len(list_args[0]) if list_args else 1
    for a in list_args:
        if len(a) != n:
            raise ValueError('Arg: {} does not have the same length as others'.format(a))
    return [func(*[a[i] if isinstance(a, list) else a for a in args]) for i in range(n)]


def unpack_multi_result(multi_res):
    assert len(multi_res) != 0
    return [[res[i] for res in multi_res] for i in range(len(multi_res[0]))]


def to_pair(val):
    if isinstance(val, int):
        return (val, val)
    val = list(val)
    assert len(val) == 2
    return tuple(val)


def wh_from_xyxy(bbox):
    return bbox[2] - bbox[0] + 1, bbox[3] - bbox[1] + 1


def center_of(bbox):
    return (bbox[2] + bbox[0]) / 2, (bbox[3] + bbox[1]) / 2


def simplify_label(label):
    out = label.clone().detach()
    out[label > 0] = 1
    return out


# ---- a3 ------------------------------------------------------------------------
def calc_iou(a, b):
    """lib/utils.py:151-172 -> [N, K] IoU table (+1 areas), bit-exact incl. signed zeros."""
    assert a.shape[0] == 4 and b.shape[0] == 4
    _C.require_cuda(a, b)
    a2, b2 = _C.f32c(a.reshape(4, -1)), _C.f32c(b.reshape(4, -1))
    N, K = a2.shape[1], b2.shape[1]
    out = torch.empty((N, K), dtype=torch.float32, device=a.device)
    _C.call("b2d_calc_iou", _C.ptr(out), _C.ptr(a2), N, _C.ptr(b2), K, _C.stream())
    return out


def elem_iou(a, b):
    """lib/utils.py:174-182 (un-paired, no +1)."""
    assert a.shape[0] == 4 and b.shape[0] == 4 and a.shape == b.shape
    _C.require_cuda(a, b)
    a2, b2 = _C.f32c(a.reshape(4, -1)), _C.f32c(b.reshape(4, -1))
    out = torch.empty(a2.shape[1], dtype=torch.float32, device=a.device)
    _C.call("b2d_elem_iou", _C.ptr(out), _C.ptr(a2), _C.ptr(b2), a2.shape[1], _C.stream())
    return out.reshape(a.shape[1:])


# ---- a7 / a8 ---------------------------------------------------------------------
def bbox2param(base, bbox, means=_Z4, stds=_O4):
    """lib/utils.py:47-70."""
    assert base.shape == bbox.shape
    _C.require_cuda(base, bbox)
    b1, b2 = _C.f32c(base), _C.f32c(bbox)
    out = torch.empty_like(b1)
    _C.call("b2d_bbox2param", _C.ptr(out), _C.ptr(b1), _C.ptr(b2), b1.shape[1], _C.host_f4(means, _Z4),
            _C.host_f4(stds, _O4), _C.stream())
    return out


def param2bbox(base, param, means=_Z4, stds=_O4, img_size=None):
    """lib/utils.py:83-92 (+ clamp_bbox :109-120 when img_size is given)."""
    assert base.shape == param.shape
    assert base.shape[0] == 4
    _C.require_cuda(base, param)
    b1, p1 = _C.f32c(base), _C.f32c(param)
    out = torch.empty_like(b1)
    clamp = img_size is not None
    h, w = (float(img_size[0]), float(img_size[1])) if clamp else (0.0, 0.0)
    _C.call("b2d_param2bbox", _C.ptr(out), _C.ptr(b1), _C.ptr(p1), b1.shape[1], _C.host_f4(means, _Z4),
            _C.host_f4(stds, _O4), int(clamp), h, w, _C.stream())
    return out


def batched_param2bbox(base, param, means=_Z4, stds=_O4, img_size=None):
    """lib/utils.py:96-106: param [4*cls, n] viewed (4, cls, n) -> bbox [4*cls, n]."""
    assert param.shape[0] % 4 == 0
    cls = param.shape[0] // 4
    if cls == 1:
        return param2bbox(base, param, means, stds, img_size)
    n = param.shape[1]
    # (4, cls, n) -> one [4, cls*n] decode against the base tiled cls times
    p = _C.f32c(param).view(4, cls * n)
    b = _C.f32c(base).unsqueeze(1).expand(4, cls, n).reshape(4, cls * n)
    return param2bbox(b, p, means, stds, img_size).view(4 * cls, n)


def clamp_bbox(bbox, img_size):
    """lib/utils.py:109-120."""
    _C.require_cuda(bbox)
    b1 = _C.f32c(bbox)
    out = torch.empty_like(b1)
    _C.call("b2d_clamp_bbox", _C.ptr(out), _C.ptr(b1), b1.shape[1], float(img_size[0]), float(img_size[1]),
            _C.stream())
    return out


# ---- a11 / a12 ---------------------------------------------------------------------
_ws_cache = {}


def _workspace(nbytes, device, tag):
    key = (tag, device)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def nms(boxes, scores, iou_threshold):
    """Drop-in for torchvision.ops.nms (CPU semantics): boxes [n,4], scores [n] ->
    int64 indices of kept boxes, in decreasing score order (ties: lower index first)."""
    _C.require_cuda(boxes, scores)
    n = int(scores.numel())
    if n == 0:
        return torch.zeros(0, dtype=torch.int64, device=boxes.device)
    b, s = _C.f32c(boxes.reshape(n, 4)), _C.f32c(scores.reshape(n))
    keep = torch.empty(n, dtype=torch.int64, device=boxes.device)
    cnt = torch.empty(1, dtype=torch.int32, device=boxes.device)
    wsb = _C.lib().b2d_nms_workspace_bytes(n, 1)
    ws = _workspace(wsb, boxes.device, "nms")
    _C.call("b2d_nms", _C.ptr(keep), _C.ptr(cnt), _C.ptr(b), _C.ptr(s), n, None, n, 1,
            _C.floor_f32(float(iou_threshold)), 0, 0, _C.ptr(ws), ws.numel(), _C.stream())
    return keep[: int(cnt.item())]


def batched_nms(bbox, score, label, nms_iou, class_agnostic=False):
    """lib/utils.py:211-221 (class-offset trick; the fp32 add is part of the result)."""
    numel = score.numel()
    if numel == 0:
        return bbox, score, label
    if class_agnostic:
        nms_bbox = bbox
    else:
        max_range = bbox.max()
        nms_bbox = bbox + (label * max_range).to(bbox).view(numel, 1)
    keep = nms(nms_bbox, score, nms_iou)
    return bbox[keep, :], score[keep], label[keep]


def multiclass_nms(bbox, score, nms_channel, nms_iou, min_score=-1, max_num=None, score_factor=None,
                   mode='official'):
    """lib/utils.py:224-269.  Candidate enumeration is index glue; the NMS is K4."""
    assert mode in ['official', 'strict']
    assert score.dim() == 2, 'multiclass_nms only applies to multi-channel score'
    cls_channel = score.shape[1]
    num_bbox = bbox.shape[0]
    simple_bbox = bbox.shape[1] == 4
    nms_channel = list(nms_channel)
    if mode == 'official':
        chan = torch.zeros(cls_channel, dtype=torch.bool, device=score.device)
        chan[torch.as_tensor(nms_channel, dtype=torch.long, device=score.device)] = True
        label = torch.arange(cls_channel, device=score.device).view(1, -1).expand(num_bbox, -1)
        if simple_bbox:
            bbox = bbox.unsqueeze(2).expand(-1, -1, cls_channel)
        else:
            bbox = bbox.view(num_bbox, 4, cls_channel)
        bbox = bbox.permute(0, 2, 1)
        chosen = (score >= min_score) & chan.view(1, -1)
        if score_factor is not None:
            if score_factor.dim() == 1:
                score_factor = score_factor.unsqueeze(1)
            score = score * score_factor
        nms_bbox, nms_score, nms_label = bbox[chosen], score[chosen], label[chosen]
    else:
        score, label = score.max(1)
        chosen = torch.zeros_like(label, dtype=torch.bool)
        for cha in nms_channel:
            chosen = chosen | (label == cha)
        if not simple_bbox:
            bbox = bbox.view(num_bbox, 4, cls_channel)[torch.arange(num_bbox, device=bbox.device), :, label]
        chosen = (score >= min_score) & chosen
        if score_factor is not None:
            score = score * score_factor
        nms_bbox, nms_score, nms_label = bbox[chosen, :], score[chosen], label[chosen]
    keep_bbox, keep_score, keep_label = batched_nms(nms_bbox, nms_score, nms_label, nms_iou)
    if max_num is not None and keep_score.numel() > max_num:
        keep_bbox, keep_score, keep_label = keep_bbox[:max_num], keep_score[:max_num], keep_label[:max_num]
    return keep_bbox, keep_score, keep_label
