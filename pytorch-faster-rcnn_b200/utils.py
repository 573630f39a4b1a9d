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
NMS_MAX_BOXES = 16384        # b2d_nms: boxes per call (dense mask above 2048 boxes); see INTEGRATION.md


def _workspace(nbytes, device, tag):
    """Scratch buffer for one entry point, cached per (tag, device, CURRENT STREAM): two streams never share scratch
    memory, and a buffer is only ever replaced by a larger one allocated on the same stream (the caching allocator
    then keeps the old block alive until that stream has passed the kernels that use it)."""
    stream = torch.cuda.current_stream(device).cuda_stream if device.type == "cuda" else 0
    key = (tag, device, stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _tv_nms(boxes, scores, iou_threshold):
    import torchvision
    return torchvision.ops.nms(boxes, scores, iou_threshold)


def nms(boxes, scores, iou_threshold):
    """Drop-in for torchvision.ops.nms (CPU semantics): boxes [n,4], scores [n] ->
    int64 indices of kept boxes, in decreasing score order (ties: lower index first).
    More than NMS_MAX_BOXES boxes (torchvision has no limit) are handed to torchvision's CUDA kernel."""
    _C.require_cuda(boxes, scores)
    n = int(scores.numel())
    if n == 0:
        return torch.zeros(0, dtype=torch.int64, device=boxes.device)
    if n > NMS_MAX_BOXES:
        return _tv_nms(boxes.reshape(n, 4).float(), scores.reshape(n).float(), float(iou_threshold))
    b, s = _C.f32c(boxes.reshape(n, 4)), _C.f32c(scores.reshape(n))
    keep = torch.empty(n, dtype=torch.int64, device=boxes.device)
    cnt = torch.empty(1, dtype=torch.int32, device=boxes.device)
    wsb = _C.lib().b2d_nms_workspace_bytes(n, 1)
    ws = _workspace(wsb, boxes.device, "nms")
    _C.call("b2d_nms", _C.ptr(keep), _C.ptr(cnt), _C.ptr(b), _C.ptr(s), n, None, n, 1,
            _C.floor_f32(float(iou_threshold)), 0, 0, _C.ptr(ws), ws.numel(), _C.stream())
    return keep[: int(cnt.item())]


def batched_nms(bbox, score, label, nms_iou, class_agnostic=False):
    """lib/utils.py:211-221: class-aware NMS through the coordinate offset label * bbox.max() (the fp32 add is part
    of the reference result); bbox [n,4], score [n], label int64 [n] -> the kept rows of all three, score order.
    The maximum and the shifted boxes come from one kernel (b2d_batched_nms_boxes)."""
    n = int(score.numel())
    if n == 0:
        return bbox, score, label
    _C.require_cuda(bbox, score, label)
    boxes = _C.f32c(bbox.reshape(n, 4))
    if not class_agnostic:
        shifted = torch.empty_like(boxes)
        lab64 = label.to(torch.int64).contiguous()          # (kept in a local: a temporary would be freed before the launch)
        _C.call("b2d_batched_nms_boxes", _C.ptr(shifted), _C.ptr(boxes), _C.ptr(lab64), n, _C.stream())
        boxes = shifted
    keep = nms(boxes, score, nms_iou)
    return bbox.index_select(0, keep), score.index_select(0, keep), label.index_select(0, keep)


def multiclass_nms(bbox, score, nms_channel, nms_iou, min_score=-1, max_num=None, score_factor=None,
                   mode='official'):
    """lib/utils.py:224-269 as ONE library call (b2d_multiclass_nms): bbox [n,4] or [n,4*C], score [n,C] ->
    (bbox [k,4], score [k], label int64 [k]).  Candidate test, ordered compaction, class offsets, NMS and the
    max_num cut run on the device; the only host read is the survivor count."""
    assert mode in ['official', 'strict']
    assert score.dim() == 2, 'multiclass_nms only applies to multi-channel score'
    _C.require_cuda(bbox, score)
    n, C = int(score.shape[0]), int(score.shape[1])
    dev = score.device
    if n == 0:
        return bbox.new_zeros((0, 4)), score.new_zeros((0,)), torch.zeros(0, dtype=torch.int64, device=dev)
    box_classes = 1 if bbox.shape[1] == 4 else C
    assert bbox.shape[1] == 4 * box_classes and C <= 128
    chan = [0, 0]
    for c in nms_channel:
        c = int(c)
        if 0 <= c < C:
            chan[c >> 6] |= 1 << (c & 63)
    cand_max = n * (len(list(nms_channel)) if mode == 'official' else 1)
    # contiguous fp32 copies are held in locals until the call returns: a temporary created inside the argument list is
    # freed at once and the next temporary may reuse (and overwrite) its block before the kernel runs
    box_c, score_c = _C.f32c(bbox), _C.f32c(score)
    factor = None
    if score_factor is not None:
        factor = _C.f32c(score_factor.reshape(-1))
        assert factor.numel() == n
    chan_arr = (_C.c_ull * 2)(*chan)
    min_f = float(torch.tensor(min_score, dtype=torch.float32))          # the fp32 torch compares the scores with
    thr = _C.floor_f32(float(nms_iou))
    if cand_max > NMS_MAX_BOXES:
        # more (box, class) pairs than b2d_nms takes could pass the test: candidate kernel alone, then the NMS by size
        cbox = torch.empty((cand_max, 4), dtype=torch.float32, device=dev)
        nbox = torch.empty((cand_max, 4), dtype=torch.float32, device=dev)
        cscore = torch.empty(cand_max, dtype=torch.float32, device=dev)
        clabel = torch.empty(cand_max, dtype=torch.int32, device=dev)
        meta = torch.empty(2, dtype=torch.int32, device=dev)
        _C.call("b2d_multiclass_candidates", _C.ptr(cbox), _C.ptr(nbox), _C.ptr(cscore), _C.ptr(clabel), _C.ptr(meta[0:1]),
                _C.ptr(meta[1:2]), _C.ptr(box_c), box_classes, _C.ptr(score_c), n, C, chan_arr, min_f, _C.ptr(factor),
                int(mode == 'strict'), cand_max, _C.stream())
        m = int(meta[0])
        keep = nms(nbox[:m], cscore[:m], nms_iou)
        if max_num is not None:
            keep = keep[:int(max_num)]
        return cbox.index_select(0, keep), cscore.index_select(0, keep), clabel.index_select(0, keep).to(torch.int64)
    cap = max(1, cand_max)
    max_keep = cap if max_num is None else max(1, min(int(max_num), cap))
    out_box = torch.empty((max_keep, 4), dtype=torch.float32, device=dev)
    out_score = torch.empty(max_keep, dtype=torch.float32, device=dev)
    out_label = torch.empty(max_keep, dtype=torch.int64, device=dev)
    meta = torch.empty(2, dtype=torch.int32, device=dev)              # survivor count, overflow flag
    ws = _workspace(_C.lib().b2d_multiclass_nms_workspace_bytes(cap), dev, "multiclass_nms")
    _C.call("b2d_multiclass_nms", _C.ptr(out_box), _C.ptr(out_score), _C.ptr(out_label), _C.ptr(meta[0:1]),
            _C.ptr(box_c), box_classes, _C.ptr(score_c), n, C, chan_arr, min_f, _C.ptr(factor), int(mode == 'strict'),
            thr, max_keep, cap, _C.ptr(meta[1:2]), _C.ptr(ws), ws.numel(), _C.stream())
    k = int(meta[0])
    return out_box[:k], out_score[:k], out_label[:k]
