"""Anchor grid + anchor targets with the signatures of the reference's lib/anchor.py
(AnchorCreator :80-129, anchor_target :11-76), on sm_100a kernels."""
import numpy as np
import torch

from . import _C
from . import utils
from .registry import build_module


class AnchorCreator(object):
    """lib/anchor.py:80-129.  __call__(stride, grid) -> fp32 [4, A, H, W]; anchor index
    a = scale_idx * len(ratios) + ratio_idx; sizes computed in float64 then cast (:92-99)."""

    def __init__(self, base=16, scales=[8, 16, 32], aspect_ratios=[0.5, 1.0, 2.0], center_lt=False,
                 device=None):
        self.device = torch.device("cuda") if device is None else torch.device(device)
        self.center_lt = center_lt
        self.base = base
        self.scales = scales
        self.aspect_ratios = aspect_ratios
        self.num_anchors = len(scales) * len(aspect_ratios)
        ws, hs = [], []
        for s in scales:
            for ar in aspect_ratios:
                ws.append(base * s * np.sqrt(ar))
                hs.append(base * s / np.sqrt(ar))
        self._ws = np.asarray(ws, dtype=np.float32)
        self._hs = np.asarray(hs, dtype=np.float32)
        self.anchor_ws = torch.from_numpy(self._ws.copy())
        self.anchor_hs = torch.from_numpy(self._hs.copy())

    def to(self, device):
        device = torch.device(device)
        if self.device == device:
            return True
        self.device = device
        self.anchor_ws = self.anchor_ws.to(device)
        self.anchor_hs = self.anchor_hs.to(device)

    def level(self, stride, grid):
        """Closed-form description of this level for the fused kernels (no tensor)."""
        return dict(stride=float(stride), H=int(grid[0]), W=int(grid[1]), ws=self._ws.tolist(),
                    hs=self._hs.tolist(), center_lt=bool(self.center_lt))

    def __call__(self, stride, grid):
        dev = self.device if self.device.type == "cuda" else torch.device("cuda")
        H, W = int(grid[0]), int(grid[1])
        A = self.num_anchors
        out = torch.empty((4, A, H, W), dtype=torch.float32, device=dev)
        ws = (_C.c_float * A)(*self._ws.tolist())
        hs = (_C.c_float * A)(*self._hs.tolist())
        _C.call("b2d_anchor_grid", _C.ptr(out), ws, hs, A, H, W, float(stride), int(bool(self.center_lt)),
                _C.stream())
        return out


def anchor_target(cls_out, reg_out, cls_channels, in_anchors, in_mask, gt_bbox, gt_label=None,
                  assigner=None, sampler=None, target_means=None, target_stds=None):
    """lib/anchor.py:11-76.  Returns (tar_cls_out, tar_reg_out, tar_labels, tar_anchors,
    tar_bbox, tar_param); the two head-output gathers stay differentiable."""
    assert assigner is not None
    if isinstance(assigner, dict):
        assigner = build_module(assigner)
    labels, _ = assigner(in_anchors, gt_bbox)
    if sampler is not None:
        if isinstance(sampler, dict):
            sampler = build_module(sampler)
        labels = sampler(labels)
    keep = torch.nonzero(labels >= 0).view(-1)                 # ascending, like boolean masking
    kept = labels[keep]
    gt_idx = (kept - 1).clamp(min=0)                           # negatives point at GT 0 (:45-47)
    inside_arg = torch.nonzero(in_mask).view(-1)
    chosen = inside_arg[keep]
    cls_out_ = cls_out.view(cls_channels, -1)
    reg_out_ = reg_out.view(4, -1)
    assert cls_out_.shape[-1] == reg_out_.shape[-1]
    tar_cls_out = cls_out_[:, chosen]
    tar_reg_out = reg_out_[:, chosen]
    if gt_label is None:
        tar_labels = (kept > 0).to(labels.dtype)
    else:
        tar_labels = gt_label[gt_idx]
        tar_labels = torch.where(kept > 0, tar_labels, torch.zeros_like(tar_labels))
    tar_anchors = in_anchors[:, keep]
    tar_bbox = gt_bbox.to(torch.float32)[:, gt_idx]
    if target_means is not None and target_stds is not None:
        # the second normalisation of lib/anchor.py:70-73 folded into the kernel: (t - 0)/1 is exact
        tar_param = utils.bbox2param(tar_anchors, tar_bbox, target_means, target_stds)
    else:
        tar_param = utils.bbox2param(tar_anchors, tar_bbox)
    return tar_cls_out, tar_reg_out, tar_labels, tar_anchors, tar_bbox, tar_param
