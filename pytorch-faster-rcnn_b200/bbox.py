"""Proposal (RoI) targets with the signature of the reference's lib/bbox.py:6-82."""
import torch

from . import utils
from .registry import build_module


def _bbox_target_(props_bbox, gt_bbox, gt_label, assigner, sampler, target_means=None, target_stds=None):
    if isinstance(assigner, dict):
        assigner = build_module(assigner)
    if isinstance(sampler, dict):
        sampler = build_module(sampler)
    gt_bbox = gt_bbox.to(props_bbox.dtype)
    labels, ious = assigner(props_bbox, gt_bbox)
    n_gts = gt_label.numel()
    # GT is prepended AFTER assignment with labels 1..K and IoU 1 (lib/bbox.py:27-29)
    props_bbox = torch.cat([gt_bbox, props_bbox], dim=1)
    labels = torch.cat([torch.arange(1, n_gts + 1, dtype=labels.dtype, device=labels.device), labels])
    ious = torch.cat([ious.new_full((n_gts,), 1), ious])
    labels = sampler(labels, ious, props_bbox, gt_bbox)
    keep = torch.nonzero(labels >= 0).view(-1)
    kept = labels[keep]
    gt_idx = (kept - 1).clamp(min=0)
    tar_is_gt = (keep < n_gts).to(labels.dtype)
    tar_props = props_bbox[:, keep]
    tar_label = torch.where(kept > 0, gt_label[gt_idx], torch.zeros_like(kept))
    tar_bbox = gt_bbox[:, gt_idx]
    if target_means is not None and target_stds is not None:
        tar_param = utils.bbox2param(tar_props, tar_bbox, target_means, target_stds)
    else:
        tar_param = utils.bbox2param(tar_props, tar_bbox)
    return tar_props, tar_bbox, tar_label, tar_param, tar_is_gt


def bbox_target(props, gt_bbox, gt_label, assigner, sampler, target_means=None, target_stds=None):
    with torch.no_grad():
        return _bbox_target_(props, gt_bbox, gt_label, assigner, sampler, target_means, target_stds)
