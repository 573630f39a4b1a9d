"""The reference's OWN call sequence for the train hot path, on the drop-in's reference-signature functions.

The reference's head / detector classes cannot travel to the GPU box (they need mmcv / mmdet), so this module restates
the per-image Python loops in which they call the path -- and nothing else -- to measure and test what a reference
user gets after `install()`:

  CascadeRCNN.forward_train            lib/detectors/cascade_rcnn.py:106-131
    rpn_head.loss(...)                 lib/heads/anchor_head.py:152-199  -> create_anchors, per image
        single_image_targets           lib/heads/anchor_head.py:69-111   -> inside masks, anchor_target
    rpn_head.predict_bboxes_from_output  lib/heads/anchor_head.py:268-289 -> create_anchors, per image
        RPNHead.predict_single_image   lib/heads/rpn_head.py:68-120
    rcnn_head.bbox_targets             lib/heads/bbox_head.py:47-52     -> multi_apply(bbox_target)
    roi_extractor(feats, tar_props)    lib/region.py:301-306

Every call below goes through anchor.py / region.py / bbox.py / heads.py / utils.py exactly as the rebound names would.
"""
import types

import torch

from . import anchor, batched, bbox, heads, region, utils


class _Cfg(dict):
    """Attribute + .get access, like the reference's mmcv ConfigDict."""
    __getattr__ = dict.__getitem__


class TrainCallSequence(object):
    def __init__(self, strides, device, rpn_proposal=None, rpn_assigner=None, rpn_sampler=None, rcnn_assigner=None,
                 rcnn_sampler=None, rpn_stds=(1.0, 1.0, 1.0, 1.0), rcnn_stds=(0.1, 0.1, 0.2, 0.2), allowed_border=0,
                 scales=(8,), ratios=(0.5, 1.0, 2.0), sampler_rng="device"):
        self.strides, self.device, self.border = list(strides), device, allowed_border
        z4 = [0.0, 0.0, 0.0, 0.0]
        self.creators = [anchor.AnchorCreator(base=s, scales=list(scales), aspect_ratios=list(ratios)) for s in strides]
        for c in self.creators:
            c.to(device)
        self.num_anchors = len(scales) * len(ratios)
        self.head = types.SimpleNamespace(anchor_strides=self.strides, anchor_scales=list(scales), anchor_ratios=list(ratios),
                                          target_means=z4, target_stds=list(rpn_stds), use_sigmoid=True, cls_channels=1,
                                          anchor_creators=self.creators)
        self.rpn_proposal = rpn_proposal or dict(pre_nms=2000, post_nms=2000, max_num=2000, nms_iou=0.7, min_bbox_size=0)
        a = rpn_assigner or dict(pos_iou=0.7, neg_iou=0.3, min_pos_iou=0.3)
        s = rpn_sampler or dict(max_num=256, pos_num=128)
        self.rpn_assigner = region.MaxIoUAssigner(a["pos_iou"], a["neg_iou"], a["min_pos_iou"])
        self.rpn_sampler = region.RandomSampler(s["max_num"], s["pos_num"], rng=sampler_rng)
        a = rcnn_assigner or dict(pos_iou=0.5, neg_iou=0.5, min_pos_iou=0.5)
        s = rcnn_sampler or dict(max_num=512, pos_num=128)
        self.rcnn_assigner = region.MaxIoUAssigner(a["pos_iou"], a["neg_iou"], a["min_pos_iou"])
        self.rcnn_sampler = region.RandomSampler(s["max_num"], s["pos_num"], rng=sampler_rng)
        self.rpn_ms, self.rcnn_ms = (z4, list(rpn_stds)), (z4, list(rcnn_stds))
        self.rcnn_head = types.SimpleNamespace(target_means=z4, target_stds=list(rcnn_stds))
        self.rpn_train_cfg = _Cfg(assigner=self.rpn_assigner, sampler=self.rpn_sampler, allowed_border=allowed_border)
        self.rcnn_train_cfg = _Cfg(assigner=self.rcnn_assigner, sampler=self.rcnn_sampler)
        self.extractor = region.BasicRoIExtractor(
            [dict(type="RoIAlign", spatial_scale=1.0 / st, sampling_ratio=2) for st in self.strides[:4]], output_size=(7, 7))

    def create_anchors(self, grid_sizes):
        return [ac(self.strides[i], grid_sizes[i]) for i, ac in enumerate(self.creators)]

    def rpn_targets_single_image(self, level_cls_outs, level_reg_outs, gt_bbox, level_anchors, grid_sizes, img_meta):
        """AnchorHead.single_image_targets (lib/heads/anchor_head.py:69-111), RPN form (gt_label None)."""
        cls_out = torch.cat([x.view(1, -1) for x in level_cls_outs], dim=1)
        reg_out = torch.cat([x.view(4, -1) for x in level_reg_outs], dim=1)
        anchors = torch.cat([a.view(4, -1) for a in level_anchors], dim=1)
        img_size = img_meta['img_shape'][:2]
        in_img = region.inside_anchor_mask(anchors, img_size, self.border)
        in_grid = torch.cat([region.inside_grid_mask(self.num_anchors, img_size, grid_sizes[l], st, self.device)
                             for l, st in enumerate(self.strides)])
        in_mask = in_img & in_grid.bool()
        in_anchors = anchors[:, in_mask]
        return anchor.anchor_target(cls_out, reg_out, 1, in_anchors, in_mask, gt_bbox, None, self.rpn_assigner,
                                    self.rpn_sampler, *self.rpn_ms)

    def step(self, cls_outs, reg_outs, feats, gt_bboxes, gt_labels, img_metas):
        """cls_outs[l] [B,A,H,W], reg_outs[l] [B,4A,H,W], feats[l] [B,C,H,W] (as the neck hands them over), gt_bboxes:
        list of [4,K], gt_labels: list of int64 [K].  Returns what the heads would pass on."""
        grid_sizes = [tuple(int(v) for v in c.shape[-2:]) for c in cls_outs]
        B = len(img_metas)
        # rpn_head.loss: targets
        level_anchors = self.create_anchors(grid_sizes)
        rpn_tars = []
        for i in range(B):
            rpn_tars.append(self.rpn_targets_single_image([c[i] for c in cls_outs], [r[i] for r in reg_outs], gt_bboxes[i],
                                                          level_anchors, grid_sizes, img_metas[i]))
        # rpn_head.predict_bboxes_from_output (the reference regenerates the anchors here)
        level_anchors = self.create_anchors(grid_sizes)
        props = []
        for i in range(B):
            b, _, _ = heads.rpn_predict_single_image(self.head, [c[i] for c in cls_outs], [r[i] for r in reg_outs],
                                                     level_anchors, img_metas[i], self.rpn_proposal)
            props.append(b)
        # rcnn_head.bbox_targets
        tars = utils.multi_apply(bbox.bbox_target, props, list(gt_bboxes), list(gt_labels), self.rcnn_assigner,
                                 self.rcnn_sampler, tuple(self.rcnn_ms[0]), tuple(self.rcnn_ms[1]))   # tuples: not per-image lists
        tar_props = [t[0] for t in tars]
        # roi_extractor
        roi_outs = self.extractor(list(feats), tar_props)
        return dict(rpn_targets=rpn_tars, props=props, rcnn_targets=tars, roi_feats=roi_outs)


class BatchedCallSequence(TrainCallSequence):
    """The same forward_train sequence with the three loops rebound at the METHOD level, as install() does
    (batched.py): rpn_head.loss's targets, rpn_head.predict_bboxes_from_output and rcnn_head.bbox_targets are one
    batched pass each; the RoI extractor call is unchanged."""

    def step(self, cls_outs, reg_outs, feats, gt_bboxes, gt_labels, img_metas):
        rpn_gt_labels = [torch.full_like(l, 1) for l in gt_labels]                       # cascade_rcnn.py:109
        rpn_tars = batched.anchor_head_targets(self.head, cls_outs, reg_outs, list(gt_bboxes), rpn_gt_labels, img_metas,
                                               self.rpn_train_cfg)
        assert rpn_tars is not None, "batched anchor targets: call not covered"
        props = batched.rpn_predict_bboxes_from_output(self.head, cls_outs, reg_outs, img_metas, _Cfg(self.rpn_proposal))[0]
        tars = batched.bbox_head_bbox_targets(self.rcnn_head, props, list(gt_bboxes), list(gt_labels), self.rcnn_train_cfg)
        roi_outs = self.extractor(list(feats), tars[0])
        return dict(rpn_targets=rpn_tars, props=props, rcnn_targets=tars, roi_feats=roi_outs)
