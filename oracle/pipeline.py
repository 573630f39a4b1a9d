"""CPU ORACLE (test / baseline infrastructure): the config-2 train path for ONE image,
composed from the C restatement (oracle/det_oracle.c) exactly as the reference's
CascadeRCNN.forward_train wires it (lib/detectors/cascade_rcnn.py:106-131), minus
convolutions and losses.  Used by bench.py's cpu_baseline / --impl reference legs and
by tests; never by the product."""
import numpy as np

import oracle
from oracle import sampler_spec

RPN_PROPOSAL = dict(pre_nms=2000, post_nms=2000, max_num=2000, nms_iou=0.7, min_bbox_size=0)


class ImagePath(object):
    def __init__(self, grids, strides, img_shape, roi_sampler=(512, 128), rpn_sampler=(256, 128),
                 rpn_proposal=RPN_PROPOSAL):
        self.grids, self.strides, self.img = grids, strides, tuple(img_shape[:2])
        self.anchors4 = [oracle.anchor_grid(s, g, scales=[8]) for s, g in zip(strides, grids)]   # a1 (shared by images)
        self.anchors = [a.reshape(4, -1) for a in self.anchors4]
        self.mask = np.concatenate([oracle.valid_mask(a, self.img, g, s, 0)
                                    for a, g, s in zip(self.anchors4, grids, strides)])        # a2
        self.all_anchors = np.concatenate(self.anchors, 1)
        self.in_anchors = np.ascontiguousarray(self.all_anchors[:, self.mask])
        self.in_index = np.nonzero(self.mask)[0]
        self.roi_sampler, self.rpn_sampler, self.rpn_proposal = roi_sampler, rpn_sampler, rpn_proposal

    def run(self, cls, reg, feats, gt, gt_label, seed=0):
        """cls[l]: [A,H,W], reg[l]: [4A,H,W], feats[l]: [C,H,W] (4 levels), gt [4,K]."""
        # a9: proposals
        props, scores, _, _ = oracle.rpn_proposals([c.reshape(-1) for c in cls], [r.reshape(4, -1) for r in reg],
                                                   self.anchors, self.rpn_proposal, [0, 0, 0, 0], [1, 1, 1, 1], self.img)
        # a4-a7: RPN targets
        lab, _ = oracle.assign_max_iou(self.in_anchors, gt, 0.7, 0.3, 0.3)
        full = np.full(self.all_anchors.shape[1], -1, np.int64)
        full[self.mask] = lab
        ch = sampler_spec.sample(full, self.rpn_sampler[0], self.rpn_sampler[1], seed)
        gi = np.maximum(full[ch] - 1, 0)
        rpn_param = oracle.bbox2param(np.ascontiguousarray(self.all_anchors[:, ch]), np.ascontiguousarray(gt[:, gi]))
        flat_cls = np.concatenate([c.reshape(-1) for c in cls])
        rpn_tar_cls = flat_cls[ch]
        # a13: RoI targets (assign on proposals, GT prepended afterwards)
        plab, _ = oracle.assign_max_iou(props, gt, 0.5, 0.5, 0.5)
        K = gt.shape[1]
        cand = np.concatenate([gt, props], 1)
        clab = np.concatenate([np.arange(1, K + 1), plab])
        rc = sampler_spec.sample(clab, self.roi_sampler[0], self.roi_sampler[1], seed + 1)
        rois = np.ascontiguousarray(cand[:, rc])
        gj = np.maximum(clab[rc] - 1, 0)
        roi_param = oracle.bbox2param(rois, np.ascontiguousarray(gt[:, gj]), [0, 0, 0, 0], [0.1, 0.1, 0.2, 0.2])
        roi_label = np.where(clab[rc] > 0, gt_label[gj], 0)
        # a14: RoIAlign
        roi_feats = oracle.roi_extract(feats, rois, strides=self.strides[:4])
        return dict(props=props, scores=scores, rpn_labels=full, rpn_chosen=ch, rpn_param=rpn_param,
                    rpn_tar_cls=rpn_tar_cls, roi_labels=clab, roi_chosen=rc, rois=rois, roi_param=roi_param,
                    roi_label=roi_label, roi_feats=roi_feats)
