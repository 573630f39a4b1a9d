"""b200det -- B200-native (sm_100a) detection post-backbone hot path: a drop-in for the
box / anchor / region ops of pengfeidip/pytorch-faster-rcnn (lib/anchor.py, lib/bbox.py,
lib/region.py and the box part of lib/utils.py).

The directory is named `pytorch-faster-rcnn_b200` (not an importable identifier);
import it as `b200det` through the alias module at the repository root, or with
importlib.import_module("pytorch-faster-rcnn_b200").
"""
from . import _C  # noqa: F401
from . import registry
from . import utils, region, anchor, bbox, heads, fused, workload, dropin, dist, refpath  # noqa: F401
from .anchor import AnchorCreator, anchor_target  # noqa: F401
from .bbox import bbox_target  # noqa: F401
from .region import (MaxIoUAssigner, RandomSampler, IoUBalancedNegSampler, BasicRoIExtractor,  # noqa: F401
                     SingleRoIExtractor, RoIAlign, RoIPool, ScalableRoIPool, ScalableRoIAlign, ProposalCreator,
                     inside_grid_mask, inside_anchor_mask)
from .dropin import install, uninstall  # noqa: F401

for _cls in (MaxIoUAssigner, RandomSampler, IoUBalancedNegSampler, BasicRoIExtractor, SingleRoIExtractor, RoIAlign,
             RoIPool, ScalableRoIPool, ScalableRoIAlign):
    registry.register(_cls)

__version__ = "0.1.0"
