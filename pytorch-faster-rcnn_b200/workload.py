"""Seeded synthetic inputs of SURVEY 8(d) (numpy RNG, so the bytes are identical on the
CPU oracle and the GPU path regardless of torch version)."""
import numpy as np

SEED = 2019                       # test/anchor_target_test.py:12-13
IMG_SHAPE = (800, 1333)
PAD_SHAPE = (800, 1344)
STRIDES = (4, 8, 16, 32, 64)


def fpn_grids(pad_shape=PAD_SHAPE, strides=STRIDES):
    return [(-(-pad_shape[0] // s), -(-pad_shape[1] // s)) for s in strides]


def synth_gt(rng, K, img_h, img_w):
    x1 = rng.uniform(0, 0.75 * img_w, K)
    y1 = rng.uniform(0, 0.75 * img_h, K)
    w = rng.uniform(20, 0.4 * img_w, K)
    h = rng.uniform(20, 0.4 * img_h, K)
    x2 = np.minimum(x1 + w, img_w - 1)
    y2 = np.minimum(y1 + h, img_h - 1)
    return np.stack([x1, y1, x2, y2]).astype(np.float32), rng.integers(1, 21, K).astype(np.int64)


def config2(B=8, K=8, seed=SEED, channels=256, img_shape=IMG_SHAPE, pad_shape=PAD_SHAPE, strides=STRIDES,
            num_anchors=3, with_feats=True):
    """faster_rcnn_r50_fpn train-path inputs: per level cls [B,A,H,W] ~ N(0,1), reg [B,4A,H,W]
    ~ N(0,0.5), FPN features [B,C,H,W] ~ N(0,1) (4 levels), K GT boxes per image."""
    rng = np.random.default_rng(seed)
    grids = fpn_grids(pad_shape, strides)
    out = dict(grids=grids, img_shape=img_shape, pad_shape=pad_shape, strides=strides)
    out["cls"] = [rng.standard_normal((B, num_anchors) + g, dtype=np.float32) for g in grids]
    out["reg"] = [(0.5 * rng.standard_normal((B, 4 * num_anchors) + g, dtype=np.float32)) for g in grids]
    gts = [synth_gt(rng, K, *img_shape) for _ in range(B)]
    out["gt"] = np.stack([g[0] for g in gts])            # [B,4,K]
    out["gt_label"] = np.stack([g[1] for g in gts])      # [B,K]
    if with_feats:
        out["feats"] = [rng.standard_normal((B, channels) + g, dtype=np.float32) for g in grids[:4]]
    return out
