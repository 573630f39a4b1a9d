"""Host-side check of the pruning bound behind the x-sweep NMS mask kernel (csrc/nms.cu k_nms_sweep).

The kernel only tests box j against box i (x1_i <= x1_j) when x1_j <= x1_i + (1 - 0.9 thr) w_i + 1e-4 (|x2_i| + 1),
computed in fp32.  The claim is that every pair torchvision's NMS decision accepts -- fp32 `inter / union > thr`,
areas without +1 (torchvision/csrc/ops/cpu/nms_kernel.cpp, called from lib/heads/rpn_head.py:103) -- lies inside that
window and overlaps in y.  Checked here in numpy fp32 on random, near-duplicate, sliver and large-coordinate boxes."""
import numpy as np
import pytest

f32 = np.float32


def _decision(a, b, thr):
    """fp32 restatement of the pair decision (same operation order as csrc/nms.cu suppresses())."""
    xx1, yy1 = np.maximum(a[:, 0], b[:, 0]), np.maximum(a[:, 1], b[:, 1])
    xx2, yy2 = np.minimum(a[:, 2], b[:, 2]), np.minimum(a[:, 3], b[:, 3])
    w, h = np.maximum(f32(0), xx2 - xx1), np.maximum(f32(0), yy2 - yy1)
    inter = w * h
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    u = (aa + ab) - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / u) > f32(thr)


def _boxes(rng, n, kind):
    if kind == "random":
        c = rng.uniform(0, 1333, (n, 2)); s = np.exp(rng.uniform(np.log(0.5), np.log(800), (n, 2)))
    elif kind == "sliver":          # widths down to 1e-4 px: the absolute slack of the window matters
        c = rng.uniform(0, 1333, (n, 2)); s = np.stack([np.exp(rng.uniform(np.log(1e-4), np.log(2.0), n)),
                                                        np.exp(rng.uniform(np.log(1.0), np.log(300), n))], 1)
    else:                           # "far": coordinates around 1e6, where an fp32 ulp is 0.06 px
        c = rng.uniform(1e6, 1e6 + 2000, (n, 2)); s = np.exp(rng.uniform(np.log(4.0), np.log(600), (n, 2)))
    b = np.concatenate([c - s / 2, c + s / 2], 1).astype(f32)
    return b


@pytest.mark.parametrize("thr", [0.05, 0.3, 0.5, 0.7, 0.9, 0.999])
@pytest.mark.parametrize("kind", ["random", "sliver", "far"])
def test_accepted_pairs_lie_inside_the_sweep_window(thr, kind):
    rng = np.random.default_rng(int(thr * 1000) + len(kind))
    n = 200_000
    a = _boxes(rng, n, kind)
    # partner: a jittered copy (IoU spread over the whole (0, 1] range, many pairs close to the threshold)
    scale = rng.choice([1e-4, 1e-3, 1e-2, 0.05, 0.2, 0.5], (n, 1)) * np.maximum(a[:, 2:] - a[:, :2], 1e-6).repeat(2, 1)[:, [0, 2, 1, 3]][:, :4]
    b = (a + rng.normal(0, 1, (n, 4)) * scale).astype(f32)
    ok = (b[:, 2] >= b[:, 0]) & (b[:, 3] >= b[:, 1])
    a, b = a[ok], b[ok]
    # i = the box with the smaller x1 (ties: either order satisfies the window trivially)
    swap = b[:, 0] < a[:, 0]
    bi = np.where(swap[:, None], b, a); bj = np.where(swap[:, None], a, b)
    acc = _decision(bi, bj, thr)
    assert acc.sum() > 1000                                  # the sample does exercise accepted pairs
    wi = bi[:, 2] - bi[:, 0]
    prune = f32(1.0) - f32(0.9) * f32(thr)
    xhi = bi[:, 0] + prune * wi + f32(1.0e-4) * (np.abs(bi[:, 2]) + f32(1.0))
    inside = (bj[:, 0] <= xhi) & (wi > 0) & ((bi[:, 3] - bi[:, 1]) > 0) & (bj[:, 1] < bi[:, 3]) & (bi[:, 1] < bj[:, 3])
    assert not np.any(acc & ~inside), int(np.sum(acc & ~inside))
    # the decision is symmetric in the two boxes (the kernel evaluates every unordered pair once)
    assert np.array_equal(acc, _decision(bj, bi, thr))
