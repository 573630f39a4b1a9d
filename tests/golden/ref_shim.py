"""Import harness for the UNMODIFIED reference at /root/reference (authoring
container only -- the GPU box has no /root/reference and never imports this).

The reference's heads import ``mmcv.cnn`` / ``mmdet.ops.dcn`` at module import
time (lib/heads/*.py:3-8, lib/tester.py:4); neither package is installed, and
only a handful of names are touched, so they are stubbed in ``sys.modules``.
Nothing under /root/reference is edited.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("B200DET_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "lib"))


class AttrDict(dict):
    """Minimal stand-in for mmcv.Config nodes: attribute access + .get()."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v


def install():
    """Put the stubs in place and return the reference's ``lib`` package."""
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    if "mmcv" not in sys.modules:
        import torch

        mmcv = types.ModuleType("mmcv")
        cnn = types.ModuleType("mmcv.cnn")
        for n in ("normal_init", "xavier_init", "constant_init", "kaiming_init", "bias_init_with_prob"):
            setattr(cnn, n, lambda *a, **k: None)
        mmcv.cnn = cnn

        class ProgressBar:
            def __init__(self, *a, **k):
                pass

            def update(self):
                pass

        mmcv.ProgressBar = ProgressBar
        mmdet = types.ModuleType("mmdet")
        ops = types.ModuleType("mmdet.ops")
        dcn = types.ModuleType("mmdet.ops.dcn")

        class DeformConv(torch.nn.Module):
            pass

        dcn.DeformConv = DeformConv
        sys.modules.update({"mmcv": mmcv, "mmcv.cnn": cnn, "mmdet": mmdet, "mmdet.ops": ops,
                            "mmdet.ops.dcn": dcn})
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import lib  # noqa: F401  (the reference package)
    import lib.builder  # noqa: F401

    return sys.modules["lib"]
