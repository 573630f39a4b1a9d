"""Seeded inputs that are too large to store as fixtures: regenerated identically (numpy Generator streams) by
tests/golden/make_golden.py in the authoring container and by the tests on the GPU box."""
import numpy as np

SEED = 2019
RETINA_STRIDES = (8, 16, 32, 64, 128)
RETINA_GRIDS = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]          # 800 x 1344 pad, BASELINE config 4
RETINA_SCALES = [4 * 2 ** (i / 3) for i in range(3)]                         # octave_base_scale 4, 3 scales per octave


def retina_inputs(num_cls=20):
    """RetinaNet head outputs at config-4 sizes: per level cls [A*C, H, W] ~ N(-2, 2), reg [4A, H, W] ~ N(0, 0.3), A = 9."""
    rng = np.random.default_rng(SEED + 41)
    cls = [rng.normal(-2, 2, (9 * num_cls,) + g).astype(np.float32) for g in RETINA_GRIDS]
    reg = [rng.normal(0, 0.3, (36,) + g).astype(np.float32) for g in RETINA_GRIDS]
    return cls, reg
