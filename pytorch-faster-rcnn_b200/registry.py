"""Name -> class registry for the hot-path components, mirroring the entries of the
reference's lib.builder.MODULES that belong to the path (lib/builder.py:22-37)."""
import copy

MODULES = {}


def register(cls):
    MODULES[cls.__name__] = cls
    return cls


def build_module(cfg, *args, **kwargs):
    cfg = copy.copy(cfg)
    assert 'type' in cfg
    m_type = cfg.pop('type')
    if m_type not in MODULES:
        raise ValueError("'{}' is not registered".format(m_type))
    return MODULES[m_type](*args, **cfg, **kwargs)
