"""Import alias: `import b200det` loads the package directory `pytorch-faster-rcnn_b200/`
(whose name is not a valid Python identifier) and registers it, and all of its
sub-modules, under the name `b200det`."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_REAL = "pytorch-faster-rcnn_b200"
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name == _REAL or _name.startswith(_REAL + "."):
        sys.modules["b200det" + _name[len(_REAL):]] = _mod
sys.modules["b200det"] = _pkg
