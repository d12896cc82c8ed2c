"""Importable alias of the `kinetic-gan_b200/` package (its directory name is not a Python identifier):
`import kgan_b200`, `from kgan_b200.models.generator import Generator`, ... resolve to the very same module objects."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_REAL = "kinetic-gan_b200"
_pkg = importlib.import_module(_REAL)
for _sub in ("wgan_gp", "ddp", "generate", "feeder", "train", "evaluation"):
    importlib.import_module(_REAL + "." + _sub)
for _name, _mod in list(sys.modules.items()):
    if _name == _REAL or _name.startswith(_REAL + "."):
        sys.modules[__name__ + _name[len(_REAL):]] = _mod
