"""Import shim for the UNMODIFIED reference at /root/reference (test infrastructure only).

Used in the build container to (a) validate the oracle restatement and (b) generate the
golden fixtures under tests/golden/.  Never used by the product path, the `-m gpu` tests,
smoke() or bench.py: /root/reference does not exist on the GPU box.

What the shim neutralises:
  * `import matplotlib.pyplot` at models/init_gan/graph_ntu.py:3, graph_h36m.py:3 (unused, not installed)
  * `.cuda()` on the adjacency tensors, models/generator.py:47, models/discriminator.py:19
  * `device='cuda:0'` of the per-block noise, models/generator.py:179
  * `torch.cuda.FloatTensor` of the truncation latents, models/generator.py:98
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("KGAN_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def load():
    """Returns (generator_module, discriminator_module) of the untouched reference."""
    import torch

    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        if not getattr(torch.randn, "_kgan_shim", False):
            _randn = torch.randn

            def randn(*a, **k):
                k.pop("device", None)
                return _randn(*a, **k)

            randn._kgan_shim = True
            torch.randn = randn
        # torch.cuda.FloatTensor(...) in Generator.truncate, models/generator.py:98
        if torch.cuda.FloatTensor not in (torch.FloatTensor, torch.DoubleTensor):
            torch.cuda.FloatTensor = torch.FloatTensor
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference packages are called `models`; keep them out of the way of anything else
    import importlib

    gen = importlib.import_module("models.generator")
    dis = importlib.import_module("models.discriminator")
    return gen, dis


def set_float_type(dtype):
    """fp64 ground-truth runs: make the truncation latents (models/generator.py:98) follow the module dtype."""
    import torch

    if not torch.cuda.is_available():
        torch.cuda.FloatTensor = torch.DoubleTensor if dtype == torch.float64 else torch.FloatTensor
