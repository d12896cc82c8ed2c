"""The C-ABI library must load without a GPU and export every symbol include/kgan.h declares."""
import os
import re
from importlib import import_module

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_build_and_exports():
    import __graft_entry__ as ge

    ge.build()
    _lib = import_module("kinetic-gan_b200._lib")
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "kgan.h")).read()
    declared = set(re.findall(r"\b(kgan_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.kgan_version() >= 100


def test_desc_struct_matches_header():
    """Field order of the ctypes mirror == field order of `kgan_tapconv_desc` in the header."""
    _lib = import_module("kinetic-gan_b200._lib")
    header = open(os.path.join(ROOT, "include", "kgan.h")).read()
    body = header[header.index("typedef struct kgan_tapconv_desc"):header.index("} kgan_tapconv_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(",") if "[" not in decl else [decl]:
            names.append(re.sub(r"\[.*\]", "", part.strip().split()[-1]))
    assert names == [f[0] for f in _lib.TapConvDesc._fields_]


def test_no_cpu_fallback():
    """Without a CUDA device every operator must raise, never compute on the CPU."""
    import torch

    import kgan_b200 as kgan

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises((RuntimeError, AssertionError)):
        kgan.ops.adjmix_fwd(torch.zeros(1, 1, 1, 1), torch.zeros(1, 1, 1))
    D = kgan.Discriminator(3, 4, 16, 512)
    with pytest.raises((RuntimeError, AssertionError)):
        D(torch.zeros(1, 3, 16, 25), torch.zeros(1, dtype=torch.long))
