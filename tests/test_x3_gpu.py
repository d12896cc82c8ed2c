"""The fp32-accurate tensor-core mode (precision 'fp32x3', KGAN_PREC_TF32X3): operands are split hi + lo inside the TMA-fed kernels and
every product is three tcgen05 kind::tf32 MMAs (lo*hi + hi*lo + hi*hi) accumulated in fp32; activations stay full fp32 in HBM.
Checked against the float64 statement of the descriptor semantics on ARBITRARY fp32 inputs (nothing pre-rounded).
Stated tolerance: rel-L2 <= 5e-6 per kernel.  What sets it (measured, profiles/r2_x3_accuracy.txt): the split itself costs ~3e-7 (the
dropped lo*lo term, the tensor core's truncation of lo), the rest is the tensor core's fp32 accumulation, which truncates - the error of a
chain of K steps grows linearly, ~2.4e-9 per contraction element: 1e-6 at K = 384, 3.7e-6 at K = 1536 (the deepest layers); the weight
gradient's chains are cut at 32 K tiles (<= 2.8e-6 at any batch).  The exact FMA kernels sit at 2e-7 .. 2e-6 on the same cases.  Through
the whole network the mode meets the fp32 path's 1e-5 (tests/test_parity_gpu.py runs in both modes)."""
from importlib import import_module

import numpy as np
import pytest
import torch

import emu_backend as emu
import kgan_b200 as kgan
from test_tf32_gpu import GEOMS, TMA_EXPECTED, WGRAD_TMA_EXPECTED, rel

pytestmark = pytest.mark.gpu
ops, G = kgan.ops, kgan.geometry
TOL = 5e-6


@pytest.fixture(autouse=True)
def x3_path():
    kgan.set_precision("fp32x3")
    yield
    kgan.set_precision("fp32")


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


def _geom(name):
    kw, n = GEOMS[name]
    kw = dict(kw)
    if kw.pop("unfold", False):
        geom = G.UnfoldedTcnGeom(**kw)
        x = rnd(n, geom.c_in, geom.kt * geom.t_out, geom.v_in, seed=1)
    else:
        geom = G.TapConvGeom(**kw)
        x = rnd(n, geom.K * geom.c_in, geom.t_in, geom.v_in, seed=1)
    return geom, kw, n, x


@pytest.mark.parametrize("name", [k for k in GEOMS if TMA_EXPECTED.get(k, (0, 0))[0] or WGRAD_TMA_EXPECTED.get(k, 0)])
def test_tapconv_x3(name):
    geom, kw, n, x = _geom(name)
    lib = import_module("kinetic-gan_b200._lib").lib()
    w = rnd(geom.K * geom.c_out, kw.get("w_cin", geom.c_in), geom.kt, 1, seed=2) / np.sqrt(geom.c_in * geom.kt * geom.K)
    bias = rnd(geom.c_out, seed=3)
    add = rnd(n, geom.c_out, geom.t_out, geom.v_out, seed=4)
    go = rnd(n, geom.c_out, geom.t_out, geom.v_out, seed=5)
    xc, wc = x.cuda(), w.cuda()
    prof = ops.profile_start()
    try:
        fwd = ops.tapconv_fwd(xc, wc, geom.fwd)
        fwd2 = ops.tapconv_fwd(xc, wc, geom.fwd, bias.cuda(), add.cuda(), ops.ACT_LRELU)
        dg = ops.tapconv_fwd(go.cuda(), wc, geom.dgrad)
        dw = ops.tapconv_wgrad(xc, go.cuda(), geom.fwd, tuple(w.shape))
        acc = torch.ones(tuple(w.shape), device="cuda")
        ops.tapconv_wgrad(xc, go.cuda(), geom.fwd, tuple(w.shape), out=acc)
    finally:
        ops.profile_stop(prof)
    fams = [p[0] for p in prof]
    # the split kernels are the ones that ran wherever the tf32 mode runs the TMA-fed ones
    if TMA_EXPECTED.get(name, (0, 0))[0]:
        assert lib.kgan_tapconv_tf32_workspace(geom.fwd.cstruct(n, 0, 2)) > lib.kgan_tapconv_tf32_workspace(geom.fwd.cstruct(n, 0, 1))      # hi + lo images
        assert fams.count("tapconv_fwd_x3") == 3, fams
    if WGRAD_TMA_EXPECTED.get(name, 0):
        assert fams.count("tapconv_wgrad_x3") == 2, fams
    errs = {
        "fwd": rel(fwd, emu.tapconv_fwd(x.double(), w.double(), geom.fwd)),
        "fwd+epilogue": rel(fwd2, emu.tapconv_fwd(x.double(), w.double(), geom.fwd, bias.double(), add.double(), ops.ACT_LRELU)),
        "dgrad": rel(dg, emu.tapconv_fwd(go.double(), w.double(), geom.dgrad)),
        "wgrad": rel(dw, emu.tapconv_wgrad(x.double(), go.double(), geom.fwd, tuple(w.shape))),
        "wgrad accumulate": rel(acc - 1.0, emu.tapconv_wgrad(x.double(), go.double(), geom.fwd, tuple(w.shape))),
    }
    print("x3 %-22s " % name + "  ".join("%s %.1e" % kv for kv in errs.items()))
    assert max(errs.values()) < TOL, errs
    # full fp32 outputs: nothing is stored tf32-rounded in this mode
    assert (fwd.view(torch.int32) & 0x1FFF).ne(0).any()


def test_x3_streamed_and_resident_weights_large_batch():
    """The two weight-staging variants (resident image / streamed hi + lo stages) and many tiles per CTA, at a batch where both occur."""
    for name, n in (("d1_tcn_v12", 96), ("d2_gcn", 64), ("d4_gcn_p80_big", 400), ("mlp_632", 2048)):
        kw = dict(GEOMS[name][0])
        geom = G.TapConvGeom(**kw)
        x = rnd(n, geom.K * geom.c_in, geom.t_in, geom.v_in, seed=7)
        w = rnd(geom.K * geom.c_out, geom.c_in, geom.kt, 1, seed=8) / np.sqrt(geom.c_in * geom.kt * geom.K)
        go = rnd(n, geom.c_out, geom.t_out, geom.v_out, seed=9)
        got = ops.tapconv_fwd(x.cuda(), w.cuda(), geom.fwd)
        dw = ops.tapconv_wgrad(x.cuda(), go.cuda(), geom.fwd, tuple(w.shape))
        kgan.set_precision("fp32")
        exact = ops.tapconv_fwd(x.cuda(), w.cuda(), geom.fwd)
        dwe = ops.tapconv_wgrad(x.cuda(), go.cuda(), geom.fwd, tuple(w.shape))
        kgan.set_precision("fp32x3")
        e1, e2 = rel(got, exact.cpu()), rel(dw, dwe.cpu())
        print("x3 vs exact FMA kernel, %s n=%d: fwd %.1e wgrad %.1e" % (name, n, e1, e2))
        assert e1 < TOL and e2 < TOL, (name, e1, e2)
