"""Operand-building tensor-core tap convolution (csrc/tapconv_build.cu): raw tiles staged by TMA, tap operands gathered in shared
memory through the position map.  Every shape class of the two networks - aligned and unaligned shifts, stride-2 frame selection,
joint selection, small planes with several samples per tile, ragged channel counts, two staged boxes per tile - against the fp64
statement of the descriptor semantics (tests/emu_backend.py), forward and data gradient, tf32 tolerance 1e-3; and exactly on
tf32-representable data."""
from importlib import import_module

import numpy as np
import pytest
import torch

import emu_backend as emu
import kgan_b200 as kgan

pytestmark = pytest.mark.gpu
ops, G = kgan.ops, kgan.geometry
TOL = 1e-3


@pytest.fixture(autouse=True)
def staged_everywhere():
    kgan.set_precision("tf32")
    old = G.STAGED_POLICY
    G.STAGED_POLICY = "all"
    yield
    G.STAGED_POLICY = old
    kgan.set_precision("fp32")


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32)


def rel(a, b):
    b = b.double()
    return ((a.detach().cpu().double() - b).norm() / b.norm().clamp_min(1e-30)).item()


# name: (geometry kwargs, batch, forward staged?, data gradient staged?)
GEOMS = {
    "d0_tcn_v12": (dict(c_in=32, c_out=32, t_in=64, v_in=12, kt=3, pad=1), 5, 1, 1),                           # aligned shifts, 6 tiles per plane
    "d1_tcn_v11": (dict(c_in=64, c_out=64, t_in=64, v_in=11, kt=3, pad=1), 4, 1, 1),                           # unaligned shifts (+-11), ragged last tile
    "d1_gcn_v11": (dict(c_in=32, c_out=64, t_in=64, v_in=11, K=3), 4, 1, 1),                                   # three channel blocks
    "d2_tcn_stride2_v5": (dict(c_in=128, c_out=128, t_in=64, v_in=5, kt=3, pad=1, t_sel=list(range(0, 64, 2))), 5, 1, 1),      # 2 boxes per tile
    "d2_tcn_select": (dict(c_in=128, c_out=128, t_in=64, v_in=11, kt=3, pad=1, t_sel=list(range(0, 64, 2)), v_keep=[2, 4, 6, 8, 10]), 5, 0, 1),
    "d3_tcn_v5": (dict(c_in=256, c_out=256, t_in=32, v_in=5, kt=3, pad=1, t_sel=list(range(0, 32, 2))), 6, 1, 1),              # p_out = 80: one sample per tile
    "d4_tcn_512": (dict(c_in=512, c_out=512, t_in=16, v_in=5, kt=3, pad=1, t_sel=list(range(0, 16, 2)), v_keep=[4]), 160, 1, 1),   # p_out = 8: 16 samples per tile
    "d5_tcn_p4": (dict(c_in=512, c_out=512, t_in=8, v_in=1, kt=3, pad=1, t_sel=[0, 2, 4, 6]), 130, 1, 1),                        # p_out = 4: 32 samples per tile
    "d4_res_select": (dict(c_in=256, c_out=512, t_in=16, v_in=5, kt=1, t_sel=list(range(0, 16, 2)), v_keep=[4]), 40, 1, 1),
    "d1_res_select": (dict(c_in=64, c_out=128, t_in=64, v_in=12, kt=1, t_sel=list(range(0, 64, 2)), v_keep=[1, 3, 5, 7, 9]), 6, 0, 1),
    "ragged_k": (dict(c_in=50, c_out=24, t_in=12, v_in=7, kt=3, pad=1), 21, 1, 1),
    "g2_gcn": (dict(c_in=256, c_out=128, t_in=4, v_in=5, K=3), 64, 1, 1),
    "d1_tcn_v25": (dict(c_in=32, c_out=48, t_in=64, v_in=25, kt=3, pad=1), 3, 1, 1),                           # 1600 positions: 12.5 tiles per plane
}


@pytest.mark.parametrize("name", list(GEOMS))
def test_tapconv_staged(name):
    kw, n, fwd_ok, dg_ok = GEOMS[name]
    geom = G.TapConvGeom(**kw)
    lib = import_module("kinetic-gan_b200._lib").lib()
    assert lib.kgan_tapconv_staged_ok(geom.fwd.cstruct(n, 0, 1)) == fwd_ok, geom.fwd.stage_span
    assert lib.kgan_tapconv_staged_ok(geom.dgrad.cstruct(n, 0, 1)) == dg_ok, geom.dgrad.stage_span
    x = rnd(n, geom.K * geom.c_in, geom.t_in, geom.v_in, seed=1)
    w = rnd(geom.K * geom.c_out, geom.c_in, geom.kt, 1, seed=2) / np.sqrt(geom.c_in * geom.kt * geom.K)
    bias = rnd(geom.c_out, seed=3)
    add = rnd(n, geom.c_out, geom.t_out, geom.v_out, seed=4)
    go = rnd(n, geom.c_out, geom.t_out, geom.v_out, seed=5)
    xc, wc = x.cuda(), w.cuda()
    got = ops.tapconv_fwd(xc, wc, geom.fwd)
    assert rel(got, emu.tapconv_fwd(x.double(), w.double(), geom.fwd)) < TOL
    got = ops.tapconv_fwd(xc, wc, geom.fwd, bias.cuda(), add.cuda(), ops.ACT_LRELU)
    assert rel(got, emu.tapconv_fwd(x.double(), w.double(), geom.fwd, bias.double(), add.double(), ops.ACT_LRELU)) < TOL
    got = ops.tapconv_fwd(go.cuda(), wc, geom.dgrad)
    assert rel(got, emu.tapconv_fwd(go.double(), w.double(), geom.dgrad)) < TOL
    # exact on tf32-representable data (small integers; the builder rounds to nearest, which is the identity there)
    gen = torch.Generator().manual_seed(7)
    xi = torch.randint(-3, 4, x.shape, generator=gen).float()
    wi = torch.randint(-2, 3, w.shape, generator=gen).float()
    assert torch.equal(ops.tapconv_fwd(xi.cuda(), wi.cuda(), geom.fwd).cpu().double(), emu.tapconv_fwd(xi.double(), wi.double(), geom.fwd))
    gi = torch.randint(-3, 4, go.shape, generator=gen).float()
    assert torch.equal(ops.tapconv_fwd(gi.cuda(), wi.cuda(), geom.dgrad).cpu().double(), emu.tapconv_fwd(gi.double(), wi.double(), geom.dgrad))


def test_staged_rounds_unrounded_inputs_to_nearest():
    """The builder rounds what it gathers (round to nearest): on inputs that did NOT come from a libkgan kernel the result equals the
    fp64 product of the tf32-rounded operands - the TMA-fed kernel would truncate such inputs."""
    geom = G.TapConvGeom(64, 64, 64, 12, kt=3, pad=1)
    n = 8
    x = rnd(n, 64, 64, 12, seed=11) * 1.37
    w = rnd(64, 64, 3, 1, seed=12) / 14
    rn = lambda t: ((t.view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    got = ops.tapconv_fwd(x.cuda(), w.cuda(), geom.fwd)
    want = emu.tapconv_fwd(rn(x).double(), rn(w).double(), geom.fwd)
    assert rel(got, want) < 2.5e-4      # only the rounding of the stored output (2^-11 / sqrt(3)) is left


@pytest.mark.parametrize("c_in,c_out,t,v,w,n", [(32, 64, 64, 12, 12, 5), (64, 128, 64, 12, 5, 5), (128, 256, 32, 5, 5, 7), (256, 512, 16, 5, 1, 40),
                                                (512, 512, 8, 1, 1, 70), (32, 64, 64, 11, 11, 4), (24, 40, 16, 25, 11, 9)])
def test_gcn_with_adjacency_product_inside_the_gemm(c_in, c_out, t, v, w, n):
    """kgan_gcn_fwd_tf32 (conv1x1 + einsum('nkctv,kvw->nctw') of tgcn.py:61-66 as ONE kernel, the A (.) importance product taken in shared
    memory by the operand builders) against the fp64 statement and against the two-kernel formulation it replaces."""
    gen = torch.Generator().manual_seed(3)
    K = 3
    A = torch.zeros(K, v, w)
    for k in range(K):                                        # skeleton-like sparsity: 1-3 non-zeros per column
        for col in range(w):
            for _ in range(1 + (k + col) % 3):
                A[k, int(torch.randint(0, v, (1,), generator=gen)), col] = float(torch.rand(1, generator=gen)) + 0.1
    nnz = int((A != 0).sum(1).max())
    fused = G.GcnFusedGeom(c_in, c_out, t, v, w, K, nnz)
    two = G.TapConvGeom(c_in, c_out, t, w, K=K)
    x = rnd(n, c_in, t, v, seed=1)
    wt = rnd(K * c_out, c_in, 1, 1, seed=2) / np.sqrt(K * c_in)
    got = ops.gcn_fused_fwd(x.cuda(), A.cuda(), wt.cuda(), fused)
    assert got is not None
    want = emu.tapconv_fwd(emu.adjmix_fwd(x.double(), A.double()), wt.double(), two.fwd)
    assert rel(got, want) < TOL
    sep = ops.tapconv_fwd(ops.adjmix_fwd(x.cuda(), A.cuda()), wt.cuda(), two.fwd)
    assert rel(got, sep.cpu()) < 5e-4
    # exact on tf32-representable data
    xi = torch.randint(-3, 4, x.shape, generator=gen).float()
    Ai = (A != 0).float() * 2
    wi = torch.randint(-2, 3, wt.shape, generator=gen).float()
    assert torch.equal(ops.gcn_fused_fwd(xi.cuda(), Ai.cuda(), wi.cuda(), fused).cpu().double(),
                       emu.tapconv_fwd(emu.adjmix_fwd(xi.double(), Ai.double()), wi.double(), two.fwd))


@pytest.mark.parametrize("c_in,c_out,t,v,w,n", [(64, 128, 64, 12, 5, 5), (128, 256, 32, 5, 5, 7), (256, 512, 16, 5, 1, 40)])
def test_fused_gcn_with_scatter_store(c_in, c_out, t, v, w, n):
    """The one-kernel graph conv writing its result directly in the time-unfolded layout of the following stride-2 temporal conv
    == adjacency kernel + tap convolution + gather kernel."""
    gen = torch.Generator().manual_seed(4)
    A = (torch.rand(3, v, w, generator=gen) < 0.25).float() * torch.rand(3, v, w, generator=gen)
    A[:, 0, :] += 0.5
    nnz = int((A != 0).sum(1).max())
    fused = G.GcnFusedGeom(c_in, c_out, t, v, w, 3, nnz)
    two = G.TapConvGeom(c_in, c_out, t, w, K=3)
    unf = G.UnfoldedTcnGeom(c_out, c_out, t, w, 3, 1, 1, 1, list(range(0, t, 2))).unfold
    x = rnd(n, c_in, t, v, seed=1)
    wt = rnd(3 * c_out, c_in, 1, 1, seed=2) / np.sqrt(3 * c_in)
    poison = torch.full((n, c_out, unf.t_out, unf.v_out), float("nan"), device="cuda")
    del poison
    got = ops.gcn_fused_fwd(x.cuda(), A.cuda(), wt.cuda(), fused, unf)
    assert got is not None
    want = emu.plane_spmm(emu.tapconv_fwd(emu.adjmix_fwd(x.double(), A.double()), wt.double(), two.fwd), unf)
    assert rel(got, want) < TOL


@pytest.mark.parametrize("c,t,v,n,act", [(32, 32, 11, 24, ops.ACT_LRELU), (64, 16, 11, 40, ops.ACT_LRELU), (128, 8, 5, 60, ops.ACT_LRELU), (32, 64, 25, 6, ops.ACT_TANH)])
def test_generator_block_with_noise_epilogue(c, t, v, n, act):
    """kgan_tapconv_fwd_tf32_noise: the eval-mode generator block (generator.py:168-182, BatchNorm folded) - temporal conv + bias + residual +
    noise_weight * noise + activation - as one launch of the operand-building kernel, against the float64 statement and against the two-pass
    formulation (convolution, then the noise / activation pass)."""
    geom = G.TapConvGeom(c, c, t, v, kt=3, pad=1)
    g, r = rnd(n, c, t, v, seed=1), rnd(n, c, t, v, seed=2)
    noise = rnd(n, 1, t, v, seed=3)
    w = rnd(c, c, 3, 1, seed=4) / np.sqrt(3 * c)
    b, nw = rnd(c, seed=5), rnd(c, seed=6)
    got = ops.tapconv_fwd_noise(g.cuda(), w.cuda(), geom.fwd, noise.cuda(), nw.cuda(), b.cuda(), r.cuda(), act)
    assert got is not None, "no plan for a shape of the generator"
    z = emu.tapconv_fwd(g.double(), w.double(), geom.fwd, b.double())
    want = emu.epilogue_fwd(z, r.double(), None, nw.double(), noise.double(), act)
    assert rel(got, want) < TOL
    two = ops.epilogue_fwd(ops.tapconv_fwd(g.cuda(), w.cuda(), geom.fwd, b.cuda()), r.cuda(), None, nw.cuda(), noise.cuda(), act)
    assert rel(got, two.cpu()) < 6e-4
    # without residual
    got = ops.tapconv_fwd_noise(g.cuda(), w.cuda(), geom.fwd, noise.cuda(), nw.cuda(), b.cuda(), None, act)
    assert rel(got, emu.epilogue_fwd(z, None, None, nw.double(), noise.double(), act)) < TOL
