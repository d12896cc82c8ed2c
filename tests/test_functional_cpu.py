"""Host-side autograd logic on CPU (C-ABI primitives emulated, tests/emu_backend.py): every Function family must be
correct to second order, because the gradient penalty differentiates the critic twice (kinetic-gan.py:104-113,154)."""
import numpy as np
import pytest
import torch
from torch.autograd import gradcheck, gradgradcheck

import kgan_b200 as kgan

KF = kgan.functional
G = kgan.geometry


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64).requires_grad_(True)


GEOMS = {
    "gcn_k3": dict(c_in=3, c_out=4, t_in=5, v_in=4, K=3),
    "tcn_pad": dict(c_in=3, c_out=2, t_in=6, v_in=3, kt=3, pad=1),
    "tcn_select": dict(c_in=2, c_out=3, t_in=8, v_in=5, kt=3, pad=1, t_sel=[0, 2, 4, 6], v_keep=[1, 3, 4]),
    "res_select": dict(c_in=2, c_out=3, t_in=8, v_in=5, kt=1, t_sel=[0, 2, 4, 6], v_keep=[0, 2]),
    "gcn_temporal": dict(c_in=2, c_out=2, t_in=7, v_in=3, K=2, kt=3, pad=1, stride=2, dil=1),
    "linear": dict(c_in=5, c_out=4, t_in=1, v_in=1),
}


@pytest.mark.parametrize("name", list(GEOMS))
def test_tapconv_family_second_order(emu, name):
    kw = GEOMS[name]
    geom = G.TapConvGeom(**kw)
    x = rnd(2, geom.K * geom.c_in, geom.t_in, geom.v_in, seed=1)
    w = rnd(geom.K * geom.c_out, geom.c_in, geom.kt, 1, seed=2)
    f = lambda x, w: KF.TapConv.apply(x, w, geom)
    assert gradcheck(f, (x, w))
    assert gradgradcheck(f, (x, w))
    # reference semantics: an ordinary conv over (k, ci, dt) with the selection applied afterwards
    import torch.nn.functional as F
    K, ci, co, kt = geom.K, geom.c_in, geom.c_out, geom.kt
    pad, stride, dil = kw.get("pad", 0), kw.get("stride", 1), kw.get("dil", 1)
    ref = 0
    for k in range(K):
        ref = ref + F.conv2d(x[:, k * ci:(k + 1) * ci], w[k * co:(k + 1) * co], padding=(pad, 0), stride=(stride, 1), dilation=(dil, 1))
    ref = ref[:, :, geom.t_sel][:, :, :, geom.v_keep]
    assert torch.allclose(f(x, w), ref, atol=1e-12)


def test_tapconv_epilogue_second_order(emu):
    geom = G.TapConvGeom(3, 4, 6, 3, kt=3, pad=1)
    x, w, b = rnd(2, 3, 6, 3, seed=1), rnd(4, 3, 3, 1, seed=2), rnd(4, seed=3)
    add = rnd(2, 4, 6, 3, seed=4)
    f = lambda x, w, b, add: KF.TapConvEp.apply(x, w, b, add, geom, KF.ACT_LRELU)
    assert gradcheck(f, (x, w, b, add))
    assert gradgradcheck(f, (x, w, b, add))


def test_adjmix_family_second_order(emu):
    x, A = rnd(2, 3, 4, 5, seed=1), rnd(3, 5, 4, seed=2)          # rectangular A: V=5 -> W=4
    f = lambda x, A: KF.AdjMix.apply(x, A)
    assert gradcheck(f, (x, A))
    assert gradgradcheck(f, (x, A))


def test_graph_conv_equals_reference_formula(emu):
    """AdjMix + TapConv == conv1x1 then einsum('nkctv,kvw->nctw') (tgcn.py:58-68)."""
    from oracle.networks import conv_temporal_graphical
    m = kgan.ConvTemporalGraphical(5, 4, 3).double()
    x, A = rnd(2, 5, 6, 7, seed=1), rnd(3, 7, 7, seed=2)
    out, A2 = m(x, A)
    ref = conv_temporal_graphical(x, m.conv.weight, A)
    assert A2 is A and out.is_contiguous()
    assert torch.allclose(out, ref, atol=1e-12)
    assert gradgradcheck(lambda x, A, w: KF.TapConv.apply(KF.AdjMix.apply(x, A), w, m._geom(6, 7)), (x, A, m.conv.weight))


def test_pointwise_functions(emu):
    go, y = rnd(2, 3, 2, 2, seed=1), rnd(2, 3, 2, 2, seed=2)
    assert gradcheck(lambda g: KF.ActGrad.apply(g, y.detach(), KF.ACT_LRELU), (go,))
    assert gradcheck(lambda g, y: KF.ActGrad.apply(g, y, KF.ACT_TANH), (go, y))
    assert gradcheck(lambda g: KF.ChanSum.apply(g), (go,))
    noise = rnd(2, 1, 2, 2, seed=3).detach()
    nw = rnd(1, 3, 1, 1, seed=4)
    for act in (KF.ACT_LRELU, KF.ACT_TANH):
        assert gradcheck(lambda a, b, nw: KF.NoiseAct.apply(a, b, noise, nw, act), (go, y, nw))
    tab = G.resample_table(2, 2, 4, G.upsample_matrix([np.array([1, 0, 1])], 2, halve=False))
    x = rnd(2, 3, 2, 2, seed=5)
    assert gradcheck(lambda x: KF.PlaneSpmm.apply(x, tab), (x,))
    assert gradgradcheck(lambda x: KF.PlaneSpmm.apply(x, tab), (x,))
    e = rnd(2, 4, seed=6)
    assert gradcheck(lambda e, x: KF.LabelConcat.apply(e, x), (e, x))
    assert gradgradcheck(lambda e, x: KF.LabelConcat.apply(e, x), (e, x))


def test_batchnorm_matches_torch(emu):
    x = rnd(3, 4, 5, 2, seed=1)
    gamma, beta = rnd(4, seed=2), rnd(4, seed=3)
    rm, rv = torch.zeros(4, dtype=torch.float64), torch.ones(4, dtype=torch.float64)
    rm2, rv2 = rm.clone(), rv.clone()
    y = KF.BatchNormTrain.apply(x, gamma, beta, rm, rv, 1e-5, 0.1)
    ref = torch.nn.functional.batch_norm(x, rm2, rv2, gamma, beta, True, 0.1, 1e-5)
    assert torch.allclose(y, ref, atol=1e-12) and torch.allclose(rm, rm2) and torch.allclose(rv, rv2)
    cot = rnd(3, 4, 5, 2, seed=4).detach()
    g1 = torch.autograd.grad((y * cot).sum(), (x, gamma, beta))
    g2 = torch.autograd.grad((ref * cot).sum(), (x, gamma, beta))
    for a, b in zip(g1, g2):
        assert torch.allclose(a, b, atol=1e-10)
    ye = KF.BatchNormEval.apply(x, gamma, beta, rm, rv, 1e-5)
    assert torch.allclose(ye, torch.nn.functional.batch_norm(x, rm, rv, gamma, beta, False, 0.1, 1e-5), atol=1e-12)


def test_geometry_tables():
    # upsample_s as a matrix (generator.py:185-200) vs the oracle's literal insertion loop
    from oracle.networks import upsample_s
    from oracle.graph import SkeletonTables
    for ds in ("ntu", "h36m"):
        t = SkeletonTables(ds)
        for lvl in (2, 1, 0):
            vc = t.num_node[lvl + 1]
            x = torch.randn(2, 3, 4, vc, dtype=torch.float64)
            U = G.upsample_matrix(t.mapping[lvl], vc, halve=(lvl == 2))
            assert U.shape == (vc, t.num_node[lvl])
            assert torch.allclose(x @ torch.from_numpy(U), upsample_s(x, t.mapping[lvl], lvl == 2), atol=1e-14)
    assert G.nearest_src(4, 8) == [0, 0, 1, 1, 2, 2, 3, 3] and G.nearest_src(8, 4) == [0, 2, 4, 6] and G.nearest_src(1, 4) == [0] * 4
    g = G.TapConvGeom(2, 2, 4, 3, kt=3, pad=1)
    assert g.fwd.pmap[0, 0] == -1 and g.fwd.pmap[1, 0] == 0 and g.fwd.pmap[2, 0] == 3   # zero padding at t = -1
    assert (g.dgrad.pmap[0] == np.array([3, 4, 5, 6, 7, 8, 9, 10, 11, -1, -1, -1])).all()


@pytest.mark.parametrize("res", ["conv", "identity"])
def test_gcn_res_joint_node_second_order(emu, res):
    """functional.GcnRes (graph conv + residual branch of a critic block as one node, with the previous block's LeakyReLU slope
    applied in its backward): equals the composition of the separate Functions to first and second order, in both recording
    modes of its backward (plain sweep: stored intermediates; create_graph: recomputed as graph nodes)."""
    c_in, c_out, T, V, W = 3, (4 if res == "conv" else 3), 6, 5, 4
    keep = [0, 2, 3, 4]
    t_sel = [0, 2, 4]
    gcn_geom = G.TapConvGeom(c_in, c_out, T, W, K=3)
    res_geom = G.TapConvGeom(c_in, c_out, len(t_sel), W, kt=1) if res == "conv" else None
    sel = G.select_table(T, V, t_sel, keep)
    z = rnd(2, c_in, T, V, seed=1)                       # pre-activation of the "previous block"
    A, wg = rnd(3, V, W, seed=2), rnd(3 * c_out, c_in, 1, 1, seed=3)
    wr, br = (rnd(c_out, c_in, 1, 1, seed=4), rnd(c_out, seed=5)) if res == "conv" else (None, None)
    cg, cr = rnd(2, c_out, T, W, seed=6).detach(), rnd(2, c_out, len(t_sel), W, seed=7).detach()
    lrelu = lambda t: torch.where(t > 0, t, 0.2 * t)

    def fused(z, A, wg, wr=None, br=None):
        x = _StraightThrough.apply(z)
        g, r = KF.GcnRes.apply(x, A, wg, wr, br, gcn_geom, res_geom, sel, None, True)
        return (g * cg).sum() + (r * cr).sum()

    def plain(z, A, wg, wr=None, br=None):
        x = lrelu(z)
        g = KF.TapConv.apply(KF.AdjMix.apply(x, A), wg, gcn_geom)
        xs = KF.PlaneSpmm.apply(x, sel)
        r = KF.TapConvEp.apply(xs, wr, br, None, res_geom, KF.ACT_NONE) if res == "conv" else xs
        return (g * cg).sum() + (r * cr).sum()

    args = (z, A, wg) + ((wr, br) if res == "conv" else ())
    g1 = torch.autograd.grad(fused(*args), args, create_graph=True)
    g2 = torch.autograd.grad(plain(*args), args, create_graph=True)
    for a, b in zip(g1, g2):
        assert torch.allclose(a, b, atol=1e-11)
    # second order: gradient of a function of the first-order input gradient (what the penalty does), w.r.t. everything
    h1 = torch.autograd.grad((g1[0] ** 2).sum(), args, allow_unused=True)
    h2 = torch.autograd.grad((g2[0] ** 2).sum(), args, allow_unused=True)
    for a, b in zip(h1, h2):
        if a is None or b is None:                      # no dependence on one side: the other must be (numerically) absent too
            assert (a is None or a.abs().max() == 0) and (b is None or b.abs().max() == 0)
        else:
            assert torch.allclose(a, b, atol=1e-10)
    # a plain (unrecorded) sweep uses the stored intermediates
    p1 = torch.autograd.grad(fused(*args), args)
    for a, b in zip(p1, g2):
        assert torch.allclose(a, b, atol=1e-11)


class _StraightThrough(torch.autograd.Function):
    """lrelu(z) in value, identity in gradient: stands for `TapConvEp(..., act_bwd=False)`, whose LeakyReLU slope is applied by
    the consumer (GcnRes with mask_input=True)."""

    @staticmethod
    def forward(ctx, z):
        return torch.where(z > 0, z, 0.2 * z)

    @staticmethod
    def backward(ctx, g):
        return g
