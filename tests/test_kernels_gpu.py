"""Every C-ABI kernel (called through the real libkgan.so on the GPU) against the torch-CPU statement of the same
descriptor semantics in float64 (tests/emu_backend.py).  fp32 path: rel-L2 <= 1e-5."""
import numpy as np
import pytest
import torch

import emu_backend as emu
import kgan_b200 as kgan

pytestmark = pytest.mark.gpu
ops, G = kgan.ops, kgan.geometry
TOL = 1e-5


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32)


def cu(t):
    return None if t is None else t.cuda().contiguous()


def dbl(t):
    return None if t is None else t.double()


def rel(a, b):
    b = b.double()
    return ((a.detach().cpu().double() - b).norm() / b.norm().clamp_min(1e-30)).item()


GEOMS = {
    "gcn_k3_small": (dict(c_in=3, c_out=4, t_in=5, v_in=4, K=3), 2),
    "tcn_select": (dict(c_in=2, c_out=3, t_in=8, v_in=5, kt=3, pad=1, t_sel=[0, 2, 4, 6], v_keep=[1, 3, 4]), 3),
    "gcn_temporal": (dict(c_in=2, c_out=2, t_in=7, v_in=3, K=2, kt=3, pad=1, stride=2), 2),
    "linear_mlp": (dict(c_in=572, c_out=572, t_in=1, v_in=1), 37),
    "d0_gcn": (dict(c_in=63, c_out=32, t_in=64, v_in=25, K=3), 3),
    "d2_gcn": (dict(c_in=64, c_out=128, t_in=64, v_in=11, K=3), 2),
    "d2_tcn": (dict(c_in=128, c_out=128, t_in=64, v_in=11, kt=3, pad=1, t_sel=list(range(0, 64, 2)), v_keep=[2, 4, 6, 8, 10]), 2),
    "d4_tcn": (dict(c_in=512, c_out=512, t_in=16, v_in=5, kt=3, pad=1, t_sel=list(range(0, 16, 2)), v_keep=[4]), 5),
    "g6_tcn": (dict(c_in=3, c_out=3, t_in=64, v_in=25, kt=3, pad=1), 4),
    "head": (dict(c_in=512, c_out=1, t_in=1, v_in=1), 9),
}


@pytest.mark.parametrize("name", list(GEOMS))
def test_tapconv_fwd_dgrad_wgrad(name):
    kw, n = GEOMS[name]
    geom = G.TapConvGeom(**kw)
    x = rnd(n, geom.K * geom.c_in, geom.t_in, geom.v_in, seed=1)
    w = rnd(geom.K * geom.c_out, geom.c_in, geom.kt, 1, seed=2) / np.sqrt(geom.c_in * geom.kt * geom.K)
    bias = rnd(geom.c_out, seed=3)
    add = rnd(n, geom.c_out, geom.t_out, geom.v_out, seed=4)
    go = rnd(n, geom.c_out, geom.t_out, geom.v_out, seed=5)
    assert rel(ops.tapconv_fwd(cu(x), cu(w), geom.fwd), emu.tapconv_fwd(dbl(x), dbl(w), geom.fwd)) < TOL
    for act in (ops.ACT_NONE, ops.ACT_LRELU, ops.ACT_TANH):
        got = ops.tapconv_fwd(cu(x), cu(w), geom.fwd, cu(bias), cu(add), act)
        assert rel(got, emu.tapconv_fwd(dbl(x), dbl(w), geom.fwd, dbl(bias), dbl(add), act)) < TOL, act
    assert rel(ops.tapconv_fwd(cu(go), cu(w), geom.dgrad), emu.tapconv_fwd(dbl(go), dbl(w), geom.dgrad)) < TOL
    assert rel(ops.tapconv_wgrad(cu(x), cu(go), geom.fwd, tuple(w.shape)), emu.tapconv_wgrad(dbl(x), dbl(go), geom.fwd, tuple(w.shape))) < TOL


ADJ_SHAPES = [(3, 5, 7, 25, 25, 3), (2, 4, 3, 11, 25, 3), (2, 3, 4, 16, 16, 3), (5, 7, 1, 1, 1, 3), (2, 2, 5, 5, 11, 2),
              # 16-byte addressable row blocks -> the bulk-copy pipelined kernels: several tiles per sample with a ragged tail,
              # the padded 12-joint level, joint selection 12 -> 5, 25 -> 12, single-joint level, H36M 16 joints
              (3, 32, 64, 12, 12, 3), (2, 64, 64, 12, 5, 3), (2, 8, 16, 25, 12, 3), (4, 16, 8, 5, 1, 3), (3, 64, 4, 1, 1, 3),
              (2, 12, 32, 16, 16, 3), (70, 4, 4, 5, 5, 3), (2, 200, 64, 12, 12, 3)]


def sparse_adjacency(k, v, w, seed, density=0.15):
    g = torch.Generator().manual_seed(seed)
    A = (torch.rand(k, v, w, generator=g) < density).float() * (torch.rand(k, v, w, generator=g) + 0.5)
    A[0, :min(v, w), :min(v, w)] += torch.eye(min(v, w))
    return A


@pytest.mark.parametrize("shape", ADJ_SHAPES)
@pytest.mark.parametrize("sparse", [False, True])
def test_adjmix(shape, sparse):
    n, c, t, v, w, k = shape
    x, g = rnd(n, c, t, v, seed=1), rnd(n, k * c, t, w, seed=3)
    A = sparse_adjacency(k, v, w, 2) if sparse else rnd(k, v, w, seed=2)
    assert rel(ops.adjmix_fwd(cu(x), cu(A)), emu.adjmix_fwd(dbl(x), dbl(A))) < TOL
    assert rel(ops.adjmix_bwd_x(cu(g), cu(A)), emu.adjmix_bwd_x(dbl(g), dbl(A))) < TOL
    assert rel(ops.adjmix_bwd_a(cu(x), cu(g), k), emu.adjmix_bwd_a(dbl(x), dbl(g), k)) < TOL
    # masked form: exactly zero outside the support of the mask, the dense values inside
    got = ops.adjmix_bwd_a(cu(x), cu(g), k, cu(A))
    ref = emu.adjmix_bwd_a(dbl(x), dbl(g), k, dbl(A))
    assert rel(got, ref) < TOL
    assert (got.cpu()[A == 0] == 0).all()


def test_adjmix_large_rows():
    n, c, t, v, k = 16, 64, 64, 11, 3
    x, g = rnd(n, c, t, v, seed=1), rnd(n, k * c, t, v, seed=3)
    assert rel(ops.adjmix_bwd_a(cu(x), cu(g), k), emu.adjmix_bwd_a(dbl(x), dbl(g), k)) < TOL


@pytest.mark.parametrize("shape", [(3, 5, 4, 7), (2, 512, 4, 1), (33, 3, 64, 25), (9, 7, 1, 1), (4, 8, 4, 5)])
def test_pointwise(shape):
    n, c, t, v = shape
    a, b, o = rnd(*shape, seed=1), rnd(*shape, seed=2), rnd(*shape, seed=3)
    bias, nw, noise = rnd(c, seed=4), rnd(c, seed=5), rnd(n, 1, t, v, seed=6)
    for act in (0, 1, 2):
        assert rel(ops.epilogue_fwd(cu(a), cu(b), cu(bias), cu(nw), cu(noise), act), emu.epilogue_fwd(dbl(a), dbl(b), dbl(bias), dbl(nw), dbl(noise), act)) < TOL
        assert rel(ops.epilogue_fwd(cu(a), None, None, None, None, act), emu.epilogue_fwd(dbl(a), act=act)) < TOL
        assert rel(ops.act_bwd(cu(a), cu(o), act), emu.act_bwd(dbl(a), dbl(o), act)) < TOL
    assert rel(ops.chan_reduce(cu(a)), emu.chan_reduce(dbl(a))) < TOL
    assert rel(ops.chan_reduce(cu(a), cu(noise)), emu.chan_reduce(dbl(a), dbl(noise))) < TOL
    e = rnd(n, 6, seed=7)
    cat = ops.label_concat(cu(e), cu(a))
    assert torch.equal(cat.cpu(), emu.label_concat(e, a))
    ge, gx = ops.label_split(cat, 6)
    rge, rgx = emu.label_split(dbl(cat.cpu()), 6)
    assert rel(ge, rge) < TOL and torch.equal(gx.cpu(), a)
    alpha = torch.rand(n)
    assert rel(ops.interpolate(cu(alpha), cu(a), cu(b)), emu.interpolate(dbl(alpha), dbl(a), dbl(b))) < TOL


def test_plane_spmm_tables():
    from oracle.graph import SkeletonTables
    t = SkeletonTables("ntu")
    x = rnd(3, 5, 4, 11, seed=1)
    U = G.upsample_matrix(t.mapping[0], 11, halve=False)
    tab = G.resample_table(4, 11, 8, U)
    for tb, inp in ((tab, x), (tab.T, rnd(3, 5, 8, 25, seed=2)), (G.mean_table(4, 11), x),
                    (G.select_table(4, 11, [0, 2], [1, 5, 7]), x)):
        assert rel(ops.plane_spmm(cu(inp), tb), emu.plane_spmm(dbl(inp), tb)) < TOL


@pytest.mark.parametrize("shape", [(4, 6, 5, 3), (32, 3, 32, 11), (2, 256, 4, 1)])
def test_batchnorm(shape):
    n, c, t, v = shape
    x, gy = rnd(*shape, seed=1) * 2 + 0.5, rnd(*shape, seed=2)
    gamma, beta = rnd(c, seed=3), rnd(c, seed=4)
    rm, rv = rnd(c, seed=5), rnd(c, seed=6).abs() + 0.5
    rm_c, rv_c = cu(rm), cu(rv)
    mean, rstd = ops.bn_stats(cu(x), rm_c, rv_c, 1e-5, 0.1)
    rm_r, rv_r = dbl(rm).clone(), dbl(rv).clone()
    mean_r, rstd_r = emu.bn_stats(dbl(x), rm_r, rv_r, 1e-5, 0.1)
    assert rel(mean, mean_r) < TOL and rel(rstd, rstd_r) < TOL and rel(rm_c, rm_r) < TOL and rel(rv_c, rv_r) < TOL
    assert rel(ops.bn_apply(cu(x), mean, rstd, cu(gamma), cu(beta)), emu.bn_apply(dbl(x), mean_r, rstd_r, dbl(gamma), dbl(beta))) < TOL
    got = ops.bn_bwd(cu(gy), cu(x), mean, rstd, cu(gamma))
    ref = emu.bn_bwd(dbl(gy), dbl(x), mean_r, rstd_r, dbl(gamma))
    for a, b in zip(got, ref):
        assert rel(a, b) < 5 * TOL


def test_adam():
    p, g = rnd(100003, seed=1), rnd(100003, seed=2)
    m, v = rnd(100003, seed=3) * 0.1, rnd(100003, seed=4).abs() * 0.01
    pc, mc, vc = cu(p), cu(m), cu(v)
    pr, mr, vr = dbl(p).clone(), dbl(m).clone(), dbl(v).clone()
    for step in (1, 2, 7):
        ops.adam_step(pc, cu(g), mc, vc, 2e-4, 0.5, 0.999, 1e-8, step, 0.5)
        emu.adam_step(pr, dbl(g), mr, vr, 2e-4, 0.5, 0.999, 1e-8, step, 0.5)
    assert rel(pc, pr) < 1e-6 and rel(mc, mr) < TOL and rel(vc, vr) < TOL


def test_errors_are_reported_not_thrown():
    _lib = __import__("importlib").import_module("kinetic-gan_b200._lib")
    lib = _lib.lib()
    assert lib.kgan_adjmix_fwd(0, 0, 0, 1, 1, 1, 1, 1, 1, 0, 0) != 0
    assert b"null" in lib.kgan_last_error()


def test_plane_spmm_small_planes_and_long_lists():
    """The one-/five-joint ends of the critic: planes below 64 positions (several planes per CTA, table entries in registers),
    their adjoints (two-entry lists), and long lists (frame sums, pooling) on the generic kernel."""
    unf = G.UnfoldedTcnGeom(7, 7, 16, 1, 3, 1, 1, 1, list(range(0, 16, 2))).unfold          # (16, 1) -> (24, 1)
    unf5 = G.UnfoldedTcnGeom(7, 7, 8, 5, 3, 1, 1, 1, list(range(0, 8, 2))).unfold            # (8, 5) -> (12, 5): 60 positions
    cases = [(unf, rnd(301, 7, 16, 1, seed=1)), (unf.T, rnd(301, 7, 24, 1, seed=2)),
             (unf5, rnd(130, 9, 8, 5, seed=3)), (unf5.T, rnd(130, 9, 12, 5, seed=4)),
             (G.select_table(16, 5, list(range(0, 16, 2)), [4]), rnd(77, 33, 16, 5, seed=5)),          # (16, 5) -> (8, 1)
             (G.select_table(16, 5, list(range(0, 16, 2)), [4]).T, rnd(77, 33, 8, 1, seed=6)),
             (G.select_table(8, 1, [0, 2, 4, 6], [0]), rnd(2, 3, 8, 1, seed=7)),                        # fewer planes than one CTA holds
             (G.sum_t_table(64, 12), rnd(40, 32, 64, 12, seed=8)), (G.sum_t_table(64, 12).T, rnd(40, 32, 1, 12, seed=9)),
             (G.mean_table(4, 1), rnd(50, 512, 4, 1, seed=10)), (G.mean_table(4, 1).T, rnd(50, 512, 1, 1, seed=11))]
    for tb, inp in cases:
        assert rel(ops.plane_spmm(cu(inp), tb), emu.plane_spmm(dbl(inp), tb)) < TOL, (tb.p_in, tb.p_out, tb.J)


def test_plane_sum_t_and_tiny_adjmix():
    """kgan_plane_sum_t (frame sums of (T, V) planes, V <= 32) against the fp64 statement, incl. V that does not divide 32; and the
    adjacency product on planes of a few floats (the label term of the critic's first layer: 32 rows of 3 floats per sample), which
    is routed to the non-pipelined kernel."""
    for shape in ((9, 32, 64, 12), (5, 7, 16, 5), (3, 4, 8, 1), (2, 3, 64, 25), (4, 5, 3, 32)):
        x = rnd(*shape, seed=sum(shape))
        assert rel(ops.plane_sum_t(cu(x)), dbl(x).sum(2, keepdim=True)) < TOL, shape
    b = rnd(300, 32, 1, 3, seed=1)
    A = rnd(1, 3, 12, seed=2)
    assert rel(ops.adjmix_fwd(cu(b), cu(A)), emu.adjmix_fwd(dbl(b), dbl(A))) < TOL
    g = rnd(300, 32, 1, 12, seed=3)
    assert rel(ops.adjmix_bwd_x(cu(g), cu(A)), emu.adjmix_bwd_x(dbl(g), dbl(A))) < TOL


@pytest.mark.parametrize("case", [
    # (n, c, t, v, w, k, dense adjacency, t_sel step, kept joints)
    (5, 64, 64, 12, 5, 3, False, 2, [1, 4, 6, 9, 11]),      # D2-like: pipelined kernel, short non-zero lists
    (3, 128, 32, 5, 5, 3, True, 2, [4]),                     # dense A (lists of 15 entries: the generic row loop), one kept joint
    (2, 2, 4, 12, 5, 3, False, 2, [0, 2, 3]),                # tiny plane: the plain kernel
    (4, 32, 16, 11, 11, 3, False, 1, [0, 5, 10]),            # joints selected, every frame kept
])
def test_adjmix_bwd_x_fused_epilogues(case):
    """kgan_adjmix_bwd_x_fused (add + LeakyReLU slope of the block input) and kgan_adjmix_bwd_x_fused_sel (the residual branch's gradient
    given in the selection's compact layout, its adjoint taken inside the kernel) against the float64 statement."""
    n, c, t, v, w, k, dense, step, keep = case
    G = kgan.geometry
    g = rnd(n, k * c, t, w, seed=1)
    A = rnd(k, v, w, seed=2) if dense else sparse_adjacency(k, v, w, 3)
    add = rnd(n, c, t, v, seed=4)
    src = rnd(n, c, t, v, seed=5)
    ref = emu.adjmix_bwd_x(dbl(g), dbl(A), dbl(add), dbl(src))
    assert rel(ops.adjmix_bwd_x(cu(g), cu(A), cu(add), cu(src)), ref) < TOL
    assert rel(ops.adjmix_bwd_x(cu(g), cu(A), cu(add), None), emu.adjmix_bwd_x(dbl(g), dbl(A), dbl(add), None)) < TOL
    sel = G.select_table(t, v, list(range(0, t, step)), keep)
    assert sel.inverse_gather() is not None
    addc = rnd(n, c, sel.t_out, sel.v_out, seed=6)
    for m in (src, None):
        got = ops.adjmix_bwd_x(cu(g), cu(A), cu(addc), None if m is None else cu(m), sel)
        want = emu.adjmix_bwd_x(dbl(g), dbl(A), emu.plane_spmm(dbl(addc), sel.T), None if m is None else dbl(m))
        assert rel(got, want) < TOL
    # the same result as the two-kernel formulation it replaces (scatter, then fused add) up to the order of the fp32 additions
    full = ops.plane_spmm(cu(addc), sel.T)
    assert rel(ops.adjmix_bwd_x(cu(g), cu(A), cu(addc), cu(src), sel), ops.adjmix_bwd_x(cu(g), cu(A), full, cu(src)).cpu()) < 1e-6


@pytest.mark.parametrize("kw,n", [
    (dict(c_in=3, c_out=3, t_in=64, v_in=25, kt=3, pad=1), 5),         # generator tail: taps shifted by an odd number of joints (scalar source loads)
    (dict(c_in=32, c_out=3, t_in=32, v_in=11, kt=1), 13),               # 32 -> 3 channels, 352 positions: contiguous sources (16-byte loads)
    (dict(c_in=8, c_out=6, t_in=16, v_in=12, kt=3, pad=1), 30),         # aligned shifts, 8 accumulators
    (dict(c_in=4, c_out=2, t_in=32, v_in=12, kt=3, pad=1, t_sel=list(range(0, 32, 2))), 30),      # strided: non-contiguous sources
    (dict(c_in=3, c_out=3, t_in=64, v_in=25, K=3), 5),                  # channel-block taps
    (dict(c_in=32, c_out=9, t_in=32, v_in=11, kt=1), 13),               # the generator's convolution-first graph conv (K * 3 = 9 outputs): 16 accumulators
    (dict(c_in=6, c_out=14, t_in=16, v_in=12, kt=3, pad=1), 30),        # 14 outputs, 16 accumulators
])
def test_thin_four_positions_per_thread(kw, n):
    """The small-contraction streaming kernel in its four-positions-per-thread form (planes that are multiples of 4 positions):
    forward with every epilogue variant - bias, full and per-joint-periodic `add`, the three activations - and the data
    gradient, against the float64 statement; fp32 FMA arithmetic (<= 1e-6)."""
    geom = G.TapConvGeom(**kw)
    assert (geom.t_out * geom.v_out) % 4 == 0 and n * geom.t_out * geom.v_out >= 4096
    x = rnd(n, geom.K * geom.c_in, geom.t_in, geom.v_in, seed=1)
    w = rnd(geom.K * geom.c_out, geom.c_in, geom.kt, 1, seed=2) / np.sqrt(geom.c_in * geom.kt * geom.K)
    bias = rnd(geom.c_out, seed=3)
    add = rnd(n, geom.c_out, geom.t_out, geom.v_out, seed=4)
    addp = rnd(n, geom.c_out, 1, geom.v_out, seed=6)
    go = rnd(n, geom.c_out, geom.t_out, geom.v_out, seed=5)
    assert rel(ops.tapconv_fwd(cu(x), cu(w), geom.fwd), emu.tapconv_fwd(dbl(x), dbl(w), geom.fwd)) < 1e-6
    for act in (ops.ACT_NONE, ops.ACT_LRELU, ops.ACT_TANH):
        got = ops.tapconv_fwd(cu(x), cu(w), geom.fwd, cu(bias), cu(add), act)
        assert rel(got, emu.tapconv_fwd(dbl(x), dbl(w), geom.fwd, dbl(bias), dbl(add), act)) < 1e-6, act
    got = ops.tapconv_fwd(cu(x), cu(w), geom.fwd, cu(bias), cu(addp), ops.ACT_LRELU)
    assert rel(got, emu.tapconv_fwd(dbl(x), dbl(w), geom.fwd, dbl(bias), dbl(addp), ops.ACT_LRELU)) < 1e-6
    if geom.c_in * geom.K <= 16:                                        # the data gradient is thin as well
        assert rel(ops.tapconv_fwd(cu(go), cu(w), geom.dgrad), emu.tapconv_fwd(dbl(go), dbl(w), geom.dgrad)) < 1e-6


@pytest.mark.parametrize("kw,n", [
    (dict(c_in=3, c_out=3, t_in=64, v_in=25, kt=3, pad=1), 5),          # the generator's last temporal conv: the unrolled 3 x 3 x 3 instance
    (dict(c_in=2, c_out=4, t_in=32, v_in=12, kt=3, pad=1), 20),         # generic instance (24 sums)
    (dict(c_in=3, c_out=3, t_in=32, v_in=11, K=3), 13),                 # channel-block taps
    (dict(c_in=4, c_out=2, t_in=32, v_in=12, kt=3, pad=1, t_sel=list(range(0, 32, 2))), 30),      # strided
])
def test_thin_weight_gradient(kw, n):
    """Weight gradient of layers with at most 32 weights per group (register-resident partial sums, one block-level reduction), also with
    `accumulate`, against the float64 statement - in fp32 and in tf32 mode (the kernel is exact fp32 in both)."""
    geom = G.TapConvGeom(**kw)
    x = rnd(n, geom.K * geom.c_in, geom.t_in, geom.v_in, seed=1)
    go = rnd(n, geom.c_out, geom.t_out, geom.v_out, seed=5)
    wshape = (geom.K * geom.c_out, geom.c_in, geom.kt, 1)
    ref = emu.tapconv_wgrad(dbl(x), dbl(go), geom.fwd, wshape)
    for mode in ("fp32", "tf32"):
        kgan.set_precision(mode)
        try:
            assert rel(ops.tapconv_wgrad(cu(x), cu(go), geom.fwd, wshape), ref) < 2e-6, mode
            acc = torch.full(wshape, 2.0, device="cuda")
            ops.tapconv_wgrad(cu(x), cu(go), geom.fwd, wshape, out=acc)
            assert rel(acc - 2.0, ref) < 2e-5, mode
        finally:
            kgan.set_precision("fp32")


@pytest.mark.parametrize("case", [
    # (n, c, t, v, w, k, frame step, kept joints)
    (5, 64, 64, 12, 5, 3, 2, [1, 4, 6, 9, 11]),       # D2: several whole channel planes per tile
    (6, 128, 32, 5, 5, 3, 2, [0, 1, 2, 3, 4]),        # D3: ragged last tile (still whole planes)
    (9, 256, 16, 5, 1, 3, 2, [2]),                    # D4: one kept joint
    (3, 8, 64, 12, 12, 3, 1, [0, 3, 7]),              # joints only
])
def test_adjmix_fwd_with_selection_byproduct(case):
    """kgan_adjmix_fwd_sel: the adjacency product and, from the same staged tile, the residual branch's input x[:, :, t_sel][..., keep] ==
    the product kernel and the gather kernel run one after the other (bit for bit: both are copies / the same sums)."""
    n, c, t, v, w, k, step, keep = case
    x = rnd(n, c, t, v, seed=1)
    A = sparse_adjacency(k, v, w, 2)
    sel = kgan.geometry.select_table(t, v, list(range(0, t, step)), keep)
    out, xs = ops.adjmix_fwd(cu(x), cu(A), sel)
    assert xs is not None, "no plan for the by-product at a shape of the critic"
    assert torch.equal(out, ops.adjmix_fwd(cu(x), cu(A)))
    assert torch.equal(xs, ops.plane_spmm(cu(x), sel))
    assert rel(out, emu.adjmix_fwd(dbl(x), dbl(A))) < TOL
