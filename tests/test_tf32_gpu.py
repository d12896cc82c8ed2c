"""tcgen05 / TMEM tap convolution (precision 'tf32') against the float64 statement of the same descriptor semantics
and against the exact fp32 SIMT kernel.  Inputs are rounded to tf32 (10-bit mantissa, round-to-nearest), products are
exact and accumulated in fp32: rel-L2 ~3e-4 per layer; the stated tolerance for this path is 1e-3 (BASELINE.json).
Activations are generated tf32-rounded, as every libkgan kernel stores them in tf32 mode (the tensor cores then read them exactly;
tests/test_bench_parity_gpu.py::test_tf32_exact_on_representable_data pins that); weights are arbitrary fp32 (rounded by the packer)."""
import numpy as np
import pytest
import torch

from importlib import import_module

import emu_backend as emu
import kgan_b200 as kgan

pytestmark = pytest.mark.gpu
ops, G = kgan.ops, kgan.geometry
TOL = 1e-3


@pytest.fixture(autouse=True)
def tf32_path():
    kgan.set_precision("tf32")
    yield
    kgan.set_precision("fp32")


def rnd(*shape, seed=0):
    """N(0,1) activations AS THE LIBRARY STORES THEM in tf32 mode: rounded to tf32, round to nearest (include/kgan.h `out_tf32`)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*shape, generator=g, dtype=torch.float32)
    return ((x.view(torch.int32) + 0x1000) & -8192).view(torch.float32)


def rel(a, b):
    b = b.double()
    return ((a.detach().cpu().double() - b).norm() / b.norm().clamp_min(1e-30)).item()


GEOMS = {
    "d1_gcn": (dict(c_in=32, c_out=64, t_in=64, v_in=11, K=3), 4),
    "d2_gcn": (dict(c_in=64, c_out=128, t_in=64, v_in=11, K=3), 3),
    "d1_tcn_v12": (dict(c_in=64, c_out=64, t_in=64, v_in=12, kt=3, pad=1), 4),                     # aligned shifts: TMA-fed kernel
    "d1_tcn_v11": (dict(c_in=64, c_out=64, t_in=64, v_in=11, kt=3, pad=1), 4),                     # unaligned shifts: gather kernel
    "d4_gcn_p80": (dict(c_in=256, c_out=512, t_in=16, v_in=5, K=3), 7),                            # boxes of 16 positions x 2 samples
    "d5_gcn_p8": (dict(c_in=512, c_out=512, t_in=8, v_in=1, K=3), 70),                             # boxes of 8 positions x 4 samples
    "d2_tcn_unfolded": (dict(unfold=True, c_in=128, c_out=128, t_in=64, v_in=5, kt=3, pad=1, stride=1, dil=1, t_sel=list(range(0, 64, 2))), 5),
    "d5_tcn_unfolded": (dict(unfold=True, c_in=512, c_out=512, t_in=8, v_in=1, kt=3, pad=1, stride=1, dil=1, t_sel=[0, 2, 4, 6]), 130),
    "d2_tcn_select": (dict(c_in=128, c_out=128, t_in=64, v_in=11, kt=3, pad=1, t_sel=list(range(0, 64, 2)), v_keep=[2, 4, 6, 8, 10]), 5),
    "d3_tcn": (dict(c_in=256, c_out=256, t_in=32, v_in=5, kt=3, pad=1, t_sel=list(range(0, 32, 2))), 6),
    "d4_tcn_512": (dict(c_in=512, c_out=512, t_in=16, v_in=5, kt=3, pad=1, t_sel=list(range(0, 16, 2)), v_keep=[4]), 160),
    "d4_res": (dict(c_in=256, c_out=512, t_in=16, v_in=5, kt=1, t_sel=list(range(0, 16, 2)), v_keep=[4]), 40),
    "ragged_k": (dict(c_in=50, c_out=24, t_in=9, v_in=7, kt=3, pad=1), 21),
    "mlp_632": (dict(c_in=632, c_out=632, t_in=1, v_in=1), 300),                                   # one-position planes: K-major TMA boxes
    "mlp_522": (dict(c_in=522, c_out=522, t_in=1, v_in=1), 300),                                   # row stride not a multiple of 16 bytes: gather kernel
    "g0_gcn_p1": (dict(c_in=632, c_out=512, t_in=1, v_in=1, K=3), 260),                            # channel-block taps of a K-major operand, ragged K tail
    "g0_tcn_p1": (dict(c_in=512, c_out=512, t_in=1, v_in=1, kt=3, pad=1), 520),                    # 2 of 3 taps only read padding: pruned
    "g2_gcn": (dict(c_in=256, c_out=128, t_in=4, v_in=5, K=3), 64),
    "d0_gcn_3ch": (dict(c_in=3, c_out=32, t_in=64, v_in=25, K=3, w_cin=123, w_ic0=120), 4),
    "g6_tcn_3ch": (dict(c_in=3, c_out=3, t_in=64, v_in=25, kt=3, pad=1), 4),
    "d3_tcn_p160": (dict(c_in=128, c_out=256, t_in=32, v_in=5, K=3), 70),                          # several N / M tiles, split-K over 2+ samples
    "d1_res_v12": (dict(c_in=32, c_out=64, t_in=64, v_in=12, kt=1), 9),
    # planes below 32 positions with >= 1024 positions in total
    "d4_gcn_p80_big": (dict(c_in=256, c_out=512, t_in=16, v_in=5, K=3), 15),                       # 64-byte swizzled sub-tiles (16 positions)
    "d5_gcn_p8_big": (dict(c_in=512, c_out=512, t_in=8, v_in=1, K=3), 141),                        # 32-byte swizzled sub-tiles, ragged last stage
    "d5_tcn_unfolded_big": (dict(unfold=True, c_in=512, c_out=512, t_in=8, v_in=1, kt=3, pad=1, stride=1, dil=1, t_sel=[0, 2, 4, 6]), 301),
}


TMA_EXPECTED = {"d1_gcn": (1, 1), "d2_gcn": (1, 1), "d1_tcn_v12": (1, 1), "d1_tcn_v11": (0, 0), "d4_gcn_p80": (1, 1), "d5_gcn_p8": (1, 1),
                "d2_tcn_unfolded": (1, 1), "d5_tcn_unfolded": (1, 1), "d2_tcn_select": (0, 0), "ragged_k": (0, 0),
                "mlp_632": (1, 1), "mlp_522": (0, 0), "g0_gcn_p1": (1, 1), "g0_tcn_p1": (1, 1)}


THIN = {"g6_tcn_3ch"}
WGRAD_TMA_EXPECTED = {"d1_gcn": 1, "d2_gcn": 1, "d1_tcn_v12": 1, "d1_tcn_v11": 0, "d0_gcn_3ch": 1, "d3_tcn_p160": 1, "d1_res_v12": 1,
                      "d2_tcn_unfolded": 0, "g2_gcn": 0, "d4_gcn_p80_big": 1, "d5_gcn_p8_big": 1, "d5_tcn_unfolded_big": 0, "g6_tcn_3ch": 0}


@pytest.mark.parametrize("name", list(GEOMS))
def test_tapconv_tf32(name):
    kw, n = GEOMS[name]
    kw = dict(kw)
    if kw.pop("unfold", False):
        geom = G.UnfoldedTcnGeom(**kw)                # operates on the time-unfolded input (kt blocks of t_out frames)
        x = rnd(n, geom.c_in, geom.kt * geom.t_out, geom.v_in, seed=1)
    else:
        geom = G.TapConvGeom(**kw)
        x = rnd(n, geom.K * geom.c_in, geom.t_in, geom.v_in, seed=1)
    if name in TMA_EXPECTED:                          # the TMA-fed kernel is the one that runs (no silent gather fallback)
        lib = import_module("kinetic-gan_b200._lib").lib()
        assert lib.kgan_tapconv_tma_ok(geom.fwd.cstruct(n, 0, 1)) == TMA_EXPECTED[name][0]
        assert lib.kgan_tapconv_tma_ok(geom.dgrad.cstruct(n, 0, 1)) == TMA_EXPECTED[name][1]
    if name in WGRAD_TMA_EXPECTED:                    # ... and likewise the TMA-fed weight-gradient kernel
        lib = import_module("kinetic-gan_b200._lib").lib()
        assert lib.kgan_tapconv_wgrad_tma_ok(geom.fwd.cstruct(n, 0, 1)) == WGRAD_TMA_EXPECTED[name]
    w = rnd(geom.K * geom.c_out, kw.get("w_cin", geom.c_in), geom.kt, 1, seed=2) / np.sqrt(geom.c_in * geom.kt * geom.K)
    bias = rnd(geom.c_out, seed=3)
    add = rnd(n, geom.c_out, geom.t_out, geom.v_out, seed=4)
    go = rnd(n, geom.c_out, geom.t_out, geom.v_out, seed=5)
    xc, wc = x.cuda(), w.cuda()
    l0 = ops.launches
    got = ops.tapconv_fwd(xc, wc, geom.fwd)
    assert rel(got, emu.tapconv_fwd(x.double(), w.double(), geom.fwd)) < TOL
    got = ops.tapconv_fwd(xc, wc, geom.fwd, bias.cuda(), add.cuda(), ops.ACT_LRELU)
    assert rel(got, emu.tapconv_fwd(x.double(), w.double(), geom.fwd, bias.double(), add.double(), ops.ACT_LRELU)) < TOL
    got = ops.tapconv_fwd(go.cuda(), wc, geom.dgrad)                       # data gradient = same kernel, transposed roles
    assert rel(got, emu.tapconv_fwd(go.double(), w.double(), geom.dgrad)) < TOL
    got = ops.tapconv_wgrad(xc, go.cuda(), geom.fwd, tuple(w.shape))         # weight gradient (tensor cores when eligible)
    assert rel(got, emu.tapconv_wgrad(x.double(), go.double(), geom.fwd, tuple(w.shape))) < TOL
    # the tensor-core entry point really ran (pack + fwd) and differs from the exact kernel only by tf32 rounding
    kgan.set_precision("fp32")
    exact = ops.tapconv_fwd(xc, wc, geom.fwd)
    kgan.set_precision("tf32")
    again = ops.tapconv_fwd(xc, wc, geom.fwd)
    d = rel(again, exact.cpu())
    if name in THIN:            # small contractions run on the exact streaming kernel in both modes (csrc/tapconv_simt.cu);
        # in tf32 mode its output is stored tf32-rounded like every activation (include/kgan.h `out_tf32`)
        assert torch.equal(again, ops.round_tf32(exact)), d
    else:
        assert 1e-6 < d < TOL, d


def test_critic_tf32_vs_oracle():
    from importlib import import_module
    from oracle import networks as onet
    from oracle.graph import SkeletonTables
    from helpers import inputs
    cfg = onet.Config()
    n = 6
    tables = SkeletonTables("ntu")
    D = kgan.Discriminator(3, 60, 64, 512)
    pd = onet.synth_params(onet.d_param_shapes(cfg), 2)
    D.load_state_dict(pd)
    D = D.cuda()
    x = inputs(cfg, n, 3)
    pd64 = {k: v.double().requires_grad_(True) for k, v in pd.items()}
    dv = D(x["real"].cuda(), x["labels"].cuda())
    dv_ref = onet.discriminator_forward(pd64, x["real"].double(), x["labels"], cfg, tables)
    assert rel(dv, dv_ref.detach()) < 2e-3
    wg = import_module("kinetic-gan_b200.wgan_gp")
    fake = torch.tanh(torch.randn(n, 3, 64, 25, generator=torch.Generator().manual_seed(5)))
    gp = wg.compute_gradient_penalty(D, x["real"].cuda(), fake.cuda(), x["labels"].cuda(), alpha=x["alpha"].cuda())
    (-dv.mean() + 10 * gp).backward()
    gp_ref = onet.gradient_penalty(pd64, x["real"].double(), fake.double(), x["labels"], x["alpha"].double(), cfg, tables)
    assert abs(gp.item() - gp_ref.item()) < 2e-3 * max(1.0, abs(gp_ref.item()))
    keys = list(pd64)
    gref = dict(zip(keys, torch.autograd.grad(-dv_ref.mean() + 10 * gp_ref, [pd64[k] for k in keys], allow_unused=True)))
    # end-to-end parameter gradients chain ~25 tf32 GEMMs (forward, backward and double backward through 6 blocks):
    # per-layer error <= 1e-3 (test above) compounds to ~1e-2 at the first block; stated tolerance 2e-2
    for k, p in D.named_parameters():
        assert rel(p.grad, gref[k]) < 2e-2, k


RES_CASES = {
    # name: (tcn geometry, unfolded?, residual in-channels, batch, fused expected)
    "d1": (dict(c_in=64, c_out=64, t_in=64, v_in=12, kt=3, pad=1), False, 32, 5, 1),                                # TMA boxes (768 positions)
    "d2_unfolded": (dict(c_in=128, c_out=128, t_in=64, v_in=5, kt=3, pad=1, stride=1, dil=1, t_sel=list(range(0, 64, 2))), True, 64, 5, 1),
    "d3_unfolded": (dict(c_in=256, c_out=256, t_in=32, v_in=5, kt=3, pad=1, stride=1, dil=1, t_sel=list(range(0, 32, 2))), True, 128, 7, 1),   # 80 positions: cp.async producers
    "d4_unfolded": (dict(c_in=512, c_out=512, t_in=16, v_in=1, kt=3, pad=1, stride=1, dil=1, t_sel=list(range(0, 16, 2))), True, 256, 70, 1),  # 8 positions, N split
    "unaligned": (dict(c_in=64, c_out=64, t_in=64, v_in=11, kt=3, pad=1), False, 32, 4, 0),                         # no TMA-fed plan: separate launches
}


@pytest.mark.parametrize("name", list(RES_CASES))
def test_tapconv_with_fused_residual(name):
    """kgan_tapconv_fwd_tf32_res: act(tcn(g) + b + res(xs) + b_res) in one accumulator, against the fp64 statement of the two
    convolutions; where the pair is not eligible ops.tapconv_fwd_res returns None (the Function then launches them separately)."""
    kw, unfolded, c_res, n, fused = RES_CASES[name]
    geom = G.UnfoldedTcnGeom(**kw) if unfolded else G.TapConvGeom(**kw)
    res = G.TapConvGeom(c_res, geom.c_out, geom.t_out, geom.v_out, kt=1)
    g = rnd(n, geom.c_in, (geom.kt * geom.t_out) if unfolded else geom.t_in, geom.v_in, seed=1)
    xs = rnd(n, c_res, geom.t_out, geom.v_out, seed=2)
    w = rnd(geom.c_out, geom.c_in, geom.kt, 1, seed=3) / np.sqrt(geom.c_in * geom.kt)
    wr = rnd(geom.c_out, c_res, 1, 1, seed=4) / np.sqrt(c_res)
    b, br = rnd(geom.c_out, seed=5), rnd(geom.c_out, seed=6)
    got = ops.tapconv_fwd_res(g.cuda(), w.cuda(), geom.fwd, xs.cuda(), wr.cuda(), res.fwd, b.cuda(), br.cuda(), ops.ACT_LRELU)
    assert (got is not None) == bool(fused)
    want = emu.tapconv_fwd_res(g.double(), w.double(), geom.fwd, xs.double(), wr.double(), res.fwd, b.double(), br.double(), ops.ACT_LRELU)
    if got is not None:
        assert rel(got, want) < TOL
        # and it equals the two separate launches up to fp32 accumulation order + one output rounding
        r = ops.tapconv_fwd(xs.cuda(), wr.cuda(), res.fwd, br.cuda())
        sep = ops.tapconv_fwd(g.cuda(), w.cuda(), geom.fwd, b.cuda(), r, ops.ACT_LRELU)
        assert rel(got, sep.cpu()) < 5e-4
    # the Function covers both cases
    out = kgan.functional.TcnRes.apply(g.cuda(), w.cuda(), b.cuda(), xs.cuda(), wr.cuda(), br.cuda(), geom, res, ops.ACT_LRELU)
    assert rel(out, want) < TOL


@pytest.mark.parametrize("c_in,c_out,t,v,n", [(64, 128, 64, 5, 6), (128, 256, 32, 5, 9), (256, 512, 16, 1, 70), (512, 512, 8, 1, 130)])
def test_tapconv_scatter_store(c_in, c_out, t, v, n):
    """kgan_tapconv_fwd_tf32_scatter: a graph conv whose epilogue stores its result directly in the time-unfolded layout of the following
    stride-2 temporal conv (two copies of the odd frames, zero padding slots) == tap convolution followed by the gather kernel, bit for
    bit (same accumulators, same rounding), and both against the fp64 statement."""
    geom = G.TapConvGeom(c_in, c_out, t, v, K=3)
    unf = G.UnfoldedTcnGeom(c_out, c_out, t, v, 3, 1, 1, 1, list(range(0, t, 2))).unfold
    m = unf.scatter_map()
    assert m is not None and sorted(int(q) for q in m.reshape(-1) if q >= 0) == list(range(unf.p_out))      # every slot has exactly one role
    x = rnd(n, 3 * c_in, t, v, seed=1)
    w = rnd(3 * c_out, c_in, 1, 1, seed=2) / np.sqrt(3 * c_in)
    poison = torch.full((n, c_out, unf.t_out, unf.v_out), float('nan'), device='cuda')       # the result must not depend on what the allocator hands out
    del poison
    got = ops.tapconv_fwd_scatter(x.cuda(), w.cuda(), geom.fwd, unf)
    assert got is not None
    two = ops.plane_spmm(ops.tapconv_fwd(x.cuda(), w.cuda(), geom.fwd), unf)
    assert torch.equal(got, two)
    assert rel(got, emu.plane_spmm(emu.tapconv_fwd(x.double(), w.double(), geom.fwd), unf)) < TOL


@pytest.mark.parametrize("c_in,c_out,n", [(632, 632, 4096), (632, 1536, 2048), (512, 512, 1024), (572, 572, 600)])
def test_linear_weight_gradient_on_the_forward_kernel(c_in, c_out, n):
    """Weight gradient of a Linear layer in tf32 mode: dW = gout^T x computed by the tensor-core FORWARD kernel on the transposed problem
    (x as a one-sample tensor of N channels, gout as the weight matrix) - against the float64 statement, with and without `accumulate`,
    and the launch is a tensor-core one."""
    geom = G.TapConvGeom(c_in=c_in, c_out=c_out, t_in=1, v_in=1)
    x, go = rnd(n, c_in, 1, 1, seed=1), rnd(n, c_out, 1, 1, seed=2)
    ref = emu.tapconv_wgrad(x.double(), go.double(), geom.fwd, (c_out, c_in, 1, 1))
    prof = ops.profile_start()
    try:
        dw = ops.tapconv_wgrad(x.cuda(), go.cuda(), geom.fwd, (c_out, c_in, 1, 1))
        acc = torch.full((c_out, c_in, 1, 1), 3.0, device="cuda")
        ops.tapconv_wgrad(x.cuda(), go.cuda(), geom.fwd, (c_out, c_in, 1, 1), out=acc)
    finally:
        ops.profile_stop(prof)
    fams = [p[0] for p in prof]
    assert fams.count("tapconv_wgrad_tf32") == 2 and "tapconv_wgrad" not in fams, fams
    assert rel(dw, ref) < TOL
    assert rel(acc - 3.0, ref) < TOL
