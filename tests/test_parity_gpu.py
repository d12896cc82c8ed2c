"""Parity of the CUDA path (product modules on the GPU, real libkgan.so) with the reference:
 (a) the golden fixtures dumped from the unmodified reference (fp64 ground truth), small shapes;
 (b) the fp64 CPU oracle at the full NTU shape (25 joints x 64 frames, 60 classes) for a small batch;
 (c) size-independent properties at BASELINE batch sizes.
Tolerance (fp32 path): rel-L2 <= 1e-5 on outputs; gradients <= 2e-5 with the LeakyReLU activation pattern pinned to the CUDA pass's
(test_full_ntu_shape_gradients_with_pinned_activation_pattern: arithmetic against arithmetic) and 5e-4 / 2e-3 against the oracle's
own pattern (slope flips of pre-activations within 1e-6 of zero; the reference's own fp32-vs-fp64 distance is of that size)."""
from importlib import import_module

import numpy as np
import pytest
import torch

import kgan_b200 as kgan
from oracle import networks as onet
from oracle.graph import SkeletonTables
from helpers import CASES, draw_noises, inputs, load_golden, parity_ok, rel_l2, sub, within_noise_floor

pytestmark = pytest.mark.gpu
TOL = 1e-5


def build(cfg, dev="cuda"):
    G = kgan.Generator(cfg.latent_dim, cfg.channels, cfg.n_classes, cfg.t_size, cfg.mlp_dim, dataset=cfg.dataset)
    D = kgan.Discriminator(cfg.channels, cfg.n_classes, cfg.t_size, cfg.latent_dim, dataset=cfg.dataset)
    pg = onet.synth_params(onet.g_param_shapes(cfg), 1)
    pd = onet.synth_params(onet.d_param_shapes(cfg), 2)
    G.load_state_dict(pg)
    D.load_state_dict(pd)
    return G.to(dev), D.to(dev), pg, pd


@pytest.fixture(autouse=True, params=["fp32", "fp32x3"])
def fp32_path(request):
    """Both fp32-accurate modes: the exact FMA kernels, and the 3xTF32 tensor-core split (KGAN_PREC_TF32X3) wherever a TMA-fed plan
    exists - same tolerances."""
    kgan.set_precision(request.param)
    yield
    kgan.set_precision("fp32")


@pytest.mark.parametrize("case", list(CASES))
def test_generator_vs_golden(case):
    gold, cfg, n = load_golden(case), CASES[case]["cfg"], CASES[case]["n"]
    G, _, _, _ = build(cfg)
    x = {k: v.cuda() for k, v in inputs(cfg, n, 0).items()}
    G.train()
    blocks = []
    hooks = [m.register_forward_hook(lambda m, i, o: blocks.append(o[0].detach())) for m in G.st_gcn_networks]
    fake = G(x["z"], x["labels"], noises=[t.cuda() for t in draw_noises(cfg, n, 11)])
    for h in hooks:
        h.remove()
    for i, b in enumerate(blocks):
        assert parity_ok(b, gold, "/g_block%d" % i, TOL), (i, rel_l2(b, gold["f64/g_block%d" % i]))
    assert parity_ok(fake, gold, "/g_out", TOL), rel_l2(fake, gold["f64/g_out"])
    (fake * x["cot_g"]).sum().backward()
    for k, p in G.named_parameters():
        assert within_noise_floor(sub(p.grad), gold, "/g_grad/" + k, "f32", 5e-4), k
    for k, b in G.named_buffers():
        if "running" in k:
            assert parity_ok(b, gold, "/g_bn_after/" + k, TOL), k
    G.eval()
    G.load_state_dict(onet.synth_params(onet.g_param_shapes(cfg), 1))
    ev = G(x["z"], x["labels"], noises=[t.cuda() for t in draw_noises(cfg, n, 12)])
    assert parity_ok(ev, gold, "/g_out_eval", TOL), rel_l2(ev, gold["f64/g_out_eval"])


@pytest.mark.parametrize("case", list(CASES))
def test_discriminator_and_gp_vs_golden(case):
    gold, cfg, n = load_golden(case), CASES[case]["cfg"], CASES[case]["n"]
    _, D, _, _ = build(cfg)
    x = {k: v.cuda() for k, v in inputs(cfg, n, 0).items()}
    xr = x["real"].clone().requires_grad_(True)
    blocks = []
    hooks = [m.register_forward_hook(lambda m, i, o: blocks.append(o[0].detach())) for m in D.st_gcn_networks]
    dv = D(xr, x["labels"])
    for h in hooks:
        h.remove()
    for i, b in enumerate(blocks):
        ref = gold["f64/d_block%d" % i]
        assert rel_l2(b[..., :ref.shape[-1]], ref) < TOL, i       # dummy joints (axis padded to a multiple of 4) are dropped
    assert rel_l2(dv, gold["f64/d_out"]) < TOL
    (dv * x["cot_d"]).sum().backward()
    assert rel_l2(xr.grad, gold["f64/d_grad_x"]) < 10 * TOL
    for k, p in D.named_parameters():
        assert within_noise_floor(sub(p.grad), gold, "/d_grad/" + k, "f32", 5e-4), k
    wg = import_module("kinetic-gan_b200.wgan_gp")
    D.zero_grad()
    fake = torch.as_tensor(gold["f64/g_out"]).float().cuda()
    gp, grads = wg.compute_gradient_penalty(D, x["real"], fake, x["labels"], alpha=x["alpha"], return_gradients=True)
    assert abs(gp.item() - float(gold["f64/gp"])) < 1e-4 * max(1.0, abs(float(gold["f64/gp"])))
    assert rel_l2(grads, gold["f64/gp_grads_x"]) < 10 * TOL
    gp.backward()
    for k, p in D.named_parameters():
        if k.endswith("bias") or k == "label_emb.weight":
            assert p.grad is None or p.grad.abs().max().item() == 0, k
        else:
            assert within_noise_floor(sub(p.grad), gold, "/gp_grad/" + k, "f32", 2e-3), k


def test_full_ntu_shape_vs_oracle():
    """NTU-60 shape (25 x 64 x 3, 60 classes, mlp4), batch 6: critic forward, first-order grads, gradient penalty and
    its double-backward grads, generator forward - against the fp64 CPU oracle."""
    cfg = onet.Config()
    n = 6
    tables = SkeletonTables("ntu")
    G, D, pg, pd = build(cfg)
    x = inputs(cfg, n, 3)
    noises = draw_noises(cfg, n, 21)
    pg64 = {k: (v.double() if v.is_floating_point() else v) for k, v in pg.items()}
    pd64 = {k: v.double().requires_grad_(True) for k, v in pd.items()}
    G.train()
    fake = G(x["z"].cuda(), x["labels"].cuda(), noises=[t.cuda() for t in noises])
    fake_ref = onet.generator_forward(pg64, x["z"].double(), x["labels"], cfg, tables, [t.double() for t in noises], True, {})
    assert rel_l2(fake, fake_ref) < TOL
    dv = D(x["real"].cuda(), x["labels"].cuda())
    dv_ref = onet.discriminator_forward(pd64, x["real"].double(), x["labels"], cfg, tables)
    assert rel_l2(dv, dv_ref) < TOL
    wg = import_module("kinetic-gan_b200.wgan_gp")
    gp = wg.compute_gradient_penalty(D, x["real"].cuda(), fake.detach(), x["labels"].cuda(), alpha=x["alpha"].cuda())
    loss = -dv.mean() + 10 * gp
    loss.backward()
    gp_ref = onet.gradient_penalty(pd64, x["real"].double(), fake_ref.detach(), x["labels"], x["alpha"].double(), cfg, tables)
    loss_ref = -dv_ref.mean() + 10 * gp_ref
    assert abs(gp.item() - gp_ref.item()) < 1e-4 * max(1.0, abs(gp_ref.item()))
    keys = list(pd64)
    gref = dict(zip(keys, torch.autograd.grad(loss_ref, [pd64[k] for k in keys], allow_unused=True)))
    for k, p in D.named_parameters():
        assert rel_l2(p.grad, gref[k]) < 5e-4, k          # the oracle's OWN LeakyReLU pattern: slopes flip within the forward error of zero


def test_full_ntu_shape_gradients_with_pinned_activation_pattern():
    """The fp32 path's gradients at the measured floor.  Against the oracle's own activation pattern the parameter gradients agree
    to ~1e-4 .. 5e-4 although every forward quantity agrees to 1e-6: a pre-activation within the forward error of zero has slope
    0.2 in one evaluation and 1 in the other, and a fraction f of flipped slopes costs ~0.8 sqrt(f) in rel-L2 (f ~ 1e-7 here).
    With the pattern PINNED to the CUDA pass's (oracle `masks=`) arithmetic is compared with arithmetic: <= 2e-5 on every
    parameter gradient of loss = -mean(D(x)) + 10 * GP, first-order and double-backward terms together (achieved: printed)."""
    cfg = onet.Config()
    n = 6
    tables = SkeletonTables("ntu")
    _, D, _, pd = build(cfg)
    x = inputs(cfg, n, 3)
    wg = import_module("kinetic-gan_b200.wgan_gp")
    fake = torch.tanh(torch.randn(n, 3, 64, 25, generator=torch.Generator().manual_seed(9)))
    blocks = []
    hooks = [m.register_forward_hook(lambda m, i, o: blocks.append(o[0].detach())) for m in D.st_gcn_networks]
    dv = D(x["real"].cuda(), x["labels"].cuda())
    gp = wg.compute_gradient_penalty(D, x["real"].cuda(), fake.cuda(), x["labels"].cuda(), alpha=x["alpha"].cuda())
    for h in hooks:
        h.remove()
    (-dv.mean() + 10 * gp).backward()
    pd64 = {k: v.double().requires_grad_(True) for k, v in pd.items()}
    ref_blocks = []
    onet.discriminator_forward(pd64, x["real"].double(), x["labels"], cfg, tables, collect=ref_blocks)
    slopes = lambda bs: [torch.where(b[..., :r.shape[-1]].cpu() > 0, 1.0, 0.2).double() for b, r in zip(bs, ref_blocks)]
    dv_ref = onet.discriminator_forward(pd64, x["real"].double(), x["labels"], cfg, tables, masks=slopes(blocks[:6]))
    gp_ref = onet.gradient_penalty(pd64, x["real"].double(), fake.double(), x["labels"], x["alpha"].double(), cfg, tables, masks=slopes(blocks[6:]))
    keys = list(pd64)
    gref = dict(zip(keys, torch.autograd.grad(-dv_ref.mean() + 10 * gp_ref, [pd64[k] for k in keys], allow_unused=True)))
    worst = 0.0
    for k, p in D.named_parameters():
        e = rel_l2(p.grad, gref[k])
        print("fp32 path, pinned pattern: %-45s rel-L2 %.2e" % (k, e))
        worst = max(worst, e)
    assert abs(gp.item() - gp_ref.item()) < 1e-5 * max(1.0, abs(gp_ref.item()))
    assert worst < 2e-5, worst


def test_properties_at_baseline_batch():
    """Size-independent properties at batch 32 (BASELINE config 1): critic is per-sample independent (no BN), so a batch
    equals the concatenation of its halves; dead lvl-3 partitions get exactly-zero gradient; output is finite and tanh-bounded."""
    cfg = onet.Config()
    G, D, _, _ = build(cfg)
    x = {k: v.cuda() for k, v in inputs(cfg, 32, 5).items()}
    dv = D(x["real"], x["labels"])
    halves = torch.cat([D(x["real"][:16], x["labels"][:16]), D(x["real"][16:], x["labels"][16:])])
    # (fp32x3: a batch of 32 and its halves take different kernel plans - tensor-core split above 256 GEMM rows, FMA below)
    assert rel_l2(dv, halves) < (1e-6 if kgan.get_precision() == "fp32" else 1e-5)
    dv.sum().backward()
    w = D.st_gcn_networks[5].gcn.conv.weight.grad.view(3, -1)
    assert w[1:].abs().max().item() == 0 and w[0].abs().max().item() > 0
    G.eval()
    with torch.no_grad():
        out = G(x["z"], x["labels"])
    assert out.shape == (32, 3, 64, 25) and torch.isfinite(out).all() and out.abs().max() <= 1


def test_properties_at_bench_size_tf32():
    """Size-independent properties at a bench size (BASELINE configs[2]: NTU-120 mlp8, per-GPU batch 1024 - the batch of round 1's bench
    line; the default is 4096 now, where the same plans run on four times the tiles - tf32 path):
    (1) adjoint identities <F(x, w), g> = <x, Dgrad(g, w)> = <w, Wgrad(x, g)> on the largest layers (the three kernels must be
    transposes of one another whatever the tiling); (2) the critic treats samples independently (a batch equals its slices);
    (3) one full WGAN-GP iteration through the CUDA-graph trainer leaves finite losses, finite parameters and a tanh-bounded
    generator."""
    if kgan.get_precision() != "fp32":
        pytest.skip("sets its own precision mode: run once")
    geo = import_module("kinetic-gan_b200.geometry")
    wg = import_module("kinetic-gan_b200.wgan_gp")
    n = 1024
    kgan.set_precision("tf32")
    try:
        gen = torch.Generator(device="cuda").manual_seed(7)
        sites = {"d1_gcn": dict(c_in=32, c_out=64, t_in=64, v_in=12, K=3), "d1_tcn": dict(c_in=64, c_out=64, t_in=64, v_in=12, kt=3, pad=1),
                 "d3_gcn": dict(c_in=128, c_out=256, t_in=32, v_in=5, K=3), "map": dict(c_in=632, c_out=632, t_in=1, v_in=1)}
        for name, kw in sites.items():
            g = geo.TapConvGeom(**kw)
            x = torch.randn(n, g.K * g.c_in, g.t_in, g.v_in, device="cuda", generator=gen)
            w = torch.randn(g.K * g.c_out, g.c_in, g.kt, 1, device="cuda", generator=gen) / (g.c_in * g.kt * g.K) ** 0.5
            go = torch.randn(n, g.c_out, g.t_out, g.v_out, device="cuda", generator=gen)
            y = kgan.ops.tapconv_fwd(x, w, g.fwd)
            a = (y.double() * go.double()).sum().item()
            b = (x.double() * kgan.ops.tapconv_fwd(go, w, g.dgrad).double()).sum().item()
            c = (w.double() * kgan.ops.tapconv_wgrad(x, go, g.fwd, tuple(w.shape)).double()).sum().item()
            # the inner product sums M = y.numel() terms of random sign: |a| ~ ||y|| ||g|| / sqrt(M), and so does the rounding noise
            # (~3e-4 per term); a wrong tap, tile or tail would show up at the scale of |a| itself
            tol = 3e-3 * (y.double().norm() * go.double().norm()).item() / y.numel() ** 0.5
            assert abs(a - b) < tol and abs(a - c) < tol, (name, a, b, c, tol)
            del x, w, go, y
        torch.manual_seed(0)
        G = kgan.Generator(512, 3, 120, 64, mlp_dim=8).cuda()
        D = kgan.Discriminator(3, 120, 64, 512).cuda()
        gc = torch.Generator().manual_seed(11)
        real = (torch.rand(n, 3, 64, 25, generator=gc) * 2 - 1).cuda()
        labels = torch.randint(0, 120, (n,), generator=gc).cuda()
        z = torch.randn(n, 512, generator=gc).cuda()
        alpha = torch.rand(n, 1, 1, 1, generator=gc).cuda()
        with torch.no_grad():
            dv = D(real, labels)
            part = torch.cat([D(real[:96], labels[:96]), D(real[-160:], labels[-160:])])
        # slices of 96 / 160 samples take other kernel plans (exact SIMT kernels below 256 GEMM rows): tf32-level agreement
        assert rel_l2(torch.cat([dv[:96], dv[-160:]]), part) < 5e-3
        tr = wg.WGANGPTrainer(G, D)
        tr.capture_graphs(real, labels, z, alpha)
        d_loss, g_loss, gp = tr.iteration(0, real, labels, z, alpha)
        d2, g2, _ = tr.iteration(1, real, labels, z, alpha)
        assert g_loss is not None and g2 is None
        vals = torch.stack([d_loss, g_loss, gp, d2]).cpu()
        assert torch.isfinite(vals).all() and gp.item() >= 0
        assert torch.isfinite(tr.fd.flat).all() and torch.isfinite(tr.fg.flat).all()
        assert tr.fd.step == 2 and tr.fg.step == 1
        G.eval()
        with torch.no_grad():
            out = G(z, labels)
        assert out.shape == (n, 3, 64, 25) and torch.isfinite(out).all() and out.abs().max() <= 1
    finally:
        kgan.set_precision("fp32")
        kgan.ops._persist.clear()
        kgan.ops._batches.clear()
        kgan.ops.clear_temporary_packs()
