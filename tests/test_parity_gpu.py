"""Parity of the CUDA path (product modules on the GPU, real libkgan.so) with the reference:
 (a) the golden fixtures dumped from the unmodified reference (fp64 ground truth), small shapes;
 (b) the fp64 CPU oracle at the full NTU shape (25 joints x 64 frames, 60 classes) for a small batch;
 (c) size-independent properties at BASELINE batch sizes.
Tolerance (fp32 path): rel-L2 <= 1e-5 on outputs, 5e-4 on gradients (reference's own fp32 noise floor allowed)."""
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


@pytest.fixture(autouse=True)
def fp32_path():
    kgan.set_precision("fp32")
    yield


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
        assert rel_l2(p.grad, gref[k]) < 5e-4, k


def test_properties_at_baseline_batch():
    """Size-independent properties at batch 32 (BASELINE config 1): critic is per-sample independent (no BN), so a batch
    equals the concatenation of its halves; dead lvl-3 partitions get exactly-zero gradient; output is finite and tanh-bounded."""
    cfg = onet.Config()
    G, D, _, _ = build(cfg)
    x = {k: v.cuda() for k, v in inputs(cfg, 32, 5).items()}
    dv = D(x["real"], x["labels"])
    halves = torch.cat([D(x["real"][:16], x["labels"][:16]), D(x["real"][16:], x["labels"][16:])])
    assert rel_l2(dv, halves) < 1e-6
    dv.sum().backward()
    w = D.st_gcn_networks[5].gcn.conv.weight.grad.view(3, -1)
    assert w[1:].abs().max().item() == 0 and w[0].abs().max().item() > 0
    G.eval()
    with torch.no_grad():
        out = G(x["z"], x["labels"])
    assert out.shape == (32, 3, 64, 25) and torch.isfinite(out).all() and out.abs().max() <= 1
