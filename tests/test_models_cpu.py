"""Product modules (host logic, CPU, C-ABI primitives emulated) against the oracle and the golden fixtures:
same synthetic weights loaded through `load_state_dict`, same inputs, same noise."""
import numpy as np
import pytest
import torch

import kgan_b200 as kgan
from oracle import networks as onet
from oracle.graph import SkeletonTables
from helpers import CASES, draw_noises, inputs, load_golden, rel_l2, sub, within_noise_floor


def build(cfg, dtype=torch.float32):
    G = kgan.Generator(cfg.latent_dim, cfg.channels, cfg.n_classes, cfg.t_size, cfg.mlp_dim, dataset=cfg.dataset)
    D = kgan.Discriminator(cfg.channels, cfg.n_classes, cfg.t_size, cfg.latent_dim, dataset=cfg.dataset)
    pg = onet.synth_params(onet.g_param_shapes(cfg), 1)
    pd = onet.synth_params(onet.d_param_shapes(cfg), 2)
    G.load_state_dict(pg)          # strict: key names and shapes must be the reference's
    D.load_state_dict(pd)
    if dtype == torch.float64:
        G, D = G.double(), D.double()
        for m in (G, D):
            for i, a in enumerate(m.graph.As):
                setattr(m, "_A%d" % i, torch.tensor(a, dtype=torch.float64))
    return G, D


def test_state_dict_surface():
    cfg = onet.Config()
    G = kgan.Generator(512, 3, 60, 64, 4)
    D = kgan.Discriminator(3, 60, 64, 512)
    assert {k: tuple(v.shape) for k, v in G.state_dict().items()} == onet.g_param_shapes(cfg)
    assert {k: tuple(v.shape) for k, v in D.state_dict().items()} == onet.d_param_shapes(cfg)
    # reference init rules: N(0,1) mapping weights, zero noise weights, unit edge importance
    assert abs(G.mlp.mlp[0].weight.std().item() - 1.0) < 0.02 and G.mlp.mlp[0].bias.abs().max() == 0
    assert all(b.noise.weight.abs().max() == 0 for b in G.st_gcn_networks)
    assert all((e == 1).all() for e in G.edge_importance) and all((e == 1).all() for e in D.edge_importance)
    assert not any(k.startswith("_A") for k in G.state_dict())


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_generator_vs_golden(emu, case, tag):
    gold, cfg, n = load_golden(case), CASES[case]["cfg"], CASES[case]["n"]
    dtype = torch.float64 if tag == "f64" else torch.float32
    tol = 1e-10 if tag == "f64" else 2e-5
    G, _ = build(cfg, dtype)
    x = inputs(cfg, n, 0, dtype)
    G.train()
    blocks = []
    hooks = [m.register_forward_hook(lambda m, i, o: blocks.append(o[0].detach())) for m in G.st_gcn_networks]
    fake = G(x["z"], x["labels"], noises=draw_noises(cfg, n, 11, dtype))
    for h in hooks:
        h.remove()
    for i, b in enumerate(blocks):
        assert rel_l2(b, gold[tag + "/g_block%d" % i]) < tol, i
    assert rel_l2(fake, gold[tag + "/g_out"]) < tol
    (fake * x["cot_g"]).sum().backward()
    for k, p in G.named_parameters():
        assert within_noise_floor(sub(p.grad), gold, "/g_grad/" + k, tag, 50 * tol), k
    for k, b in G.named_buffers():
        if "running" in k:
            assert rel_l2(b, gold[tag + "/g_bn_after/" + k]) < tol, k
        if "num_batches_tracked" in k:
            assert int(b) == 1
    # default noise path: same CPU RNG stream as the reference (generator.py:179)
    G2, _ = build(cfg, dtype)
    G2.eval()
    torch.manual_seed(12)
    assert rel_l2(G2(x["z"], x["labels"]), gold[tag + "/g_out_eval"]) < tol
    np.random.seed(5)
    torch.manual_seed(13)
    assert rel_l2(G2(x["z"], x["labels"], 0.95), gold[tag + "/g_out_trunc"]) < 5 * tol


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_discriminator_and_gp_vs_golden(emu, case, tag):
    gold, cfg, n = load_golden(case), CASES[case]["cfg"], CASES[case]["n"]
    dtype = torch.float64 if tag == "f64" else torch.float32
    tol = 1e-10 if tag == "f64" else 2e-5
    _, D = build(cfg, dtype)
    x = inputs(cfg, n, 0, dtype)
    xr = x["real"].clone().requires_grad_(True)
    blocks = []
    hooks = [m.register_forward_hook(lambda m, i, o: blocks.append(o[0].detach())) for m in D.st_gcn_networks]
    dv = D(xr, x["labels"])
    for h in hooks:
        h.remove()
    for i, b in enumerate(blocks):
        ref = gold[tag + "/d_block%d" % i]
        # inside Discriminator.forward a block may carry dummy joints beyond the graph's V (joint axis padded to a multiple of 4)
        assert rel_l2(b[..., :ref.shape[-1]], ref) < tol, i
    assert rel_l2(dv, gold[tag + "/d_out"]) < tol
    (dv * x["cot_d"]).sum().backward()
    assert rel_l2(xr.grad, gold[tag + "/d_grad_x"]) < 10 * tol
    for k, p in D.named_parameters():
        assert within_noise_floor(sub(p.grad), gold, "/d_grad/" + k, tag, 50 * tol), k
    # gradient penalty (kinetic-gan.py:94-114): value, d/dx and the double-backward parameter gradients
    from importlib import import_module
    wg = import_module("kinetic-gan_b200.wgan_gp")
    D.zero_grad()
    fake = torch.as_tensor(gold[tag + "/g_out"]).to(dtype)
    gp, grads = wg.compute_gradient_penalty(D, x["real"], fake, x["labels"], alpha=x["alpha"], return_gradients=True)
    assert abs(gp.item() - float(gold[tag + "/gp"])) < 20 * tol * max(1.0, abs(float(gold[tag + "/gp"])))
    assert rel_l2(grads, gold[tag + "/gp_grads_x"]) < 10 * tol
    gp.backward()
    for k, p in D.named_parameters():
        if k.endswith("bias") or k == "label_emb.weight":
            assert p.grad is None or p.grad.abs().max().item() == 0, k      # the penalty cannot see them
        else:
            assert within_noise_floor(sub(p.grad), gold, "/gp_grad/" + k, tag, 100 * tol), k


def test_label_folding_equals_label_planes(emu):
    """Critic block 0 with the label channels folded analytically == the same block on cat(label planes, x)."""
    cfg = CASES["ntu_small"]["cfg"]
    _, D = build(cfg, torch.float64)
    x = inputs(cfg, 3, 0, torch.float64)
    c = D.label_emb(x["labels"])
    blk, A = D.st_gcn_networks[0], D.A[0] * D.edge_importance[0]
    a, _ = blk(x["real"], A, label_emb=c)
    A = D.A[0] * D.edge_importance[0]
    b, _ = blk(kgan.functional.LabelConcat.apply(D.label_emb(x["labels"]), x["real"]), A)
    assert torch.allclose(a, b, atol=1e-12)
    ga = torch.autograd.grad(a.sum(), [blk.gcn.conv.weight, D.label_emb.weight, D.edge_importance[0]])
    gb = torch.autograd.grad(b.sum(), [blk.gcn.conv.weight, D.label_emb.weight, D.edge_importance[0]])
    for u, v in zip(ga, gb):
        assert torch.allclose(u, v, atol=1e-10)


def test_generator_eval_batchnorm_folding(emu):
    """Generator.fold_batchnorm (inference only): same eval-mode output with the BatchNorms folded into the convolutions;
    `.train()` and `load_state_dict()` drop the folds."""
    cfg, n = CASES["ntu_small"]["cfg"], 3
    G = kgan.Generator(cfg.latent_dim, cfg.channels, cfg.n_classes, cfg.t_size, cfg.mlp_dim, dataset=cfg.dataset).double()
    G.load_state_dict(onet.synth_params(onet.g_param_shapes(cfg), 1))
    for i, a in enumerate(G.graph.As):
        setattr(G, "_A%d" % i, torch.tensor(a, dtype=torch.float64))
    G.eval()
    xi = inputs(cfg, n, 4, torch.float64)
    nz = draw_noises(cfg, n, 7, torch.float64)
    with torch.no_grad():
        ref = G(xi["z"], xi["labels"], noises=nz)
        G.fold_batchnorm()
        assert all((blk._fold is not None) for blk in G.st_gcn_networks)
        calls = []
        orig = kgan.ops.bn_apply
        kgan.ops.bn_apply = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
        try:
            out = G(xi["z"], xi["labels"], noises=nz)
        finally:
            kgan.ops.bn_apply = orig
    assert not calls                                               # no BatchNorm launch left in the eval pass
    assert (out - ref).abs().max().item() < 1e-10
    G.train()
    assert all(blk._fold is None for blk in G.st_gcn_networks)
    G.eval().fold_batchnorm()
    G.load_state_dict(G.state_dict())
    assert all(blk._fold is None for blk in G.st_gcn_networks)


def test_edge_importance_gradient_where_importance_is_zero(emu):
    """the adjacency gradient is evaluated on the support of the constant BASE adjacency, not of the current
    A * edge_importance - an importance entry that is exactly 0 (or a cancelling column sum) still gets its gradient; and a
    standalone ConvTemporalGraphical with a dense learnable A gets a dense gradient.  Checked against the fp64 oracle."""
    cfg = CASES["ntu_small"]["cfg"]
    tables = SkeletonTables(cfg.dataset)
    _, D = build(cfg, torch.float64)
    pd = {k: v.double() for k, v in onet.synth_params(onet.d_param_shapes(cfg), 2).items()}
    for i in (0, 1, 3):                                   # zero one importance entry that sits ON the skeleton's support, per partition
        base = tables.As[onet.d_block_table(cfg)[i][2]]
        for k in range(base.shape[0]):
            v, w = [int(t[0]) for t in np.nonzero(base[k])]
            pd["edge_importance.%d" % i][k, v, w] = 0.0
    D.load_state_dict(pd)
    x = inputs(cfg, 3, 7, torch.float64)
    (D(x["real"], x["labels"]) * x["cot_d"]).sum().backward()
    pr = {k: v.clone().requires_grad_(True) for k, v in pd.items()}
    ref = (onet.discriminator_forward(pr, x["real"], x["labels"], cfg, tables) * x["cot_d"]).sum()
    gref = torch.autograd.grad(ref, [pr["edge_importance.%d" % i] for i in range(6)])
    for i in range(6):
        g = D.edge_importance[i].grad
        assert rel_l2(g, gref[i]) < 1e-9, i
    base = tables.As[onet.d_block_table(cfg)[1][2]]
    v, w = [int(t[0]) for t in np.nonzero(base[0])]
    assert pd["edge_importance.1"][0, v, w] == 0 and D.edge_importance[1].grad[0, v, w].abs() > 0
    # standalone operator, dense learnable A: every entry of dA
    m = kgan.ConvTemporalGraphical(4, 6, 3).double()
    xs = torch.randn(2, 4, 5, 7, dtype=torch.float64)
    A = torch.randn(3, 7, 7, dtype=torch.float64)
    A[0, 2, 3] = 0.0
    A.requires_grad_(True)
    out, _ = m(xs, A)
    out.sum().backward()
    y = torch.nn.functional.conv2d(xs, m.conv.weight.detach()).view(2, 3, 6, 5, 7)
    Ar = A.detach().clone().requires_grad_(True)
    torch.einsum("nkctv,kvw->nctw", y, Ar).sum().backward()
    assert rel_l2(A.grad, Ar.grad) < 1e-12 and A.grad[0, 2, 3].abs() > 0
