"""Pins the oracle (oracle/) against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only.  fp64 fixtures are matched to ~1e-12, fp32 to ~1e-5."""
import numpy as np
import pytest
import torch

from oracle import networks as onet
from oracle.graph import SkeletonTables
from helpers import CASES, draw_noises, inputs, load_golden, rel_l2, sub, within_noise_floor

TOL = {"f64": 1e-11, "f32": 2e-5}
DT = {"f64": torch.float64, "f32": torch.float32}


@pytest.mark.parametrize("name", ["ntu", "h36m"])
def test_graph_tables(name):
    g = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "graph_%s.npz" % name))
    t = SkeletonTables(name)
    assert t.num_node == g["num_node"].tolist() and t.center == g["center"].tolist()
    for l in range(4):
        assert np.array_equal(t.As[l], g["As%d" % l])
        assert np.array_equal(t.map[l], g["map%d" % l])
        assert np.array_equal(np.asarray(t.edge[l]), g["edge%d" % l])
    for l in range(3):
        assert len(t.mapping[l]) == int(g["mapping%d_len" % l])
        for j, h in enumerate(t.mapping[l]):
            assert np.array_equal(h, g["mapping%d_%d" % (l, j)])
        # partitions are column-stochastic
    for l in range(4):
        assert np.allclose(t.As[l].sum((0, 1)), 1.0)


def _params(cfg, dtype):
    # the golden run synthesises in float32 and casts (make_golden.build_reference)
    pg = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in onet.synth_params(onet.g_param_shapes(cfg), 1).items()}
    pd = {k: v.to(dtype) for k, v in onet.synth_params(onet.d_param_shapes(cfg), 2).items()}
    return pg, pd


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_generator(case, tag):
    gold, cfg, n, dtype = load_golden(case), CASES[case]["cfg"], CASES[case]["n"], DT[tag]
    tables = SkeletonTables(cfg.dataset)
    pg, _ = _params(cfg, dtype)
    for k in pg:
        if onet.is_trainable(k):
            pg[k].requires_grad_(True)
    x = inputs(cfg, n, 0, dtype)
    blocks, st = [], {}
    fake = onet.generator_forward(pg, x["z"], x["labels"], cfg, tables, draw_noises(cfg, n, 11, dtype), True, st,
                                  collect=blocks)
    assert rel_l2(fake, gold[tag + "/g_out"]) < TOL[tag]
    for i, b in enumerate(blocks):
        assert rel_l2(b, gold[tag + "/g_block%d" % i]) < TOL[tag], i
    keys = [k for k in pg if onet.is_trainable(k)]
    grads = torch.autograd.grad((fake * x["cot_g"]).sum(), [pg[k] for k in keys])
    for k, g in zip(keys, grads):
        assert within_noise_floor(sub(g), gold, "/g_grad/" + k, tag, 50 * TOL[tag]), k
    for k, v in st.items():
        if "running" in k:
            assert rel_l2(v, gold[tag + "/g_bn_after/" + k]) < TOL[tag], k
    # per-sample mapping loop (generator.py:84-85) == batched mapping
    fake2 = onet.generator_forward(pg, x["z"], x["labels"], cfg, tables, draw_noises(cfg, n, 11, dtype), True, {},
                                   per_sample_loop=True)
    assert rel_l2(fake2, fake) < TOL[tag]
    # eval mode and W-space truncation
    ev = onet.generator_forward(pg, x["z"], x["labels"], cfg, tables, draw_noises(cfg, n, 12, dtype), False)
    assert rel_l2(ev, gold[tag + "/g_out_eval"]) < TOL[tag]
    np.random.seed(5)
    lat = torch.as_tensor(np.random.normal(0, 1, (1000, cfg.latent_dim + cfg.n_classes))).to(dtype)
    tr = onet.generator_forward(pg, x["z"], x["labels"], cfg, tables, draw_noises(cfg, n, 13, dtype), False,
                                trunc=0.95, trunc_latents=lat)
    assert rel_l2(tr, gold[tag + "/g_out_trunc"]) < 5 * TOL[tag]


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_discriminator_and_gp(case, tag):
    gold, cfg, n, dtype = load_golden(case), CASES[case]["cfg"], CASES[case]["n"], DT[tag]
    tables = SkeletonTables(cfg.dataset)
    _, pd = _params(cfg, dtype)
    for k in pd:
        pd[k].requires_grad_(True)
    x = inputs(cfg, n, 0, dtype)
    xr = x["real"].clone().requires_grad_(True)
    blocks = []
    dv = onet.discriminator_forward(pd, xr, x["labels"], cfg, tables, collect=blocks)
    assert rel_l2(dv, gold[tag + "/d_out"]) < TOL[tag]
    for i, b in enumerate(blocks):
        assert rel_l2(b, gold[tag + "/d_block%d" % i]) < TOL[tag], i
    keys = list(pd)
    grads = torch.autograd.grad((dv * x["cot_d"]).sum(), [xr] + [pd[k] for k in keys])
    assert rel_l2(grads[0], gold[tag + "/d_grad_x"]) < 10 * TOL[tag]
    for k, g in zip(keys, grads[1:]):
        assert within_noise_floor(sub(g), gold, "/d_grad/" + k, tag, 50 * TOL[tag]), k
    # gradient penalty with the fake batch of the golden run
    fake = torch.as_tensor(gold[tag + "/g_out"]).to(dtype)
    gp, gx = onet.gradient_penalty(pd, x["real"], fake, x["labels"], x["alpha"], cfg, tables, return_grad=True)
    assert abs(gp.item() - float(gold[tag + "/gp"])) < 20 * TOL[tag] * max(1.0, abs(float(gold[tag + "/gp"])))
    assert rel_l2(gx, gold[tag + "/gp_grads_x"]) < 10 * TOL[tag]
    grads = torch.autograd.grad(gp, [pd[k] for k in keys], allow_unused=True)
    for k, g in zip(keys, grads):
        ref = gold[tag + "/gp_grad/" + k]
        g = torch.zeros_like(pd[k]) if g is None else g
        if k.endswith("bias") or k == "label_emb.weight":   # the penalty cannot see them
            assert np.abs(ref).max() == 0 and g.abs().max().item() == 0, k
        else:
            assert within_noise_floor(sub(g), gold, "/gp_grad/" + k, tag, 100 * TOL[tag]), k


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_training_iterations(case, tag):
    """Two iterations of kinetic-gan.py:137-174 (i=0: D and G step, i=1: D step only) incl. Adam."""
    gold, cfg, n, dtype = load_golden(case), CASES[case]["cfg"], CASES[case]["n"], DT[tag]
    pg, pd = _params(cfg, dtype)
    tr = onet.Trainer(cfg, pg, pd)
    for i in range(2):
        xi = inputs(cfg, n, 10 + i, dtype)
        nd = draw_noises(cfg, n, 100 + 2 * i, dtype)
        ng = draw_noises(cfg, n, 101 + 2 * i, dtype)
        d_loss, g_loss, _ = tr.iteration(i, xi["real"], xi["labels"], xi["z"], xi["alpha"], nd, ng)
        # after an Adam step the fp32 trajectory carries the reference's own fp32 noise (Adam turns
        # rounding-level gradients into +-lr steps): fp32 is held to the reference's fp32-vs-fp64 distance
        assert within_noise_floor([d_loss.item()], gold, "/train/d_loss%d" % i, tag, 50 * TOL[tag]), (i, d_loss)
        if i == 0:
            assert within_noise_floor([g_loss.item()], gold, "/train/g_loss0", tag, 50 * TOL[tag]), g_loss
    # Adam's first steps move every weight by ~lr*sign(grad): compare parameter *deltas* loosely and values tightly
    for net, p in (("g", tr.pg), ("d", tr.pd)):
        for k, v in p.items():
            ref = gold[tag + "/train/%s_after/%s" % (net, k)]
            mine = sub(v) if v.numel() > 4096 else v.detach().double().numpy()
            if k.endswith("num_batches_tracked"):
                assert int(mine) == int(ref), k
                continue
            assert (np.abs(mine - ref).max() < (1e-7 if tag == "f64" else 4.5e-4)
                    or within_noise_floor(mine, gold, "/train/%s_after/%s" % (net, k), tag, 1e-4)), k
