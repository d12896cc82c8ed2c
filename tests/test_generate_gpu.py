"""generate.py's generator call (generate.py:85-93) through GeneratorRunner: CUDA-graph replay == eager eval-mode forward,
W-space truncation == the module's own `forward(z, labels, trunc)`, eval-mode output vs the fp64 oracle."""
from importlib import import_module

import numpy as np
import pytest
import torch

import kgan_b200 as kgan
from oracle import networks as onet
from oracle.graph import SkeletonTables
from helpers import CASES

pytestmark = pytest.mark.gpu


def build(cfg, zero_noise=True):
    G = kgan.Generator(cfg.latent_dim, cfg.channels, cfg.n_classes, cfg.t_size, cfg.mlp_dim, dataset=cfg.dataset)
    pg = onet.synth_params(onet.g_param_shapes(cfg), 1)
    if zero_noise:                  # the runner draws its own per-block noise: make the output independent of it
        pg = {k: (torch.zeros_like(v) if k.endswith("noise.weight") else v) for k, v in pg.items()}
    G.load_state_dict(pg)
    return G.cuda().eval(), pg


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_runner_graph_matches_eager_and_oracle(precision):
    gen = import_module("kinetic-gan_b200.generate")
    cfg = CASES["ntu_small"]["cfg"]
    kgan.set_precision(precision)
    try:
        G, pg = build(cfg)
        z, labels = gen.class_conditioned_batch(cfg.n_classes, 8, cfg.latent_dim, seed=3)
        n = z.shape[0]
        runner = gen.GeneratorRunner(G, n, cfg.latent_dim)
        out_g = runner(z.pin_memory(), labels.pin_memory()).clone()
        out_g2 = runner(z.cuda(), labels.cuda()).clone()              # second replay, device inputs
        with torch.no_grad():
            out_e = G(z.cuda(), labels.cuda())
        assert torch.equal(out_g, out_g2)
        assert (out_g - out_e).abs().max().item() < (1e-6 if precision == "fp32" else 1e-5)
        assert runner.launches_per_call > 0
        host = runner.to_host()
        torch.cuda.synchronize()
        assert torch.equal(host, out_g2.cpu())
        tables = SkeletonTables(cfg.dataset)
        pg64 = {k: (v.double() if v.is_floating_point() else v) for k, v in pg.items()}
        nz = [torch.zeros(*s, dtype=torch.float64) for s in onet.noise_shapes(cfg, n, tables)]
        ref = onet.generator_forward(pg64, z.double(), labels, cfg, tables, nz, training=False)
        rel = ((out_g.cpu().double() - ref).norm() / ref.norm()).item()
        assert rel < (2e-5 if precision == "fp32" else 5e-3), rel
    finally:
        kgan.set_precision("fp32")
        kgan.ops._persist.clear()
        kgan.ops._batches.clear()


def test_runner_w_truncation():
    gen = import_module("kinetic-gan_b200.generate")
    cfg = CASES["h36m_small"]["cfg"]
    G, _ = build(cfg)
    z, labels = gen.class_conditioned_batch(cfg.n_classes, 4, cfg.latent_dim, seed=5)
    runner = gen.GeneratorRunner(G, z.shape[0], cfg.latent_dim, trunc=0.7, graphs=False)
    np.random.seed(11)                                                 # truncate() draws its 1000 latents from the host RNG (generator.py:98)
    a = runner(z.cuda(), labels.cuda()).clone()
    np.random.seed(11)
    with torch.no_grad():
        b = G(z.cuda(), labels.cuda(), 0.7)
    assert (a - b).abs().max().item() < 1e-6
    np.random.seed(11)
    with torch.no_grad():
        c = G(z.cuda(), labels.cuda())
    assert (a - c).abs().max().item() > 1e-4                           # the truncation really moved the latents


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_w_truncation_vs_oracle(precision):
    """generator.py:97-108 on the GPU against the fp64 oracle fed the SAME 1000 host-RNG latents (not against the module itself):
    runner and module output vs oracle `generator_forward(..., trunc=0.7, trunc_latents=...)`, eval mode, noise weights zero."""
    gen = import_module("kinetic-gan_b200.generate")
    cfg = CASES["ntu_small"]["cfg"]
    kgan.set_precision(precision)
    try:
        G, pg = build(cfg)
        z, labels = gen.class_conditioned_batch(cfg.n_classes, 8, cfg.latent_dim, seed=7)
        n = z.shape[0]
        runner = gen.GeneratorRunner(G, n, cfg.latent_dim, trunc=0.7, graphs=False)
        np.random.seed(21)
        out = runner(z.cuda(), labels.cuda()).clone()
        np.random.seed(21)
        t_lat = torch.as_tensor(np.random.normal(0, 1, (1000, G.mlp.mlp[0].in_features)))          # the draw of generator.py:98
        tables = SkeletonTables(cfg.dataset)
        pg64 = {k: (v.double() if v.is_floating_point() else v) for k, v in pg.items()}
        nz = [torch.zeros(*s, dtype=torch.float64) for s in onet.noise_shapes(cfg, n, tables)]
        ref = onet.generator_forward(pg64, z.double(), labels, cfg, tables, nz, training=False, trunc=0.7, trunc_latents=t_lat.double())
        plain = onet.generator_forward(pg64, z.double(), labels, cfg, tables, nz, training=False)
        rel = ((out.cpu().double() - ref).norm() / ref.norm()).item()
        moved = ((plain - ref).norm() / ref.norm()).item()
        print("W-space truncation vs fp64 oracle (%s): rel-L2 %.2e (truncation moves the output by %.2e)" % (precision, rel, moved))
        assert rel < (2e-5 if precision == "fp32" else 5e-3), rel
        assert moved > 100 * rel
    finally:
        kgan.set_precision("fp32")
        kgan.ops._persist.clear()
        kgan.ops._batches.clear()


def test_runner_cached_w_mean_replays_as_graph():
    """cache_mean=True: the W-space mean is estimated once, the truncated pass is a CUDA-graph replay and equals the module's
    own truncated forward with that mean; the default keeps the reference's per-call estimate (test above)."""
    gen = import_module("kinetic-gan_b200.generate")
    cfg = CASES["h36m_small"]["cfg"]
    G, _ = build(cfg)
    z, labels = gen.class_conditioned_batch(cfg.n_classes, 4, cfg.latent_dim, seed=6)
    np.random.seed(12)
    runner = gen.GeneratorRunner(G, z.shape[0], cfg.latent_dim, trunc=0.7, cache_mean=True)
    try:
        a = runner(z.cuda(), labels.cuda()).clone()
        a2 = runner(z.cuda(), labels.cuda()).clone()
        assert runner.graphs and runner.launches_per_call > 0 and torch.equal(a, a2)
        with torch.no_grad():
            b = G(z.cuda(), labels.cuda(), 0.7)
            c = G(z.cuda(), labels.cuda())
        assert (a - b).abs().max().item() < 1e-6
        assert (a - c).abs().max().item() > 1e-4
        np.random.seed(12)
        assert torch.allclose(G._w_mean, G.estimate_w_mean(), atol=1e-6)
    finally:
        G._w_mean = None
