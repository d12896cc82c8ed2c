"""WGAN-GP trainer host logic on CPU (C-ABI primitives emulated): two iterations of kinetic-gan.py:137-174 against the
golden fixtures of the unmodified reference, and the 2-rank gloo data-parallel path against the single-process result."""
import os
import subprocess
import sys
from importlib import import_module

import numpy as np
import pytest
import torch

import kgan_b200 as kgan
from oracle import networks as onet
from helpers import CASES, draw_noises, inputs, load_golden, sub, within_noise_floor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(cfg, dtype):
    G = kgan.Generator(cfg.latent_dim, cfg.channels, cfg.n_classes, cfg.t_size, cfg.mlp_dim, dataset=cfg.dataset)
    D = kgan.Discriminator(cfg.channels, cfg.n_classes, cfg.t_size, cfg.latent_dim, dataset=cfg.dataset)
    G.load_state_dict(onet.synth_params(onet.g_param_shapes(cfg), 1))
    D.load_state_dict(onet.synth_params(onet.d_param_shapes(cfg), 2))
    if dtype == torch.float64:
        G, D = G.double(), D.double()
        for m in (G, D):
            for i, a in enumerate(m.graph.As):
                setattr(m, "_A%d" % i, torch.tensor(a, dtype=torch.float64))
    return G, D


@pytest.mark.parametrize("case", list(CASES))
def test_two_iterations_vs_golden(emu, case, monkeypatch):
    gold, cfg, n = load_golden(case), CASES[case]["cfg"], CASES[case]["n"]
    wg = import_module("kinetic-gan_b200.wgan_gp")
    tag, dtype = "f64", torch.float64
    # FlatParams allocates float32 buffers; run the host logic in float64 for a tight comparison
    monkeypatch.setattr(torch, "float32", torch.float64)
    G, D = build(cfg, dtype)
    G.train()
    tr = wg.WGANGPTrainer(G, D, cfg.lr, cfg.b1, cfg.b2, cfg.n_critic, cfg.lambda_gp)
    for i in range(2):
        xi = inputs(cfg, n, 10 + i, dtype)
        d_loss, g_loss, _ = tr.iteration(i, xi["real"], xi["labels"], xi["z"], xi["alpha"],
                                         draw_noises(cfg, n, 100 + 2 * i, dtype), draw_noises(cfg, n, 101 + 2 * i, dtype))
        assert abs(d_loss.item() - float(gold[tag + "/train/d_loss%d" % i])) < 1e-8
        if i == 0:
            assert abs(g_loss.item() - float(gold[tag + "/train/g_loss0"])) < 1e-8
        else:
            assert g_loss is None
    for net, m in (("g", G), ("d", D)):
        for k, v in m.state_dict().items():
            ref = gold[tag + "/train/%s_after/%s" % (net, k)]
            mine = sub(v) if v.numel() > 4096 else v.detach().double().numpy()
            assert np.abs(mine - ref).max() < 1e-7, k


def test_ddp_two_ranks_gloo(tmp_path):
    """world_size=2 over gloo: rank-sharded batch, flat-gradient all-reduce + 1/world in Adam == single process on the
    full batch for the critic (D has no BatchNorm, so its update is exactly the global-batch update)."""
    script = os.path.join(ROOT, "tests", "ddp_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", OMP_NUM_THREADS="2")
    procs = [subprocess.Popen([sys.executable, script, str(tmp_path)], env=dict(env, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r)))
             for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    single = subprocess.run([sys.executable, script, str(tmp_path)], env=dict(env, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0"), timeout=600)
    assert single.returncode == 0
    a = torch.load(os.path.join(str(tmp_path), "d_flat_w2_r0.pt"))
    b = torch.load(os.path.join(str(tmp_path), "d_flat_w2_r1.pt"))
    c = torch.load(os.path.join(str(tmp_path), "d_flat_w1_r0.pt"))
    assert torch.equal(a, b)                                   # replicas stay identical
    assert (a - c).abs().max().item() < 1e-9                   # == global-batch update


def test_critic_step_runs_no_unread_backward_work(emu):
    """Launch census of one critic step (kinetic-gan.py:137-154).  The gradient penalty's first-order pass asks for d/dx_hat
    only, and the forward nodes of D(x_hat) receive no gradient from the second-order graph (LeakyReLU masks are piecewise
    constant): no weight gradient, bias reduction or adjacency gradient may run for them, and no backward sweep on
    materialised zeros (functional.py: set_materialize_grads(False) / data_grads_only)."""
    import collections

    wg = import_module("kinetic-gan_b200.wgan_gp")
    ops = kgan.ops
    cfg, n = CASES["ntu_small"]["cfg"], 3
    G, D = build(cfg, torch.float32)
    tr = wg.WGANGPTrainer(G, D)
    cnt = collections.Counter()
    for name in ("tapconv_fwd", "tapconv_wgrad", "adjmix_bwd_a", "adjmix_bwd_x", "chan_reduce", "act_bwd"):
        f = getattr(ops, name)

        def wrap(*a, _f=f, _n=name, **k):
            cnt[(_n, a[0].shape[0])] += 1
            return _f(*a, **k)

        setattr(ops, name, wrap)
    xi = inputs(cfg, n, 3)
    with torch.no_grad():                 # the generator's own launches (batch n as well): counted alone, then subtracted
        G(xi["z"], xi["labels"])
    gen_only = collections.Counter(cnt)
    cnt.clear()
    tr._d_grads(xi["real"], xi["labels"], xi["z"], xi["alpha"])
    cnt.subtract(gen_only)
    convs = 6 + 6 + 4 + 1                 # graph convs, temporal convs, residual convs, head (discriminator.py:29-34,47-50)
    # concatenated real+fake pass (batch 2n): one weight gradient per conv (+ the label-fold GEMM of the first layer)
    assert cnt[("tapconv_wgrad", 2 * n)] == convs + 1
    # x_hat path (batch n): one weight gradient per conv, from the second-order graph only
    assert cnt[("tapconv_wgrad", n)] == convs
    assert cnt[("adjmix_bwd_a", n)] == 6 and cnt[("adjmix_bwd_x", n)] == 6
    assert cnt[("chan_reduce", n)] == 0   # d(penalty)/d(bias) == 0 exactly: never computed
    # LeakyReLU slopes: the first-order pass applies those of blocks 0-4 inside the next block's fused input-gradient kernel
    # (functional.GcnRes -> kgan_adjmix_bwd_x_fused) and only the last block's with a mask kernel; the second-order pass applies each once
    assert cnt[("act_bwd", n)] == 1 + 6
