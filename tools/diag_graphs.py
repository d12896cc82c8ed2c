"""Diagnostic: CUDA-graph replay of the critic step vs eager launches, per iteration and per parameter."""
import os
import sys
from importlib import import_module

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kgan_b200 as kgan  # noqa: E402
from oracle import networks as onet  # noqa: E402
from helpers import CASES, inputs  # noqa: E402

wg = import_module("kinetic-gan_b200.wgan_gp")
cfg, n = CASES["ntu_small"]["cfg"], 8
kgan.set_precision(sys.argv[1] if len(sys.argv) > 1 else "fp32")


def build():
    G = kgan.Generator(cfg.latent_dim, cfg.channels, cfg.n_classes, cfg.t_size, cfg.mlp_dim, dataset=cfg.dataset)
    D = kgan.Discriminator(cfg.channels, cfg.n_classes, cfg.t_size, cfg.latent_dim, dataset=cfg.dataset)
    G.load_state_dict(onet.synth_params(onet.g_param_shapes(cfg), 1))
    D.load_state_dict(onet.synth_params(onet.d_param_shapes(cfg), 2))
    return G.cuda().train(), D.cuda()


def run(mode):
    G, D = build()
    tr = wg.WGANGPTrainer(G, D, cfg.lr, cfg.b1, cfg.b2, cfg.n_critic, cfg.lambda_gp)
    x0 = {k: v.cuda() for k, v in inputs(cfg, n, 20, torch.float32).items()}
    if mode == "graph":
        tr.capture_graphs(x0["real"], x0["labels"], x0["z"], x0["alpha"])
    out = []
    for i in range(1, 4):
        xi = {k: v.cuda() for k, v in inputs(cfg, n, 20 + i, torch.float32).items()}
        d_loss, _, gp = tr.iteration(i, xi["real"], xi["labels"], xi["z"], xi["alpha"])
        torch.cuda.synchronize()
        out.append((d_loss.item(), gp.item(), tr.fd.grad.clone(), tr.fd.flat.clone()))
    names = []
    off = 0
    for (k, p) in D.named_parameters():
        names.append((k, p.data_ptr() - tr.fd.flat.data_ptr() >> 2, p.numel()))
    return out, names


a, names = run("eager")
b, _ = run("eager")
c, _ = run("graph")
for i in range(3):
    print("iter %d  d_loss eager %.7f eager2 %.7f graph %.7f | gp %.7f %.7f %.7f" % (i + 1, a[i][0], b[i][0], c[i][0], a[i][1], b[i][1], c[i][1]))
    ga, gb, gc = a[i][2], b[i][2], c[i][2]
    print("   grad rel diff eager-eager2 %.3e  eager-graph %.3e" % (((ga - gb).norm() / ga.norm()).item(), ((ga - gc).norm() / ga.norm()).item()))
    if i == 0:
        for k, o, m in names:
            da = (ga[o:o + m] - gc[o:o + m])
            print("     %-45s |g| %.3e  diff %.3e  min|g| %.3e" % (k, ga[o:o + m].norm().item(), da.norm().item(), ga[o:o + m].abs().min().item()))
    pa, pb, pc = a[i][3], b[i][3], c[i][3]
    print("   param max diff eager-eager2 %.3e  eager-graph %.3e  mean %.3e %.3e" % ((pa - pb).abs().max().item(), (pa - pc).abs().max().item(),
                                                                                   (pa - pb).abs().mean().item(), (pa - pc).abs().mean().item()))
