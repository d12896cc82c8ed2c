"""Diagnostic: is one critic step deterministic?  eager vs eager, replay vs replay, eager vs replay, and the critic forward alone,
at the bench configuration (NTU-120 mlp8, tf32), batch 64."""
import os
import sys
from importlib import import_module

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kgan_b200 as kgan  # noqa: E402
from oracle import networks as onet  # noqa: E402
from helpers import inputs  # noqa: E402

wg = import_module("kinetic-gan_b200.wgan_gp")
CFG = onet.Config(dataset="ntu", n_classes=120, t_size=64, mlp_dim=8, channels=3)
n = 64
kgan.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32")
G = kgan.Generator(CFG.latent_dim, CFG.channels, CFG.n_classes, CFG.t_size, CFG.mlp_dim)
D = kgan.Discriminator(CFG.channels, CFG.n_classes, CFG.t_size, CFG.latent_dim)
pg = onet.synth_params(onet.g_param_shapes(CFG), 1)
for k in pg:
    if k.endswith("noise.weight"):
        pg[k] = torch.zeros_like(pg[k])
G.load_state_dict(pg)
D.load_state_dict(onet.synth_params(onet.d_param_shapes(CFG), 2))
G, D = G.cuda().train(), D.cuda()
tr = wg.WGANGPTrainer(G, D)
x = {k: v.cuda() for k, v in inputs(CFG, n, 51).items()}


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


with torch.no_grad():
    f1 = G(x["z"], x["labels"])
    f2 = G(x["z"], x["labels"])
    print("G fwd twice:", rel(f1, f2))
    d1 = D(f1, x["labels"])
    d2 = D(f1, x["labels"])
    print("D fwd twice:", rel(d1, d2))
l1 = tr._d_grads(x["real"], x["labels"], x["z"], x["alpha"])
g1 = tr.fd.grad.clone()
l2 = tr._d_grads(x["real"], x["labels"], x["z"], x["alpha"])
g2 = tr.fd.grad.clone()
print("eager vs eager: loss", l1[0].item(), l2[0].item(), "grad", rel(g2, g1))
tr.capture_graphs(x["real"], x["labels"], x["z"], x["alpha"])
gr = tr._graphs
outs = []
for _ in range(2):
    gr["d"].replay()
    torch.cuda.synchronize()
    outs.append((gr["d_out"][0].item(), tr.fd.grad.clone()))
print("replay vs replay: loss", outs[0][0], outs[1][0], "grad", rel(outs[1][1], outs[0][1]))
print("replay vs eager: grad", rel(outs[0][1], g1))
l3 = tr._d_grads(x["real"], x["labels"], x["z"], x["alpha"])
g3 = tr.fd.grad.clone()
print("eager after capture vs eager before: grad", rel(g3, g1), " vs replay", rel(g3, outs[0][1]))
off = {}
for k, p in D.named_parameters():
    o = (p.data_ptr() - tr.fd.flat.data_ptr()) >> 2
    a, b = outs[0][1][o:o + p.numel()], g1[o:o + p.numel()]
    print("   %-45s |g| %.3e  rel diff %.3e" % (k, b.norm().item(), ((a - b).norm() / b.norm().clamp_min(1e-30)).item()))
