"""Diagnostic: the trainer's critic step on the GPU (fp32 path) against the fp64 oracle, piece by piece."""
import os
import sys
from importlib import import_module

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kgan_b200 as kgan  # noqa: E402
from oracle import networks as onet  # noqa: E402
from oracle.graph import SkeletonTables  # noqa: E402
from helpers import CASES, draw_noises, inputs  # noqa: E402

wg = import_module("kinetic-gan_b200.wgan_gp")
case = sys.argv[1] if len(sys.argv) > 1 else "ntu_small"
cfg, n = CASES[case]["cfg"], CASES[case]["n"]
tables = SkeletonTables(cfg.dataset)
kgan.set_precision("fp32")
pg = onet.synth_params(onet.g_param_shapes(cfg), 1)
pd = onet.synth_params(onet.d_param_shapes(cfg), 2)
G = kgan.Generator(cfg.latent_dim, cfg.channels, cfg.n_classes, cfg.t_size, cfg.mlp_dim, dataset=cfg.dataset)
D = kgan.Discriminator(cfg.channels, cfg.n_classes, cfg.t_size, cfg.latent_dim, dataset=cfg.dataset)
G.load_state_dict(pg)
D.load_state_dict(pd)
G, D = G.cuda().train(), D.cuda()
pg64 = {k: (v.double() if v.is_floating_point() else v) for k, v in pg.items()}
pd64 = {k: v.double().requires_grad_(True) for k, v in pd.items()}


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


x = inputs(cfg, n, 10, torch.float32)
nz = draw_noises(cfg, n, 100)
xc = {k: v.cuda() for k, v in x.items()}
with torch.no_grad():
    fake = G(xc["z"], xc["labels"], noises=[t.cuda() for t in nz])
fake_ref = onet.generator_forward(pg64, x["z"].double(), x["labels"], cfg, tables, [t.double() for t in nz], True, {})
print("G out rel", rel(fake, fake_ref))
fr = fake_ref.detach()
v_real = D(xc["real"], xc["labels"])
v_fake = D(fake, xc["labels"])
v_cat = D(torch.cat((xc["real"], fake), 0), torch.cat((xc["labels"], xc["labels"]), 0))
r_real = onet.discriminator_forward(pd64, x["real"].double(), x["labels"], cfg, tables)
r_fake = onet.discriminator_forward(pd64, fr, x["labels"], cfg, tables)
print("D(real) rel", rel(v_real, r_real), " D(fake) rel", rel(v_fake, r_fake))
print("D(cat) rel  real half", rel(v_cat[:n], r_real), " fake half", rel(v_cat[n:], r_fake))
print("values real", v_real.flatten().tolist(), r_real.flatten().tolist())
gp = wg.compute_gradient_penalty(D, xc["real"], fake, xc["labels"], xc["alpha"])
gp_ref = onet.gradient_penalty(pd64, x["real"].double(), fr, x["labels"], x["alpha"].double(), cfg, tables)
print("gp", gp.item(), gp_ref.item())
d_loss = -v_cat[:n].mean() + v_cat[n:].mean() + cfg.lambda_gp * gp
d_ref = -r_real.mean() + r_fake.mean() + cfg.lambda_gp * gp_ref
print("d_loss", d_loss.item(), d_ref.item())
for p in D.parameters():
    p.grad = None
d_loss.backward()
gref = torch.autograd.grad(d_ref, [pd64[k] for k, _ in D.named_parameters()], allow_unused=True)
for (k, p), g in zip(D.named_parameters(), gref):
    if g is None:
        continue
    e = rel(p.grad, g) if g.norm() > 0 else p.grad.abs().max().item()
    print("  grad %-45s rel %.2e  |ref| %.2e" % (k, e, g.norm().item()))
