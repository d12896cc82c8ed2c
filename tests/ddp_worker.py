"""Worker of tests/test_trainer_cpu.py::test_ddp_two_ranks_gloo (CPU, gloo, C-ABI primitives emulated)."""
import os
import sys
from importlib import import_module

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
torch.set_num_threads(2)
torch.float32 = torch.float64          # host logic in float64 so the comparison is tight

import emu_backend  # noqa: E402
import kgan_b200 as kgan  # noqa: E402
from oracle import networks as onet  # noqa: E402
from helpers import draw_noises, inputs  # noqa: E402


class _MP:
    def setattr(self, obj, name, val):
        setattr(obj, name, val)


emu_backend.install(_MP())
wg = import_module("kinetic-gan_b200.wgan_gp")
ddp = import_module("kinetic-gan_b200.ddp")
comm = ddp.Comm(backend="gloo")
cfg = onet.Config(dataset="h36m", n_classes=4, t_size=16, mlp_dim=1, channels=2)
G = kgan.Generator(cfg.latent_dim, cfg.channels, cfg.n_classes, cfg.t_size, cfg.mlp_dim, dataset=cfg.dataset).double()
D = kgan.Discriminator(cfg.channels, cfg.n_classes, cfg.t_size, cfg.latent_dim, dataset=cfg.dataset).double()
torch.manual_seed(100 + comm.rank)         # different init per rank: broadcast must fix it
for p in list(G.parameters()) + list(D.parameters()):
    p.data.normal_(0, 0.05)
if comm.world_size == 1:                   # single-process reference run uses rank 0's init
    torch.manual_seed(100)
    for p in list(G.parameters()) + list(D.parameters()):
        p.data.normal_(0, 0.05)
for m in (G, D):
    for i, a in enumerate(m.graph.As):
        setattr(m, "_A%d" % i, torch.tensor(a, dtype=torch.float64))
tr = wg.WGANGPTrainer(G, D, comm=comm)
n = 4
x = inputs(cfg, n, 7, torch.float64)
lo, hi = comm.shard(n)
fake = torch.randn(n, cfg.channels, cfg.t_size, 16, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
# critic-only update on a fixed fake batch: loss = mean over the GLOBAL batch, so each rank scales by its share
tr.fd.zero_grad()
rv, fv = D(x["real"][lo:hi], x["labels"][lo:hi]), D(fake[lo:hi], x["labels"][lo:hi])
gp = wg.compute_gradient_penalty(D, x["real"][lo:hi], fake[lo:hi], x["labels"][lo:hi], alpha=x["alpha"][lo:hi])
(-rv.mean() + fv.mean() + 10 * gp).backward()
tr._reduce(tr.fd)
tr.fd.adam(2e-4, 0.5, 0.999, grad_scale=1.0 / comm.world_size)
torch.save(tr.fd.flat.clone(), os.path.join(sys.argv[1], "d_flat_w%d_r%d.pt" % (comm.world_size, comm.rank)))
comm.close()
