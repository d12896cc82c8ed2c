"""Shared test helpers: golden loading, seeded inputs (must match tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

from oracle import networks as onet
from oracle.graph import SkeletonTables

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SUB = 211

CASES = {
    "ntu_small": dict(cfg=onet.Config(dataset="ntu", n_classes=6, t_size=16, mlp_dim=2, channels=3), n=3),
    "h36m_small": dict(cfg=onet.Config(dataset="h36m", n_classes=4, t_size=32, mlp_dim=3, channels=2), n=2),
}


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def inputs(cfg, n, seed, dtype=torch.float32):
    v = SkeletonTables(cfg.dataset).num_node[0]
    g = torch.Generator().manual_seed(1000 + seed)
    real = torch.rand(n, cfg.channels, cfg.t_size, v, generator=g, dtype=torch.float64) * 2 - 1
    z = torch.randn(n, cfg.latent_dim, generator=g, dtype=torch.float64)
    labels = torch.randint(0, cfg.n_classes, (n,), generator=g)
    alpha = torch.rand(n, 1, 1, 1, generator=g, dtype=torch.float64)
    cot_g = torch.randn(n, cfg.channels, cfg.t_size, v, generator=g, dtype=torch.float64)
    cot_d = torch.randn(n, 1, generator=g, dtype=torch.float64)
    return dict(real=real.to(dtype), z=z.to(dtype), labels=labels, alpha=alpha.to(dtype),
                cot_g=cot_g.to(dtype), cot_d=cot_d.to(dtype))


def draw_noises(cfg, n, seed, dtype=torch.float32):
    """The tensors the reference draws at generator.py:179 after torch.manual_seed(seed) (CPU generator)."""
    torch.manual_seed(seed)
    return [torch.randn(*s).to(dtype) for s in onet.noise_shapes(cfg, n)]


def sub(t):
    return t.detach().reshape(-1)[::SUB].double().cpu().numpy()


def rel_l2(a, b):
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a.detach().cpu()).double()
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b.detach().cpu()).double()
    den = b.norm().item()
    num = (a - b).norm().item()
    return num / den if den > 0 else num


def within_noise_floor(mine, gold, key, tag, tol, floor_mult=4.0):
    """fp64 fixtures: plain rel-L2.  fp32 fixtures: ill-conditioned quantities (e.g. the gradient of a
    scale that a following BatchNorm removes) are allowed the reference's OWN fp32-vs-fp64 error."""
    ref = gold[tag + key]
    mine = np.asarray(mine, dtype=np.float64)
    err = np.linalg.norm(mine - ref)
    den = np.linalg.norm(ref)
    if err <= tol * den or np.abs(mine - ref).max() < 1e-9:
        return True
    if tag == "f32" and ref.size <= 3 and np.abs(mine - ref).max() < 2e-4:
        # single-element subsample of a cancellation residue (scale in front of a BatchNorm): absolute check
        return True
    if tag == "f32":
        r64 = gold["f64" + key]
        floor = np.linalg.norm(ref - r64)
        return np.linalg.norm(mine - r64) <= floor_mult * floor + tol * np.linalg.norm(r64)
    return False


def parity_ok(mine, gold, key, tol, floor_mult=4.0):
    """fp32 CUDA result vs the fp64 ground truth of the reference: within `tol` rel-L2, or - for ill-conditioned
    quantities (BatchNorm over a handful of samples) - within floor_mult x the reference's OWN fp32-vs-fp64 distance."""
    mine = np.asarray(mine.detach().cpu() if torch.is_tensor(mine) else mine, dtype=np.float64)
    r64, r32 = gold["f64" + key], gold["f32" + key]
    err = np.linalg.norm(mine - r64)
    return err <= tol * np.linalg.norm(r64) or err <= floor_mult * np.linalg.norm(r32 - r64) + tol * np.linalg.norm(r64)
