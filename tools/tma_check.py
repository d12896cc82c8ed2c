"""Bring-up check of the TMA-fed tap convolution against the float64 statement of the descriptor semantics.
    python tools/tma_check.py            (prints one line per geometry; never stops at the first mismatch)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emu_backend as emu  # noqa: E402
import kgan_b200 as kgan  # noqa: E402
from importlib import import_module  # noqa: E402

ops, G = kgan.ops, kgan.geometry
_lib = import_module("kinetic-gan_b200._lib")
kgan.set_precision("tf32")
GEOMS = {
    "d1_gcn": (dict(c_in=32, c_out=64, t_in=64, v_in=11, K=3), 4),
    "d1_tcn_v11": (dict(c_in=64, c_out=64, t_in=64, v_in=11, kt=3, pad=1), 4),
    "d1_tcn_v12": (dict(c_in=64, c_out=64, t_in=64, v_in=12, kt=3, pad=1), 4),
    "d2_tcn_unf": (dict(unfold=True, c_in=128, c_out=128, t_in=64, v_in=5, kt=3, pad=1, stride=1, dil=1, t_sel=list(range(0, 64, 2))), 5),
    "d3_tcn_unf": (dict(unfold=True, c_in=256, c_out=256, t_in=32, v_in=5, kt=3, pad=1, stride=1, dil=1, t_sel=list(range(0, 32, 2))), 6),
    "d4_tcn_unf": (dict(unfold=True, c_in=512, c_out=512, t_in=16, v_in=1, kt=3, pad=1, stride=1, dil=1, t_sel=list(range(0, 16, 2))), 160),
    "d5_tcn_unf": (dict(unfold=True, c_in=512, c_out=512, t_in=8, v_in=1, kt=3, pad=1, stride=1, dil=1, t_sel=[0, 2, 4, 6]), 130),
    "tcn_v4": (dict(c_in=32, c_out=32, t_in=16, v_in=4, kt=3, pad=1), 8),
    "tcn_v8": (dict(c_in=32, c_out=32, t_in=16, v_in=8, kt=3, pad=1), 8),
    "d0_tcn": (dict(c_in=32, c_out=32, t_in=64, v_in=11, kt=3, pad=1), 3),
    "d2_gcn": (dict(c_in=64, c_out=128, t_in=64, v_in=5, K=3), 3),
    "d3_gcn": (dict(c_in=128, c_out=256, t_in=32, v_in=5, K=3), 5),
    "d4_gcn_p80": (dict(c_in=256, c_out=512, t_in=16, v_in=5, K=3), 7),
    "d4_gcn_p16": (dict(c_in=256, c_out=512, t_in=16, v_in=1, K=3), 37),
    "d5_gcn_p8": (dict(c_in=512, c_out=512, t_in=8, v_in=1, K=3), 70),
    "d5_tcn_p8": (dict(c_in=512, c_out=512, t_in=8, v_in=1, kt=3, pad=1), 70),
    "g6_tcn_3ch": (dict(c_in=3, c_out=3, t_in=64, v_in=25, kt=3, pad=1), 4),
    "d0_gcn_3ch": (dict(c_in=3, c_out=32, t_in=64, v_in=25, K=3, w_cin=123, w_ic0=120), 4),
    "g2_gcn_p20": (dict(c_in=256, c_out=128, t_in=4, v_in=5, K=3), 64),
    "d2_tcn_s2": (dict(c_in=128, c_out=128, t_in=64, v_in=5, kt=3, pad=1, t_sel=list(range(0, 64, 2))), 5),
    "d4_tcn_s2": (dict(c_in=512, c_out=512, t_in=16, v_in=1, kt=3, pad=1, t_sel=list(range(0, 16, 2))), 160),
    "d2_res_sel": (dict(c_in=64, c_out=128, t_in=64, v_in=11, t_sel=list(range(0, 64, 2)), v_keep=[2, 4, 6, 8, 10]), 5),
}


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32)


def rel(a, b):
    b = b.double()
    return ((a.detach().cpu().double() - b).norm() / b.norm().clamp_min(1e-30)).item()


only = sys.argv[1] if len(sys.argv) > 1 else ""
bad = 0
for name, (kw, n) in GEOMS.items():
    if only and only not in name:
        continue
    kw = dict(kw)
    if kw.pop("unfold", False):
        geom = G.UnfoldedTcnGeom(**kw)
        x = rnd(n, geom.c_in, geom.kt * geom.t_out, geom.v_in, seed=1)
    else:
        geom = G.TapConvGeom(**kw)
        x = rnd(n, geom.K * geom.c_in, geom.t_in, geom.v_in, seed=1)
    w = rnd(geom.K * geom.c_out, kw.get("w_cin", geom.c_in), geom.kt, 1, seed=2) / np.sqrt(geom.c_in * geom.kt * geom.K)
    bias, add = rnd(geom.c_out, seed=3), rnd(n, geom.c_out, geom.t_out, geom.v_out, seed=4)
    go = rnd(n, geom.c_out, geom.t_out, geom.v_out, seed=5)
    got = ops.tapconv_wgrad(x.cuda(), go.cuda(), geom.fwd, tuple(w.shape))
    ew = rel(got, emu.tapconv_wgrad(x.double(), go.double(), geom.fwd, tuple(w.shape)))
    print("%-12s wgrad rel=%.2e %s" % (name, ew, "ok" if ew < 1e-3 else "MISMATCH"), flush=True)
    bad += not ew < 1e-3
    for what, desc, inp in (("fwd", geom.fwd, x), ("dgrad", geom.dgrad, go)):
        tma = bool(_lib.lib().kgan_tapconv_tma_ok(desc.cstruct(n, 0, 1)))
        got = ops.tapconv_fwd(inp.cuda(), w.cuda(), desc)
        torch.cuda.synchronize()
        e = rel(got, emu.tapconv_fwd(inp.double(), w.double(), desc))
        e2 = float("nan")
        if what == "fwd":
            got = ops.tapconv_fwd(inp.cuda(), w.cuda(), desc, bias.cuda(), add.cuda(), ops.ACT_LRELU)
            e2 = rel(got, emu.tapconv_fwd(inp.double(), w.double(), desc, bias.double(), add.double(), ops.ACT_LRELU))
        ok = e < 1e-3 and not e2 >= 1e-3
        bad += not ok
        print("%-12s %-5s tma_mode=%d tma=%d  rel=%.2e  rel_epilogue=%.2e  %s" % (name, what, desc.tma_mode, tma, e, e2, "ok" if ok else "MISMATCH"), flush=True)
print("mismatches:", bad)
