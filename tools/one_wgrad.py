"""One weight-gradient launch of a temporal conv (for ncu): python tools/one_wgrad.py [--c 32] [--n 8192] [--once]"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import kgan_b200 as kgan  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--c", type=int, default=32)
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--v", type=int, default=12)
ap.add_argument("--once", action="store_true")
a = ap.parse_args()
ops, G = kgan.ops, kgan.geometry
kgan.set_precision("tf32")
geom = G.TapConvGeom(a.c, a.c, 64, a.v, kt=3, pad=1)
x = torch.randn(a.n, a.c, 64, a.v, device="cuda")
go = torch.randn(a.n, a.c, 64, a.v, device="cuda")
ops.tapconv_wgrad(x, go, geom.fwd, (a.c, a.c, 3, 1))
torch.cuda.synchronize()
if a.once:
    torch.cuda.profiler.start()
    ops.tapconv_wgrad(x, go, geom.fwd, (a.c, a.c, 3, 1))
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
