"""The critic's deep temporal conv (D4 / D5: 512 -> 512 channels, 3 taps, planes of 4-8 positions) alone: timing or one launch for ncu.
    python tools/one_layer_deep.py [--once] [--n 1024] [--t 8]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kgan_b200 as kgan  # noqa: E402

ops, G = kgan.ops, kgan.geometry
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1024)
ap.add_argument("--t", type=int, default=8)
ap.add_argument("--c", type=int, default=512)
ap.add_argument("--once", action="store_true")
ap.add_argument("--staged", type=int, default=0)
a = ap.parse_args()
kgan.set_precision("tf32")
G.STAGED_POLICY = "all" if a.staged else "fallback"
t_sel = list(range(0, a.t, 2))
unf = G.UnfoldedTcnGeom(a.c, a.c, a.t, 1, 3, 1, 1, 1, t_sel)
u = torch.randn(a.n, a.c, 3 * len(t_sel), 1, device="cuda")
w = torch.randn(a.c, a.c, 3, 1, device="cuda") / (3 * a.c) ** 0.5
go = torch.randn(a.n, a.c, len(t_sel), 1, device="cuda")
fns = {"fwd": lambda: ops.tapconv_fwd(u, w, unf.fwd), "dgrad": lambda: ops.tapconv_fwd(go, w, unf.dgrad),
       "wgrad": lambda: ops.tapconv_wgrad(u, go, unf.fwd, tuple(w.shape))}
for f in fns.values():
    f()
torch.cuda.synchronize()
if a.once:
    torch.cuda.profiler.start()
    for f in fns.values():
        f()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    for name, f in fns.items():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            f()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        fl = 2.0 * a.n * len(t_sel) * a.c * a.c * 3
        print("%s n=%d c=%d T=%d->%d: %.1f us  %.0f TFLOP/s" % (name, a.n, a.c, a.t, len(t_sel), us, fl / (us * 1e-6) / 1e12))
