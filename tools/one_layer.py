"""One tap convolution, both tensor-core producers (TMA-fed vs operand-building), CUDA-event timing or a single launch for ncu.
    python tools/one_layer.py [--staged 0|1] [--n 1024] [--once]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kgan_b200 as kgan  # noqa: E402

ops, G = kgan.ops, kgan.geometry
ap = argparse.ArgumentParser()
ap.add_argument("--staged", type=int, default=1)
ap.add_argument("--n", type=int, default=1024)
ap.add_argument("--c", type=int, default=64)
ap.add_argument("--v", type=int, default=12)
ap.add_argument("--kt", type=int, default=3)
ap.add_argument("--once", action="store_true")
a = ap.parse_args()
kgan.set_precision("tf32")
G.STAGED_POLICY = "all" if a.staged else "fallback"
geom = G.TapConvGeom(a.c, a.c, 64, a.v, kt=a.kt, pad=a.kt // 2)
x = torch.randn(a.n, a.c, 64, a.v, device="cuda")
w = torch.randn(a.c, a.c, a.kt, 1, device="cuda") / 14
ops.tapconv_fwd(x, w, geom.fwd)
torch.cuda.synchronize()
if a.once:
    torch.cuda.profiler.start()
    ops.tapconv_fwd(x, w, geom.fwd)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.tapconv_fwd(x, w, geom.fwd)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    gb = (x.numel() + a.n * a.c * 64 * a.v) * 4 / 1e9
    print("staged=%d n=%d c=%d v=%d kt=%d: %.1f us  %.0f GB/s" % (a.staged, a.n, a.c, a.v, a.kt, us, gb / (us * 1e-6)))
