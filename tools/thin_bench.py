"""Times of the small-contraction forward kernel on the two 9-output layers of the step (D0 graph conv data gradient, the generator's
convolution-first graph conv): python tools/thin_bench.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import kgan_b200 as kgan  # noqa: E402

ops, G = kgan.ops, kgan.geometry
kgan.set_precision("tf32")


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, geom, n, dgrad in (("D0 gcn dgrad 32 -> 3x3, 64x12", G.TapConvGeom(c_in=3, c_out=32, t_in=64, v_in=12, K=3), 8192, True),
                             ("G conv-first 32 -> 9, 32x11", G.TapConvGeom(c_in=32, c_out=9, t_in=32, v_in=11, kt=1), 4096, False),
                             ("G tail 3 -> 3, 3 taps, 64x25", G.TapConvGeom(c_in=3, c_out=3, t_in=64, v_in=25, kt=3, pad=1), 4096, False)):
    w = torch.randn(geom.K * geom.c_out, geom.c_in, geom.kt, 1, device="cuda")
    if dgrad:
        x = torch.randn(n, geom.c_out, geom.t_out, geom.v_out, device="cuda")
        us = timeit(lambda: ops.tapconv_fwd(x, w, geom.dgrad))
    else:
        x = torch.randn(n, geom.K * geom.c_in, geom.t_in, geom.v_in, device="cuda")
        us = timeit(lambda: ops.tapconv_fwd(x, w, geom.fwd))
    print("%-34s n=%d: %.1f us" % (name, n, us))
