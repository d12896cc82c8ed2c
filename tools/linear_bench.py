"""Times of the Linear layers (mapping network, first generator block, critic head) and of small-batch layers where the channel split
matters: python tools/linear_bench.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import kgan_b200 as kgan  # noqa: E402

ops, G = kgan.ops, kgan.geometry
kgan.set_precision("tf32")


def timeit(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


tot = 0.0
for name, kw, n in (("mapping 632->632", dict(c_in=632, c_out=632, t_in=1, v_in=1), 4096), ("g0 gcn 632->3x512", dict(c_in=632, c_out=512, t_in=1, v_in=1, K=3), 4096),
                    ("g0 tcn 512->512", dict(c_in=512, c_out=512, t_in=1, v_in=1, kt=3, pad=1), 4096), ("g1 gcn 512->3x256 4x1", dict(c_in=512, c_out=256, t_in=4, v_in=1, K=3), 4096),
                    ("D5 gcn 3x512->512 8x1 n=1024", dict(c_in=512, c_out=512, t_in=8, v_in=1, K=3), 1024), ("D4 gcn 3x256->512 16x1 n=512", dict(c_in=256, c_out=512, t_in=16, v_in=1, K=3), 512),
                    ("D1 tcn 64->64 64x12 n=16", dict(c_in=64, c_out=64, t_in=64, v_in=12, kt=3, pad=1), 16)):
    geom = G.TapConvGeom(**kw)
    x = ops.round_tf32(torch.randn(n, geom.K * geom.c_in if "gcn 3x" in name else geom.c_in, geom.t_in, geom.v_in, device="cuda"))
    if x.shape[1] != geom.fwd.c_in_total:
        x = ops.round_tf32(torch.randn(n, geom.fwd.c_in_total, geom.t_in, geom.v_in, device="cuda"))
    w = torch.randn(geom.K * geom.c_out, geom.c_in, geom.kt, 1, device="cuda") / 10
    us = timeit(lambda: ops.tapconv_fwd(x, w, geom.fwd))
    tot += us
    print("%-32s n=%-5d %7.1f us" % (name, n, us))
print("sum %.1f us" % tot)
