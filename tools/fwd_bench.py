"""Times of the TMA-fed forward kernel on the un-strided layers of the critic at n = 8192: python tools/fwd_bench.py"""
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


n = 8192
tot = 0.0
for name, kw in (("D0 tcn 32->32 3 taps 64x12", dict(c_in=32, c_out=32, t_in=64, v_in=12, kt=3, pad=1)),
                 ("D1 tcn 64->64 3 taps 64x12", dict(c_in=64, c_out=64, t_in=64, v_in=12, kt=3, pad=1)),
                 ("D1 gcn 3x32->64 64x12", dict(c_in=32, c_out=64, t_in=64, v_in=12, K=3)),
                 ("D2 gcn 3x64->128 64x5", dict(c_in=64, c_out=128, t_in=64, v_in=5, K=3)),
                 ("D3 gcn 3x128->256 32x5", dict(c_in=128, c_out=256, t_in=32, v_in=5, K=3)),
                 ("D0 gcn 3x3->32 64x12", dict(c_in=3, c_out=32, t_in=64, v_in=12, K=3)),
                 ("D5 gcn 3x512->512 8x1", dict(c_in=512, c_out=512, t_in=8, v_in=1, K=3))):
    geom = G.TapConvGeom(**kw)
    x = ops.round_tf32(torch.randn(n, geom.K * geom.c_in, geom.t_in, geom.v_in, device="cuda"))
    w = torch.randn(geom.K * geom.c_out, geom.c_in, geom.kt, 1, device="cuda") / 10
    go = ops.round_tf32(torch.randn(n, geom.c_out, geom.t_out, geom.v_out, device="cuda"))
    f = timeit(lambda: ops.tapconv_fwd(x, w, geom.fwd))
    dg = timeit(lambda: ops.tapconv_fwd(go, w, geom.dgrad))
    wg = timeit(lambda: ops.tapconv_wgrad(x, go, geom.fwd, tuple(w.shape)))
    gb = (x.numel() + go.numel()) * 4 / 1e9
    tot += f + dg + wg
    print("%-28s fwd %7.1f us %5.0f GB/s | dgrad %7.1f us %5.0f GB/s | wgrad %7.1f us %5.0f GB/s" % (name, f, gb / f * 1e6, dg, gb / dg * 1e6, wg, gb / wg * 1e6))
    del x, w, go
print("sum %.1f us" % tot)
